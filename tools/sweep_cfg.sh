for cfg in "256 320" "128 320" "128 160" "192 256" "128 224" "64 160"; do
  set -- $cfg
  MOT3D_SOLVE_THREADS=$1 MOT3D_SOLVE_RECCAP=$2 timeout 300 python bench.py --steps 3 --warmup 3 --seqs 60 --no-cpu > gpurun_out/b.log 2>&1
  python -c "
import json
d=json.loads(open('gpurun_out/b.log').read().strip().splitlines()[-1]); print('$cfg', round(d['value']/1e6,1), 'Mdet/s solve_ms', round(d['stage_ms']['solve'],2), {k:int(v/1e5)/10 for k,v in d['roofline']['phase_cycles_per_graph_mean'].items()})
" 2>&1 | tail -1
done
