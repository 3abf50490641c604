for cfg in "128 128" "32 128" "32 96" "32 64" "32 48" "64 64"; do
  set -- $cfg
  MOT3D_SOLVE_THREADS=$1 MOT3D_SOLVE_RECCAP=$2 timeout 300 python bench.py --steps 3 --warmup 3 --seqs 60 --no-cpu > gpurun_out/b.log 2>&1
  python -c "
import json
d=json.loads(open('gpurun_out/b.log').read().strip().splitlines()[-1]); print('$cfg', round(d['value']/1e6,1), 'Mdet/s solve_ms', round(d['stage_ms']['solve'],2), {k:int(v/1e5)/10 for k,v in d['roofline']['phase_cycles_per_graph_mean'].items()})
" 2>&1 | tail -1
done
