#!/bin/bash
# solver launch-configuration sweep (bench experiments): prints solve ms per config
run() {
  echo "== $*"
  env "$@" timeout 200 python bench.py --no-cpu --steps 4 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.1f M/s  step %.2f ms  solve %.2f ms' % (d['value'] / 1e6, d['ms_per_step'], d['stage_ms']['solve']))"
}
run MOT3D_X=0
run MOT3D_SOLVE_THREADS=64
run MOT3D_SOLVE_THREADS=32
run MOT3D_SOLVE_THREADS=64 MOT3D_SOLVE_SMEM=13312 MOT3D_SOLVE_RESIDENT=128
run MOT3D_SOLVE_THREADS=32 MOT3D_SOLVE_SMEM=13312 MOT3D_SOLVE_RESIDENT=128
run MOT3D_SOLVE_THREADS=32 MOT3D_SOLVE_SMEM=6144 MOT3D_SOLVE_RESIDENT=64
run MOT3D_SOLVE_THREADS=64 MOT3D_SOLVE_SMEM=18432 MOT3D_SOLVE_RESIDENT=160
run MOT3D_SOLVE_THREADS=256
