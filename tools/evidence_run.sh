#!/bin/bash
# The evidence run of a round, on the GPU box (one gpurun call): GPU tests, the default bench line, the ncu launch list
# of the bench command and one ncu --set full capture of the hot path. Every step under its own hard time limit.
# SKIP_TESTS=1 / SKIP_FULL=1 leave a step out.
#   gpurun --timeout 900 -- 'bash tools/evidence_run.sh r02e'      then here: python tools/refresh_profiles.py r02e
TAG=${1:-evidence}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$SKIP_TESTS" ]; then
    timeout -s KILL 240 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > $OUT/${TAG}_pytest.log
    cat $OUT/${TAG}_pytest.log
fi
timeout -s KILL 420 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python tools/bench_summary.py $OUT/${TAG}_bench.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-latency \
    > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list: $(wc -l < $OUT/${TAG}_launches_bench.csv) lines"
# (gpurun brings back at most 64 MiB: 45 launches = the build of the driver + one pipeline pass, no source pages)
if [ -z "$SKIP_FULL" ]; then
    timeout -s KILL 300 ncu --set full --clock-control none -c 45 -f -o $OUT/${TAG}_full \
        python tools/profile_solve.py 1200 > $OUT/${TAG}_full.log 2>&1
    ls -la $OUT/${TAG}_full.ncu-rep
fi
