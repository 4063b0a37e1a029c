#!/bin/bash
# Round-2 measurement set (1 GPU): GPU tests, default bench, per-launch-class A/B timeline, ncu launch list, ncu --set full of the
# dominant kernels.  Outputs under gpurun_out/ with the tag given as $1.
set -u
T=${1:-r2b}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${T}_pytest_gpu.log
tail -n 3 $O/${T}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err
tail -c 600 $O/${T}_bench_cfg2.err
AB_SPLIT_RES=1 python tools/ab_timeline.py "" > $O/${T}_timeline.log 2>&1
head -n 12 $O/${T}_timeline.log
if [ "${2:-}" = "ncu" ]; then
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
      --log-file $O/${T}_launches.csv python tools/profile_solve.py cfg2 1 > $O/${T}_ncu.log 2>&1
  tail -n 2 $O/${T}_ncu.log
fi
