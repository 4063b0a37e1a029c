#!/bin/bash
# Round-2 final measurement set (1 GPU).  Outputs under gpurun_out/ with tag $1.
set -u
T=${1:-r2e}
O=gpurun_out
mkdir -p $O
python __graft_entry__.py --smoke 2>&1 | tail -4
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${T}_pytest_gpu.log
tail -n 3 $O/${T}_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_cfg2_reference.json 2>> $O/${T}_bench_cfg2.err
python bench.py --workload cfg3 --steps 3 --no-cpu-baseline > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
python bench.py --workload cfg4 --steps 5 --no-cpu-baseline > $O/${T}_bench_cfg4.json 2> $O/${T}_bench_cfg4.err
python bench.py --workload cfg1 --steps 20 --no-cpu-baseline > $O/${T}_bench_cfg1.json 2> $O/${T}_bench_cfg1.err
for n in 2 4 10 32; do
  python bench.py --workload cfg5 --spks 96 --ode-steps $n --steps 3 --no-cpu-baseline > $O/${T}_bench_cfg5_spks_n$n.json 2> $O/${T}_bench_cfg5.err
done
python bench.py --precision fp32_tc --steps 5 --no-cpu-baseline --no-gpu-baseline > $O/${T}_bench_cfg2_fp32_tc.json 2> $O/${T}_bench_fp32tc.err
python bench.py --precision fp32 --steps 1 --no-cpu-baseline --no-gpu-baseline > $O/${T}_bench_cfg2_fp32.json 2> $O/${T}_bench_fp32.err
python bench.py --solver midpoint --ode-steps 4 --steps 5 --no-cpu-baseline --no-gpu-baseline > $O/${T}_bench_cfg2_midpoint4.json 2>> $O/${T}_bench_fp32.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1400 --csv \
    --log-file $O/${T}_launches.csv python tools/profile_solve.py cfg2 1 > $O/${T}_ncu.log 2>&1
tail -n 1 $O/${T}_ncu.log
for f in $O/${T}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    g=(d.get('roofline',{}) or {}).get('gemm',{}) or {}
    print(sys.argv[1].split('/')[-1], "value=%.4g"%d.get('value',0), "ms=%.2f"%d.get('ms_per_step',0), "frac=%.3f"%(d.get('roofline',{}) or {}).get('frac',0), "gemm_frac=%.3f"%g.get('frac_in_graph_est',0), "e2e=%.4g"%(d.get('e2e',{}) or {}).get('value',0), (d.get('clocks') or {}).get('sm_mhz'))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
du -sm $O
