#!/bin/bash
# Round-end measurement set (1 GPU).  Outputs under gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err
python bench.py --impl reference --steps 2 --warmup 0 > $O/bench_cfg2_reference.json 2>> $O/bench_cfg2.err
python bench.py --workload cfg3 --steps 3 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err
python bench.py --workload cfg4 --steps 5 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python bench.py --workload cfg1 --steps 20 --no-cpu-baseline > $O/bench_cfg1.json 2> $O/bench_cfg1.err
for n in 2 4 10 32; do
  python bench.py --workload cfg5 --spks 96 --ode-steps $n --steps 3 --no-cpu-baseline > $O/bench_cfg5_spks_n$n.json 2> $O/bench_cfg5.err
done
python bench.py --workload cfg5 --ode-steps 10 --steps 3 --no-cpu-baseline > $O/bench_cfg5_nospk_n10.json 2>> $O/bench_cfg5.err
python bench.py --precision fp32 --steps 1 --no-cpu-baseline > $O/bench_cfg2_fp32.json 2> $O/bench_fp32.err
tail -n 3 $O/pytest_gpu.log
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "value=%.4g"%d.get('value',0), "ms=%.2f"%d.get('ms_per_step',0), "frac=%.3f"%(d.get('roofline',{}) or {}).get('frac',0), "e2e=%.4g"%(d.get('e2e',{}) or {}).get('value',0), d.get('clocks'))
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
done
