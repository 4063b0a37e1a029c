#!/bin/bash
# bench.py at N GPUs of one node (weak cfg2 + strong cfg3), as the driver launches it.
N=$1
O=gpurun_out
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_cfg2_n$N.json 2> $O/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload cfg3 --steps 5 --warmup 3 > $O/bench_cfg3_n$N.json 2>> $O/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --workload cfg4 --steps 5 --warmup 3 > $O/bench_cfg4_n$N.json 2>> $O/bench_n$N.err
for f in cfg2 cfg3 cfg4; do python - $O/bench_${f}_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["n_gpus"], d["scaling"], "value=%.4g ms=%.2f e2e=%.4g"%(d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -2 $O/bench_n$N.err
