#!/bin/bash
# Repeats short benches and prints device-timed vs host-buffer ms per step: a run whose device time exceeds the other by several
# ms had a stall inside the timed region (see ptx.cuh CFM_SUSPEND_HINT_NS).  $1 = repetitions.
for i in $(seq 1 ${1:-4}); do
python bench.py --workload cfg4 --steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4  ', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], d['run']['step_ms'])"
python bench.py --workload cfg5 --spks 96 --ode-steps 4 --steps 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5n4', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], d['run']['step_ms'])"
python bench.py --steps 10 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2  ', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), d['clocks']['sm_mhz'], d['run']['step_ms'])"
done
