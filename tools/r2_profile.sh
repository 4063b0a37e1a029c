#!/bin/bash
# ncu --set full captures for profiles/ (1 GPU): the first 19 GEMM launches of an estimator evaluation (full-resolution stage,
# strided conv, first half-resolution stage), attention (cfg2 full / half, cfg4), the norm / apply passes.  $1 = tag.
# Reports must stay under gpurun's 64 MiB pull limit: sources are imported for the attention kernel only.
set -u
T=${1:-r2c}
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none"
$NCU -k regex:gemm_tc -c 19 -o $O/${T}_gemm -f python tools/profile_solve.py cfg2 1 > $O/${T}_ncu_gemm.log 2>&1; tail -n 1 $O/${T}_ncu_gemm.log
$NCU --import-source on -k regex:attn_tc -c 3 -o $O/${T}_attn_cfg2 -f python tools/profile_solve.py cfg2 1 > $O/${T}_ncu_attn.log 2>&1; tail -n 1 $O/${T}_ncu_attn.log
$NCU -k regex:attn_tc -c 1 -o $O/${T}_attn_cfg4 -f python tools/profile_solve.py cfg4 1 > $O/${T}_ncu_attn4.log 2>&1; tail -n 1 $O/${T}_ncu_attn4.log
$NCU -k "regex:layernorm|gn_apply" -c 6 -o $O/${T}_norm -f python tools/profile_solve.py cfg2 1 > $O/${T}_ncu_norm.log 2>&1; tail -n 1 $O/${T}_ncu_norm.log
ls -la $O/${T}_*.ncu-rep
du -sm $O
