#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_step_ab.log; : > $L
run() { echo "== $1" >> $L; shift; env "$@" timeout 400 python bench.py --num_blocks 8 --steps 6 --warmup 3 --no_e2e --no_full_ckpt_probe 2>&1 | tail -1 | python -c "
import sys,json
r=json.loads(sys.stdin.read()); print(round(r['ms_per_step'],1), 'ms', round(r['value'],1),'img/s', r['config']['activation_ckpt'][:90], r['clocks']['sm_mhz'], 'peak', round(r['peak_mem_gb'],1))" >> $L 2>&1; }
run "default (flash attention pair, LN stream)" X=1
run "un-fused attention" B200_FUSED_ATTN_BWD=0
run "old LN bwd" B200_LN_STREAM=0
run "keep 0 (full recompute) default" X=1 B200_DUMMY=1
cat $L
