#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_persist' -s 4 -c 2 -f -o gpurun_out/prof_bwd python tools/prof_attn_ln.py > gpurun_out/r2_ncu_bwd.log 2>&1
tail -5 gpurun_out/r2_ncu_bwd.log
