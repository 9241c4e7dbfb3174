#!/bin/bash
mkdir -p gpurun_out
timeout 14 python tools/exp_attn_bwd.py 2>&1 | tail -4 | tee gpurun_out/attn_bwd_timing.log
