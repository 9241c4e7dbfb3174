#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16|ln_bwd|ln_fwd|adamw_split|softmax_fwd" -s 12 -c 7 -o gpurun_out/kernels_prof python tools/prof_kernels.py > gpurun_out/ncu3.log 2>&1
echo "ncu exit $?"
tail -3 gpurun_out/ncu3.log
