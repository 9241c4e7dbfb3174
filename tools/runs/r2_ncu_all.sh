#!/bin/bash
# 1 GPU: ncu --set full of every hot kernel variant (one pass of tools/prof_kernels.py) + compute-sanitizer logs
mkdir -p gpurun_out; L=gpurun_out/r2_ncu_all.log; : > $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|ln_|attn_|adamw|p2p_all_gather|gelu_fwd' -s 45 -c 15 -f -o gpurun_out/prof_r2 python tools/prof_kernels.py >> $L 2>&1
ls -la gpurun_out/prof_r2.ncu-rep >> $L
echo "== compute-sanitizer memcheck" >> $L
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/sanitize_memcheck.log python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_attention.py tests/test_gpu_gemm.py -x -q -m gpu -k "not real_vit10b" 2>&1 | tail -3 >> $L
tail -4 gpurun_out/sanitize_memcheck.log >> $L
echo "== compute-sanitizer racecheck (attention + LayerNorm stream kernels)" >> $L
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 1 --log-file gpurun_out/sanitize_racecheck.log python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_attention.py -x -q -m gpu -k "layernorm or persistent or fused_attention_forward" 2>&1 | tail -3 >> $L
tail -4 gpurun_out/sanitize_racecheck.log >> $L
tail -30 $L
