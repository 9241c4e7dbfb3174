#!/bin/bash
# Final 1-GPU validation of the default path + ncu capture of every hot kernel variant + compute-sanitizer memcheck
mkdir -p gpurun_out; L=gpurun_out/r2_final1.log; : > $L
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 >> $L
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $L
echo "== bench vit10b, 4 blocks, 1 GPU" >> $L
timeout 300 python bench.py --num_blocks 4 --steps 4 --warmup 3 2>&1 | tail -1 | cut -c1-1700 >> $L
echo "== ncu" >> $L
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|ln_|attn_|adamw|p2p_all_gather|gelu_fwd' -s 45 -c 15 -f -o gpurun_out/prof_r2 python tools/prof_kernels.py 2>&1 | grep -E "launches per pass|Report|Error|error" >> $L
echo "== compute-sanitizer memcheck (elementwise + attention kernels)" >> $L
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 1 --log-file gpurun_out/sanitize_memcheck.log python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_attention.py -x -q -m gpu -k "layernorm or softmax or persistent_attention_forward or cross_entropy or adamw" 2>&1 | tail -2 >> $L
tail -3 gpurun_out/sanitize_memcheck.log >> $L
cat $L
