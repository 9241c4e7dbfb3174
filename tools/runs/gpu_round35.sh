#!/bin/bash
# single-shot bring-up of the fused attention backward (unverified kernel): numerics vs the fp32 reference
mkdir -p gpurun_out
B200_TEST_UNVERIFIED=1 timeout 40 python -m pytest tests/test_gpu_attention.py -k fused_attention_backward -x -q 2>&1 | tail -25 | tee gpurun_out/attn_bwd_bringup.log
