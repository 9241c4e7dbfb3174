#!/bin/bash
# Round 2, call 3 (1 GPU): first hardware run of the persistent / long-sequence attention kernels + timings.
#   gpurun --timeout 1200 -- 'bash tools/runs/r2_attn_bringup.sh'
mkdir -p gpurun_out; L=gpurun_out/r2_attn_bringup.log; : > $L
export B200_TEST_UNVERIFIED=1
for t in test_persistent_attention_forward test_fused_attention_long_sequence test_persistent_attention_backward; do
  echo "== $t" >> $L
  timeout 240 python -m pytest tests/test_gpu_attention.py -q -m gpu -k $t 2>&1 | tail -30 >> $L
done
echo "== timings (persist fwd only)" >> $L
timeout 300 python tools/exp_attn2.py --shapes 10b,l --skip persist_bwd --json gpurun_out/attn_times_a.json 2>&1 | tail -8 >> $L
echo "== timings (with persist bwd)" >> $L
timeout 300 python tools/exp_attn2.py --shapes 10b,l --json gpurun_out/attn_times_b.json 2>&1 | tail -8 >> $L
echo "== timings 336" >> $L
timeout 300 python tools/exp_attn2.py --shapes 336 --skip persist_bwd --json gpurun_out/attn_times_336.json 2>&1 | tail -8 >> $L
cat $L
echo "== overlap experiment (L2 hint on)" >> $L
timeout 300 python tools/exp_overlap.py --json gpurun_out/overlap_hint1.json 2>&1 | grep -v "^{" | tail -12 >> $L
echo "== overlap experiment (L2 hint off)" >> $L
B200_COMM_L2_HINT=0 timeout 300 python tools/exp_overlap.py --json gpurun_out/overlap_hint0.json 2>&1 | grep -v "^{" | tail -12 >> $L
cat $L
