#!/bin/bash
# Round 2, call 2 (2 GPUs): light-CTA collectives with the in-kernel flag protocol -- correctness, then the same
# timeline A/B as call 1.   gpurun --gpus 2 --timeout 1200 -- 'bash tools/runs/r2_comm_v2.sh'
mkdir -p gpurun_out; L=gpurun_out/r2_comm_v2.log; : > $L
BL=${BL:-8}
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((20000 + RANDOM % 20000)) "$@"; }
echo "== pytest tests/test_gpu_multi.py" >> $L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -25 >> $L
echo "== W=2 default (light collectives, fused AG)" >> $L
timeout 300 bash -c "$(declare -f run2); run2 tools/step_timeline.py --blocks $BL --out gpurun_out/tl2_w2.json" 2>&1 | grep -E "^\[rank|^   r|Error|error|timeout" | head -40 >> $L
echo "== W=2 B200_FUSE_AG=0" >> $L
B200_FUSE_AG=0 timeout 300 bash -c "$(declare -f run2); run2 tools/step_timeline.py --blocks $BL --out gpurun_out/tl2_w2_nofuse.json" 2>&1 | grep -E "^\[rank|^   r|Error|error|timeout" | head -40 >> $L
for c in 16 148; do
echo "== W=2 B200_FUSE_AG=0 B200_COMM_CTAS=$c" >> $L
B200_COMM_CTAS=$c B200_FUSE_AG=0 timeout 300 bash -c "$(declare -f run2); run2 tools/step_timeline.py --blocks $BL --out gpurun_out/tl2_w2_c$c.json" 2>&1 | grep -E "^\[rank|Error|error|timeout" >> $L
done
echo "== W=2 B200_NVLS=0 B200_FUSE_AG=0" >> $L
B200_NVLS=0 B200_FUSE_AG=0 timeout 300 bash -c "$(declare -f run2); run2 tools/step_timeline.py --blocks $BL --out gpurun_out/tl2_w2_p2p.json" 2>&1 | grep -E "^\[rank|Error|error|timeout" >> $L
echo "== collectives: custom vs NCCL at W=2" >> $L
timeout 240 bash -c "$(declare -f run2); run2 tools/bench_comm.py --out gpurun_out/bench_comm2_n2.json" 2>&1 | grep -E '^\{|Error|error' >> $L
cat $L
