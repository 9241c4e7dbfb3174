#!/bin/bash
# Round 2, call 1 (2 GPUs): attribute the N=1 -> N=2 step-time delta at equal activation policy.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/runs/r2_timeline.sh'
mkdir -p gpurun_out; L=gpurun_out/r2_timeline.log; : > $L
BL=${BL:-8}
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((20000 + RANDOM % 20000)) "$@"; }
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.limit --format=csv >> $L 2>&1
echo "== W=1 on GPU0 and GPU1 concurrently (independent jobs)" >> $L
( timeout 300 python tools/step_timeline.py --blocks $BL --device_index 0 --out gpurun_out/tl_w1_gpu0.json > gpurun_out/tl_a.log 2>&1 ) &
( timeout 300 python tools/step_timeline.py --blocks $BL --device_index 1 --out gpurun_out/tl_w1_gpu1.json > gpurun_out/tl_b.log 2>&1 ) &
wait
grep -E "^\[rank|^   r" gpurun_out/tl_a.log >> $L; grep -E "^\[rank|^   r" gpurun_out/tl_b.log >> $L
tail -3 gpurun_out/tl_a.log | grep -v "^   r" >> $L
echo "== W=1 on GPU0 alone" >> $L
timeout 300 python tools/step_timeline.py --blocks $BL --device_index 0 --out gpurun_out/tl_w1_alone.json 2>&1 | grep -E "^\[rank" >> $L
echo "== W=2 default" >> $L
timeout 300 bash -c "$(declare -f run2); run2 tools/step_timeline.py --blocks $BL --out gpurun_out/tl_w2.json" 2>&1 | grep -E "^\[rank|^   r|Error|error" >> $L
echo "== W=2 B200_COMM_LOCAL=1 (same kernels, no NVLink, no barriers; numerics invalid)" >> $L
B200_COMM_LOCAL=1 B200_FUSE_AG=0 timeout 300 bash -c "$(declare -f run2); run2 tools/step_timeline.py --blocks $BL --out gpurun_out/tl_w2_local.json" 2>&1 | grep -E "^\[rank|Error|error" >> $L
echo "== W=2 B200_FUSE_AG=0" >> $L
B200_FUSE_AG=0 timeout 300 bash -c "$(declare -f run2); run2 tools/step_timeline.py --blocks $BL --out gpurun_out/tl_w2_nofuse.json" 2>&1 | grep -E "^\[rank|Error|error" >> $L
echo "== W=2 B200_COMM_PRIORITY=-1" >> $L
B200_COMM_PRIORITY=-1 timeout 300 bash -c "$(declare -f run2); run2 tools/step_timeline.py --blocks $BL --out gpurun_out/tl_w2_prio.json" 2>&1 | grep -E "^\[rank|Error|error" >> $L
echo "== collectives: custom vs NCCL at W=2" >> $L
timeout 240 bash -c "$(declare -f run2); run2 tools/bench_comm.py --out gpurun_out/bench_comm_n2.json" 2>&1 | grep '^{' >> $L
cat $L
