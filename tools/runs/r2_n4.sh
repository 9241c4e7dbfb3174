#!/bin/bash
# Round 2, 4-GPU call: BASELINE config 5 (ViT-10B, --shard_on_cpu, checkpoint / resume / consolidate / re-shard through the
# CLI), sm100-vs-NCCL engine trajectories at W=4, collective bandwidths at W=4, ViT-L (config 2) both arms.
mkdir -p gpurun_out; L=gpurun_out/r2_n4.log; : > $L
run4() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((20000 + RANDOM % 20000)) "$@"; }
free -g | head -2 >> $L; df -h /tmp | tail -1 >> $L
echo "== config 5" >> $L
timeout 900 bash tools/runs/r2_config5.sh 4 > /dev/null 2>&1; cat gpurun_out/r2_config5.log >> $L
echo "== engine trajectories sm100 vs NCCL at W=4 (random data)" >> $L
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "matches_nccl" 2>&1 | tail -4 >> $L
echo "== collectives custom vs NCCL, W=4" >> $L
timeout 200 bash -c "$(declare -f run4); run4 tools/bench_comm.py --out gpurun_out/bench_comm_n4.json" 2>&1 | grep '^{' | cut -c1-420 >> $L
echo "== ViT-L N=4 ours" >> $L
timeout 300 bash -c "$(declare -f run4); run4 bench.py --gpus 4 --model vitl --steps 20 --warmup 5 --no_full_ckpt_probe" 2>&1 | tail -1 | cut -c1-1800 >> $L
echo "== ViT-L N=4 reference" >> $L
timeout 300 bash -c "$(declare -f run4); run4 bench.py --gpus 4 --model vitl --steps 20 --warmup 5 --impl reference" 2>&1 | tail -1 | cut -c1-1200 >> $L
cat $L
