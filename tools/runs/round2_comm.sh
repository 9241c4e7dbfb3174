#!/bin/bash
# Round 2, multi-GPU call: collective roofline fractions + comm-stream priority A/B + exposed-comm readings.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/runs/round2_comm.sh 8'
N=${1:-8}
mkdir -p gpurun_out; L=gpurun_out/round2_comm.log; : > $L
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((20000 + RANDOM % 20000)) "$@"; }
echo "== custom AG / RS vs NCCL" >> $L
timeout 300 bash -c "$(declare -f run); N=$N; run tools/bench_comm.py --out gpurun_out/bench_comm_n$N.json" 2>&1 | grep '^{' >> $L
for prio in 0 -1; do
  echo "== ViT-10B, B200_COMM_PRIORITY=$prio" >> $L
  B200_COMM_PRIORITY=$prio timeout 400 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps 4 --warmup 3 --no_e2e --no_full_ckpt_probe" 2>&1 | tail -1 |
    python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['ms_per_step'], r['value'], r['exposed_comm_ms_per_step'], r['clocks'])" >> $L 2>&1
done
cat $L
