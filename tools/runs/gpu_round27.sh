#!/bin/bash
# N=4 A/B of the communication knobs on one box
mkdir -p gpurun_out; rm -f gpurun_out/ab4.log
run() { echo "== $1" >> gpurun_out/ab4.log; shift
  P=$((20000 + RANDOM % 20000))
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 4 --steps 3 --warmup 3 --no_e2e 2>&1 | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['ms_per_step'], r['value'], r['clocks']['sm_mhz'])" >> gpurun_out/ab4.log 2>&1; }
run "default" A=1
run "comm_ctas=8" B200_COMM_CTAS=8
run "no AG fusion" B200_FUSE_AG=0
run "p2p reduce-scatter (no NVLS)" B200_NVLS=0
run "default again" A=1
cat gpurun_out/ab4.log
