#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/r2_ag4.log; : > $L
run4() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((20000 + RANDOM % 20000)) "$@"; }
timeout 200 bash -c "$(declare -f run4); run4 tools/bench_comm.py --out gpurun_out/bench_comm_n4c.json" 2>&1 | grep '^{' | cut -c1-420 >> $L
timeout 300 bash -c "$(declare -f run4); run4 tools/step_timeline.py --blocks 8 --steps 8 --out gpurun_out/tl4c_w4.json" 2>&1 | grep -E "^\[rank|p2p_all_gather|reduce_scatter|Error|error" | head -14 >> $L
cat $L
