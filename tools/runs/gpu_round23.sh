#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/ab2.log
timeout 600 python -m pytest tests/test_gpu_attention.py -x -q -m gpu > gpurun_out/tests23.log 2>&1
echo "attention tests exit $?" >> gpurun_out/summary.txt
tail -12 gpurun_out/tests23.log
timeout 600 python -m pytest tests/test_gpu_engine.py tests/test_gpu_elementwise.py -x -q -m gpu > gpurun_out/tests23b.log 2>&1
echo "engine tests exit $?" >> gpurun_out/summary.txt
tail -4 gpurun_out/tests23b.log
ab() { echo "== $1" >> gpurun_out/ab2.log; shift; env "$@" >> gpurun_out/ab2.log 2>&1; }
one() { python bench.py --model $1 --steps $2 --warmup 3 --no_e2e 2>&1 | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['ms_per_step'], r['clocks']['sm_mhz'], r['gpu_launches'])"; }
export -f one
ab "vitl fused attn" bash -c 'one vitl 10'
ab "vitl unfused attn" B200_FUSED_ATTN=0 bash -c 'one vitl 10'
ab "10b fused attn" bash -c 'one vit10b 3'
ab "10b unfused attn" B200_FUSED_ATTN=0 bash -c 'one vit10b 3'
cat gpurun_out/ab2.log; cat gpurun_out/summary.txt
