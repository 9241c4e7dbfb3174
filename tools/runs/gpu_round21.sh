#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
ab() { echo "== $1" >> gpurun_out/ab.log; shift; env "$@" >> gpurun_out/ab.log 2>&1; }
one() { python bench.py --model $1 --steps $2 --warmup 3 --no_e2e 2>&1 | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['ms_per_step'], r['clocks']['sm_mhz'], r['gpu_launches'])"; }
export -f one
ab "vitl new" bash -c 'one vitl 10'
ab "vitl old LN" B200_LN_SMALL=0 bash -c 'one vitl 10'
ab "vitl fused act" B200_FUSE_ACT_MIN_K=0 bash -c 'one vitl 10'
ab "vitl new again" bash -c 'one vitl 10'
ab "10b CLC on" bash -c 'one vit10b 3'
ab "10b CLC off" B200_GEMM_CLC=0 bash -c 'one vit10b 3'
ab "10b CLC on again" bash -c 'one vit10b 3'
cat gpurun_out/ab.log
