#!/bin/bash
# Round 2, 8-GPU call: collectives at W=8, headline config short run, BASELINE config 4 (ViT-10B at 336 px) both arms.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/runs/r2_n8.sh'
mkdir -p gpurun_out; L=gpurun_out/r2_n8.log; : > $L
run8() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((20000 + RANDOM % 20000)) "$@"; }
echo "== collectives at W=8 (and 4, 2)" >> $L
timeout 240 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "symmetric_memory_collectives" 2>&1 | tail -6 >> $L
echo "== ViT-10B 224, N=8, ours (short)" >> $L
timeout 300 bash -c "$(declare -f run8); run8 bench.py --gpus 8 --steps 4 --warmup 3 --no_e2e" 2>&1 | tail -1 | cut -c1-2500 >> $L
echo "== ViT-10B 336, N=8, ours" >> $L
timeout 400 bash -c "$(declare -f run8); run8 bench.py --gpus 8 --model vit10b_336 --steps 4 --warmup 3 --no_full_ckpt_probe" 2>&1 | tail -1 | cut -c1-2500 >> $L
echo "== ViT-10B 336, N=8, reference" >> $L
timeout 500 bash -c "$(declare -f run8); run8 bench.py --gpus 8 --model vit10b_336 --steps 4 --warmup 3 --impl reference" 2>&1 | tail -1 | cut -c1-2000 >> $L
cat $L
