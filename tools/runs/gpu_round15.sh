#!/bin/bash
# CLI end-to-end on GPUs: train, checkpoint, resume, eval, consolidate (reference flags), 1 and 2 GPUs
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
M="--fake_data --image_size 224 --patch_size 14 --embed_dim 640 --num_heads 4 --num_blocks 4 --mlp_ratio 4.0 --num_classes 1000"
rm -rf /tmp/ck1 /tmp/ck2
timeout 600 python run_vit_training.py $M --batch_size 64 --num_epochs 2 --max_steps 6 --log_step_interval 2 --warmup_steps 4 \
   --ckpt_dir /tmp/ck2 --ckpt_epoch_interval 1 --test_epoch_interval 1 --shard_on_cpu > gpurun_out/cli_2gpu.log 2>&1
echo "cli 2gpu exit $?" >> gpurun_out/summary.txt
grep -E "epoch|accuracy|checkpoint|completed|sharded" gpurun_out/cli_2gpu.log | tail -16
ls -la /tmp/ck2 >> gpurun_out/summary.txt
timeout 600 python run_vit_training.py $M --batch_size 64 --num_epochs 3 --resume_epoch 2 --max_steps 4 --log_step_interval 2 --warmup_steps 4 \
   --ckpt_dir /tmp/ck2 --ckpt_epoch_interval 1 --test_epoch_interval 1 --flatten_parameters=0 > gpurun_out/cli_2gpu_resume.log 2>&1 || \
timeout 600 python run_vit_training.py $M --batch_size 64 --num_epochs 3 --resume_epoch 2 --max_steps 4 --log_step_interval 2 --warmup_steps 4 \
   --ckpt_dir /tmp/ck2 --ckpt_epoch_interval 1 --test_epoch_interval 1 > gpurun_out/cli_2gpu_resume.log 2>&1
echo "cli resume exit $?" >> gpurun_out/summary.txt
grep -E "resumed|epoch 3|accuracy|completed" gpurun_out/cli_2gpu_resume.log | tail -8
timeout 300 python -m vit_10b_fsdp_example_b200.consolidate_sharded_ckpts --ckpt_prefix /tmp/ck2/epoch_2 --save_path /tmp/ck2/epoch_2_full.pth > gpurun_out/consolidate.log 2>&1
echo "consolidate exit $?" >> gpurun_out/summary.txt
tail -2 gpurun_out/consolidate.log
# other modes on 2 GPUs: DDP comparison mode, ZeRO-2, no grad ckpt, flatten, clip off (fused AdamW in reduce-scatter)
for extra in "--run_without_fsdp" "--no_reshard_after_forward --no_grad_ckpt" "--flatten_parameters" "--clip_grad_norm 0"; do
  timeout 300 python run_vit_training.py $M --batch_size 64 --num_epochs 1 --max_steps 4 --log_step_interval 2 --warmup_steps 2 --ckpt_dir /tmp/ck3 $extra > gpurun_out/cli_mode.log 2>&1
  echo "mode [$extra] exit $? : $(grep -E 'step 4' gpurun_out/cli_mode.log | tail -1 | cut -c1-110)" >> gpurun_out/summary.txt
done
# single GPU + nccl backend variant
CUDA_VISIBLE_DEVICES=0 timeout 300 python run_vit_training.py $M --batch_size 32 --num_epochs 1 --max_steps 4 --log_step_interval 2 --warmup_steps 2 --ckpt_dir /tmp/ck1 > gpurun_out/cli_1gpu.log 2>&1
echo "cli 1gpu exit $? : $(grep -E 'step 4' gpurun_out/cli_1gpu.log | tail -1 | cut -c1-110)" >> gpurun_out/summary.txt
timeout 300 python run_vit_training.py $M --batch_size 64 --num_epochs 1 --max_steps 4 --log_step_interval 2 --warmup_steps 2 --ckpt_dir /tmp/ck4 --backend nccl > gpurun_out/cli_nccl.log 2>&1
echo "cli nccl-backend exit $? : $(grep -E 'step 4' gpurun_out/cli_nccl.log | tail -1 | cut -c1-110)" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
