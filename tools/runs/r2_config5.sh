#!/bin/bash
# BASELINE config 5: --shard_on_cpu init + per-rank sharded checkpoint save / resume / consolidate / re-shard, through the
# product CLI (run_vit_training.py), ViT-10B on N GPUs.     gpurun --gpus 4 --timeout 1500 -- 'bash tools/runs/r2_config5.sh 4'
N=${1:-4}; HALF=$((N / 2)); [ $HALF -lt 1 ] && HALF=1
EXTRA=${C5_EXTRA:-}            # e.g. tiny dims for a CPU dry run
CK=${C5_DIR:-/dev/shm/ck5}; OUT=${C5_OUT:-gpurun_out}
mkdir -p $OUT; L=$OUT/r2_config5.log; : > $L; rm -rf $CK
BS=$((128 * N))
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((20000 + RANDOM % 20000)) run_vit_training.py --fake_data --log_step_interval 1 --ckpt_dir $CK --test_epoch_interval 100 --warmup_steps 20 $EXTRA "$@"; }
t0=$(date +%s)
echo "== 1. fresh run: --shard_on_cpu, 1 epoch capped at 3 steps, every rank saves its shards  (N=$N, batch $BS)" >> $L
tr $N --shard_on_cpu --batch_size $BS --num_epochs 1 --max_steps 3 --bench_json $OUT/c5_run1.jsonl 2>&1 | grep -E "^epoch|built ViT block 31|sharded\)|checkpoint saved|accuracy|training completed|Error|error|Traceback" | tail -12 >> $L
t1=$(date +%s); echo "   wall $((t1 - t0)) s; files:" >> $L; ls -la $CK | tail -n +2 | awk '{print "   ", $5, $9}' >> $L
echo "== 2. resume from epoch 1 (--resume_epoch 1), epoch 2 capped at 3 steps" >> $L
tr $N --shard_on_cpu --batch_size $BS --num_epochs 2 --resume_epoch 1 --max_steps 3 --bench_json $OUT/c5_run2.jsonl 2>&1 | grep -E "^epoch|resumed from|checkpoint saved|accuracy|training completed|Error|error|Traceback" | tail -12 >> $L
t2=$(date +%s); echo "   wall $((t2 - t1)) s" >> $L
echo "== 3. consolidate the epoch-2 shards into one timm-style state_dict" >> $L
python -m vit_10b_fsdp_example_b200.consolidate_sharded_ckpts --ckpt_prefix $CK/epoch_2 --save_path $CK/epoch_2_full.pth 2>&1 | tail -2 >> $L
t3=$(date +%s); echo "   wall $((t3 - t2)) s; $(ls -la $CK/epoch_2_full.pth | awk '{print $5}') bytes" >> $L
echo "== 4. continue on $HALF GPU(s) from the consolidated checkpoint (--init_from_full_ckpt), 2 steps" >> $L
tr $HALF --batch_size $((128 * HALF)) --num_epochs 1 --max_steps 2 --init_from_full_ckpt $CK/epoch_2_full.pth --bench_json $OUT/c5_run3.jsonl 2>&1 | grep -E "^epoch|initialised from|training completed|Error|error|Traceback" | tail -6 >> $L
t4=$(date +%s); echo "   wall $((t4 - t3)) s" >> $L
echo "== loss trajectories (smoothed, as logged)" >> $L
for f in c5_run1 c5_run2 c5_run3; do echo "   $f: $(python -c "
import json,sys
print([round(json.loads(l)['loss'],4) for l in open('$OUT/$f.jsonl')])" 2>&1)" >> $L; done
rm -rf $CK
cat $L
