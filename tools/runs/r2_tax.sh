#!/bin/bash
# Round 2, call 4 (2 GPUs): what part of the W=1 -> W=2 slowdown needs NVLink traffic at all?
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/runs/r2_tax.sh'
mkdir -p gpurun_out; L=gpurun_out/r2_tax.log; : > $L
BL=${BL:-8}
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((20000 + RANDOM % 20000)) "$@"; }
smi() { nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 100 > $1 & echo $!; }
summ() { python - "$1" <<'P'
import sys,statistics as st
rows=[l.split(',') for l in open(sys.argv[1]) if l.count(',')>=3]
for g in ('0','1'):
    c=[float(r[1].split()[0]) for r in rows if r[0].strip()==g and float(r[2].split()[0])>600]
    p=[float(r[2].split()[0]) for r in rows if r[0].strip()==g and float(r[2].split()[0])>600]
    if c: print(f"   gpu{g}: under load n={len(c)} sm_mhz median {st.median(c):.0f} mean {st.mean(c):.0f}  power mean {st.mean(p):.0f} W")
P
}
one() { # name, env..., then timeline args
  name=$1; shift
  echo "== $name" >> $L
  pid=$(smi gpurun_out/clk_$name.csv)
  env "$@" timeout 300 bash -c "$(declare -f run2); run2 tools/step_timeline.py --blocks $BL --steps 12 --out gpurun_out/tl3_$name.json \$TLARGS" 2>&1 | grep -E "^\[rank|Error|error|timeout" >> $L
  kill $pid; summ gpurun_out/clk_$name.csv >> $L
}
echo "== W=1 alone on GPU0 (12 timed steps)" >> $L
pid=$(smi gpurun_out/clk_w1.csv)
timeout 300 python tools/step_timeline.py --blocks $BL --steps 12 --device_index 0 --out gpurun_out/tl3_w1.json 2>&1 | grep -E "^\[rank" >> $L
kill $pid; summ gpurun_out/clk_w1.csv >> $L
one w2_default B200_X=1
one w2_local B200_COMM_LOCAL=1 B200_FUSE_AG=0
one w2_nofuse B200_FUSE_AG=0
one w2_nccl TLARGS="--backend torchdist"
one w2_ctas8 B200_COMM_CTAS=8 B200_FUSE_AG=0
cat $L
