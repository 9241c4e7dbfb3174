#!/bin/bash
mkdir -p gpurun_out
timeout 240 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 4 -f -o gpurun_out/attn_prof python tools/prof_attn.py > gpurun_out/attn_prof.log 2>&1
tail -3 gpurun_out/attn_prof.log
