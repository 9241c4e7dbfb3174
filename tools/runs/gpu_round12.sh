#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/wgrad.log
for c in wgrad_fc2 wgrad_fc1 dgrad_fc2 dgrad_fc2_dgelu; do
  export EXP_CASE=$c
  timeout 120 python tools/exp_wgrad.py >> gpurun_out/wgrad.log 2>&1
  timeout 200 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__inst_executed.sum \
     -k regex:gemm_bf16 -s 4 -c 1 --csv python tools/exp_wgrad.py 2>/dev/null | grep -E "dram__bytes|hit_rate|duration|tensor|bank|inst_exec" | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr '\n' ';' >> gpurun_out/wgrad.log
  echo >> gpurun_out/wgrad.log
done
cat gpurun_out/wgrad.log
