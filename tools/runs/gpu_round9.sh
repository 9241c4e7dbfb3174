#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/exp_sustained.py > gpurun_out/sustained.log 2>&1
cat gpurun_out/sustained.log
# cuBLAS DRAM traffic on the same shapes
cat > /tmp/cb.py <<'PY'
import torch
T,D=32768,5120
x=torch.randn(T,4*D,device='cuda').to(torch.bfloat16); w=(torch.randn(D,4*D,device='cuda')*0.02).to(torch.bfloat16)
x1=torch.randn(T,D,device='cuda').to(torch.bfloat16); wq=(torch.randn(3*D,D,device='cuda')*0.02).to(torch.bfloat16)
for _ in range(3):
    torch.nn.functional.linear(x1,wq); torch.nn.functional.linear(x,w)
torch.cuda.synchronize()
PY
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,launch__cluster_size,launch__grid_size,launch__block_size -k regex:"gemm|cutlass|nvjet|sm100" -s 4 -c 2 --csv python /tmp/cb.py 2>/dev/null | grep -E "dram__bytes|hit_rate|duration|cluster|grid_size|block_size" | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | cut -c1-220 > gpurun_out/cublas_dram.log
cat gpurun_out/cublas_dram.log
