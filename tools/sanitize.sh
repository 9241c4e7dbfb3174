#!/bin/bash
# compute-sanitizer passes over the single-GPU kernel tests.  Run on a GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'      (or racecheck / synccheck / initcheck)
# The multi-GPU flag protocols are covered differently: every spin is bounded and traps (ptx.cuh, comm.cu), sequence
# numbers only grow, and tests/test_gpu_multi.py reuses the flags for >1000 iterations.
tool=${1:-memcheck}
mkdir -p gpurun_out
timeout 1400 compute-sanitizer --tool "$tool" --error-exitcode 1 --log-file gpurun_out/sanitize_$tool.log \
    python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_attention.py tests/test_gpu_gemm.py -x -q -m gpu \
    -k "not sustained" 2>&1 | tail -5
echo "exit: $?"; tail -5 gpurun_out/sanitize_$tool.log
