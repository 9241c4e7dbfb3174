// Host API of the sm_100a tcgen05 GEMM (see gemm_sm100.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace b200 {

enum GemmAct : int { kActNone = 0, kActGelu = 1, kActDGelu = 2 };

// A strided bf16 matrix with up to two batch dimensions.  `ld` and the batch strides are in elements.
// The *inner* (contiguous) extent and the *outer* extent are implied by the GEMM shape and the major.
struct GemmOperand {
    const void* ptr = nullptr;
    int64_t ld = 0;
    int64_t nb_inner = 1, stride_b_inner = 0;
    int64_t nb_outer = 1, stride_b_outer = 0;
};

struct GemmEpilogue {
    const __nv_bfloat16* bias = nullptr;      // [N]           v += bias[n]
    const __nv_bfloat16* residual = nullptr;  // [M, ld_res]   v += residual[m, n]   (after activation)
    const __nv_bfloat16* aux_in = nullptr;    // [M, ld_aux]   pre-activation for kActDGelu
    float* colsum = nullptr;                  // [N] fp32      atomically += sum_m out[m, n]
    int64_t ld_res = 0;
    int64_t ld_aux = 0;
    int64_t colsum_bi_stride = 0;
    int res_row_mod = 0;  // > 0: residual row = m % res_row_mod (broadcast a [rows, N] table, e.g. pos_embed)
    int act = kActNone;
    int has_aux_out = 0;  // also store the pre-activation (after bias) through the aux tensor map
};

// D[b][M, N] = epi(A[b] (M x K) * B[b] (N x K)^T).  major_x: 0 = K contiguous, 1 = M/N contiguous.
// block_n: 0 = auto, else 128 / 256.  max_ctas: 0 = all SMs (used to carve SMs out for comm kernels).
void gemm_bf16(const GemmOperand& A, int major_a, const GemmOperand& B, int major_b, const GemmOperand& D,
               const GemmOperand* aux_out, int M, int N, int K, const GemmEpilogue& epi, int block_n, int max_ctas,
               cudaStream_t stream);

}  // namespace b200
