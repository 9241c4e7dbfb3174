// Host API of the fused tcgen05 attention forward (see attention_sm100.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace b200 {

bool attention_fwd_supported(int N, int hd);

// qkv: packed [B*N, 3*H*hd] (row stride ld_qkv).  out: [B*N, H*hd].  lse: [B*H, N] fp32 or null.
// probs: normalised softmax [B*H, N, ldp] bf16 or null (only written when the un-fused backward needs it).
void attention_fwd(const __nv_bfloat16* qkv, int64_t ld_qkv, __nv_bfloat16* out, float* lse, __nv_bfloat16* probs,
                   int64_t ldp, int B, int N, int H, int hd, cudaStream_t stream);

}  // namespace b200
