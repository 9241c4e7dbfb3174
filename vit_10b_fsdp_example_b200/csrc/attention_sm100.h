// Host API of the fused tcgen05 attention kernels (attention_sm100.cu, attention_bwd_sm100.cu).
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

// Long-sequence forward (any even N; two passes over 64-key tiles, attention_bwd_sm100.cu): out + row log-sum-exp.
bool attention_fwd_long_supported(int N, int hd);
void attention_fwd_long(const __nv_bfloat16* qkv, int64_t ld_qkv, __nv_bfloat16* out, float* lse, int B, int N, int H,
                        int hd, cudaStream_t stream);

// Persistent, software-pipelined forward for 128 < N <= 256 (attention_persist_sm100.cu): out, optional lse.
bool attention_fwd_persist_supported(int N, int hd);
// debug: clock64 stamps of CTA 0 (16 int64 slots per work item) written by the persistent forward; nullptr = off
void attention_set_trace(long long* buf, int items);
// same for the persistent backward: 16 slots per global tile of CTA 0; role 0 = dK/dV kernel, 1 = dQ kernel
void attention_bwd_set_trace(long long* buf, int tiles, int role);
void attention_fwd_persist(const __nv_bfloat16* qkv, int64_t ld_qkv, __nv_bfloat16* out, float* lse, int B, int N, int H,
                           int hd, cudaStream_t stream);

bool attention_bwd_supported(int N, int hd);

// Fused backward (attention_bwd_sm100.cu).  dout / out: [B*N, H*hd] gradient and forward output of the attention core,
// lse: [B*H, N] from attention_fwd, delta: [B*H, N] fp32 workspace (written here), dqkv: packed [B*N, 3*H*hd].
// persist = true runs the persistent 8-softmax-warp variant (attention_bwd_persist_sm100.cu) instead of the one-shot kernels.
void attention_bwd(const __nv_bfloat16* qkv, int64_t ld_qkv, const __nv_bfloat16* dout, int64_t ld_do,
                   const __nv_bfloat16* out, int64_t ld_o, const float* lse, float* delta, __nv_bfloat16* dqkv,
                   int B, int N, int H, int hd, cudaStream_t stream, bool persist = false, float* colsum = nullptr);
// colsum (optional, persistent kernels only): zero-initialised fp32 [3 * D]; receives the column sums of dq | dk | dv
// (the qkv bias gradient) straight from the epilogue's staging tiles.
void attention_bwd_persist_core(const __nv_bfloat16* qkv, int64_t ld_qkv, const __nv_bfloat16* dout, int64_t ld_do,
                                const float* lse, const float* delta, __nv_bfloat16* dqkv, float* colsum, int B, int N,
                                int H, int hd, cudaStream_t stream);

}  // namespace b200
