// Host API of the memory-bound sm_100a kernels (see elementwise.cu).
// LayerNorm, softmax, GELU, cross-entropy, AdamW, clipping: the ATen/XLA ops behind run_vit_training.py:134-162,229,237,270.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace b200 {

void layernorm_fwd(const __nv_bfloat16* x, const __nv_bfloat16* gamma, const __nv_bfloat16* beta, __nv_bfloat16* y,
                   float* mean, float* rstd, int rows, int D, float eps, cudaStream_t stream);

// dgamma / dbeta / dxsum are fp32 [D] accumulators (atomically added to; zero them first).
void layernorm_bwd(const __nv_bfloat16* dy, const __nv_bfloat16* x, const __nv_bfloat16* gamma, const float* mean,
                   const float* rstd, const __nv_bfloat16* dres, __nv_bfloat16* dx, float* dgamma, float* dbeta,
                   float* dxsum, int rows, int D, cudaStream_t stream);

void softmax_fwd(__nv_bfloat16* s, int64_t rows, int n, int64_t ld, float scale, cudaStream_t stream);
void softmax_bwd(__nv_bfloat16* dp, const __nv_bfloat16* p, int64_t rows, int n, int64_t ld, float scale,
                 cudaStream_t stream);

// loss (fp32 scalar, atomically added) = mean CE; dlogits may be null (eval); correct may be null.
void cross_entropy(const __nv_bfloat16* logits, const int64_t* target, __nv_bfloat16* dlogits, float* loss,
                   int* correct, int B, int C, cudaStream_t stream);

void im2col(const void* img, bool img_is_bf16, __nv_bfloat16* cols, int B, int S, int P, int Kpad,
            cudaStream_t stream);

// Wide-row LayerNorm backward as a cp.async.bulk row pipeline (layernorm_stream.cu); same contract as layernorm_bwd.
bool layernorm_bwd_stream_supported(int D);
void layernorm_bwd_stream(const __nv_bfloat16* dy, const __nv_bfloat16* x, const __nv_bfloat16* gamma, const float* mean,
                          const float* rstd, const __nv_bfloat16* dres, __nv_bfloat16* dx, float* dgamma, float* dbeta,
                          float* dxsum, int rows, int D, cudaStream_t stream);
void gelu_fwd(const __nv_bfloat16* u, __nv_bfloat16* g, int64_t n, cudaStream_t stream);
void dgelu_mul(const __nv_bfloat16* dg, const __nv_bfloat16* u, __nv_bfloat16* du, int64_t n, cudaStream_t stream);
// y = x * keep / (1 - p), keep = Philox-4x32-10(key, vector index): a pure function of (key, position), so recompute and
// backward regenerate the mask.  p is quantised to 1/65536.  y may alias x.
void dropout(const __nv_bfloat16* x, __nv_bfloat16* y, int64_t n, float p, uint64_t key, cudaStream_t stream);
// pooled[b] = mean over the N tokens of image b; backward broadcasts dpooled[b] / N to every token row.
void meanpool_fwd(const __nv_bfloat16* xn, __nv_bfloat16* pooled, int B, int N, int D, cudaStream_t stream);
void meanpool_bwd(const __nv_bfloat16* dpooled, __nv_bfloat16* dxn, int B, int N, int D, cudaStream_t stream);
void colsum(const __nv_bfloat16* x, float* out, int64_t rows, int C, cudaStream_t stream);
void sumsq(const void* x, bool is_bf16, int64_t n, float* out, cudaStream_t stream);

// hyper (optional, device): [lr, step] override the host values (CUDA-graph friendly)
void adamw_split(uint16_t* hi, int16_t* lo, float* m, float* v, const void* grad, bool grad_is_bf16, int64_t n,
                 const float* clip_coef, float lr, float beta1, float beta2, float eps, float wd, int step,
                 cudaStream_t stream, const float* hyper = nullptr);
void adamw_fp32(float* w, float* m, float* v, const void* grad, bool grad_is_bf16, int64_t n, const float* clip_coef,
                float lr, float beta1, float beta2, float eps, float wd, int step, cudaStream_t stream);
void split_fp32(const float* w, uint16_t* hi, int16_t* lo, int64_t n, cudaStream_t stream);
void merge_fp32(const uint16_t* hi, const int16_t* lo, float* w, int64_t n, cudaStream_t stream);
void clip_coef(const float* sumsq_in, float max_norm, float* coef, float* norm_out, cudaStream_t stream);

}  // namespace b200
