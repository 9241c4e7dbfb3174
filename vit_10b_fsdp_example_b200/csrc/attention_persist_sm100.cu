// Persistent, software-pipelined attention forward for sm_100a (128 < N <= 256 tokens, head dim 64 / 128 / 160).
//
// The one-shot kernel in attention_sm100.cu runs load -> S MMA -> softmax -> PV MMA -> store strictly in sequence and
// only one CTA fits an SM at hd = 160 (200 KB of shared memory), so nothing hides its loads (ncu: tensor pipe 9 %,
// DRAM 21 %, 773 us against a 210 us DRAM floor; profiles/r1e_attention_fwd_fused.md).  This kernel keeps ONE CTA per
// SM alive for the whole launch and lets its warps run ahead across work items (image, head, 128-query block):
//   warp 0   TMA producer: Q + K of item i+1 are requested as soon as the S MMAs of item i have drained them; V is
//            streamed in 64-key tiles through a 2-stage ring;
//   warp 1   tcgen05.mma issuer: S = Q K^T (N = 256) for item i+1 is issued right after the PV MMAs of item i, i.e. it
//            runs under the epilogue of item i;  O += P_j V_j per 64-key tile as soon as that tile of P is staged;
//   warps 2-5 one thread per query row: row max, exp2, un-normalised bf16 P into a 2-deep shared-memory ring
//            (K-major SWIZZLE_128B, the A operand of the PV MMA), 1/sum applied in the epilogue, optional log-sum-exp.
// TMEM: S [128 x 256] fp32 + O [128 x hd] fp32 (416 columns at hd = 160).  Shared memory: Q 40 KB + K 80 KB +
// V ring 40 KB + P ring 32 KB = 192 KB at hd = 160.  Scores never reach HBM.
// Replaces timm Attention's materialised softmax (reference run_vit_training.py:134 -> timm Block -> Attention).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "attention_sm100.h"
#include "gemm_sm100.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kPThreads = 192;
constexpr int kVT = 64;    // keys per V / P tile
constexpr int kNK = 256;   // key columns of the score tile (padded)

struct PersistParams {
    int N, H, B, D;
    int nq;          // 128-query blocks per (image, head)
    int nkt;         // 64-key tiles per item
    int total;       // work items = B * H * nq
    float scale_log2, scale;
    __nv_bfloat16* out;
    float* lse;      // optional
    long long* trace;  // optional in-kernel timeline of CTA 0: clock64 stamps, 16 slots per work item (see kTr*)
    int trace_items;
};

// trace slots (per item): who / what
enum { kTrQkIssued = 0, kTrQkFull = 1, kTrSIssue = 2, kTrAccEmpty = 3, kTrPv0 = 4, kTrSFull = 8, kTrPass1 = 9,
       kTrPass2 = 10, kTrAccFull = 11, kTrEpiDone = 12 };
__device__ __forceinline__ void stamp(const PersistParams& p, int item, int slot) {
    if (p.trace != nullptr && blockIdx.x == 0 && item < p.trace_items) p.trace[item * 16 + slot] = clock64();
}

template <int HD>
struct PersistCfg {
    static constexpr int W = (HD % 64 == 0) ? 64 : 32;
    static constexpr int kAtoms = HD / W;
    static constexpr uint32_t kLayout = (W == 64) ? 2u : 4u;
    static constexpr int kRowBytes = W * 2;
    static constexpr int kQBytes = 128 * HD * 2;
    static constexpr int kKBytes = kNK * HD * 2;
    static constexpr int kVBytes = kVT * HD * 2;
    static constexpr int kEBytes = 128 * kVT * 2;
    static constexpr int kColAcc = kNK;
    static constexpr int kSmem = kQBytes + kKBytes + 2 * kVBytes + 2 * kEBytes + 256;
    static_assert(kNK + HD <= 512, "TMEM budget exceeded");
    static_assert(kSmem <= 232448, "shared memory budget exceeded");
};

template <int HD>
__global__ void __launch_bounds__(kPThreads) attn_fwd_persist_sm100_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                                          const __grid_constant__ CUtensorMap tmap_k,
                                                                          const __grid_constant__ CUtensorMap tmap_v,
                                                                          const PersistParams p) {
    using C = PersistCfg<HD>;
    constexpr int W = C::W, kAtoms = C::kAtoms, kRowBytes = C::kRowBytes;
    constexpr uint32_t kLayout = C::kLayout;
    constexpr uint32_t kSbo = 8 * kRowBytes;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + C::kQBytes;
    uint8_t* sV = sK + C::kKBytes;               // [2 stages]
    uint8_t* sE = sV + 2 * C::kVBytes;           // [2 buffers]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sE + 2 * C::kEBytes);
    uint64_t* qk_full = bars;
    uint64_t* qk_empty = bars + 1;
    uint64_t* s_full = bars + 2;
    uint64_t* s_empty = bars + 3;
    uint64_t* acc_full = bars + 4;
    uint64_t* acc_empty = bars + 5;
    uint64_t* v_full = bars + 6;    // [2]
    uint64_t* v_empty = bars + 8;   // [2]
    uint64_t* e_full = bars + 10;   // [2]
    uint64_t* e_empty = bars + 12;  // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

    const uint32_t warp_idx = threadIdx.x / 32;
    const uint32_t lane = lane_id();
    const int nkt = p.nkt;
    // static striding: neighbouring CTAs work on the two query blocks of the same (image, head) -> K / V hit in L2
    const int n_items = (p.total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    auto decode = [&](int i, int& qb, int& h, int& b) {
        const int w = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
        qb = w % p.nq;
        const int bh = w / p.nq;
        h = bh % p.H;
        b = bh / p.H;
    };

    if (warp_idx == 0 && elect_one()) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_k);
        prefetch_tmap(&tmap_v);
        mbar_init(qk_full, 1);
        mbar_init(qk_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(s_empty, 4);
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&e_full[i], 4);
            mbar_init(&e_empty[i], 1);
        }
        fence_mbar_init();
    }
    if (warp_idx == 1) tmem_alloc<1>(tmem_ptr_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            auto load_qk = [&](int i) {
                int qb, h, b;
                decode(i, qb, h, b);
                mbar_arrive_expect_tx(qk_full, C::kQBytes + C::kKBytes);
#pragma unroll
                for (int a = 0; a < kAtoms; ++a) {
                    tma_load_4d(&tmap_q, qk_full, sQ + a * (128 * kRowBytes), a * W, qb * 128, h, b);
                    tma_load_4d(&tmap_k, qk_full, sK + a * (kNK * kRowBytes), a * W, 0, h, b);
                }
            };
            auto load_v = [&](int i, int j) {
                int qb, h, b;
                decode(i, qb, h, b);
                const int t = i * nkt + j, st = t & 1;
                if (t >= 2) mbar_wait(&v_empty[st], ((t >> 1) - 1) & 1);
                uint8_t* dst = sV + st * C::kVBytes;
                mbar_arrive_expect_tx(&v_full[st], C::kVBytes);
#pragma unroll
                for (int a = 0; a < kAtoms; ++a)
                    tma_load_4d(&tmap_v, &v_full[st], dst + a * (kVT * kRowBytes), a * W, j * kVT, h, b);
            };
            if (n_items > 0) load_qk(0);
            for (int i = 0; i < n_items; ++i) {
                const int first = nkt < 2 ? nkt : 2;
                for (int j = 0; j < first; ++j) load_v(i, j);
                if (i + 1 < n_items) {  // Q / K of the next item as soon as this item's S MMAs have drained them
                    mbar_wait(qk_empty, i & 1);
                    load_qk(i + 1);
                    stamp(p, i + 1, kTrQkIssued);
                }
                for (int j = first; j < nkt; ++j) load_v(i, j);
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer =====================================
        if (elect_one()) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, kNK, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
            for (int i = 0; i < n_items; ++i) {
                mbar_wait(qk_full, i & 1);
                stamp(p, i, kTrQkFull);
                if (i > 0) mbar_wait(s_empty, (i - 1) & 1);
                stamp(p, i, kTrSIssue);
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const int atom = (k * 16) / W, within = (k * 16) % W;
                    const uint64_t da = make_smem_desc(smem_u32(sQ) + atom * (128 * kRowBytes) + within * 2, 0, kSbo, kLayout);
                    const uint64_t db = make_smem_desc(smem_u32(sK) + atom * (kNK * kRowBytes) + within * 2, 0, kSbo, kLayout);
                    umma_bf16<1>(tmem_base, da, db, idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit<1>(s_full);
                umma_commit<1>(qk_empty);
                if (i > 0) mbar_wait(acc_empty, (i - 1) & 1);
                stamp(p, i, kTrAccEmpty);
                for (int j = 0; j < nkt; ++j) {
                    const int t = i * nkt + j, st = t & 1;
                    mbar_wait(&v_full[st], (t >> 1) & 1);
                    mbar_wait(&e_full[st], (t >> 1) & 1);
                    if (j < 4) stamp(p, i, kTrPv0 + j);
                    tc_fence_after();
                    const uint8_t* se = sE + st * C::kEBytes;
                    const uint8_t* sv = sV + st * C::kVBytes;
#pragma unroll
                    for (int k = 0; k < kVT / 16; ++k) {
                        const uint64_t da = make_smem_desc(smem_u32(se) + k * 32, 0, 1024, 2u);
                        const uint64_t db = make_smem_desc(smem_u32(sv) + k * 16 * kRowBytes, kVT * kRowBytes, kSbo, kLayout);
                        umma_bf16<1>(tmem_base + C::kColAcc, da, db, idesc_o, (j > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit<1>(&e_empty[st]);
                    umma_commit<1>(&v_empty[st]);
                }
                umma_commit<1>(acc_full);
            }
        }
    } else {
        // ===================================== softmax + epilogue =====================================
        const uint32_t quarter = warp_idx & 3;
        const uint32_t r = quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((quarter * 32) << 16);
        for (int i = 0; i < n_items; ++i) {
            int qb, h, b;
            decode(i, qb, h, b);
            const int q = qb * 128 + static_cast<int>(r);
            const bool row_ok = q < p.N;
            mbar_wait(s_full, i & 1);
            if (warp_idx == 2 && lane == 0) stamp(p, i, kTrSFull);
            tc_fence_after();
            float mx = -INFINITY;
            for (int c = 0; c < nkt * 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int x = 0; x < 32; ++x)
                    if (c * 32 + x < p.N) mx = fmaxf(mx, __uint_as_float(v[x]));
            }
            const float m_scaled = mx * p.scale_log2;
            float sum = 0.f;
            if (warp_idx == 2 && lane == 0) stamp(p, i, kTrPass1);
            for (int j = 0; j < nkt; ++j) {
                const int t = i * nkt + j, eb = t & 1;
                if (t >= 2) mbar_wait(&e_empty[eb], ((t >> 1) - 1) & 1);
                const uint32_t erow = smem_u32(sE) + eb * C::kEBytes + r * 128;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + j * kVT + c * 32, v);
                    tmem_ld_wait();
                    uint32_t pk[16];
#pragma unroll
                    for (int x = 0; x < 32; x += 2) {
                        const int col = j * kVT + c * 32 + x;
                        const float e0 = col < p.N ? exp2f(fmaf(__uint_as_float(v[x]), p.scale_log2, -m_scaled)) : 0.f;
                        const float e1 = col + 1 < p.N ? exp2f(fmaf(__uint_as_float(v[x + 1]), p.scale_log2, -m_scaled)) : 0.f;
                        sum += e0 + e1;
                        pk[x / 2] = pack_bf16x2(e0, e1);
                    }
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        const uint32_t chunk = c * 4 + j8;
                        st_shared_v4(erow + ((chunk ^ (r & 7)) << 4), pk[j8 * 4], pk[j8 * 4 + 1], pk[j8 * 4 + 2],
                                     pk[j8 * 4 + 3]);
                    }
                }
                fence_proxy_async_smem();  // generic-proxy P writes -> async-proxy (tensor core) reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&e_full[eb]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty);  // S of the next item may overwrite the score columns
            if (warp_idx == 2 && lane == 0) stamp(p, i, kTrPass2);
            const float inv = 1.0f / sum;
            const int64_t bh = static_cast<int64_t>(b) * p.H + h;
            if (p.lse != nullptr && row_ok) p.lse[bh * p.N + q] = mx * p.scale + __logf(sum);

            mbar_wait(acc_full, i & 1);
            if (warp_idx == 2 && lane == 0) stamp(p, i, kTrAccFull);
            tc_fence_after();
            __nv_bfloat16* orow = p.out + (static_cast<int64_t>(b) * p.N + q) * p.D + h * HD;
#pragma unroll 1
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(taddr + C::kColAcc + c * 32, v);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[j8 * 8]) * inv, __uint_as_float(v[j8 * 8 + 1]) * inv);
                        o.y = pack_bf16x2(__uint_as_float(v[j8 * 8 + 2]) * inv, __uint_as_float(v[j8 * 8 + 3]) * inv);
                        o.z = pack_bf16x2(__uint_as_float(v[j8 * 8 + 4]) * inv, __uint_as_float(v[j8 * 8 + 5]) * inv);
                        o.w = pack_bf16x2(__uint_as_float(v[j8 * 8 + 6]) * inv, __uint_as_float(v[j8 * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(orow + c * 32 + j8 * 8) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);  // the PV MMAs of the next item may overwrite O
            if (warp_idx == 2 && lane == 0) stamp(p, i, kTrEpiDone);
        }
    }
    __syncthreads();
    if (warp_idx == 1) tmem_dealloc<1>(tmem_base, 512);
}

template <int HD>
void launch_persist(const GemmOperand& q, const GemmOperand& k, const GemmOperand& v, const PersistParams& p,
                    cudaStream_t stream) {
    using C = PersistCfg<HD>;
    auto kern = attn_fwd_persist_sm100_kernel<HD>;
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem);
        if (err != cudaSuccess)
            throw std::runtime_error(std::string("attention persist smem attr: ") + cudaGetErrorString(err));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        attr_set = true;
    }
    const int sw = C::W * 2;
    CUtensorMap tq = make_tensor_map_4d(q, HD, p.N, C::W, 128, sw);
    CUtensorMap tk = make_tensor_map_4d(k, HD, p.N, C::W, kNK, sw);
    CUtensorMap tv = make_tensor_map_4d(v, HD, p.N, C::W, kVT, sw);
    const int grid = p.total < num_sms ? p.total : num_sms;
    kern<<<grid, kPThreads, C::kSmem, stream>>>(tq, tk, tv, p);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        throw std::runtime_error(std::string("attention persist launch: ") + cudaGetErrorString(err));
}

}  // namespace

bool attention_fwd_persist_supported(int N, int hd) {
    return N > 128 && N <= 256 && N % 2 == 0 && (hd == 64 || hd == 128 || hd == 160);
}

// Optional in-kernel timeline: clock64 stamps of CTA 0's producer / MMA / softmax roles, 16 per work item.
static long long* g_trace = nullptr;
static int g_trace_items = 0;
void attention_set_trace(long long* buf, int items) { g_trace = buf, g_trace_items = items; }

void attention_fwd_persist(const __nv_bfloat16* qkv, int64_t ld_qkv, __nv_bfloat16* out, float* lse, int B, int N, int H,
                           int hd, cudaStream_t stream) {
    if (!attention_fwd_persist_supported(N, hd))
        throw std::runtime_error("attention_fwd_persist: unsupported (N, head_dim)");
    const int D = H * hd;
    GemmOperand q, k, v;
    q.ptr = qkv, k.ptr = qkv + D, v.ptr = qkv + 2 * D;
    for (GemmOperand* o : {&q, &k, &v}) {
        o->ld = ld_qkv;
        o->nb_inner = H, o->stride_b_inner = hd;
        o->nb_outer = B, o->stride_b_outer = static_cast<int64_t>(N) * ld_qkv;
    }
    PersistParams p;
    p.N = N, p.H = H, p.B = B, p.D = D;
    p.nq = (N + 127) / 128;
    p.nkt = (N + kVT - 1) / kVT;
    p.total = B * H * p.nq;
    p.scale = 1.0f / sqrtf(static_cast<float>(hd));
    p.scale_log2 = p.scale * 1.4426950408889634f;
    p.out = out, p.lse = lse;
    p.trace = g_trace, p.trace_items = g_trace_items;
    if (hd == 64) launch_persist<64>(q, k, v, p, stream);
    else if (hd == 128) launch_persist<128>(q, k, v, p, stream);
    else launch_persist<160>(q, k, v, p, stream);
}

}  // namespace b200
