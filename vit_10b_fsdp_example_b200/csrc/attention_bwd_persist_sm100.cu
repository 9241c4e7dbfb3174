// Persistent variant of the fused attention backward (attention_bwd_sm100.cu) for sm_100a.
//
// Same math, same two roles (kT = true: a CTA owns 128 keys and produces dK / dV; kT = false: a CTA owns 128 queries
// and produces dQ), same TMA / tcgen05 operand layouts.  What changes is the schedule, aimed at the hd = 160 case
// where only one CTA fits an SM and the one-shot kernel spends ~40 % of its 13.5 us per CTA outside the tile pipeline
// (launch + barrier / TMEM setup, the first resident-tile load, the first score MMA, the epilogue):
//   * one CTA per SM stays alive and loops over work items (image, head, 128-row block); the Y-tile ring, the T_s / T_p
//     buffers and the P / dS staging buffers simply keep cycling across item boundaries;
//   * the resident tiles of item i+1 are requested as soon as the last score MMA of item i has drained them, so that
//     load, the first score MMA of item i+1 and the epilogue of item i overlap;
//   * EIGHT softmax warps (two per TMEM lane quarter, each taking 32 of a tile's 64 columns) halve the per-tile
//     exp2 / dS critical path, and the epilogue columns are split the same way.
// Opt-in (B200_ATTN_PERSIST=1) until it has been run on hardware; the one-shot kernel stays the validated reference.
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

constexpr int kThreads = 320;  // warp 0: TMA, warp 1: MMA + TMEM, warps 2-9: softmax-backward math + epilogue
constexpr int kTileC = 64;
constexpr int kMaxSeqP = 1024;

struct BwdPParams {
    int N, H, B, D;
    int nblk;   // 128-row blocks per (image, head)
    int total;  // work items
    float scale, scale_log2;
    const float* lse;
    const float* delta;
    __nv_bfloat16* out1;
    __nv_bfloat16* out2;
    int64_t ld_out;
};

template <int HD, bool kT>
struct BwdPCfg {
    static constexpr int W = (HD % 64 == 0) ? 64 : 32;
    static constexpr int kAtoms = HD / W;
    static constexpr uint32_t kLayout = (W == 64) ? 2u : 4u;
    static constexpr int kRowBytes = W * 2;
    static constexpr int kXBytes = 128 * HD * 2;
    static constexpr int kYBytes = kTileC * HD * 2;
    static constexpr int kStageBytes = 2 * kYBytes;
    static constexpr int kEBytes = 128 * kTileC * 2;
    static constexpr int kTsBufs = 2;
    static constexpr int kNumAcc = kT ? 2 : 1;
    static constexpr int kColTs = 0;
    static constexpr int kColTp = kTsBufs * kTileC;
    static constexpr int kColAcc1 = kColTp + kTileC;
    static constexpr int kColAcc2 = kColAcc1 + HD;
    static constexpr int kColsUsed = kColAcc1 + kNumAcc * HD;
    static constexpr int kTmemCols = kColsUsed <= 256 ? 256 : 512;
    __host__ __device__ static constexpr int stat_bytes(int n_tokens) {
        return kT ? 2 * 4 * ((n_tokens + kTileC - 1) / kTileC) * kTileC : 0;
    }
    __host__ __device__ static constexpr int smem_bytes(int n_tokens) {
        return 2 * kXBytes + 2 * kStageBytes + 2 * kEBytes + stat_bytes(n_tokens) + 256;
    }
    static_assert(kColsUsed <= 512, "TMEM budget exceeded");
    static_assert(2 * kXBytes + 2 * kStageBytes + 2 * kEBytes + (kT ? 8 * kMaxSeqP : 0) + 256 <= 232448,
                  "shared memory budget exceeded");
};

template <int HD, bool kT>
__global__ void __launch_bounds__(kThreads) attn_bwd_persist_sm100_kernel(const __grid_constant__ CUtensorMap tmap_x1,
                                                                         const __grid_constant__ CUtensorMap tmap_x2,
                                                                         const __grid_constant__ CUtensorMap tmap_y1,
                                                                         const __grid_constant__ CUtensorMap tmap_y2,
                                                                         const BwdPParams p) {
    using C = BwdPCfg<HD, kT>;
    constexpr int W = C::W, kAtoms = C::kAtoms, kRowBytes = C::kRowBytes, kTsBufs = C::kTsBufs;
    constexpr uint32_t kLayout = C::kLayout;
    constexpr uint32_t kSbo = 8 * kRowBytes;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sX1 = smem;
    uint8_t* sX2 = sX1 + C::kXBytes;
    uint8_t* sY = sX2 + C::kXBytes;
    uint8_t* sE = sY + 2 * C::kStageBytes;
    uint8_t* sD = sE + C::kEBytes;
    const int nt = (p.N + kTileC - 1) / kTileC;
    float* s_lse2 = reinterpret_cast<float*>(sD + C::kEBytes);
    float* s_delta = s_lse2 + (kT ? nt * kTileC : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_lse2) + C::stat_bytes(p.N));
    uint64_t* x_full = bars;
    uint64_t* x_empty = bars + 1;
    uint64_t* y_full = bars + 2;     // [2]
    uint64_t* y_empty = bars + 4;    // [2]
    uint64_t* ts_full = bars + 6;    // [2]
    uint64_t* ts_empty = bars + 8;   // [2]
    uint64_t* tp_full = bars + 10;
    uint64_t* tp_empty = bars + 11;
    uint64_t* e_full = bars + 12;
    uint64_t* e_empty = bars + 13;
    uint64_t* d_full = bars + 14;
    uint64_t* d_empty = bars + 15;
    uint64_t* acc_full = bars + 16;
    uint64_t* acc_empty = bars + 17;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

    const uint32_t warp_idx = threadIdx.x / 32;
    const uint32_t lane = lane_id();
    const int n_items =
        (p.total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int n_tiles = n_items * nt;
    auto decode = [&](int i, int& blk, int& h, int& b) {
        const int w = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
        blk = w % p.nblk;
        const int bh = w / p.nblk;
        h = bh % p.H;
        b = bh / p.H;
    };

    if (warp_idx == 0 && elect_one()) {
        prefetch_tmap(&tmap_x1);
        prefetch_tmap(&tmap_x2);
        prefetch_tmap(&tmap_y1);
        prefetch_tmap(&tmap_y2);
        mbar_init(x_full, 1);
        mbar_init(x_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&y_full[i], 1);
            mbar_init(&y_empty[i], 1);
            mbar_init(&ts_full[i], 1);
            mbar_init(&ts_empty[i], 8);
        }
        mbar_init(tp_full, 1);
        mbar_init(tp_empty, 8);
        mbar_init(e_full, 8);
        mbar_init(e_empty, 1);
        mbar_init(d_full, 8);
        mbar_init(d_empty, 1);
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
        fence_mbar_init();
    }
    if (warp_idx == 1) tmem_alloc<1>(tmem_ptr_smem, C::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp_idx == 0) {
        // ===================================== TMA producer =====================================
        if (elect_one()) {
            auto load_x = [&](int i) {
                int blk, h, b;
                decode(i, blk, h, b);
                mbar_arrive_expect_tx(x_full, 2 * C::kXBytes);
#pragma unroll
                for (int a = 0; a < kAtoms; ++a) {
                    tma_load_4d(&tmap_x1, x_full, sX1 + a * (128 * kRowBytes), a * W, blk * 128, h, b);
                    tma_load_4d(&tmap_x2, x_full, sX2 + a * (128 * kRowBytes), a * W, blk * 128, h, b);
                }
            };
            if (n_items > 0) load_x(0);
            for (int t = 0; t < n_tiles; ++t) {
                const int i = t / nt, j = t - i * nt, st = t & 1;
                int blk, h, b;
                decode(i, blk, h, b);
                if (t >= 2) mbar_wait(&y_empty[st], ((t >> 1) - 1) & 1);
                uint8_t* y1 = sY + st * C::kStageBytes;
                uint8_t* y2 = y1 + C::kYBytes;
                mbar_arrive_expect_tx(&y_full[st], C::kStageBytes);
#pragma unroll
                for (int a = 0; a < kAtoms; ++a) {
                    tma_load_4d(&tmap_y1, &y_full[st], y1 + a * (kTileC * kRowBytes), a * W, j * kTileC, h, b);
                    tma_load_4d(&tmap_y2, &y_full[st], y2 + a * (kTileC * kRowBytes), a * W, j * kTileC, h, b);
                }
                if (j == nt - 1 && i + 1 < n_items) {
                    // resident tiles of the next item: free once the last score MMAs of this item have read them
                    mbar_wait(x_empty, i & 1);
                    load_x(i + 1);
                }
            }
        }
    } else if (warp_idx == 1) {
        // ===================================== MMA issuer =====================================
        if (elect_one()) {
            constexpr uint32_t idesc_t = make_idesc_bf16(128, kTileC, 0, 0);
            constexpr uint32_t idesc_a = make_idesc_bf16(128, HD, 0, 1);
            auto issue_scores = [&](uint32_t tmem_d, const uint8_t* sx, const uint8_t* sy) {
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    const int atom = (k * 16) / W, within = (k * 16) % W;
                    const uint64_t da = make_smem_desc(smem_u32(sx) + atom * (128 * kRowBytes) + within * 2, 0, kSbo, kLayout);
                    const uint64_t db = make_smem_desc(smem_u32(sy) + atom * (kTileC * kRowBytes) + within * 2, 0, kSbo, kLayout);
                    umma_bf16<1>(tmem_d, da, db, idesc_t, k > 0 ? 1u : 0u);
                }
            };
            auto issue_acc = [&](uint32_t tmem_d, const uint8_t* se, const uint8_t* sy, bool accumulate) {
#pragma unroll
                for (int k = 0; k < kTileC / 16; ++k) {
                    const uint64_t da = make_smem_desc(smem_u32(se) + k * 32, 0, 1024, 2u);
                    const uint64_t db = make_smem_desc(smem_u32(sy) + k * 16 * kRowBytes, kTileC * kRowBytes, kSbo, kLayout);
                    umma_bf16<1>(tmem_d, da, db, idesc_a, (accumulate || k > 0) ? 1u : 0u);
                }
            };
            // score tile T_s of global tile t1 (waits for its operands and for a free T_s buffer)
            auto prefetch_ts = [&](int t1) {
                const int i1 = t1 / nt, j1 = t1 - i1 * nt, s1 = t1 & 1, tb1 = t1 % kTsBufs;
                if (j1 == 0) mbar_wait(x_full, i1 & 1);
                mbar_wait(&y_full[s1], (t1 >> 1) & 1);
                if (t1 >= kTsBufs) mbar_wait(&ts_empty[tb1], (t1 / kTsBufs - 1) & 1);
                tc_fence_after();
                issue_scores(tmem_base + C::kColTs + tb1 * kTileC, sX1, sY + s1 * C::kStageBytes);
                umma_commit<1>(&ts_full[tb1]);
            };
            if (n_tiles > 0) prefetch_ts(0);
            for (int t = 0; t < n_tiles; ++t) {
                const int i = t / nt, j = t - i * nt, st = t & 1;
                const uint8_t* y1 = sY + st * C::kStageBytes;
                const uint8_t* y2 = y1 + C::kYBytes;
                const bool last_of_item = j == nt - 1;
                if (t > 0) mbar_wait(tp_empty, (t - 1) & 1);
                tc_fence_after();
                issue_scores(tmem_base + C::kColTp, sX2, y2);
                umma_commit<1>(tp_full);
                if (last_of_item) umma_commit<1>(x_empty);  // both score MMAs of the item's last tile are issued
                // next score tile of the SAME item goes first (it runs under this tile's exponentials); across an
                // item boundary it would block on the next resident-tile load, so there the accumulation goes first
                if (!last_of_item && t + 1 < n_tiles) prefetch_ts(t + 1);
                if (j == 0 && i > 0) mbar_wait(acc_empty, (i - 1) & 1);  // previous item's epilogue has read O
                if constexpr (kT) {
                    mbar_wait(e_full, t & 1);
                    tc_fence_after();
                    issue_acc(tmem_base + C::kColAcc2, sE, y2, j > 0);
                    umma_commit<1>(e_empty);
                }
                mbar_wait(d_full, t & 1);
                tc_fence_after();
                issue_acc(tmem_base + C::kColAcc1, sD, y1, j > 0);
                umma_commit<1>(d_empty);
                umma_commit<1>(&y_empty[st]);
                if (last_of_item) {
                    umma_commit<1>(acc_full);
                    if (t + 1 < n_tiles) prefetch_ts(t + 1);
                }
            }
        }
    } else {
        // ============================ softmax-backward math + epilogue (8 warps) ============================
        const uint32_t quarter = warp_idx & 3;
        const uint32_t grp = (warp_idx - 2) >> 2;        // which 32-column half of every tile
        const uint32_t r = quarter * 32 + lane;
        const uint32_t taddr = tmem_base + ((quarter * 32) << 16);
        constexpr float kLog2e = 1.4426950408889634f;
        const uint32_t erow = smem_u32(sE) + r * 128;
        const uint32_t drow = smem_u32(sD) + r * 128;
        const int tid = static_cast<int>(threadIdx.x) - 64;  // 0 .. 255
        for (int i = 0; i < n_items; ++i) {
            int blk, h, b;
            decode(i, blk, h, b);
            const int64_t bh = static_cast<int64_t>(b) * p.H + h;
            const int row = blk * 128 + static_cast<int>(r);
            const bool row_ok = row < p.N;
            float lse2_r = 0.f, delta_r = 0.f;
            if constexpr (kT) {
                named_bar_sync(1, 256);  // every warp is done with the previous item's column statistics
                for (int q = tid; q < nt * kTileC; q += 256) {
                    const bool ok = q < p.N;
                    s_lse2[q] = ok ? p.lse[bh * p.N + q] * kLog2e : 0.f;
                    s_delta[q] = ok ? p.delta[bh * p.N + q] : 0.f;
                }
                named_bar_sync(1, 256);
            } else {
                if (row_ok) {
                    lse2_r = p.lse[bh * p.N + row] * kLog2e;
                    delta_r = p.delta[bh * p.N + row];
                }
            }
            for (int j = 0; j < nt; ++j) {
                const int t = i * nt + j, tb = t % kTsBufs;
                const int col0 = j * kTileC + static_cast<int>(grp) * 32;
                uint32_t pk[16];
                mbar_wait(&ts_full[tb], (t / kTsBufs) & 1);
                tc_fence_after();
                if constexpr (kT) {
                    if (t > 0) mbar_wait(e_empty, (t - 1) & 1);
                }
                {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + C::kColTs + tb * kTileC + grp * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int x = 0; x < 32; x += 2) {
                        const int col = col0 + x;
                        float l0, l1;
                        if constexpr (kT) {
                            l0 = s_lse2[col], l1 = s_lse2[col + 1];
                        } else {
                            l0 = l1 = lse2_r;
                        }
                        const float e0 = (row_ok && col < p.N) ? exp2f(fmaf(__uint_as_float(v[x]), p.scale_log2, -l0)) : 0.f;
                        const float e1 = (row_ok && col + 1 < p.N) ? exp2f(fmaf(__uint_as_float(v[x + 1]), p.scale_log2, -l1)) : 0.f;
                        pk[x / 2] = pack_bf16x2(e0, e1);
                    }
                    if constexpr (kT) {
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            const uint32_t chunk = grp * 4 + j8;
                            st_shared_v4(erow + ((chunk ^ (r & 7)) << 4), pk[j8 * 4], pk[j8 * 4 + 1], pk[j8 * 4 + 2],
                                         pk[j8 * 4 + 3]);
                        }
                    }
                }
                tc_fence_before();
                if constexpr (kT) fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&ts_empty[tb]);
                    if constexpr (kT) mbar_arrive(e_full);
                }

                mbar_wait(tp_full, t & 1);
                tc_fence_after();
                if (t > 0) mbar_wait(d_empty, (t - 1) & 1);
                {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + C::kColTp + grp * 32, v);
                    tmem_ld_wait();
                    uint32_t dk[16];
#pragma unroll
                    for (int x = 0; x < 32; x += 2) {
                        const int col = col0 + x;
                        float d0, d1;
                        if constexpr (kT) {
                            d0 = s_delta[col], d1 = s_delta[col + 1];
                        } else {
                            d0 = d1 = delta_r;
                        }
                        const uint32_t pw = pk[x / 2];
                        const float s0 = bf16_lo(pw) * (__uint_as_float(v[x]) - d0) * p.scale;
                        const float s1 = bf16_hi(pw) * (__uint_as_float(v[x + 1]) - d1) * p.scale;
                        dk[x / 2] = pack_bf16x2(s0, s1);
                    }
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        const uint32_t chunk = grp * 4 + j8;
                        st_shared_v4(drow + ((chunk ^ (r & 7)) << 4), dk[j8 * 4], dk[j8 * 4 + 1], dk[j8 * 4 + 2],
                                     dk[j8 * 4 + 3]);
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(tp_empty);
                    mbar_arrive(d_full);
                }
            }

            // ---- epilogue of item i: the two warp groups take alternating 32-column chunks ----
            mbar_wait(acc_full, i & 1);
            tc_fence_after();
#pragma unroll
            for (int a = 0; a < C::kNumAcc; ++a) {
                __nv_bfloat16* obase = (a == 0) ? p.out1 : p.out2;
                __nv_bfloat16* orow = obase + (static_cast<int64_t>(b) * p.N + row) * p.ld_out + h * HD;
                const uint32_t cbase = (a == 0) ? C::kColAcc1 : C::kColAcc2;
#pragma unroll 1
                for (int c = static_cast<int>(grp); c < HD / 32; c += 2) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + cbase + c * 32, v);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8) {
                            uint4 o;
                            o.x = pack_bf16x2(__uint_as_float(v[j8 * 8]), __uint_as_float(v[j8 * 8 + 1]));
                            o.y = pack_bf16x2(__uint_as_float(v[j8 * 8 + 2]), __uint_as_float(v[j8 * 8 + 3]));
                            o.z = pack_bf16x2(__uint_as_float(v[j8 * 8 + 4]), __uint_as_float(v[j8 * 8 + 5]));
                            o.w = pack_bf16x2(__uint_as_float(v[j8 * 8 + 6]), __uint_as_float(v[j8 * 8 + 7]));
                            *reinterpret_cast<uint4*>(orow + c * 32 + j8 * 8) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
        }
    }
    __syncthreads();
    if (warp_idx == 1) tmem_dealloc<1>(tmem_base, C::kTmemCols);
}

template <int HD, bool kT>
void launch_bwd_persist(const GemmOperand& x1, const GemmOperand& x2, const GemmOperand& y1, const GemmOperand& y2,
                        const BwdPParams& p, cudaStream_t stream) {
    using C = BwdPCfg<HD, kT>;
    auto kern = attn_bwd_persist_sm100_kernel<HD, kT>;
    static bool attr_set = false;
    static int num_sms = 0;
    if (!attr_set) {
        cudaError_t err =
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes(kMaxSeqP));
        if (err != cudaSuccess)
            throw std::runtime_error(std::string("attention bwd-persist smem attr: ") + cudaGetErrorString(err));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        attr_set = true;
    }
    const int sw = C::W * 2;
    CUtensorMap tx1 = make_tensor_map_4d(x1, HD, p.N, C::W, 128, sw);
    CUtensorMap tx2 = make_tensor_map_4d(x2, HD, p.N, C::W, 128, sw);
    CUtensorMap ty1 = make_tensor_map_4d(y1, HD, p.N, C::W, kTileC, sw);
    CUtensorMap ty2 = make_tensor_map_4d(y2, HD, p.N, C::W, kTileC, sw);
    const int grid = p.total < num_sms ? p.total : num_sms;
    kern<<<grid, kThreads, C::smem_bytes(p.N), stream>>>(tx1, tx2, ty1, ty2, p);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess)
        throw std::runtime_error(std::string("attention bwd-persist launch: ") + cudaGetErrorString(err));
}

template <int HD>
void run_bwd_persist(const GemmOperand& q, const GemmOperand& k, const GemmOperand& v, const GemmOperand& dO,
                     BwdPParams p, __nv_bfloat16* dqkv, cudaStream_t stream) {
    p.out1 = dqkv + p.D, p.out2 = dqkv + 2 * p.D;
    launch_bwd_persist<HD, true>(k, v, q, dO, p, stream);   // dK, dV
    p.out1 = dqkv, p.out2 = nullptr;
    launch_bwd_persist<HD, false>(q, dO, k, v, p, stream);  // dQ
}

}  // namespace

// delta must already hold rowsum(dO o O) (attention_bwd computes it; see attention_bwd_sm100.cu)
void attention_bwd_persist_core(const __nv_bfloat16* qkv, int64_t ld_qkv, const __nv_bfloat16* dout, int64_t ld_do,
                                const float* lse, const float* delta, __nv_bfloat16* dqkv, int B, int N, int H, int hd,
                                cudaStream_t stream) {
    const int D = H * hd;
    GemmOperand q, k, v, dO;
    q.ptr = qkv, k.ptr = qkv + D, v.ptr = qkv + 2 * D, dO.ptr = dout;
    for (GemmOperand* o : {&q, &k, &v, &dO}) {
        o->ld = (o == &dO) ? ld_do : ld_qkv;
        o->nb_inner = H, o->stride_b_inner = hd;
        o->nb_outer = B, o->stride_b_outer = static_cast<int64_t>(N) * o->ld;
    }
    BwdPParams p;
    p.N = N, p.H = H, p.B = B, p.D = D;
    p.nblk = (N + 127) / 128;
    p.total = B * H * p.nblk;
    p.scale = 1.0f / sqrtf(static_cast<float>(hd));
    p.scale_log2 = p.scale * 1.4426950408889634f;
    p.lse = lse, p.delta = delta;
    p.out1 = p.out2 = nullptr;
    p.ld_out = 3 * static_cast<int64_t>(D);
    if (hd == 64) run_bwd_persist<64>(q, k, v, dO, p, dqkv, stream);
    else if (hd == 128) run_bwd_persist<128>(q, k, v, dO, p, dqkv, stream);
    else run_bwd_persist<160>(q, k, v, dO, p, dqkv, stream);
}

}  // namespace b200
