// Persistent, warp-specialised bf16 GEMM for sm_100a (B200):
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B)  ->  shared-memory ring
//   tcgen05.mma.cta_group::2 (UMMA 256 x BLOCK_N x 16, one elected thread of the leader CTA)
//   fp32 accumulators double-buffered in TMEM  ->  tcgen05.ld epilogue warps
//   fused epilogue (bias, exact GELU, dGELU, residual add, pre-activation side output,
//   bias-gradient column sums)  ->  swizzled smem  ->  TMA store.
//
// One CTA pair (thread-block cluster of 2 = one TPC) owns a 256 x BLOCK_N output tile; each CTA
// loads its own 128 rows of A and half of the B tile, so every operand byte is fetched once per pair.
//
//   D[b][m, n] = epilogue( sum_k A[b][m, k] * B[b][n, k] )
//
// A and B may each be K-major (reduction dim contiguous) or MN-major (reduction dim strided), which
// covers forward (NT), dgrad (NN) and wgrad (TN) without materialising transposes.  Operands are 4-D
// TMA tensors (inner, outer, batch_inner, batch_outer) so strided per-head attention operands inside a
// packed qkv buffer are addressed in place.
//
// Capability parity: replaces the XLA-lowered dot ops under timm's Linear layers that the reference
// calls at run_vit_training.py:134-141,153 (qkv / proj / fc1 / fc2 / head) and their autograd
// backward.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>

#include "gemm_sm100.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kBlockM = 128;  // rows per CTA (cluster tile = 256)
constexpr int kBlockK = 64;   // 64 bf16 = 128 B = one swizzle atom
constexpr int kUmmaK = 16;
constexpr int kNumThreads = 384;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc + CLC scheduler, warp3 AG copier, warps4-11 epilogue
constexpr int kEpiThreads = 128;
constexpr int kCdBufs = 4;             // ring of 128x64 bf16 staging buffers for TMA stores
constexpr int kCdBufBytes = 128 * 128;  // 128 rows x 128 B

struct KernelParams {
    int M, N, K;
    int batch, nb_inner;
    int m_tiles, n_tiles;  // cluster tiles (256 x BLOCK_N)
    int n_rot;             // n-tile rotation so that tiles are visited in slab-arrival order (AG fusion)
    int group_n;           // n-tiles per raster group
    int use_clc;           // 1 = dynamic tile scheduling through cluster launch control (work stealing)
    int hint_a, hint_b;    // L2 eviction hints for the operand loads: 0 normal, 1 evict-first, 2 evict-last
    GemmEpilogue epi;
    GemmAgFuse ag;
};

constexpr int kAgChunkBytes = 16384;

__device__ __forceinline__ uint4 ld_peer_v4(const void* ptr) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(ptr)
                 : "memory");
    return r;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* ptr, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* ptr) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ int total_tiles_fwd(const KernelParams& p) { return p.m_tiles * p.n_tiles * p.batch; }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// erf via Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below bf16 resolution): 1 rcp + 1 exp + 6 FMA.
// e = exp(-z^2) is returned too: for z = x/sqrt(2) it is exactly the Gaussian factor gelu'(x) needs.
__device__ __forceinline__ float erf_as(float z, float& e) {
    const float az = fabsf(z);
    const float t = __frcp_rn(fmaf(0.3275911f, az, 1.0f));
    e = __expf(-az * az);
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float y = 1.0f - poly * t * e;
    return copysignf(y, z);
}
__device__ __forceinline__ float gelu_erf(float x) {
    float e;
    return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f, e));
}
__device__ __forceinline__ float dgelu_erf(float x) {
    float e;
    const float cdf = 0.5f * (1.0f + erf_as(x * 0.70710678118654752f, e));
    return fmaf(x * 0.3989422804014327f, e, cdf);  // cdf + x * pdf,  pdf = exp(-x^2/2)/sqrt(2 pi)
}

template <int kMajorA, int kMajorB, int BLOCK_N, int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
    gemm_bf16_sm100_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                           const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_aux,
                           const KernelParams p) {
    constexpr int LOAD_N = BLOCK_N / 2;  // B rows loaded per CTA
    constexpr int kABytes = kBlockM * kBlockK * 2;
    constexpr int kBBytes = LOAD_N * kBlockK * 2;
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr int kTmemCols = (2 * BLOCK_N <= 32)    ? 32
                              : (2 * BLOCK_N <= 64)  ? 64
                              : (2 * BLOCK_N <= 128) ? 128
                              : (2 * BLOCK_N <= 256) ? 256
                                                     : 512;
    static_assert(2 * BLOCK_N <= 512, "accumulator double buffer must fit TMEM");
    static_assert(BLOCK_N % 64 == 0, "epilogue works in 64-column chunks");
    static_assert(kMajorB == 0 || LOAD_N % 64 == 0, "MN-major B needs whole 64-wide atoms per CTA");
    static_assert(kABytes % 1024 == 0 && kBBytes % 1024 == 0, "swizzle-128B needs 1024 B aligned stages");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_cd = smem;
    uint8_t* smem_a = smem + kCdBufs * kCdBufBytes;
    uint8_t* smem_b = smem_a + kStages * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kStages * kBBytes);
    uint64_t* full_bar = bars;                     // [kStages]   TMA -> MMA   (leader CTA's copy is used)
    uint64_t* empty_bar = bars + kStages;          // [kStages]   MMA -> TMA   (both CTAs)
    uint64_t* tmem_full_bar = bars + 2 * kStages;  // [2]         MMA -> epilogue (both CTAs)
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]         epilogue -> MMA (leader CTA's copy)
    constexpr int kClcStages = 2;  // reserve at most ~1 tile ahead so co-running clusters stay aligned
    uint64_t* clc_full_bar = tmem_empty_bar + 2;            // [kClcStages] scheduler (HW) -> consumers, every CTA
    uint64_t* clc_empty_bar = clc_full_bar + kClcStages;    // [kClcStages] consumers -> scheduler (leader CTA's copy)
    uint4* clc_resp = reinterpret_cast<uint4*>(clc_empty_bar + kClcStages);  // [kClcStages] 16 B responses
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(clc_resp + kClcStages);

    const uint32_t warp_idx = threadIdx.x / 32;
    const uint32_t lane = lane_id();
    const uint32_t cta_rank = cluster_ctarank();
    const bool is_leader = cta_rank == 0;

    if (warp_idx == 0 && elect_one()) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_d);
        if (p.epi.has_aux_out) prefetch_tmap(&tmap_aux);
    }
    if (warp_idx == 1 && elect_one()) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);   // leader's expect_tx arrive (covers both CTAs' bytes)
            mbar_init(&empty_bar[i], 1);  // one multicast tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full_bar[i], 1);   // one multicast tcgen05.commit
            mbar_init(&tmem_empty_bar[i], 16);  // 8 epilogue warps x 2 CTAs
        }
        for (int i = 0; i < kClcStages; ++i) {
            mbar_init(&clc_full_bar[i], 1);    // the scheduler's expect_tx arrive (+16 B from the hardware)
            // leader: producer + MMA + 8 epilogue warps + scheduler; peer: producer + 8 epilogue warps
            mbar_init(&clc_empty_bar[i], 20);
        }
        fence_mbar_init();
    }
    cluster_sync();  // barriers visible cluster-wide before anyone touches a peer's barrier / TMEM alloc
    if (warp_idx == 2) tmem_alloc<2>(tmem_ptr_smem, kTmemCols);
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int num_clusters = gridDim.x / 2;
    const int cluster_id = blockIdx.x / 2;
    // Tile iteration.  Static mode: tile = cluster_id + i * num_clusters.  CLC mode: the grid has one cluster per
    // tile; a resident cluster computes its own tile and then keeps *cancelling* not-yet-launched clusters and
    // computing their tiles (clusterlaunchcontrol.try_cancel).  The hardware hands tiles out in launch order, so
    // the tiles in flight are always a contiguous window of the raster -> co-running CTAs share A/B panels in L2
    // (2.4x less DRAM traffic than a static round-robin whose clusters drift apart over ~100 tiles).
    struct TileIter {
        int tile;
        uint32_t stage, phase;
    };
    auto tile_begin = [&]() { return TileIter{cluster_id, 0u, 0u}; };
    auto tile_next = [&](TileIter& it, bool arrive_lane) -> bool {
        if (!p.use_clc) {
            it.tile += num_clusters;
            return it.tile < total_tiles_fwd(p);
        }
        mbar_wait(&clc_full_bar[it.stage], it.phase);
        uint32_t valid, cx, cy, cz;
        asm volatile(
            "{\n\t"
            ".reg .pred p1;\n\t"
            ".reg .b128 resp;\n\t"
            "ld.shared.b128 resp, [%4];\n\t"
            "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, resp;\n\t"
            "selp.u32 %3, 1, 0, p1;\n\t"
            "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, resp;\n\t"
            "}\n"
            : "=r"(cx), "=r"(cy), "=r"(cz), "=r"(valid)
            : "r"(smem_u32(&clc_resp[it.stage]))
            : "memory");
        fence_proxy_async_smem();  // our generic read of the slot before the async proxy may overwrite it
        if (arrive_lane) mbar_arrive_cluster(&clc_empty_bar[it.stage], 0);
        it.stage = (it.stage + 1 == kClcStages) ? 0 : it.stage + 1;
        it.phase ^= (it.stage == 0);
        it.tile = static_cast<int>(cx) / 2;
        return valid != 0;
    };
    const int tiles_per_batch = p.m_tiles * p.n_tiles;
    const int total_tiles = tiles_per_batch * p.batch;
    const int num_kb = (p.K + kBlockK - 1) / kBlockK;
    const int kGroupN = p.group_n;  // n-tiles per raster group (keeps a wave's A/B footprint L2-resident)

    auto decode_tile = [&](int t, int& b, int& mt, int& nt) {
        b = t / tiles_per_batch;
        const int r = t - b * tiles_per_batch;
        const int per_group = p.m_tiles * kGroupN;
        const int g = r / per_group;
        const int first_n = g * kGroupN;
        const int gsz = min(kGroupN, p.n_tiles - first_n);
        const int in_g = r - g * per_group;
        mt = in_g / gsz;
        nt = first_n + in_g % gsz;
        nt += p.n_rot;  // AG fusion: start with the n-tiles of the locally owned slab
        if (nt >= p.n_tiles) nt -= p.n_tiles;
    };
    const int ag_chunks_per_slab =
        p.ag.world > 1 ? static_cast<int>((p.ag.slab_bytes + kAgChunkBytes - 1) / kAgChunkBytes) : 0;

    if (warp_idx == 0) {
        // ================================= TMA producer =================================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            const uint64_t pol_a = p.hint_a == 1 ? kL2EvictFirst : (p.hint_a == 2 ? kL2EvictLast : kL2EvictNormal);
            const uint64_t pol_b = p.hint_b == 1 ? kL2EvictFirst : (p.hint_b == 2 ? kL2EvictLast : kL2EvictNormal);
            TileIter it = tile_begin();
            for (bool more = it.tile < total_tiles; more; more = tile_next(it, true)) {
                const int t = it.tile;
                int b, mt, nt;
                decode_tile(t, b, mt, nt);
                const int bi = b % p.nb_inner, bo = b / p.nb_inner;
                const int m_idx = mt * (2 * kBlockM) + cta_rank * kBlockM;
                const int n_idx = nt * BLOCK_N + cta_rank * LOAD_N;
                if (p.ag.world > 1) {
                    // B rows [n_idx, n_idx + LOAD_N) must have been pulled into the local gathered buffer
                    const int last_row = min(n_idx + LOAD_N, p.N) - 1;
                    if (last_row >= n_idx) {
                        const int s_lo = min(n_idx / p.ag.rows_per_slab, p.ag.world - 1);
                        const int s_hi = min(last_row / p.ag.rows_per_slab, p.ag.world - 1);
                        for (int sl = s_lo; sl <= s_hi; ++sl) {
                            uint32_t spins = 0;
                            while (ld_acquire_gpu(p.ag.flags + sl) < static_cast<uint32_t>(ag_chunks_per_slab)) {
                                if (++spins > (1u << 26)) {
                                    printf("[b200] AG-fused GEMM: slab %d never arrived (block %d)\n", sl, blockIdx.x);
                                    __trap();
                                }
                            }
                        }
                        fence_proxy_async_all();  // generic-proxy peer copies -> async-proxy (TMA) reads
                    }
                }
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const int k_idx = kb * kBlockK;
                    uint8_t* sa = smem_a + stage * kABytes;
                    uint8_t* sb = smem_b + stage * kBBytes;
                    if constexpr (kMajorA == 0) {
                        tma_load_4d_2cta_hint(&tmap_a, &full_bar[stage], sa, k_idx, m_idx, bi, bo, pol_a);
                    } else {
#pragma unroll
                        for (int i = 0; i < kBlockM / 64; ++i)
                            tma_load_4d_2cta_hint(&tmap_a, &full_bar[stage], sa + i * (64 * kBlockK * 2), m_idx + i * 64,
                                                  k_idx, bi, bo, pol_a);
                    }
                    if constexpr (kMajorB == 0) {
                        tma_load_4d_2cta_hint(&tmap_b, &full_bar[stage], sb, k_idx, n_idx, bi, bo, pol_b);
                    } else {
#pragma unroll
                        for (int i = 0; i < LOAD_N / 64; ++i)
                            tma_load_4d_2cta_hint(&tmap_b, &full_bar[stage], sb + i * (64 * kBlockK * 2), n_idx + i * 64,
                                                  k_idx, bi, bo, pol_b);
                    }
                    // Only the leader arms the barrier, for the bytes of BOTH CTAs.  The peer's complete_tx may
                    // land first (tx-count goes transiently negative, which is legal); it can never leak into
                    // the next phase because the peer only refills a slot after the MMA that consumed it
                    // committed.  (A remote arrive here costs a GPU-scope membar per k-block.)
                    if (is_leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageBytes);
                    stage = (stage + 1 == kStages) ? 0 : stage + 1;
                    phase ^= (stage == 0);
                }
            }
        }
    } else if (warp_idx == 1) {
        // ================================= MMA issuer (leader CTA only) =================================
        if (is_leader) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * kBlockM, BLOCK_N, kMajorA, kMajorB);
            // K-major : 8-row groups are 1024 B apart (SBO); one swizzle atom along K so LBO unused.
            // MN-major: 8-k groups are 1024 B apart (SBO); 64-wide MN atoms are BLOCK_K*128 B apart (LBO).
            constexpr uint32_t kLboA = kMajorA == 0 ? 0 : kBlockK * 128;
            constexpr uint32_t kLboB = kMajorB == 0 ? 0 : kBlockK * 128;
            constexpr uint32_t kKStepA = kMajorA == 0 ? (kUmmaK * 2) : (kUmmaK * 128);  // bytes per UMMA_K
            constexpr uint32_t kKStepB = kMajorB == 0 ? (kUmmaK * 2) : (kUmmaK * 128);
            uint32_t stage = 0, phase = 0;
            uint32_t iter = 0;
            TileIter it = tile_begin();
            for (bool more = it.tile < total_tiles; more; more = tile_next(it, lane == 0), ++iter) {
                const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BLOCK_N;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_addr = smem_u32(smem_a + stage * kABytes);
                        const uint32_t b_addr = smem_u32(smem_b + stage * kBBytes);
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            const uint64_t da = make_smem_desc_sw128(a_addr + k * kKStepA, kLboA, 1024);
                            const uint64_t db = make_smem_desc_sw128(b_addr + k * kKStepB, kLboB, 1024);
                            umma_bf16<2>(tmem_d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit<2>(&empty_bar[stage]);  // frees the smem slot in both CTAs
                        if (kb == num_kb - 1) umma_commit<2>(&tmem_full_bar[as]);
                    }
                    __syncwarp();
                    stage = (stage + 1 == kStages) ? 0 : stage + 1;
                    phase ^= (stage == 0);
                }
            }
            // Drain: make sure every epilogue (both CTAs) released the last accumulators before teardown.
            if (iter > 0) {
                const uint32_t last = iter - 1;
                mbar_wait(&tmem_empty_bar[last & 1], (last >> 1) & 1);
                if (iter > 1) {
                    const uint32_t prev = iter - 2;
                    mbar_wait(&tmem_empty_bar[prev & 1], (prev >> 1) & 1);
                }
            }
        }
    } else if (warp_idx == 2) {
        // ================================= CLC tile scheduler (leader CTA) =================================
        if (p.use_clc && is_leader) {
            uint32_t stage = 0, phase = 0;
            TileIter it = tile_begin();  // the scheduler also consumes its own responses to know when to stop
            bool more = it.tile < total_tiles;
            while (more) {
                mbar_wait(&clc_empty_bar[stage], phase ^ 1);
                if (lane < 2) {  // arm the response barrier of both CTAs of the pair (16 B each)
                    asm volatile(
                        "{\n\t"
                        ".reg .b32 raddr;\n\t"
                        "mapa.shared::cluster.u32 raddr, %0, %1;\n\t"
                        "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [raddr], 16;\n\t"
                        "}\n" ::"r"(smem_u32(&clc_full_bar[stage])),
                        "r"(lane)
                        : "memory");
                }
                __syncwarp();
                if (elect_one()) {
                    asm volatile(
                        "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes"
                        ".multicast::cluster::all.b128 [%0], [%1];" ::"r"(smem_u32(&clc_resp[stage])),
                        "r"(smem_u32(&clc_full_bar[stage]))
                        : "memory");
                }
                __syncwarp();
                stage = (stage + 1 == kClcStages) ? 0 : stage + 1;
                phase ^= (stage == 0);
                more = tile_next(it, lane == 0);
            }
        }
    } else if (warp_idx == 3) {
        // ================================= All-gather copier (AG fusion only) =================================
        if (p.ag.world > 1) {
            const int total_chunks = p.ag.world * ag_chunks_per_slab;
            uint32_t* chunk_counter = p.ag.flags + p.ag.world;  // zeroed with the flags
            for (;;) {
                int c = 0;
                if (lane == 0) c = static_cast<int>(atomicAdd(chunk_counter, 1u));
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= total_chunks) break;
                const int k = c / ag_chunks_per_slab;            // arrival index: 0 = own slab
                const int sl = (p.ag.rank + k) % p.ag.world;      // slab pulled now (ranks start at different peers)
                const int64_t off = static_cast<int64_t>(c - k * ag_chunks_per_slab) * kAgChunkBytes;
                const int64_t nbytes = min(static_cast<int64_t>(kAgChunkBytes), p.ag.slab_bytes - off);
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.ag.peer_src[sl]) + off;
                uint8_t* dst = static_cast<uint8_t*>(p.ag.dst) + static_cast<int64_t>(sl) * p.ag.slab_bytes + off;
                const int nvec = static_cast<int>(nbytes / 16);
                int i = lane;
                for (; i + 7 * 32 < nvec; i += 8 * 32) {
                    uint4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) v[u] = ld_peer_v4(src + static_cast<int64_t>(i + u * 32) * 16);
#pragma unroll
                    for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(dst + static_cast<int64_t>(i + u * 32) * 16) = v[u];
                }
                for (; i < nvec; i += 32) *reinterpret_cast<uint4*>(dst + static_cast<int64_t>(i) * 16) = ld_peer_v4(src + static_cast<int64_t>(i) * 16);
                __threadfence();
                __syncwarp();
                if (lane == 0) red_release_gpu_add(p.ag.flags + sl, 1);
            }
        }
    } else if (warp_idx >= 4) {
        // ================================= Epilogue warps =================================
        // Two groups of four warps.  A warp may only touch the TMEM lane quarter (warp_idx % 4), so both groups
        // cover all 128 rows and split the tile by 64-column chunks (group g takes chunks g, g+2): the epilogue of a
        // tile takes half as long, which matters for the fat epilogues (dGELU, two outputs) and for the short-K
        // attention GEMMs where the epilogue, not the MMA, sets the pace.  Each chunk is processed as two 32-column
        // halves written straight into the swizzled staging buffer, which keeps the register count under the
        // 168/thread a 384-thread CTA allows.
        constexpr int kChunks = BLOCK_N / 64;
        static_assert(kChunks >= 2, "both epilogue groups need at least one chunk");
        const uint32_t eg = (warp_idx - 4) >> 2;       // epilogue group
        const uint32_t ew = warp_idx & 3;              // TMEM lane quarter of this warp
        const uint32_t etid = (threadIdx.x - 128) & 127;
        const uint32_t row_in_tile = ew * 32 + lane;   // accumulator row owned by this thread
        const uint32_t bar_id = 1 + eg;
        uint8_t* const grp_buf = smem_cd + eg * 2 * kCdBufBytes;  // two staging buffers per group
        const GemmEpilogue& e = p.epi;
        const bool ext_is_aux = e.act == kActDGelu;
        uint32_t iter = 0;
        uint32_t flip = 0;
        TileIter it = tile_begin();
        for (bool more = it.tile < total_tiles; more; more = tile_next(it, lane == 0), ++iter) {
            const int t = it.tile;
            int b, mt, nt;
            decode_tile(t, b, mt, nt);
            const int bi = b % p.nb_inner, bo = b / p.nb_inner;
            const int m0 = mt * (2 * kBlockM) + cta_rank * kBlockM;
            const int n0 = nt * BLOCK_N;
            const int row = m0 + row_in_tile;
            const bool row_ok = row < p.M;
            const uint32_t as = iter & 1, aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((ew * 32) << 16) + as * BLOCK_N;
            const __nv_bfloat16* ext_row = nullptr;
            if (ext_is_aux) {
                ext_row = e.aux_in + static_cast<int64_t>(row) * e.ld_aux;
            } else if (e.residual != nullptr) {
                const int rrow = e.res_row_mod > 0 ? row % e.res_row_mod : row;
                ext_row = e.residual + static_cast<int64_t>(rrow) * e.ld_res;
            }

#pragma unroll 1
            for (int c = eg; c < kChunks; c += 2) {
                const int ncol0 = n0 + c * 64;
                const bool last_chunk = c + 2 >= kChunks;  // last TMEM read of this warp for the tile
                if (ncol0 >= p.N) {  // tile-uniform: nothing to store, but the TMEM stage must still be released
                    if (last_chunk) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
                    }
                    continue;
                }
                // staging buffer(s) of this group free again?
                uint8_t* buf0 = grp_buf + (e.has_aux_out ? 0 : (flip & 1)) * kCdBufBytes;
                uint8_t* buf1 = grp_buf + kCdBufBytes;
                if (etid == 0) {
                    if (e.has_aux_out)
                        tma_store_wait_read<0>();
                    else
                        tma_store_wait_read<1>();
                }
                named_bar_sync(bar_id, kEpiThreads);
                const uint32_t rbase = smem_u32(buf0) + row_in_tile * 128;
                const uint32_t rbase1 = smem_u32(buf1) + row_in_tile * 128;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t acc[32];
                    tmem_ld_32x32b_x32(taddr + c * 64 + h * 32, acc);
                    // per-row epilogue inputs (dGELU pre-activation or residual) are fetched while the TMEM load flies
                    uint4 ext[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col = ncol0 + h * 32 + j * 8;
                        ext[j] = (ext_row != nullptr && row_ok && col < p.N)
                                     ? *reinterpret_cast<const uint4*>(ext_row + col)
                                     : make_uint4(0, 0, 0, 0);
                    }
                    tmem_ld_wait();
                    if (last_chunk && h == 1) {
                        // accumulators are in registers: hand the TMEM stage back to the MMA warp early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int col = ncol0 + h * 32 + j * 8;
                        const bool col_ok = col < p.N;
                        float v[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(acc[j * 8 + q]);
                        if (e.bias != nullptr && col_ok) {
                            const uint4 bv = __ldg(reinterpret_cast<const uint4*>(e.bias + col));
                            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                v[2 * q] += bf16_lo(bw[q]);
                                v[2 * q + 1] += bf16_hi(bw[q]);
                            }
                        }
                        const uint32_t off = (((h * 4 + j) ^ (row_in_tile & 7)) * 16);
                        if (e.has_aux_out)
                            st_shared_v4(rbase1 + off, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                         pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                        if (e.act == kActGelu) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) v[q] = gelu_erf(v[q]);
                        } else if (ext_is_aux) {
                            const uint32_t uw[4] = {ext[j].x, ext[j].y, ext[j].z, ext[j].w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                v[2 * q] *= dgelu_erf(bf16_lo(uw[q]));
                                v[2 * q + 1] *= dgelu_erf(bf16_hi(uw[q]));
                            }
                        }
                        if (e.residual != nullptr && !ext_is_aux) {
                            const uint32_t rw[4] = {ext[j].x, ext[j].y, ext[j].z, ext[j].w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                v[2 * q] += bf16_lo(rw[q]);
                                v[2 * q + 1] += bf16_hi(rw[q]);
                            }
                        }
                        if (!(row_ok && col_ok)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) v[q] = 0.f;  // keeps the column sums clean
                        }
                        st_shared_v4(rbase + off, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                     pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    }
                }
                // ---- swizzled staging buffer -> global through TMA ----
                fence_proxy_async_smem();
                named_bar_sync(bar_id, kEpiThreads);
                if (etid == 0) {
                    tma_store_4d(&tmap_d, buf0, ncol0, m0, bi, bo);
                    tma_store_commit();
                    if (e.has_aux_out) {
                        tma_store_4d(&tmap_aux, buf1, ncol0, m0, bi, bo);
                        tma_store_commit();
                    }
                }
                if (e.colsum != nullptr) {
                    // Bias gradient: column sums of the (bf16-rounded) output tile, fp32 atomics.
                    const int ccol = etid & 63;       // column within the chunk
                    const int rhalf = etid >> 6;      // 0/1 -> rows [0,64) / [64,128)
                    const int jj = ccol >> 3, within = ccol & 7;
                    float s = 0.f;
#pragma unroll 8
                    for (int r = 0; r < 64; ++r) {
                        const int rr = rhalf * 64 + r;
                        const uint8_t* ptr = buf0 + rr * 128 + ((jj ^ (rr & 7)) * 16) + within * 2;
                        s += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(ptr));
                    }
                    if (ncol0 + ccol < p.N)
                        atomicAdd(e.colsum + static_cast<int64_t>(bi) * e.colsum_bi_stride + ncol0 + ccol, s);
                }
                ++flip;
            }
        }
        if (etid == 0) tma_store_wait<0>();
    }

    tc_fence_before();
    cluster_sync();
    if (warp_idx == 2) tmem_dealloc<2>(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
    static EncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
        if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || sym == nullptr)
            throw std::runtime_error("cuTensorMapEncodeTiled not available from the driver");
        fn = reinterpret_cast<EncodeFn>(sym);
    });
    return fn;
}

struct TmapKey {
    uint64_t ptr;
    int64_t d[4];
    int64_t s[3];
    int32_t box[2];
    bool operator==(const TmapKey& o) const { return std::memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
        size_t h = 1469598103934665603ull;
        for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
        return h;
    }
};

// 4-D bf16 tensor map: dims (inner, outer, batch_inner, batch_outer), box (box_inner, box_outer, 1, 1).
CUtensorMap make_tmap(const GemmOperand& op, int64_t inner, int64_t outer, int box_inner, int box_outer,
                      int swizzle_bytes = 128) {
    static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
    static std::mutex mu;
    TmapKey key;
    std::memset(&key, 0, sizeof(key));
    key.ptr = reinterpret_cast<uint64_t>(op.ptr);
    key.d[0] = inner, key.d[1] = outer, key.d[2] = op.nb_inner, key.d[3] = op.nb_outer;
    key.s[0] = op.ld, key.s[1] = op.stride_b_inner, key.s[2] = op.stride_b_outer;
    key.box[0] = box_inner, key.box[1] = box_outer + (swizzle_bytes << 16);
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
    }
    if ((reinterpret_cast<uint64_t>(op.ptr) & 15) != 0) throw std::runtime_error("gemm: operand base must be 16 B aligned");
    if ((op.ld * 2) % 16 != 0) throw std::runtime_error("gemm: leading dimension must be a multiple of 8 elements");
    CUtensorMap tm;
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer),
                          static_cast<cuuint64_t>(op.nb_inner), static_cast<cuuint64_t>(op.nb_outer)};
    // Strides of size-1 batch dims are irrelevant but must still be legal multiples of 16 B.
    const int64_t sbi = op.nb_inner > 1 ? op.stride_b_inner : op.ld * outer;
    const int64_t sbo = op.nb_outer > 1 ? op.stride_b_outer : sbi * op.nb_inner;
    if ((sbi * 2) % 16 != 0 || (sbo * 2) % 16 != 0) throw std::runtime_error("gemm: batch strides must be multiples of 8 elements");
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(op.ld * 2), static_cast<cuuint64_t>(sbi * 2),
                             static_cast<cuuint64_t>(sbo * 2)};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult res = get_encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(op.ptr), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                         : CU_TENSOR_MAP_SWIZZLE_32B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (res != CUDA_SUCCESS) {
        char msg[256];
        snprintf(msg, sizeof(msg),
                 "cuTensorMapEncodeTiled failed (%d): dims=(%lld,%lld,%lld,%lld) ld=%lld box=(%d,%d)", (int)res,
                 (long long)inner, (long long)outer, (long long)op.nb_inner, (long long)op.nb_outer, (long long)op.ld,
                 box_inner, box_outer);
        throw std::runtime_error(msg);
    }
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, tm);
    return tm;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

template <int kMajorA, int kMajorB, int BLOCK_N, int kStages>
void launch(const GemmOperand& A, const GemmOperand& B, const GemmOperand& D, const GemmOperand* aux, int M, int N,
            int K, const GemmEpilogue& epi, int max_ctas, cudaStream_t stream, const GemmAgFuse* ag) {
    constexpr int LOAD_N = BLOCK_N / 2;
    constexpr int kSmem = kCdBufs * kCdBufBytes + kStages * (kBlockM * kBlockK * 2 + LOAD_N * kBlockK * 2) +
                          (2 * kStages + 4 + 8) * 8 + 64 + 16;
    static_assert(kSmem <= 232448, "shared memory budget exceeded");
    auto kern = gemm_bf16_sm100_kernel<kMajorA, kMajorB, BLOCK_N, kStages>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (err != cudaSuccess) throw std::runtime_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
        attr_set = true;
    }
    // A: K-major -> (inner=K, outer=M), box (64, 128).  MN-major -> (inner=M, outer=K), box (64, 64).
    CUtensorMap ta = kMajorA == 0 ? make_tmap(A, K, M, kBlockK, kBlockM) : make_tmap(A, M, K, 64, kBlockK);
    CUtensorMap tb = kMajorB == 0 ? make_tmap(B, K, N, kBlockK, LOAD_N) : make_tmap(B, N, K, 64, kBlockK);
    CUtensorMap td = make_tmap(D, N, M, 64, kBlockM);
    CUtensorMap tx = (epi.has_aux_out && aux != nullptr) ? make_tmap(*aux, N, M, 64, kBlockM) : td;

    KernelParams p;
    p.M = M, p.N = N, p.K = K;
    p.batch = static_cast<int>(D.nb_inner * D.nb_outer);
    p.nb_inner = static_cast<int>(D.nb_inner);
    p.m_tiles = (M + 2 * kBlockM - 1) / (2 * kBlockM);
    p.n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
    p.epi = epi;
    p.n_rot = 0;
    {
        // raster / cache-hint knobs (tuned defaults; env overrides are for experiments only)
        static const int env_group = getenv("B200_GEMM_GROUP_N") ? atoi(getenv("B200_GEMM_GROUP_N")) : 0;
        static const int env_ha = getenv("B200_GEMM_HINT_A") ? atoi(getenv("B200_GEMM_HINT_A")) : -1;
        static const int env_hb = getenv("B200_GEMM_HINT_B") ? atoi(getenv("B200_GEMM_HINT_B")) : -1;
        static const int env_clc = getenv("B200_GEMM_CLC") ? atoi(getenv("B200_GEMM_CLC")) : 1;
        // dynamic scheduling pays for long tiles; for short-K (attention) tiles the try_cancel round trip would be
        // on the critical path of every ~1 us tile, so those keep the static persistent schedule
        p.use_clc = env_clc && K >= 1024;
        // n-tiles per raster group: the B panel of a group (group_n x 256 rows x K) should stay L2-resident
        // (~40 MB) while the A panels stream past it.  Measured DRAM reads, qkv (K=5120): 8->2.9 GB, 16->2.0 GB.
        // Deep-K panels (K >= 8192) cannot stay resident anyway: use a square-ish wave (74 tiles ~ 9 x 8) so that
        // every panel fetched from DRAM is shared by as many co-running clusters as possible
        // (wgrad K=32768: group 2 -> 17 GB of DRAM reads, group 8 -> ~7 GB).
        const int g = (K > 0 && 81920 / K >= 12) ? 16 : 8;
        p.group_n = env_group > 0 ? env_group : g;
        p.hint_a = env_ha >= 0 ? env_ha : 0;
        p.hint_b = env_hb >= 0 ? env_hb : 0;
    }
    if (ag != nullptr && ag->world > 1) {
        if (kMajorB != 0 || p.batch != 1) throw std::runtime_error("gemm: AG fusion needs a K-major, un-batched B");
        if (static_cast<int64_t>(ag->rows_per_slab) * K * 2 != ag->slab_bytes || ag->slab_bytes % 16 != 0)
            throw std::runtime_error("gemm: AG fusion needs whole rows per slab");
        p.ag = *ag;
        p.n_rot = static_cast<int>((static_cast<int64_t>(ag->rank) * ag->rows_per_slab) / BLOCK_N) % p.n_tiles;
        cudaMemsetAsync(ag->flags, 0, sizeof(uint32_t) * (ag->world + 1), stream);  // slab counters + chunk counter
    }
    const int64_t total = static_cast<int64_t>(p.m_tiles) * p.n_tiles * p.batch;
    int sms = num_sms();
    if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
    int clusters = static_cast<int>(std::min<int64_t>(total, sms / 2));
    if (clusters < 1) clusters = 1;
    if (max_ctas > 0) p.use_clc = 0;  // an SM carve-out needs a bounded persistent grid
    if (p.use_clc) clusters = static_cast<int>(std::min<int64_t>(total, 1 << 22));  // one cluster per tile

    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    cfg.blockDim = dim3(kNumThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 2;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, ta, tb, td, tx, p);
    if (err != cudaSuccess) throw std::runtime_error(std::string("gemm launch failed: ") + cudaGetErrorString(err));
}

}  // namespace

CUtensorMap make_tensor_map_4d(const GemmOperand& op, int64_t inner, int64_t outer, int box_inner, int box_outer,
                               int swizzle_bytes) {
    return make_tmap(op, inner, outer, box_inner, box_outer, swizzle_bytes);
}

void gemm_bf16(const GemmOperand& A, int major_a, const GemmOperand& B, int major_b, const GemmOperand& D,
               const GemmOperand* aux_out, int M, int N, int K, const GemmEpilogue& epi, int block_n, int max_ctas,
               cudaStream_t stream, const GemmAgFuse* ag) {
    if (N % 8 != 0 && (epi.residual || epi.aux_in))
        throw std::runtime_error("gemm: N must be a multiple of 8 when residual/aux_in are used");
    if (epi.act == kActDGelu && (epi.residual != nullptr || epi.aux_in == nullptr))
        throw std::runtime_error("gemm: dGELU epilogue needs aux_in and cannot be combined with a residual");
    if (D.nb_inner * D.nb_outer > 1 && (epi.bias || epi.residual || epi.aux_in))
        throw std::runtime_error("gemm: bias/residual/aux_in are not supported for batched problems");
    if (block_n == 0) block_n = (N > 128) ? 256 : 128;
#define B200_DISPATCH(MA, MB, BN, ST)                                                             \
    if (major_a == MA && major_b == MB && block_n == BN) {                                        \
        launch<MA, MB, BN, ST>(A, B, D, aux_out, M, N, K, epi, max_ctas, stream, ag);             \
        return;                                                                                   \
    }
    B200_DISPATCH(0, 0, 256, 5)
    B200_DISPATCH(0, 1, 256, 5)
    B200_DISPATCH(1, 1, 256, 5)
    B200_DISPATCH(1, 0, 256, 5)
    B200_DISPATCH(0, 0, 128, 6)
    B200_DISPATCH(0, 1, 128, 6)
    B200_DISPATCH(1, 1, 128, 6)
    B200_DISPATCH(1, 0, 128, 6)
#undef B200_DISPATCH
    throw std::runtime_error("gemm: unsupported (major_a, major_b, block_n) combination");
}

}  // namespace b200
