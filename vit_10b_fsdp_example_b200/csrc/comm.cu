// NVLink 5 / NVSwitch collectives over symmetric (peer-mapped) memory, written directly against raw
// peer / multicast pointers -- no NCCL on these paths.
//
//   p2p_all_gather      : sync-free pull of the peers' parameter shards straight into their final
//                         position in the gathered flat buffer (no copy-out pass).  Shards only change in
//                         the optimizer step, so no per-block flags are needed.
//   p2p_reduce_scatter  : every rank pulls *its* slab of each peer's gradient buffer, reduces in fp32,
//                         applies the 1/W mean, writes the fp32 shard gradient and accumulates the
//                         sum-of-squares partial for the global grad norm -- one kernel, one pass.
//   nvls_reduce_scatter : same contract, but the reduction happens inside the NVSwitch
//                         (multimem.ld_reduce on the multicast address): 1x ingress instead of (W-1)x.
//   signal_barrier      : device-side barrier through flags in symmetric memory (st.release.sys /
//                         ld.acquire.sys), monotonically increasing sequence numbers, bounded spin.
//   allreduce_scalars   : W x K floats exchanged through symmetric scratch (grad-norm^2, loss, max time).
//
// Capability parity: XLA all_gather / reduce_scatter / all_reduce emitted by XlaFullyShardedDataParallel
// (reference run_vit_training.py:177-181, 261-270) and xm.mesh_reduce (run_vit_training.py:205).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "comm.h"
#include "ptx.cuh"

namespace b200 {

namespace {

constexpr int kMaxWorld = 16;
constexpr int kCommThreads = 512;

struct PeerPtrs {
    uint64_t p[kMaxWorld];
};

__device__ __forceinline__ uint4 ld_stream_v4(const void* ptr) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(ptr)
                 : "memory");
    return r;
}
__device__ __forceinline__ void st_stream_v4(void* ptr, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_release_sys(uint32_t* ptr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* ptr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* ptr) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_f32(float* ptr, float v) {
    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(ptr), "f"(v) : "memory");
}
// In-switch reduction of 8 bf16 values (fp32 accumulate) over all GPUs bound to the multicast object.
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc_ptr) {
    uint4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(mc_ptr)
                 : "memory");
    return r;
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int64_t kAgChunkBytes = 65536;
constexpr int64_t kRsChunkElems = 16384;

// seg_table row (all-gather): [src_rank, src_off_bytes, dst_off_bytes, nbytes, chunk_prefix]
__global__ void __launch_bounds__(kCommThreads) p2p_all_gather_kernel(PeerPtrs peers, uint8_t* __restrict__ out,
                                                                      const int64_t* __restrict__ seg, int nseg,
                                                                      int64_t total_chunks) {
    int s = 0;
    for (int64_t c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        while (s + 1 < nseg && seg[(s + 1) * 5 + 4] <= c) ++s;
        const int64_t* row = seg + s * 5;
        const int64_t local_chunk = c - row[4];
        const int64_t off = local_chunk * kAgChunkBytes;
        const int64_t nbytes = min(kAgChunkBytes, row[3] - off);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(peers.p[row[0]]) + row[1] + off;
        uint8_t* dst = out + row[2] + off;
        const int64_t nvec = nbytes / 16;
        // 4 independent 16 B loads in flight per thread
        int64_t i = threadIdx.x;
        for (; i + 3 * kCommThreads < nvec; i += 4 * kCommThreads) {
            const uint4 a = ld_stream_v4(src + i * 16);
            const uint4 b = ld_stream_v4(src + (i + kCommThreads) * 16);
            const uint4 cc = ld_stream_v4(src + (i + 2 * kCommThreads) * 16);
            const uint4 d = ld_stream_v4(src + (i + 3 * kCommThreads) * 16);
            st_stream_v4(dst + i * 16, a);
            st_stream_v4(dst + (i + kCommThreads) * 16, b);
            st_stream_v4(dst + (i + 2 * kCommThreads) * 16, cc);
            st_stream_v4(dst + (i + 3 * kCommThreads) * 16, d);
        }
        for (; i < nvec; i += kCommThreads) st_stream_v4(dst + i * 16, ld_stream_v4(src + i * 16));
    }
}

// seg_table row (reduce-scatter): [full_off_bytes, shard_off_elems, nelems, chunk_prefix]
template <bool kBf16In, bool kNvls, bool kAdam>
__global__ void __launch_bounds__(kCommThreads) reduce_scatter_kernel(PeerPtrs peers, uint64_t mc_base, int rank,
                                                                      int world, float* __restrict__ out,
                                                                      const int64_t* __restrict__ seg, int nseg,
                                                                      int64_t total_chunks, float scale,
                                                                      float* __restrict__ sumsq_out, AdamFuse adam) {
    __shared__ float red[kCommThreads / 32];
    float sq = 0.f;
    int s = 0;
    constexpr int kVec = kBf16In ? 8 : 4;        // elements per 16 B
    constexpr int kElemBytes = kBf16In ? 2 : 4;
    for (int64_t c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        while (s + 1 < nseg && seg[(s + 1) * 4 + 3] <= c) ++s;
        const int64_t* row = seg + s * 4;
        const int64_t e0 = (c - row[3]) * kRsChunkElems;
        const int64_t ne = min(kRsChunkElems, row[2] - e0);
        const int64_t src_off = row[0] + e0 * kElemBytes;
        float* dst = out + row[1] + e0;
        const int64_t nvec = ne / kVec;
        for (int64_t i = threadIdx.x; i < nvec; i += kCommThreads) {
            float acc[kVec];
#pragma unroll
            for (int q = 0; q < kVec; ++q) acc[q] = 0.f;
            if constexpr (kNvls) {
                const uint4 v = multimem_ld_reduce_bf16x8(reinterpret_cast<const uint8_t*>(mc_base) + src_off + i * 16);
                acc[0] = bf16_lo(v.x), acc[1] = bf16_hi(v.x), acc[2] = bf16_lo(v.y), acc[3] = bf16_hi(v.y);
                acc[4] = bf16_lo(v.z), acc[5] = bf16_hi(v.z), acc[6] = bf16_lo(v.w), acc[7] = bf16_hi(v.w);
            } else {
                uint4 v[kMaxWorld];
#pragma unroll
                for (int r = 0; r < kMaxWorld; ++r) {
                    if (r < world) {
                        const int peer = (rank + r) % world;  // start at self, stagger egress ports
                        v[r] = ld_stream_v4(reinterpret_cast<const uint8_t*>(peers.p[peer]) + src_off + i * 16);
                    }
                }
#pragma unroll
                for (int r = 0; r < kMaxWorld; ++r) {
                    if (r < world) {
                        if constexpr (kBf16In) {
                            acc[0] += bf16_lo(v[r].x), acc[1] += bf16_hi(v[r].x);
                            acc[2] += bf16_lo(v[r].y), acc[3] += bf16_hi(v[r].y);
                            acc[4] += bf16_lo(v[r].z), acc[5] += bf16_hi(v[r].z);
                            acc[6] += bf16_lo(v[r].w), acc[7] += bf16_hi(v[r].w);
                        } else {
                            acc[0] += __uint_as_float(v[r].x), acc[1] += __uint_as_float(v[r].y);
                            acc[2] += __uint_as_float(v[r].z), acc[3] += __uint_as_float(v[r].w);
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kVec; ++q) {
                acc[q] *= scale;
                sq += acc[q] * acc[q];
            }
            if constexpr (kAdam) {
                // Sharded AdamW right here: the reduced gradient never goes to memory.
                const int64_t e = row[1] + e0 + i * kVec;  // element offset inside the shard
                const float decay = 1.f - adam.lr * adam.wd;
#pragma unroll
                for (int q = 0; q < kVec; ++q) {
                    const int32_t bits = (static_cast<int32_t>(adam.hi[e + q]) << 16) + static_cast<int32_t>(adam.lo[e + q]);
                    float w = __int_as_float(bits);
                    const float mi = adam.beta1 * adam.m[e + q] + (1.f - adam.beta1) * acc[q];
                    const float vi = adam.beta2 * adam.v[e + q] + (1.f - adam.beta2) * acc[q] * acc[q];
                    adam.m[e + q] = mi;
                    adam.v[e + q] = vi;
                    w = w * decay - adam.lr * (mi * adam.inv_bc1) / (sqrtf(vi * adam.inv_bc2) + adam.eps);
                    const int32_t nb = __float_as_int(w);
                    const int32_t h = (nb + 0x8000) >> 16;
                    adam.hi[e + q] = static_cast<uint16_t>(h & 0xFFFF);
                    adam.lo[e + q] = static_cast<int16_t>(nb - (h << 16));
                }
            } else {
                float4* d4 = reinterpret_cast<float4*>(dst + i * kVec);
                d4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                if constexpr (kVec == 8) d4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
        }
    }
    if (sumsq_out != nullptr) {
        sq = warp_sum_f(sq);
        if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = sq;
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < kCommThreads / 32 ? red[threadIdx.x] : 0.f;
            v = warp_sum_f(v);
            if (threadIdx.x == 0) atomicAdd(sumsq_out, v);
        }
    }
}

// flags layout in every rank's symmetric flag region: uint32 flags[slot][world]
// seq_dev (optional): per-slot sequence counters kept on the device (every rank advances them identically), so
// the same launch can be replayed from a CUDA graph.
__global__ void signal_barrier_kernel(PeerPtrs flag_bases, int rank, int world, int slot, uint32_t seq,
                                      uint32_t* seq_dev) {
    const int r = threadIdx.x;
    if (seq_dev != nullptr) seq = seq_dev[slot] + 1;
    __syncwarp();
    if (r == 0 && seq_dev != nullptr) seq_dev[slot] = seq;
    if (r >= world) return;
    __threadfence_system();
    uint32_t* remote = reinterpret_cast<uint32_t*>(flag_bases.p[r]) + slot * kMaxWorld + rank;
    st_release_sys(remote, seq);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(flag_bases.p[rank]) + slot * kMaxWorld + r;
    uint32_t spins = 0;
    // Sequence numbers only grow; signed distance handles wrap-around.
    while (static_cast<int32_t>(ld_acquire_sys(mine) - seq) < 0) {
        if (++spins > (1u << 28)) {
            printf("[b200] signal_barrier timeout: rank %d waiting for peer %d slot %d seq %u\n", rank, r, slot, seq);
            __trap();
        }
    }
}

// scratch layout in every rank's symmetric region: float scratch[slot][world][kMaxScalars]
constexpr int kMaxScalars = 16;
__global__ void allreduce_scalars_kernel(PeerPtrs flag_bases, PeerPtrs scratch_bases, int rank, int world, int slot,
                                         uint32_t seq, float* __restrict__ vals, int k, int op, uint32_t* seq_dev,
                                         int counter_idx) {
    const int t = threadIdx.x;
    if (seq_dev != nullptr) {  // slot alternates with the parity of the device-side sequence number
        seq = seq_dev[counter_idx] + 1;
        slot = slot + static_cast<int>(seq & 1u);
    }
    __syncthreads();
    if (t == 0 && seq_dev != nullptr) seq_dev[counter_idx] = seq;
    // phase 1: thread (r, j) pushes vals[j] into peer r's scratch[slot][rank][j]
    if (t < world * k) {
        const int r = t / k, j = t % k;
        float* dst = reinterpret_cast<float*>(scratch_bases.p[r]) + (slot * kMaxWorld + rank) * kMaxScalars + j;
        st_relaxed_sys_f32(dst, vals[j]);
    }
    __syncthreads();
    if (t < world) {
        __threadfence_system();
        uint32_t* remote = reinterpret_cast<uint32_t*>(flag_bases.p[t]) + slot * kMaxWorld + rank;
        st_release_sys(remote, seq);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(flag_bases.p[rank]) + slot * kMaxWorld + t;
        uint32_t spins = 0;
        while (static_cast<int32_t>(ld_acquire_sys(mine) - seq) < 0) {
            if (++spins > (1u << 28)) {
                printf("[b200] allreduce_scalars timeout: rank %d waiting for peer %d seq %u\n", rank, t, seq);
                __trap();
            }
        }
    }
    __syncthreads();
    if (t < k) {
        const float* src = reinterpret_cast<const float*>(scratch_bases.p[rank]) + slot * kMaxWorld * kMaxScalars + t;
        float acc = ld_relaxed_sys_f32(src);
        for (int r = 1; r < world; ++r) {  // fixed order -> bitwise identical on every rank
            const float v = ld_relaxed_sys_f32(src + r * kMaxScalars);
            acc = (op == 0) ? acc + v : fmaxf(acc, v);
        }
        vals[t] = acc;
    }
}

PeerPtrs to_peers(const std::vector<int64_t>& v) {
    if (v.size() > kMaxWorld) throw std::runtime_error("comm: world size > 16 not supported");
    PeerPtrs p;
    for (int i = 0; i < kMaxWorld; ++i) p.p[i] = i < (int)v.size() ? static_cast<uint64_t>(v[i]) : 0;
    return p;
}

inline void check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(err));
}

}  // namespace

void p2p_all_gather(const std::vector<int64_t>& peer_ptrs, int rank, void* out, const int64_t* seg_table_dev,
                    int nseg, int64_t total_chunks, int max_ctas, cudaStream_t stream) {
    (void)rank;
    if (total_chunks == 0) return;
    const int grid = static_cast<int>(std::min<int64_t>(total_chunks, max_ctas > 0 ? max_ctas : 32));
    p2p_all_gather_kernel<<<grid, kCommThreads, 0, stream>>>(to_peers(peer_ptrs), static_cast<uint8_t*>(out),
                                                            seg_table_dev, nseg, total_chunks);
    check_launch("p2p_all_gather");
}

void p2p_reduce_scatter(const std::vector<int64_t>& peer_ptrs, int rank, float* out, const int64_t* seg_table_dev,
                        int nseg, int64_t total_chunks, bool in_is_bf16, float scale, float* sumsq_out, int max_ctas,
                        cudaStream_t stream, const AdamFuse* adam) {
    if (total_chunks == 0) return;
    const int world = static_cast<int>(peer_ptrs.size());
    const int grid = static_cast<int>(std::min<int64_t>(total_chunks, max_ctas > 0 ? max_ctas : 32));
    const AdamFuse a = adam != nullptr ? *adam : AdamFuse{};
    if (in_is_bf16 && adam != nullptr)
        reduce_scatter_kernel<true, false, true><<<grid, kCommThreads, 0, stream>>>(
            to_peers(peer_ptrs), 0, rank, world, out, seg_table_dev, nseg, total_chunks, scale, sumsq_out, a);
    else if (in_is_bf16)
        reduce_scatter_kernel<true, false, false><<<grid, kCommThreads, 0, stream>>>(
            to_peers(peer_ptrs), 0, rank, world, out, seg_table_dev, nseg, total_chunks, scale, sumsq_out, a);
    else if (adam != nullptr)
        throw std::runtime_error("reduce_scatter: fused AdamW needs bf16 gradients");
    else
        reduce_scatter_kernel<false, false, false><<<grid, kCommThreads, 0, stream>>>(
            to_peers(peer_ptrs), 0, rank, world, out, seg_table_dev, nseg, total_chunks, scale, sumsq_out, a);
    check_launch("p2p_reduce_scatter");
}

void nvls_reduce_scatter(int64_t mc_ptr, int rank, int world, float* out, const int64_t* seg_table_dev, int nseg,
                         int64_t total_chunks, float scale, float* sumsq_out, int max_ctas, cudaStream_t stream,
                         const AdamFuse* adam) {
    if (total_chunks == 0) return;
    const int grid = static_cast<int>(std::min<int64_t>(total_chunks, max_ctas > 0 ? max_ctas : 32));
    PeerPtrs none{};
    const AdamFuse a = adam != nullptr ? *adam : AdamFuse{};
    if (adam != nullptr)
        reduce_scatter_kernel<true, true, true><<<grid, kCommThreads, 0, stream>>>(
            none, static_cast<uint64_t>(mc_ptr), rank, world, out, seg_table_dev, nseg, total_chunks, scale, sumsq_out, a);
    else
        reduce_scatter_kernel<true, true, false><<<grid, kCommThreads, 0, stream>>>(
            none, static_cast<uint64_t>(mc_ptr), rank, world, out, seg_table_dev, nseg, total_chunks, scale, sumsq_out, a);
    check_launch("nvls_reduce_scatter");
}

void signal_barrier(const std::vector<int64_t>& flag_ptrs, int rank, int world, int slot, uint32_t seq,
                    cudaStream_t stream, uint32_t* seq_dev) {
    signal_barrier_kernel<<<1, 32, 0, stream>>>(to_peers(flag_ptrs), rank, world, slot, seq, seq_dev);
    check_launch("signal_barrier");
}

void allreduce_scalars(const std::vector<int64_t>& flag_ptrs, const std::vector<int64_t>& scratch_ptrs, int rank,
                       int world, int slot, uint32_t seq, float* vals, int k, int op, cudaStream_t stream,
                       uint32_t* seq_dev, int counter_idx) {
    if (k > kMaxScalars) throw std::runtime_error("allreduce_scalars: at most 16 values");
    allreduce_scalars_kernel<<<1, 256, 0, stream>>>(to_peers(flag_ptrs), to_peers(scratch_ptrs), rank, world, slot, seq,
                                                   vals, k, op, seq_dev, counter_idx);
    check_launch("allreduce_scalars");
}

int64_t ag_chunk_bytes() { return kAgChunkBytes; }
int64_t rs_chunk_elems() { return kRsChunkElems; }
int comm_max_world() { return kMaxWorld; }
int comm_max_scalars() { return kMaxScalars; }

}  // namespace b200
