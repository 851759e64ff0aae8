// The one exchange step of the path (bottleneck.py:72-75: broadcast(k_rand) + all_reduce(k_sum) + all_reduce(k_elem)) done
// over NVLink peer memory instead of three (or one) NCCL collectives.  One process per GPU, one node.
//
// Every rank owns a REGION (cudaMalloc, exported with cudaIpcGetMemHandle, mapped by all peers):
//     [ flags: 64 x uint32 ][ stats[2][K*D + K] ][ k_rand[2][K*D] ]            (two slots, used alternately by step parity)
// Per training step s a rank (1) accumulates its statistics straight into stats[s & 1] of its own region (K3a) and puts its
// restart rows into k_rand[s & 1]; (2) PUBLISHES: stores s into flags[rank] of every peer's region; (3) WAITS until all R
// flags of its own region show s; (4) REDUCES: reads the R statistics slots through the peer mappings and sums them in rank
// order -- every rank gets bit-identical sums -- and copies rank 0's restart rows; then the usual finalize kernel runs.
// 526 KB per rank at the default shape: ~4 MB of NVLink reads per GPU and two flag round trips, against ~0.15-0.2 ms for an
// 8-rank NCCL all-reduce of the same buffer (latency-bound).
// Slot reuse is safe with two slots: a rank writes stats[s & 1] for step s + 2 only after its own step s + 1 exchange, which
// waited for every peer's step s + 1 flag, which a peer sets only after it finished reading step s.
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int P2P_MAX_RANKS = 16;
constexpr size_t P2P_FLAG_BYTES = 256;

struct P2PPeers { void* region[P2P_MAX_RANKS]; };

inline size_t p2p_stats_floats(int K, int D) { return size_t(K) * D + K; }
inline size_t p2p_region_bytes(int K, int D) { return P2P_FLAG_BYTES + 2 * p2p_stats_floats(K, D) * 4 + 2 * size_t(K) * D * 4; }
__host__ __device__ inline float* p2p_stats_slot(void* region, unsigned step, size_t stats_floats) {
    return reinterpret_cast<float*>(static_cast<char*>(region) + P2P_FLAG_BYTES) + size_t(step & 1u) * stats_floats;
}
__host__ __device__ inline float* p2p_krand_slot(void* region, unsigned step, size_t stats_floats, size_t krand_floats) {
    return reinterpret_cast<float*>(static_cast<char*>(region) + P2P_FLAG_BYTES) + 2 * stats_floats + size_t(step & 1u) * krand_floats;
}

__global__ void p2p_publish_kernel(P2PPeers peers, int n_ranks, int rank, unsigned step) {
    const int r = threadIdx.x;
    if (r < n_ranks) {
        __threadfence_system();                                   // this rank's statistics (earlier kernels of the stream) before the flag
        *reinterpret_cast<volatile unsigned*>(static_cast<unsigned*>(peers.region[r]) + rank) = step;
    }
}

__global__ void p2p_wait_kernel(const unsigned* my_flags, int n_ranks, unsigned step) {
    const int r = threadIdx.x;
    if (r < n_ranks) {
        unsigned long long spins = 0;
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + r) : "memory");
            if (int(v - step) >= 0) break;
            if (++spins > (1ull << 31)) __trap();                 // a peer died: fail loudly instead of hanging the GPU
            __nanosleep(200);
        }
    }
}

__global__ void __launch_bounds__(256) p2p_reduce_kernel(P2PPeers peers, int n_ranks, unsigned step, size_t stats_floats, size_t krand_floats,
                                                        float* __restrict__ stats_out, float* __restrict__ k_rand_out) {
    const size_t tid = size_t(blockIdx.x) * blockDim.x + threadIdx.x, nthreads = size_t(gridDim.x) * blockDim.x;
    // 16-byte loads (the slots are 16-byte aligned when K*D + K is a multiple of 4; the tail goes element by element)
    const bool vec = (stats_floats % 4 == 0) && (krand_floats % 4 == 0);
    const size_t n4 = vec ? stats_floats / 4 : 0;
    for (size_t i = tid; i < n4; i += nthreads) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < n_ranks; ++r) {                        // rank order: the same sum everywhere
            const float4 v = __ldcg(reinterpret_cast<const float4*>(p2p_stats_slot(peers.region[r], step, stats_floats)) + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        reinterpret_cast<float4*>(stats_out)[i] = s;
    }
    for (size_t i = n4 * 4 + tid; i < stats_floats; i += nthreads) {
        float s = 0.f;
        for (int r = 0; r < n_ranks; ++r) s += __ldcg(p2p_stats_slot(peers.region[r], step, stats_floats) + i);
        stats_out[i] = s;
    }
    const float* kr = p2p_krand_slot(peers.region[0], step, stats_floats, krand_floats);                              // rank 0's restart rows (bottleneck.py:73)
    for (size_t i = tid; i < krand_floats; i += nthreads) k_rand_out[i] = __ldcg(kr + i);
}

}  // namespace vq
