// K3: EMA codebook statistics and update.
// K3a (accumulate) replaces the dense one-hot build + [K,M]x[M,D] GEMM + row sums of bottleneck.py:64-68
//     with a scatter-add that never materialises the one-hot matrix:
//     algorithmic bytes per frame = 4D (x) + 8 (idx) + 4 (mask);  + 4K(D+1) per launch for the statistics.
// K3b (finalize) replaces bottleneck.py:78-89 (EMA lerp, usage threshold, dead-code revival, 4 metrics).
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int E_TT = 128;        // frames per tile
constexpr int E_DS = 32;         // depth slice owned by one CTA (lane == depth)
constexpr int E_THREADS = 256;   // large-K kernel

// ---- small-K kernel: private [K][32] slab in shared memory, "owner computes", cp.async tile pipeline
constexpr int EA_WARPS = 16;                 // warp w owns the codes with (c & 15) == w
constexpr int EA_THREADS = EA_WARPS * 32;
constexpr int EA_STAGES = 4;                 // tiles in flight per CTA (the kernel is a pure HBM stream)
constexpr int EA_XS = E_TT + 4;              // 16-byte aligned rows for 16-byte cp.async (column reads then take 4 wavefronts; they are rare)
constexpr int EA_STAGE_BYTES = E_DS * EA_XS * 4 + E_TT * 8 + E_TT * 4;   // x slice + idx (int64) + mask

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Lane d of warp w is the ONLY thread that ever touches acc[c][d] for (c & 15) == w, so the read-modify-write needs no
// atomics; each warp compacts the rows it owns with ballots (no per-row branch for the other 15 warps) and sums runs of
// one code in a register before touching the slab.  Tiles stream through a 4-stage cp.async ring (16-byte copies).
__global__ void __launch_bounds__(EA_THREADS, 1)
ema_accumulate_smem_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ mask,
                           int64_t N, int D, int64_t T, int K, float* __restrict__ stats) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* stage0 = smem_raw;                                                    // EA_STAGES x EA_STAGE_BYTES
    float* acc = reinterpret_cast<float*>(smem_raw + size_t(EA_STAGES) * EA_STAGE_BYTES);   // [K][E_DS]
    float* cnt = acc + size_t(K) * E_DS;                                            // [K] (slice 0 only)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slice = blockIdx.y;
    const int d0 = slice * E_DS;
    const int dn = min(E_DS, D - d0);
    const bool count_here = slice == 0;
    const int64_t tiles_per_utt = (T + E_TT - 1) / E_TT;
    const int64_t n_tiles = N * tiles_per_utt;
    const bool vec16 = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);

    for (int i = tid; i < K * E_DS; i += EA_THREADS) acc[i] = 0.f;
    for (int i = tid; i < K; i += EA_THREADS) cnt[i] = 0.f;

    auto issue = [&](int64_t tile, int st) {
        float* Xs = reinterpret_cast<float*>(stage0 + size_t(st) * EA_STAGE_BYTES);
        int64_t* s_idx = reinterpret_cast<int64_t*>(Xs + E_DS * EA_XS);
        float* s_mask = reinterpret_cast<float*>(s_idx + E_TT);
        if (tile < n_tiles) {
            const int64_t n = tile / tiles_per_utt, t0 = (tile % tiles_per_utt) * E_TT;
            const int tt = int(min(int64_t(E_TT), T - t0));
            if (vec16) {                          // 16 bytes per copy: 2 copies per thread per tile
#pragma unroll
                for (int i = tid; i < E_DS * (E_TT / 4); i += EA_THREADS) {
                    const int d = i >> 5, t = (i & 31) * 4;
                    float* dst = Xs + d * EA_XS + t;
                    if (d < dn && t < tt) cp_async16(dst, x + (size_t(n) * D + d0 + d) * T + t0 + t);   // T % 4 == 0: a chunk never straddles tt
                    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
#pragma unroll
                for (int i = tid; i < E_DS * E_TT; i += EA_THREADS) {
                    const int d = i >> 7, t = i & (E_TT - 1);
                    float* dst = Xs + d * EA_XS + t;
                    if (d < dn && t < tt) cp_async4(dst, x + (size_t(n) * D + d0 + d) * T + t0 + t);
                    else *dst = 0.f;
                }
            }
            if (tid < E_TT) {
                if (tid < tt) {
                    cp_async8(s_idx + tid, idx + n * T + t0 + tid);
                    if (mask) cp_async4(s_mask + tid, mask + n * T + t0 + tid);
                    else s_mask[tid] = 1.f;
                } else {
                    s_idx[tid] = -1;
                    s_mask[tid] = 0.f;
                }
            }
        }
        cp_async_commit();
    };

    const int64_t first = blockIdx.x, step = gridDim.x;
#pragma unroll
    for (int s = 0; s < EA_STAGES - 1; ++s) issue(first + s * step, s);

    int it = 0;
    int pc[4] = {-1, -1, -1, -1};                // pending (code, partial sum, count) of this warp's four chains
    float ps[4] = {0.f, 0.f, 0.f, 0.f}, pn[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t tile = first; tile < n_tiles; tile += step, ++it) {
        cp_async_wait<EA_STAGES - 2>();          // this tile's group has landed (for this thread's copies)
        __syncthreads();                         // ... for everyone's; and the stage freed last iteration is reusable
        issue(tile + int64_t(EA_STAGES - 1) * step, (it + EA_STAGES - 1) % EA_STAGES);
        const int st = it % EA_STAGES;
        const float* Xs = reinterpret_cast<const float*>(stage0 + size_t(st) * EA_STAGE_BYTES);
        const int64_t* s_idx = reinterpret_cast<const int64_t*>(Xs + E_DS * EA_XS);
        const float* s_mask = reinterpret_cast<const float*>(s_idx + E_TT);
        const float* xcol = Xs + lane * EA_XS;
        int code[4];
        unsigned m[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = j * 32 + lane;
            const int64_t ci = s_idx[t];
            code[j] = (s_mask[t] != 0.f && ci >= 0) ? int(min(ci, int64_t(K - 1))) : -1;
            m[j] = __ballot_sync(0xffffffffu, code[j] >= 0 && (code[j] & (EA_WARPS - 1)) == warp);
        }
        // Four independent chains (one per 32-frame group), each with its own pending (code, sum, count): the shuffles
        // and column reads of a step are issued together, so their latency is paid once per step, not once per row.
        while (m[0] | m[1] | m[2] | m[3]) {
            int c[4];
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                c[j] = -1;
                v[j] = 0.f;
                if (m[j]) {                       // warp-uniform
                    const int b = __ffs(m[j]) - 1;
                    m[j] &= m[j] - 1;
                    c[j] = __shfl_sync(0xffffffffu, code[j], b);
                    v[j] = xcol[j * 32 + b];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c[j] < 0) continue;
                if (c[j] != pc[j]) {              // runs of one code (hot codes!) collapse into a register sum
                    if (pc[j] >= 0) {
                        acc[size_t(pc[j]) * E_DS + lane] += ps[j];
                        if (count_here && lane == 0) cnt[pc[j]] += pn[j];
                    }
                    pc[j] = c[j]; ps[j] = v[j]; pn[j] = 1.f;
                } else {
                    ps[j] += v[j]; pn[j] += 1.f;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (pc[j] >= 0) {
            acc[size_t(pc[j]) * E_DS + lane] += ps[j];
            if (count_here && lane == 0) cnt[pc[j]] += pn[j];
        }
    cp_async_wait<0>();
    __syncthreads();
    // flush the private slab: one FP32 reduction per touched cell, coalesced along depth
    float* sums = stats;
    float* counts = stats + size_t(K) * D;
    for (int i = tid; i < K * E_DS; i += EA_THREADS) {
        const int c = i / E_DS, d = i % E_DS;
        const float v = acc[i];
        if (d < dn && v != 0.f) atomicAdd(&sums[size_t(c) * D + d0 + d], v);
    }
    if (count_here)
        for (int c = tid; c < K; c += EA_THREADS)
            if (cnt[c] != 0.f) atomicAdd(&counts[c], cnt[c]);
}

// Large-K variant: rows of one tile rarely share a code, so privatisation buys nothing; transpose the
// tile through shared memory and issue depth-coalesced FP32 reductions straight to L2.
__global__ void __launch_bounds__(E_THREADS)
ema_accumulate_global_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ mask,
                             int64_t N, int D, int64_t T, int K, float* __restrict__ stats) {
    __shared__ float Xs[E_DS][E_TT + 1];
    __shared__ int s_code[E_TT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* sums = stats;
    float* counts = stats + size_t(K) * D;
    const int64_t tiles_per_utt = (T + E_TT - 1) / E_TT;
    const int64_t n_tiles = N * tiles_per_utt;
    const int n_slices = (D + E_DS - 1) / E_DS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n = tile / tiles_per_utt, t0 = (tile % tiles_per_utt) * E_TT;
        const int tt = int(min(int64_t(E_TT), T - t0));
        __syncthreads();
        if (tid < E_TT) {
            int c = -1;
            if (tid < tt) {
                float m = mask ? mask[n * T + t0 + tid] : 1.f;
                if (m != 0.f) c = int(min(max(idx[n * T + t0 + tid], int64_t(0)), int64_t(K - 1)));
            }
            s_code[tid] = c;
            if (c >= 0) atomicAdd(&counts[c], 1.f);
        }
        for (int s = 0; s < n_slices; ++s) {
            const int d0 = s * E_DS, dn = min(E_DS, D - d0);
            __syncthreads();
            for (int i = tid; i < E_DS * E_TT; i += E_THREADS) {
                int d = i / E_TT, t = i % E_TT;
                float v = 0.f;
                if (d < dn && t < tt) v = ld_stream(x + (size_t(n) * D + d0 + d) * T + t0 + t);
                Xs[d][t] = v;
            }
            __syncthreads();
            for (int t = warp; t < tt; t += E_THREADS / 32) {
                int c = s_code[t];
                if (c >= 0 && lane < dn) atomicAdd(&sums[size_t(c) * D + d0 + lane], Xs[lane][t]);
            }
        }
    }
}

// sum of the (already all-reduced) counts -> scalars[VQ_S_COUNT_TOTAL]
__global__ void __launch_bounds__(256) ema_count_total_kernel(const float* __restrict__ counts, int K, double* scalars) {
    __shared__ double red[32];
    double s = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < K; c += gridDim.x * blockDim.x) s += double(counts[c]);
    s = block_sum(s, red);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(&scalars[VQ_S_COUNT_TOTAL], s);
}

// One warp per code.
__global__ void __launch_bounds__(256)
ema_finalize_kernel(const float* __restrict__ stats, const float* __restrict__ k_rand, const float* k_old,
                    float* k, float* __restrict__ k_sum, float* __restrict__ k_elem,
                    int K, int D, float mu, float one_minus_mu, float threshold, float laplace_eps,
                    double* __restrict__ scalars, float* __restrict__ results, int64_t* __restrict__ used_curr_out,
                    unsigned int total_blocks) {
    __shared__ double red[32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const float* sums = stats;
    const float* counts = stats + size_t(K) * D;
    const double total = scalars[VQ_S_COUNT_TOTAL];
    double ent = 0.0, used = 0.0, usage_n = 0.0, dk = 0.0;

    for (int c = blockIdx.x * warps_per_block + (threadIdx.x >> 5); c < K; c += gridDim.x * warps_per_block) {
        const float n_c = counts[c];
        // k_elem <- mu k_elem + (1 - mu) n_c                                   (bottleneck.py:80)
        float ke = __fadd_rn(__fmul_rn(mu, k_elem[c]), __fmul_rn(one_minus_mu, n_c));
        float ke_div = ke;
        if (laplace_eps > 0.f) {
            float ntot = float(total);
            ke_div = (ke + laplace_eps) / (ntot + float(K) * laplace_eps) * ntot;
        }
        const bool alive = ke >= threshold;                                     // :81
        float dsq = 0.f;
        for (int d = lane; d < D; d += 32) {
            const size_t o = size_t(c) * D + d;
            // k_sum <- mu k_sum + (1 - mu) sums                                 (:79)
            float ks = __fadd_rn(__fmul_rn(mu, k_sum[o]), __fmul_rn(one_minus_mu, sums[o]));
            k_sum[o] = ks;
            // k <- usage * (k_sum / k_elem) + (1 - usage) * k_rand              (:82-83), as a select
            float nk = alive ? __fdiv_rn(ks, ke_div) : k_rand[o];
            float diff = nk - k_old[o];
            dsq = fmaf(diff, diff, dsq);
            k[o] = nk;
        }
        dsq = warp_sum(dsq);
        if (lane == 0) {
            k_elem[c] = ke;
            dk += double(dsq);
            usage_n += alive ? 1.0 : 0.0;
            used += (n_c >= threshold) ? 1.0 : 0.0;                              // :87
            float p = float(double(n_c) / total);                               // :85
            ent -= double(p * logf(fmaxf(p, 1e-5f)));                            // :86 (safe_log)
        }
    }
    double e1 = block_sum(ent, red), e2 = block_sum(used, red), e3 = block_sum(usage_n, red), e4 = block_sum(dk, red);
    if (threadIdx.x == 0) {
        atomicAdd(&scalars[VQ_S_ENTROPY], e1);
        atomicAdd(&scalars[VQ_S_USED_CURR], e2);
        atomicAdd(&scalars[VQ_S_USAGE], e3);
        atomicAdd(&scalars[VQ_S_DK_SQ], e4);
        __threadfence();
        unsigned int ticket = atomicAdd(reinterpret_cast<unsigned int*>(&scalars[VQ_S_TICKET]) + 1, 1u);
        is_last = (ticket == total_blocks - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        volatile double* sc = scalars;
        results[VQ_R_ENTROPY] = float(sc[VQ_S_ENTROPY]);
        results[VQ_R_USAGE] = float(sc[VQ_S_USAGE]);
        results[VQ_R_USED_CURR] = float(sc[VQ_S_USED_CURR]);
        results[VQ_R_DK] = float(sqrt(sc[VQ_S_DK_SQ]) / sqrt(double(K) * double(D)));   // :89
        if (used_curr_out) *used_curr_out = (long long)(sc[VQ_S_USED_CURR] + 0.5);
        reinterpret_cast<unsigned int*>(&scalars[VQ_S_TICKET])[1] = 0u;
    }
}

}  // namespace vq
