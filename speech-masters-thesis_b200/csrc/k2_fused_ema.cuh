// K2 forward + K3a in ONE pass over x (training forward, K <= 512, D <= 128, T % 4 == 0):
//   x_q = (x + (e - x)) * mask, commit-loss / fit reductions        (bottleneck.py:143-145,194,197,118-124,201)
//   per-code sums and counts of the valid frames                     (bottleneck.py:64-68)
// The separate K3a kernel re-read x (152 MB at the bench shape) through 256-byte pieces at 0.24 of the HBM peak; here the
// x tile is already in shared memory for the straight-through arithmetic, so the EMA statistics cost no HBM traffic at all.
//
// Where the [K][D] FP32 accumulators live: shared memory is full (three pipeline stages of x + gathered codebook rows =
// 209 KB), but the SM's 256 KB of TENSOR MEMORY are idle in this kernel and have exactly the right shape -- 128 lanes
// (depth) x 512 columns (code) x 4 bytes.  Each 64-frame tile is sorted by code (one warp, bitonic, one tile ahead); a
// run of equal codes is summed in registers (lane == depth) and added to its TMEM column with one tcgen05.ld / add /
// tcgen05.st, so a hot code costs one update per tile, not one per frame, and no two warps ever touch the same column in
// the same tile.  Warp w works on TMEM lane quadrant w % 4 (depths 32 (w % 4) ..+31) and takes every fourth run.  The
// columns are flushed to the global statistics buffer with coalesced FP32 reductions when the CTA is done.
#pragma once
#include "k2_gather.cuh"

namespace vq {

constexpr int FE_THREADS = 512;
constexpr int FE_XS = G_TT + 4;                      // row stride of the x tile AND of the gathered rows (68 floats, 16-byte aligned rows;
                                                     // lane == depth reads of one frame are 4-way bank conflicts instead of 32-way)
constexpr int FE_NST = 3;
constexpr int FE_STAGE_FLOATS = 2 * GA_DS * FE_XS;   // x tile + codebook rows
constexpr int FE_KMAX = 512, FE_DMAX = 128;
constexpr size_t FE_SMEM = size_t(FE_NST) * FE_STAGE_FLOATS * 4 + GA_RING * G_TT * 12 + 2 * G_TT * 4 + FE_KMAX * 4;

__device__ __forceinline__ void fe_tmem_ld1(uint32_t taddr, float& v) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v = __uint_as_float(r);
}
__device__ __forceinline__ void fe_tmem_st1(uint32_t taddr, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void fe_tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void fe_tmem_zero16(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
}

__global__ void __launch_bounds__(FE_THREADS, 1)
gather_fwd_ema_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ mask,
                      const float* __restrict__ k, int N, int D, int T, int K, float* __restrict__ out,
                      double* __restrict__ scalars, float* __restrict__ results, float* __restrict__ stats,
                      unsigned int total_blocks) {
    extern __shared__ __align__(16) float smem[];
    int64_t* s_idx = reinterpret_cast<int64_t*>(smem + size_t(FE_NST) * FE_STAGE_FLOATS);   // [GA_RING][G_TT]
    float* s_mask = reinterpret_cast<float*>(s_idx + GA_RING * G_TT);                        // [GA_RING][G_TT]
    uint32_t* sorted = reinterpret_cast<uint32_t*>(s_mask + GA_RING * G_TT);                 // [2][G_TT]  (code << 8 | frame), ~0u = no row
    float* s_cnt = reinterpret_cast<float*>(sorted + 2 * G_TT);                              // [FE_KMAX]
    __shared__ double red[32];
    __shared__ bool is_last;
    __shared__ int4 s_loc[GA_RING];
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_per_utt = (T + G_TT - 1) / G_TT;
    const int n_units = N * tiles_per_utt;
    const int u0 = blockIdx.x, step = gridDim.x;
    const int dn = D;                                              // one depth slice: D <= 128

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(uint32_t(__cvta_generic_to_shared(&s_tmem))), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int c = tid; c < FE_KMAX; c += FE_THREADS) s_cnt[c] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const int quad = warp & 3, sub = warp >> 2;                    // TMEM lane quadrant of this warp; which runs / columns it takes
    const uint32_t tq = tmem + (uint32_t(quad * 32) << 16);
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 16) fe_tmem_zero16(tq + uint32_t(128 * sub + c0));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");

    auto prefetch_im = [&](int j, int u) {
        if (tid < G_TT) {
            const int slot = j & (GA_RING - 1);
            int4 loc = make_int4(0, 0, 0, 0);
            bool copied = false;
            if (u < n_units) {
                const int n = u / tiles_per_utt, t0 = (u - n * tiles_per_utt) * G_TT;
                loc = make_int4(n, t0, 0, 1);
                const int t = t0 + tid;
                if (t < T) {
                    cp_async8(s_idx + slot * G_TT + tid, idx + int64_t(n) * T + t);
                    if (mask) cp_async4(s_mask + slot * G_TT + tid, mask + int64_t(n) * T + t);
                    else s_mask[slot * G_TT + tid] = 1.f;
                    copied = true;
                }
            }
            if (!copied) { s_idx[slot * G_TT + tid] = -1; s_mask[slot * G_TT + tid] = 0.f; }
            if (tid == 0) s_loc[slot] = loc;
        }
    };
    auto issue = [&](int st, int ring) {
        const int4 loc = s_loc[ring];
        if (loc.w) {
            const int n = loc.x, t0 = loc.y;
            float* S = smem + size_t(st) * FE_STAGE_FLOATS;
#pragma unroll
            for (int r = 0; r < GA_DS * (G_TT / 4) / FE_THREADS; ++r) {
                const int i = tid + r * FE_THREADS, d = i >> 4, c4 = (i & 15) * 4;
                if (d < dn && t0 + c4 < T) cp_async16(S + d * FE_XS + c4, x + (int64_t(n) * D + d) * T + t0 + c4);
            }
        }
        cp_async_commit();
    };
    const int g_r = 4 * warp + (lane & 3), g_dd = lane >> 2;
    auto gather_ld = [&](int ring, float (&ev)[GA_DS / 8]) {
        const int4 loc = s_loc[ring];
        if (loc.w) {
            const int code = int(min(max(s_idx[ring * G_TT + g_r], int64_t(0)), int64_t(K - 1)));
            const float* src = k + size_t(code) * D + g_dd;
#pragma unroll
            for (int db = 0; db < GA_DS / 8; ++db) ev[db] = (8 * db + g_dd < dn) ? __ldg(src + 8 * db) : 0.f;
        }
    };
    auto gather_st = [&](int st, const float (&ev)[GA_DS / 8]) {
        float* dst = smem + size_t(st) * FE_STAGE_FLOATS + GA_DS * FE_XS + g_dd * FE_XS + g_r;
#pragma unroll
        for (int db = 0; db < GA_DS / 8; ++db) dst[8 * db * FE_XS] = ev[db];
    };
    // keys of the unit in ring slot `ring`: (code << 8) | frame for valid frames, ~0u otherwise; bitonic sort in one warp
    auto sort_unit = [&](int ring, uint32_t* dst) {
        uint32_t key[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int t = r * 32 + lane;
            const int64_t ci = s_idx[ring * G_TT + t];
            key[r] = (s_mask[ring * G_TT + t] != 0.f && ci >= 0 && s_loc[ring].w) ? ((uint32_t(min(ci, int64_t(K - 1))) << 8) | uint32_t(t)) : 0xFFFFFFFFu;
        }
#pragma unroll
        for (int kk = 2; kk <= G_TT; kk <<= 1) {
#pragma unroll
            for (int j = kk >> 1; j > 0; j >>= 1) {
                if (j >= 32) {
                    const uint32_t a = key[0], b = key[1];
                    key[0] = min(a, b);
                    key[1] = max(a, b);
                } else {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const uint32_t other = __shfl_xor_sync(0xffffffffu, key[r], j);
                        const int i = r * 32 + lane;
                        const bool up = (i & kk) == 0, lower = (lane & j) == 0;
                        key[r] = (lower == up) ? min(key[r], other) : max(key[r], other);
                    }
                }
            }
        }
        dst[lane] = key[0];
        dst[32 + lane] = key[1];
    };

    double sq = 0.0, sq_all = 0.0, msum_local = 0.0;

#pragma unroll
    for (int j = 0; j < GA_AHEAD; ++j) prefetch_im(j, u0 + j * step);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    float ev[GA_DS / 8];
#pragma unroll
    for (int j = 0; j < GA_DS / 8; ++j) ev[j] = 0.f;
    gather_ld(0, ev);
    gather_st(0, ev);
    if (warp == 15) sort_unit(0, sorted);
#pragma unroll
    for (int j = 0; j < FE_NST - 1; ++j) issue(j, j);

    int it = 0;
    for (int u = u0; u < n_units; u += step, ++it) {
        gather_ld((it + 1) & (GA_RING - 1), ev);
        prefetch_im(it + GA_AHEAD, u + GA_AHEAD * step);
        cp_async_wait<FE_NST - 2>();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                          // x of unit u landed; sorted[it & 1] complete; last unit's TMEM updates done
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue((it + FE_NST - 1) % FE_NST, (it + FE_NST - 1) & (GA_RING - 1));

        const int4 cur = s_loc[it & (GA_RING - 1)];
        const int n = cur.x, t0 = cur.y;
        const int tt = min(G_TT, T - t0);
        const float* S = smem + size_t(it % FE_NST) * FE_STAGE_FLOATS;
        const float* Es = S + GA_DS * FE_XS;
        const float* sm = s_mask + (it & (GA_RING - 1)) * G_TT;
        // ---- straight-through output + loss reductions (bottleneck.py:194-201)
        {
            const int t4 = (tid & 15) * 4, dg = tid >> 4;
            if (tid < G_TT) msum_local += double(sm[tid]);
            if (t4 < tt) {
                const float4 m4 = *reinterpret_cast<const float4*>(sm + t4);
                const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
                const float vv[4] = {m4.x != 0.f ? 1.f : 0.f, m4.y != 0.f ? 1.f : 0.f, m4.z != 0.f ? 1.f : 0.f, m4.w != 0.f ? 1.f : 0.f};
                float accf[4] = {0.f, 0.f, 0.f, 0.f};
                float* dst = out + (int64_t(n) * D) * T + t0 + t4;
#pragma unroll 4
                for (int d = dg; d < dn; d += FE_THREADS / 16) {
                    const float4 e4 = *reinterpret_cast<const float4*>(Es + d * FE_XS + t4);
                    const float4 x4 = *reinterpret_cast<const float4*>(S + d * FE_XS + t4);
                    const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
                    const float xx[4] = {x4.x, x4.y, x4.z, x4.w};
                    float oo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float diff = __fsub_rn(ee[j], xx[j]);                   // (x_d - x)
                        oo[j] = __fmul_rn(__fadd_rn(xx[j], diff), mm[j]);             // (x + (x_d - x)) * mask
                        accf[j] = fmaf(diff, diff, accf[j]);
                    }
                    st_stream4(reinterpret_cast<float4*>(dst + int64_t(d) * T), make_float4(oo[0], oo[1], oo[2], oo[3]));
                }
                sq_all += double((accf[0] + accf[1]) + (accf[2] + accf[3]));
                sq += double((accf[0] * vv[0] + accf[1] * vv[1]) + (accf[2] * vv[2] + accf[3] * vv[3]));
            }
        }
        // ---- EMA statistics of this unit (bottleneck.py:64-68): runs of equal codes -> one TMEM column update each
        if (warp == 15) sort_unit((it + 1) & (GA_RING - 1), sorted + ((it + 1) & 1) * G_TT);   // next unit's keys (its indices landed a wait ago)
        if (quad * 32 < dn) {
            // Runs are found lane-parallel (two ballots over the sorted keys give a 64-bit mask of run heads) and this warp
            // takes runs sub, sub + 4, ...: at most 16.  Their sums stay in registers; the 16 TMEM columns are then read,
            // updated and written back as ONE batch, so the tensor-memory round trip is paid once per tile, not per run.
            const uint32_t* keys = sorted + (it & 1) * G_TT;
            const float* col = S + (quad * 32 + lane) * FE_XS;        // this lane's depth row of the x tile
            const uint32_t k0 = keys[lane], k1 = keys[32 + lane];
            const bool v0 = k0 != 0xFFFFFFFFu, v1 = k1 != 0xFFFFFFFFu;
            const uint32_t p0 = __shfl_up_sync(0xffffffffu, k0, 1), p1 = __shfl_up_sync(0xffffffffu, k1, 1);
            const uint32_t k0_last = __shfl_sync(0xffffffffu, k0, 31);
            const bool h0 = v0 && (lane == 0 || (k0 >> 8) != (p0 >> 8));
            const bool h1 = v1 && ((k1 >> 8) != ((lane == 0 ? k0_last : p1) >> 8));
            unsigned long long heads = (unsigned long long)__ballot_sync(0xffffffffu, h0) |
                                       ((unsigned long long)__ballot_sync(0xffffffffu, h1) << 32);
            const int nvalid = __popc(__ballot_sync(0xffffffffu, v0)) + __popc(__ballot_sync(0xffffffffu, v1));
            for (int sk = 0; sk < sub; ++sk) heads &= heads - 1;      // runs 0 .. sub-1 belong to the other warps of this quadrant
            float a[16];
            uint32_t cc[16];
            int len[16];
            int nr = 0;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                a[r] = 0.f; cc[r] = 0u; len[r] = 0;
                if (heads) {                                          // (warp-uniform)
                    const int sb = __ffsll((long long)heads) - 1;
                    heads &= heads - 1;
                    const int eb = heads ? __ffsll((long long)heads) - 1 : nvalid;
                    heads &= heads - 1; heads &= heads - 1; heads &= heads - 1;      // the next three runs are other warps'
                    const uint32_t ks = sb < 32 ? __shfl_sync(0xffffffffu, k0, sb) : __shfl_sync(0xffffffffu, k1, sb - 32);
                    float acc = 0.f;
                    for (int f = sb; f < eb; ++f) {
                        const uint32_t kf = f < 32 ? __shfl_sync(0xffffffffu, k0, f) : __shfl_sync(0xffffffffu, k1, f - 32);
                        acc += col[kf & 255u];
                    }
                    a[r] = acc; cc[r] = ks >> 8; len[r] = eb - sb; nr = r + 1;
                }
            }
            uint32_t v[16];
#pragma unroll
            for (int r = 0; r < 16; ++r)
                if (r < nr) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[r]) : "r"(tq + cc[r]) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int r = 0; r < 16; ++r)
                if (r < nr) {
                    fe_tmem_st1(tq + cc[r], __uint_as_float(v[r]) + a[r]);
                    if (quad == 0 && lane == 0) s_cnt[cc[r]] += float(len[r]);
                }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        gather_st((it + 1) % FE_NST, ev);
    }
    cp_async_wait<0>();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- flush the TMEM accumulators: lane == depth, so each reduction is a coalesced 128-byte segment of one code row
    {
        float* sums = stats;
        float* counts = stats + size_t(K) * D;
        const int d = quad * 32 + lane;
        for (int c0 = 128 * sub; c0 < 128 * sub + 128 && c0 < K; c0 += 16) {
            float v[16];
            fe_tmem_ld16(tq + uint32_t(c0), v);
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c0 + i < K && d < dn && v[i] != 0.f) atomicAdd(&sums[size_t(c0 + i) * D + d], v[i]);
        }
        for (int c = tid; c < K; c += FE_THREADS)
            if (s_cnt[c] != 0.f) atomicAdd(&counts[c], s_cnt[c]);
    }
    {
        double s1 = block_sum(sq, red);
        double s2 = block_sum(msum_local, red);
        double s3 = block_sum(sq_all, red);
        if (tid == 0) {
            atomicAdd(&scalars[VQ_S_SUM_MIN_D], s3);
            atomicAdd(&scalars[VQ_S_COMMIT_SQ], s1);
            atomicAdd(&scalars[VQ_S_MASK_SUM], s2);
            __threadfence();
            unsigned int ticket = atomicAdd(reinterpret_cast<unsigned int*>(&scalars[VQ_S_TICKET]), 1u);
            is_last = (ticket == total_blocks - 1);
        }
        __syncthreads();
        if (is_last && tid == 0) {
            __threadfence();
            volatile double* sc = scalars;
            results[VQ_R_COMMIT] = float(sc[VQ_S_COMMIT_SQ] / (sc[VQ_S_MASK_SUM] * double(D)));
            results[VQ_R_FIT] = float(sc[VQ_S_SUM_MIN_D] / double(K));
            *reinterpret_cast<unsigned int*>(&scalars[VQ_S_TICKET]) = 0u;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

}  // namespace vq
