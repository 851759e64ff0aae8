// K2 forward + K3a in ONE pass over x (training forward, K <= 512, T % 4 == 0):
//   x_q = (x + (e - x)) * mask, commit-loss / fit reductions        (bottleneck.py:143-145,194,197,118-124,201)
//   per-code sums and counts of the valid frames                     (bottleneck.py:64-68)
// The separate K3a kernel re-read x (152 MB at the bench shape) at 0.24 of the HBM peak; here the x tile is already in
// shared memory for the straight-through arithmetic, so the EMA statistics cost no HBM traffic.
//
// Work unit = 64 frames x 64 depths.  A CTA owns ONE 64-deep depth slice (blockIdx.y) and a private [K][64] FP32 slab of
// per-code sums in shared memory (128 KB at K = 512) next to a three-stage ring of x tiles and a two-stage ring of gathered
// codebook rows (87 KB); the CTAs of a slice stride over the frame tiles.  4 warps do the straight-through arithmetic while
// the other 12 do the statistics ("owner computes": warp w owns the codes with code % 12 == w, so no two warps ever touch
// the same slab row and no atomics are needed).  Per tile a warp finds its frames
// with two ballots, groups equal codes with a match (ballot on code == c), sums a group in registers (lane == depth; the
// loads of a group are independent) and adds it to the slab with ONE read-modify-write -- a hot code costs one update per
// tile, not one per frame.  The slab is flushed once per CTA with FP32 reductions.
// Measured history (K = 512, D = 128, 256 utterances; K2 alone 0.098 ms, K3a alone 0.095 ms):
//   accumulators in tensor memory (128 lanes x 512 columns is exactly [D][K]): single-column tcgen05.ld / st cost ~100
//   cycles each whatever their size -- 0.32 ms;  sorted runs (one warp bitonic-sorts the next tile, as the standalone K3a
//   does): the sort is a ~2000-cycle latency chain per 64 x 64 unit that the lock-step iterations cannot hide -- 0.187 ms.
#pragma once
#include "k2_gather.cuh"

namespace vq {

constexpr int FE_THREADS = 512;
constexpr int FE_WARPS = FE_THREADS / 32;
constexpr int FE_DS = 64;                            // depths per work unit (= slab width)
constexpr int FE_XS = G_TT + 4;                      // row stride of the x tile and of the gathered rows (68 floats: 16-byte aligned rows,
                                                     // lane == depth reads of one frame are 4-way bank conflicts instead of 32-way)
constexpr int FE_NX = 3, FE_NE = 2;                  // x stages (two tiles in flight), gathered-row stages
constexpr int FE_CW = 4;                             // warps 0 .. FE_CW-1: straight-through arithmetic; the other FE_OW warps: EMA statistics
constexpr int FE_OW = FE_WARPS - FE_CW;              // (between the same two barriers, so an iteration costs max(), not sum(), of the two)
constexpr int FE_TILE_FLOATS = FE_DS * FE_XS;        // 4352 floats = 17 KB
constexpr int FE_KMAX = 512;
inline size_t fe_smem_bytes(int K) {
    return size_t(FE_NX + FE_NE) * FE_TILE_FLOATS * 4 + GA_RING * G_TT * 12 + (size_t(K) * FE_DS + K) * 4;
}

__global__ void __launch_bounds__(FE_THREADS, 1)
gather_fwd_ema_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ mask,
                      const float* __restrict__ k, int N, int D, int T, int K, float* __restrict__ out,
                      double* __restrict__ scalars, float* __restrict__ results, float* __restrict__ stats,
                      unsigned int total_blocks) {
    extern __shared__ __align__(16) float smem[];
    float* xs_base = smem;                                                                   // [FE_NX][FE_DS][FE_XS]
    float* es_base = smem + FE_NX * FE_TILE_FLOATS;                                          // [FE_NE][FE_DS][FE_XS]
    int64_t* s_idx = reinterpret_cast<int64_t*>(es_base + FE_NE * FE_TILE_FLOATS);           // [GA_RING][G_TT]
    float* s_mask = reinterpret_cast<float*>(s_idx + GA_RING * G_TT);                        // [GA_RING][G_TT]
    float* slab = s_mask + GA_RING * G_TT;                                                   // [K][FE_DS]
    float* s_cnt = slab + size_t(K) * FE_DS;                                                 // [K]  (slice 0 only)
    __shared__ double red[32];
    __shared__ bool is_last;
    __shared__ int2 s_loc[GA_RING];           // (utterance, first frame) of the units in the ring; x < 0: none

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_per_utt = (T + G_TT - 1) / G_TT;
    const int n_units = N * tiles_per_utt;
    const int u0 = blockIdx.x, step = gridDim.x;
    const int d0 = blockIdx.y * FE_DS, dn = min(FE_DS, D - d0);
    const bool first_slice = blockIdx.y == 0;                      // the scalar reductions and the counts are done once, by slice 0

    for (int i = tid; i < K * FE_DS + K; i += FE_THREADS) slab[i] = 0.f;

    auto prefetch_im = [&](int j, int u) {
        if (tid < G_TT) {
            const int slot = j & (GA_RING - 1);
            int2 loc = make_int2(-1, 0);
            bool copied = false;
            if (u < n_units) {
                const int n = u / tiles_per_utt, t0 = (u - n * tiles_per_utt) * G_TT;
                loc = make_int2(n, t0);
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
    auto issue = [&](int st, int ring) {                           // x tile of the unit in ring slot `ring` -> x stage st
        const int2 loc = s_loc[ring];
        if (loc.x >= 0) {
            float* S = xs_base + size_t(st) * FE_TILE_FLOATS;
#pragma unroll
            for (int r = 0; r < FE_DS * (G_TT / 4) / FE_THREADS; ++r) {
                const int i = tid + r * FE_THREADS, d = i >> 4, c4 = (i & 15) * 4;
                if (d < dn && loc.y + c4 < T) cp_async16(S + d * FE_XS + c4, x + (int64_t(loc.x) * D + d0 + d) * T + loc.y + c4);
            }
        }
        cp_async_commit();
    };
    // codebook rows of a unit through registers: warp w takes frames 4w .. 4w+3 and all depth blocks of 8 (lane = 4 * depth + row,
    // so the transposing stores hit 32 different banks); loads are issued one iteration before the stores
    const int g_r = 4 * warp + (lane & 3), g_dd = lane >> 2;
    auto gather_ld = [&](int ring, float (&ev)[FE_DS / 8]) {
        if (s_loc[ring].x >= 0) {
            const int code = int(min(max(s_idx[ring * G_TT + g_r], int64_t(0)), int64_t(K - 1)));
            const float* src = k + size_t(code) * D + d0 + g_dd;
#pragma unroll
            for (int db = 0; db < FE_DS / 8; ++db) ev[db] = (8 * db + g_dd < dn) ? __ldg(src + 8 * db) : 0.f;
        }
    };
    auto gather_st = [&](int st, const float (&ev)[FE_DS / 8]) {
        float* dst = es_base + size_t(st) * FE_TILE_FLOATS + g_dd * FE_XS + g_r;
#pragma unroll
        for (int db = 0; db < FE_DS / 8; ++db) dst[8 * db * FE_XS] = ev[db];
    };
    double sq = 0.0, sq_all = 0.0, msum_local = 0.0;

#pragma unroll
    for (int j = 0; j < GA_AHEAD; ++j) prefetch_im(j, u0 + j * step);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();                                               // (also: the slab is zeroed)
    float ev[FE_DS / 8];
#pragma unroll
    for (int j = 0; j < FE_DS / 8; ++j) ev[j] = 0.f;
    gather_ld(0, ev);
    gather_st(0, ev);
#pragma unroll
    for (int j = 0; j < FE_NX - 1; ++j) issue(j, j);

    int it = 0;
    for (int u = u0; u < n_units; u += step, ++it) {
        gather_ld((it + 1) & (GA_RING - 1), ev);                   // stored at the end of this iteration
        prefetch_im(it + GA_AHEAD, u + GA_AHEAD * step);           // joins the copy group committed by issue() below
        cp_async_wait<FE_NX - 2>();                                // this thread's copies of unit u have landed
        __syncthreads();                                           // ... everyone's; Es[it & 1] complete; stage of unit u-1 free
        issue((it + FE_NX - 1) % FE_NX, (it + FE_NX - 1) & (GA_RING - 1));

        const int2 cur = s_loc[it & (GA_RING - 1)];
        const int n = cur.x, t0 = cur.y;
        const int tt = min(G_TT, T - t0);
        const float* S = xs_base + size_t(it % FE_NX) * FE_TILE_FLOATS;
        const float* Es = es_base + size_t(it % FE_NE) * FE_TILE_FLOATS;
        const float* sm = s_mask + (it & (GA_RING - 1)) * G_TT;
        // ---- straight-through output + loss reductions (bottleneck.py:194-201): warps 0 .. FE_CW-1
        if (warp < FE_CW) {
            const int t4 = (tid & 15) * 4, dg = tid >> 4;
            if (first_slice && tid < G_TT) msum_local += double(sm[tid]);
            if (t4 < tt) {
                const float4 m4 = *reinterpret_cast<const float4*>(sm + t4);
                const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
                const float vv[4] = {m4.x != 0.f ? 1.f : 0.f, m4.y != 0.f ? 1.f : 0.f, m4.z != 0.f ? 1.f : 0.f, m4.w != 0.f ? 1.f : 0.f};
                float accf[4] = {0.f, 0.f, 0.f, 0.f};
                float* dst = out + (int64_t(n) * D + d0) * T + t0 + t4;
#pragma unroll
                for (int d = dg; d < dn; d += FE_CW * 2) {
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
        // ---- EMA statistics of this unit (bottleneck.py:64-68): warp FE_CW + w owns the codes with code % FE_OW == w
        else {
            const int ow = warp - FE_CW;
            const int64_t* ci = s_idx + (it & (GA_RING - 1)) * G_TT;
            const int64_t i0 = ci[lane], i1 = ci[lane + 32];
            const int c0 = (sm[lane] != 0.f && i0 >= 0) ? int(min(i0, int64_t(K - 1))) : -1;
            const int c1 = (sm[lane + 32] != 0.f && i1 >= 0) ? int(min(i1, int64_t(K - 1))) : -1;
            unsigned long long mine = (unsigned long long)__ballot_sync(0xffffffffu, c0 >= 0 && c0 % FE_OW == ow) |
                                      ((unsigned long long)__ballot_sync(0xffffffffu, c1 >= 0 && c1 % FE_OW == ow) << 32);
            const float* col0 = S + lane * FE_XS;                     // this lane's two depth rows of the x tile
            const float* col1 = S + (lane + 32) * FE_XS;
            while (mine) {                                            // (warp-uniform)
                const int f = __ffsll((long long)mine) - 1;
                const int c = f < 32 ? __shfl_sync(0xffffffffu, c0, f) : __shfl_sync(0xffffffffu, c1, f - 32);
                unsigned long long same = ((unsigned long long)__ballot_sync(0xffffffffu, c0 == c) |
                                           ((unsigned long long)__ballot_sync(0xffffffffu, c1 == c) << 32));
                mine &= ~same;
                const int cnt = __popcll(same);
                float a0 = 0.f, a1 = 0.f;
                while (same) {
                    const int g = __ffsll((long long)same) - 1;
                    same &= same - 1;
                    if (lane < dn) a0 += col0[g];
                    if (lane + 32 < dn) a1 += col1[g];
                }
                slab[size_t(c) * FE_DS + lane] += a0;
                slab[size_t(c) * FE_DS + lane + 32] += a1;
                if (first_slice && lane == 0) s_cnt[c] += float(cnt);
            }
        }
        gather_st((it + 1) % FE_NE, ev);                           // Es[(it+1) & 1] was last read in iteration it-1
    }
    cp_async_wait<0>();
    __syncthreads();
    // ---- flush the slab: one FP32 reduction per touched (code, depth)
    {
        float* sums = stats;
        float* counts = stats + size_t(K) * D;
        for (int i = tid; i < K * FE_DS; i += FE_THREADS) {
            const int c = i / FE_DS, d = i % FE_DS;
            const float v = slab[i];
            if (d < dn && v != 0.f) atomicAdd(&sums[size_t(c) * D + d0 + d], v);
        }
        if (first_slice)
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
            if (first_slice) atomicAdd(&scalars[VQ_S_MASK_SUM], s2);
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
}

}  // namespace vq
