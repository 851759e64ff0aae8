// K2: codebook gather + straight-through output + commitment-loss reduction, NCT in / NCT out.
// Replaces bottleneck.py:143-145 (F.embedding), :194 (commit loss), :197 (straight-through),
// :118-124 (row-major -> NCT transpose) and the `* mask` of :201 in ONE pass over x:
//   algorithmic bytes per frame = 4D (read x) + 4D (write x_q) + 8 (idx) + 4 (mask); codebook rows come from L2.
// HBM-bound: every global access is a 128-byte coalesced segment along the frame axis; the gathered
// codebook rows are transposed through shared memory (odd row stride => conflict-free both ways).
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int G_TT = 64;          // frames per tile (128-frame tiles measured no faster and need twice the shared memory)
constexpr int G_THREADS = 256;

enum GatherMode { GM_FWD = 0, GM_BWD = 1, GM_DECODE = 2 };

// Shared memory: gathered codebook rows + idx/mask staging.
//   VEC = false: Es[frame][Ds] (Ds odd), 4-byte accesses along frames -- any T / alignment.
//   VEC = true : Es[depth][G_TT + 4], 16-byte accesses along frames (T % 4 == 0, 16-byte aligned tensors): a warp moves
//                512 contiguous bytes per request and a thread has 8 independent 16-byte loads in flight.
template <int MODE, bool VEC>
__global__ void __launch_bounds__(G_THREADS, (VEC || MODE == GM_DECODE) ? 4 : 6)   // forward/backward: latency-bound, 6 resident CTAs (40 registers) beat 5;
                                                                                   // decode keeps 32 gathered values in registers instead
gather_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ mask,
              const float* __restrict__ k, const float* __restrict__ grad_xq, const float* __restrict__ grad_commit,
              int64_t N, int D, int64_t T, int K, int Ds,
              float* __restrict__ out, double* __restrict__ scalars, float* __restrict__ results,
              unsigned int total_blocks, int dec4) {
    extern __shared__ __align__(16) float smem[];
    float* Es = smem;                                   // [G_TT][Ds] or [D][G_TT + 4]
    const size_t es_floats = ((VEC ? size_t(D) * (G_TT + 4) : size_t(G_TT) * Ds) + 3) & ~size_t(3);
    float* s_mask = Es + es_floats;                                // [G_TT], 16-byte aligned (read as float4)
    int* s_idx = reinterpret_cast<int*>(s_mask + G_TT);            // [G_TT]
    __shared__ double red[32];
    __shared__ bool is_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tiles_per_utt = (T + G_TT - 1) / G_TT;
    const int64_t n_tiles = N * tiles_per_utt;

    float gscale = 0.f;
    if (MODE == GM_BWD) {
        // 2 * grad_commit / (sum(mask) * D)      (d commit / d x, bottleneck.py:194)
        double msum = scalars[VQ_S_MASK_SUM];
        gscale = float(2.0 * double(*grad_commit) / (msum * double(D)));
    }
    double sq = 0.0, sq_all = 0.0, msum_local = 0.0;

    // index and mask of frame `tid` of a tile, loaded one tile ahead (they stream from HBM: a synchronous load would put a
    // DRAM round trip into every tile of every block)
    int64_t pre_ci = 0;
    float pre_m = 0.f;
    auto load_im = [&](int64_t tile) {
        pre_ci = 0; pre_m = 0.f;
        if (tid < G_TT && tile < n_tiles) {
            const int64_t n = tile / tiles_per_utt, t = (tile % tiles_per_utt) * G_TT + tid;
            if (t < T) {
                pre_ci = idx[n * T + t];
                pre_m = mask ? mask[n * T + t] : 1.f;
            }
        }
    };
    constexpr bool PREFETCH = MODE == GM_DECODE;       // (measured: the extra registers cost the 40-register forward/backward more than they save)
    if (PREFETCH) load_im(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n = tile / tiles_per_utt, t0 = (tile % tiles_per_utt) * G_TT;
        const int tt = int(min(int64_t(G_TT), T - t0));
        __syncthreads();
        if (!PREFETCH) load_im(tile);
        if (tid < G_TT) {
            s_idx[tid] = int(min(max(pre_ci, int64_t(0)), int64_t(K - 1)));
            s_mask[tid] = pre_m;
            if (MODE == GM_FWD) msum_local += double(pre_m);
        }
        if (PREFETCH) load_im(tile + gridDim.x);
        __syncthreads();
        if (VEC) {
            constexpr int ES = G_TT + 4;
            // ---- gather, transposed into Es[depth][frame]: a warp copies 4 codebook rows x 8 depths per instruction
            // (lane = 4 * depth + row), so the 32 stores of an instruction fall into 32 different banks (row stride 68)
            {
                const int rr = lane & 3, dd = lane >> 2;
                for (int r4 = warp * 4; r4 < tt; r4 += (G_THREADS / 32) * 4) {
                    const int r = r4 + rr;                                       // (tt is a multiple of 4 on this path)
                    const float* src = k + size_t(s_idx[r]) * D + dd;
                    float* dst = Es + dd * ES + r;
#pragma unroll 4
                    for (int d8 = 0; d8 + dd < D; d8 += 8) dst[d8 * ES] = __ldg(src + d8);
                }
            }
            __syncthreads();
            // ---- stream: G_TT/4 threads cover the frames of one depth with float4
            constexpr int TQ = G_TT / 4;                     // float4 groups per depth row
            const int t4 = (tid % TQ) * 4, dg = tid / TQ;
            if (t4 < tt) {
                const float4 m4 = *reinterpret_cast<const float4*>(s_mask + t4);
                const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
                const float vv[4] = {m4.x != 0.f ? 1.f : 0.f, m4.y != 0.f ? 1.f : 0.f, m4.z != 0.f ? 1.f : 0.f, m4.w != 0.f ? 1.f : 0.f};
                const size_t base = (size_t(n) * D) * T + t0 + t4;
                float acc_all = 0.f, acc_valid = 0.f;
#pragma unroll 4
                for (int d = dg; d < D; d += G_THREADS / TQ) {
                    const size_t o = base + size_t(d) * T;
                    const float4 e4 = *reinterpret_cast<const float4*>(Es + d * ES + t4);
                    const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
                    float oo[4];
                    if (MODE == GM_DECODE) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) oo[j] = ee[j];
                    } else {
                        const float4 x4 = ld_stream4(reinterpret_cast<const float4*>(x + o));
                        const float xx[4] = {x4.x, x4.y, x4.z, x4.w};
                        if (MODE == GM_FWD) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float diff = __fsub_rn(ee[j], xx[j]);                   // (x_d - x)
                                oo[j] = __fmul_rn(__fadd_rn(xx[j], diff), mm[j]);             // (x + (x_d - x)) * mask
                                const float d2 = diff * diff;
                                acc_all += d2;
                                acc_valid = fmaf(d2, vv[j], acc_valid);
                            }
                        } else {
                            float gg[4] = {0.f, 0.f, 0.f, 0.f};
                            if (grad_xq) {
                                const float4 g4 = ld_stream4(reinterpret_cast<const float4*>(grad_xq + o));
                                gg[0] = g4.x; gg[1] = g4.y; gg[2] = g4.z; gg[3] = g4.w;
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float g = __fmul_rn(gg[j], mm[j]);
                                if (vv[j] != 0.f) g = fmaf(gscale, __fsub_rn(xx[j], ee[j]), g);
                                oo[j] = g;
                            }
                        }
                    }
                    st_stream4(reinterpret_cast<float4*>(out + o), make_float4(oo[0], oo[1], oo[2], oo[3]));
                }
                sq_all += double(acc_all);
                sq += double(acc_valid);
            }
        } else {
            // ---- gather: one warp per row, coalesced 128-byte reads of the codebook row (L2-resident).  Decode, D <= 128: all
            // the loads of a warp's 8 rows are issued before the first store (one L2 round trip per tile instead of eight).
            if (MODE == GM_DECODE && D <= 128) {
                float v[G_TT / (G_THREADS / 32)][4];
#pragma unroll
                for (int i = 0; i < G_TT / (G_THREADS / 32); ++i) {
                    const int r = warp + i * (G_THREADS / 32);
                    const float* src = k + size_t(s_idx[r]) * D;              // (rows beyond tt hold code 0: harmless reads)
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[i][q] = lane + 32 * q < D ? __ldg(src + lane + 32 * q) : 0.f;
                }
#pragma unroll
                for (int i = 0; i < G_TT / (G_THREADS / 32); ++i) {
                    float* dst = Es + size_t(warp + i * (G_THREADS / 32)) * Ds;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (lane + 32 * q < D) dst[lane + 32 * q] = v[i][q];
                }
            } else {
                for (int r = warp; r < tt; r += G_THREADS / 32) {
                    const float* src = k + size_t(s_idx[r]) * D;
                    float* dst = Es + size_t(r) * Ds;
                    for (int d = lane; d < D; d += 32) dst[d] = __ldg(src + d);
                }
            }
            __syncthreads();
            if (MODE == GM_DECODE && dec4) {
                // decode with 16-byte stores (T % 4 == 0, aligned output): a thread owns 4 consecutive frames and walks the
                // depth; 4 shared-memory reads (2-way conflicts) + 1 store per 4 elements.  The 4-byte loop below spent
                // ~26 thread instructions per element and was issue-bound (69 % issue utilisation, DRAM 34 % busy).
                const int t4 = (tid & 15) * 4, dg = tid >> 4;             // 16 frame quads x 16 depth groups
                if (t4 < tt) {
                    const float* e0 = Es + size_t(t4) * Ds;
                    float* dst = out + (size_t(n) * D) * T + t0 + t4;
#pragma unroll 4
                    for (int d = dg; d < D; d += G_THREADS / 16)
                        st_stream4(reinterpret_cast<float4*>(dst + size_t(d) * T), make_float4(e0[d], e0[Ds + d], e0[2 * Ds + d], e0[3 * Ds + d]));
                }
                continue;
            }
            // ---- stream: lanes along frames (coalesced), warps along depth
            const int t = tid & (G_TT - 1);
            const int dgrp = tid / G_TT;                       // 0..3
            if (t < tt) {
                const float m = s_mask[t];
                const bool valid = m != 0.f;
                const float* er = Es + size_t(t) * Ds;
                const size_t base = (size_t(n) * D) * T + t0 + t;
                float acc = 0.f;
#pragma unroll 8
                for (int d = dgrp; d < D; d += G_THREADS / G_TT) {
                    const size_t o = base + size_t(d) * T;
                    const float e = er[d];
                    if (MODE == GM_DECODE) {
                        st_stream(out + o, e);
                    } else {
                        const float xv = ld_stream(x + o);
                        if (MODE == GM_FWD) {
                            const float diff = __fsub_rn(e, xv);              // (x_d - x)
                            st_stream(out + o, __fmul_rn(__fadd_rn(xv, diff), m));   // (x + (x_d - x)) * mask
                            acc = fmaf(diff, diff, acc);
                        } else {
                            float g = grad_xq ? __fmul_rn(ld_stream(grad_xq + o), m) : 0.f;
                            if (valid) g = fmaf(gscale, __fsub_rn(xv, e), g);
                            st_stream(out + o, g);
                        }
                    }
                }
                sq_all += double(acc);                    // ||k[idx] - x||^2 of EVERY row: the numerator of `fit`
                if (valid) sq += double(acc);
            }
        }
    }
    if (MODE == GM_FWD) {
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
}

// ---- asynchronous variant (T % 4 == 0, 16-byte aligned tensors): one persistent CTA per SM; the x (and grad) tile AND the
// gathered codebook rows of the next one or two work units are in flight (cp.async) while the current unit is computed
// entirely from shared memory.  The synchronous kernel above keeps only ~half a tile per CTA in flight between its
// barriers and measured 0.47 of the HBM peak; this one is bound by the copy engine instead of by load latency.
//   work unit = 64 frames x (up to) 128 depths;  stage = [x tile 32 KB] (+ [grad tile 32 KB]) + [codebook rows 34 KB]
// Gather mapping: a warp copies 4 codebook rows x 8 depths per instruction (lane = 4*depth + row), which makes the
// transposing 4-byte writes into Es[depth][frame] (row stride 68 floats) hit 32 different banks.
constexpr int GA_THREADS = 512;
constexpr int GA_DS = 128;                   // depths per work unit
constexpr int GA_ES = G_TT + 4;              // Es row stride (floats)
constexpr int GA_RING = 8, GA_AHEAD = 4;     // index/mask ring: slots, prefetch distance in work units
template <int MODE> struct GaCfg {
    static constexpr int NST = MODE == GM_BWD ? 2 : 3;                              // stages in the ring
    static constexpr int NX = MODE == GM_DECODE ? 0 : (MODE == GM_BWD ? 2 : 1);     // streamed input tiles per stage
    static constexpr int STAGE_FLOATS = NX * GA_DS * G_TT + GA_DS * GA_ES;
    static constexpr size_t SMEM = size_t(NST) * STAGE_FLOATS * 4 + GA_RING * G_TT * 12;
    static constexpr int CTAS_PER_SM = MODE == GM_DECODE ? 2 : 1;    // decode has no streamed input: 2 x 110 KB fit, and a second CTA
                                                                      // covers the L2 latency of the gather
};

template <int MODE>
__global__ void __launch_bounds__(GA_THREADS, GaCfg<MODE>::CTAS_PER_SM)
gather_async_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ mask,
                    const float* __restrict__ k, const float* __restrict__ grad_xq, const float* __restrict__ grad_commit,
                    int N, int D, int T, int K, float* __restrict__ out, double* __restrict__ scalars,
                    float* __restrict__ results, unsigned int total_blocks) {
    constexpr int NST = GaCfg<MODE>::NST, NX = GaCfg<MODE>::NX, STAGE_FLOATS = GaCfg<MODE>::STAGE_FLOATS;
    extern __shared__ __align__(16) float smem[];
    int64_t* s_idx = reinterpret_cast<int64_t*>(smem + size_t(NST) * STAGE_FLOATS);   // [GA_RING][G_TT]  ring over work units
    float* s_mask = reinterpret_cast<float*>(s_idx + GA_RING * G_TT);                  // [GA_RING][G_TT]
    __shared__ double red[32];
    __shared__ bool is_last;
    __shared__ int4 s_loc[GA_RING];           // (utterance, first frame, first depth, exists) of the units in the ring

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_per_utt = (T + G_TT - 1) / G_TT;
    const int n_slices = (D + GA_DS - 1) / GA_DS;
    const int n_units = N * tiles_per_utt * n_slices;                              // host guarantees < 2^31
    const int u0 = blockIdx.x, step = gridDim.x;
    const bool have_grad = MODE == GM_BWD && grad_xq != nullptr;

    auto locate = [&](int u, int& n, int& t0, int& d0) {
        const int tile = u / n_slices;
        d0 = (u - tile * n_slices) * GA_DS;
        n = tile / tiles_per_utt;
        t0 = (tile - n * tiles_per_utt) * G_TT;
    };
    // indices and mask of local unit j (= global unit u) -> ring slot j % GA_RING, asynchronously, GA_AHEAD units ahead of
    // their first use (they come from HBM: a synchronous load would put a DRAM round trip into every iteration).
    // The two integer divisions of locate() run here once per unit, on 64 threads; everyone else reads s_loc.
    auto prefetch_im = [&](int j, int u) {
        if (tid < G_TT) {
            const int slot = j & (GA_RING - 1);
            int4 loc = make_int4(0, 0, 0, 0);
            bool copied = false;
            if (u < n_units) {
                int n, t0, d0;
                locate(u, n, t0, d0);
                loc = make_int4(n, t0, d0, 1);
                const int t = t0 + tid;
                if (t < T) {
                    cp_async8(s_idx + slot * G_TT + tid, idx + int64_t(n) * T + t);
                    if (mask) cp_async4(s_mask + slot * G_TT + tid, mask + int64_t(n) * T + t);
                    else s_mask[slot * G_TT + tid] = 1.f;
                    copied = true;
                }
            }
            if (!copied) { s_idx[slot * G_TT + tid] = 0; s_mask[slot * G_TT + tid] = 0.f; }
            if (tid == 0) s_loc[slot] = loc;
        }
    };
    auto issue = [&](int st, int ring) {
        const int4 loc = s_loc[ring];
        if (loc.w) {
            const int n = loc.x, t0 = loc.y, d0 = loc.z;
            const int dn = min(GA_DS, D - d0);
            float* S = smem + size_t(st) * STAGE_FLOATS;
            if (NX > 0) {
                // streamed tiles: 128 depths x 16 chunks of 4 frames, 4 chunks per thread
#pragma unroll
                for (int r = 0; r < GA_DS * (G_TT / 4) / GA_THREADS; ++r) {
                    const int i = tid + r * GA_THREADS, d = i >> 4, c4 = (i & 15) * 4;
                    if (d < dn && t0 + c4 < T) {
                        const int64_t o = (int64_t(n) * D + d0 + d) * T + t0 + c4;
                        cp_async16(S + d * G_TT + c4, x + o);
                        if (NX > 1 && have_grad) cp_async16(S + GA_DS * G_TT + d * G_TT + c4, grad_xq + o);
                    }
                }
            }
        }
        cp_async_commit();
    };
    // codebook rows of unit u through registers (4-byte cp.async is serialised per element by the hardware: measured
    // 1.6x slower than the synchronous kernel): warp w takes frames 4w .. 4w+3 and all depth blocks of 8; the loads are
    // issued one iteration before the stores, so their L2 latency hides behind the computation of the current unit.
    const int g_r = 4 * warp + (lane & 3), g_dd = lane >> 2;
    auto gather_ld = [&](int ring, float (&ev)[GA_DS / 8]) {
        const int4 loc = s_loc[ring];
        if (loc.w) {
            const int d0 = loc.z;
            const int dn = min(GA_DS, D - d0);
            const int code = int(min(max(s_idx[ring * G_TT + g_r], int64_t(0)), int64_t(K - 1)));
            const float* src = k + size_t(code) * D + d0 + g_dd;
#pragma unroll
            for (int db = 0; db < GA_DS / 8; ++db) ev[db] = (8 * db + g_dd < dn) ? __ldg(src + 8 * db) : 0.f;
        }
    };
    auto gather_st = [&](int st, const float (&ev)[GA_DS / 8]) {
        float* dst = smem + size_t(st) * STAGE_FLOATS + NX * GA_DS * G_TT + g_dd * GA_ES + g_r;
#pragma unroll
        for (int db = 0; db < GA_DS / 8; ++db) dst[8 * db * GA_ES] = ev[db];
    };

    float gscale = 0.f;
    if (MODE == GM_BWD) gscale = float(2.0 * double(*grad_commit) / (scalars[VQ_S_MASK_SUM] * double(D)));
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
#pragma unroll
    for (int j = 0; j < NST - 1; ++j) issue(j, j);                // (each call commits one group)

    int it = 0;
    for (int u = u0; u < n_units; u += step, ++it) {
        gather_ld((it + 1) & (GA_RING - 1), ev);                  // stored at the end of this iteration
        prefetch_im(it + GA_AHEAD, u + GA_AHEAD * step);          // joins the copy group committed by issue() below
        cp_async_wait<NST - 2>();                                 // this thread's copies of unit u have landed
        __syncthreads();                                          // ... everyone's; the stage of unit u-1 is free again
        issue((it + NST - 1) % NST, (it + NST - 1) & (GA_RING - 1));

        const int4 cur = s_loc[it & (GA_RING - 1)];
        const int n = cur.x, t0 = cur.y, d0 = cur.z;
        const int tt = min(G_TT, T - t0), dn = min(GA_DS, D - d0);
        const float* S = smem + size_t(it % NST) * STAGE_FLOATS;
        const float* Es = S + NX * GA_DS * G_TT;
        const float* sm = s_mask + (it & (GA_RING - 1)) * G_TT;
        const int t4 = (tid & 15) * 4, dg = tid >> 4;
        if (MODE == GM_FWD && d0 == 0 && tid < G_TT) msum_local += double(sm[tid]);
        if (t4 < tt) {
            const float4 m4 = *reinterpret_cast<const float4*>(sm + t4);
            const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
            const float vv[4] = {m4.x != 0.f ? 1.f : 0.f, m4.y != 0.f ? 1.f : 0.f, m4.z != 0.f ? 1.f : 0.f, m4.w != 0.f ? 1.f : 0.f};
            float accf[4] = {0.f, 0.f, 0.f, 0.f};                 // sum over depth of (x_d - x)^2, per frame
            float* dst = out + (int64_t(n) * D + d0) * T + t0 + t4;
#pragma unroll 4
            for (int d = dg; d < dn; d += GA_THREADS / 16) {
                const float4 e4 = *reinterpret_cast<const float4*>(Es + d * GA_ES + t4);
                const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
                float oo[4];
                if (MODE == GM_DECODE) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) oo[j] = ee[j];
                } else {
                    const float4 x4 = *reinterpret_cast<const float4*>(S + d * G_TT + t4);
                    const float xx[4] = {x4.x, x4.y, x4.z, x4.w};
                    if (MODE == GM_FWD) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float diff = __fsub_rn(ee[j], xx[j]);                   // (x_d - x)
                            oo[j] = __fmul_rn(__fadd_rn(xx[j], diff), mm[j]);             // (x + (x_d - x)) * mask
                            accf[j] = fmaf(diff, diff, accf[j]);
                        }
                    } else {
                        float gg[4] = {0.f, 0.f, 0.f, 0.f};
                        if (have_grad) {
                            const float4 g4 = *reinterpret_cast<const float4*>(S + GA_DS * G_TT + d * G_TT + t4);
                            gg[0] = g4.x; gg[1] = g4.y; gg[2] = g4.z; gg[3] = g4.w;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float g = __fmul_rn(gg[j], mm[j]);
                            if (vv[j] != 0.f) g = fmaf(gscale, __fsub_rn(xx[j], ee[j]), g);
                            oo[j] = g;
                        }
                    }
                }
                st_stream4(reinterpret_cast<float4*>(dst + int64_t(d) * T), make_float4(oo[0], oo[1], oo[2], oo[3]));
            }
            if (MODE == GM_FWD) {
                sq_all += double((accf[0] + accf[1]) + (accf[2] + accf[3]));     // every frame: the numerator of `fit`
                sq += double((accf[0] * vv[0] + accf[1] * vv[1]) + (accf[2] * vv[2] + accf[3] * vv[3]));   // valid frames: commit loss
            }
        }
        gather_st((it + 1) % NST, ev);                            // that stage's rows were last read NST-1 iterations ago
    }
    cp_async_wait<0>();
    if (MODE == GM_FWD) {
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
}

// out[j, :] = x[n_j, :, t_j]   (restart rows; bottleneck.py:40,70 touches only K rows this way)
__global__ void gather_rows_kernel(const float* __restrict__ x, const int64_t* __restrict__ rows, int64_t n_rows,
                                   int64_t N, int D, int64_t T, float* __restrict__ out) {
    int64_t j = blockIdx.x;
    if (j >= n_rows) return;
    int64_t r = rows[j];
    if (r < 0 || r >= N * T) return;
    const float* src = x + (r / T) * int64_t(D) * T + (r % T);
    for (int d = threadIdx.x; d < D; d += blockDim.x) out[j * D + d] = src[int64_t(d) * T];
}


// ---- device-side restart rows (vq_restart_rows_device): valid frames per utterance -> prefix sums -> K uniform draws
__global__ void __launch_bounds__(256) restart_count_kernel(const float* __restrict__ mask, int64_t N, int64_t T,
                                                           long long* __restrict__ prefix) {
    const int64_t n = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;          // one warp per utterance
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    long long c = 0;
    if (mask == nullptr) c = lane == 0 ? T : 0;
    else for (int64_t t = lane; t < T; t += 32) c += mask[n * T + t] != 0.f ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) prefix[n + 1] = c;
    if (n == 0 && lane == 0) prefix[0] = 0;
}
// in-place inclusive scan of prefix[1..N] by one block
__global__ void __launch_bounds__(1024) restart_scan_kernel(long long* __restrict__ prefix, int64_t N) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < N; base += 1024) {
        const int64_t i = base + threadIdx.x;
        long long v = i < N ? prefix[i + 1] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        if (lane == 31) warp_tot[warp] = v;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const long long before = carry + (warp ? warp_tot[warp - 1] : 0);
        if (i < N) prefix[i + 1] = v + before;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_tot[31];
        __syncthreads();
    }
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// one block per drawn row
__global__ void __launch_bounds__(128) restart_select_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                            const long long* __restrict__ prefix, int64_t N, int D, int64_t T,
                                                            int K, uint64_t seed, const uint64_t* __restrict__ seed_dev,
                                                            float* __restrict__ out) {
    __shared__ long long s_src;               // offset of x[n, 0, t], or -1
    const int j = blockIdx.x, lane = threadIdx.x & 31;
    if (seed_dev) seed ^= splitmix64(*seed_dev);
    const long long m = prefix[N];
    if (threadIdx.x < 32) {
        long long src = -1;
        if (m > 0) {
            const uint64_t h = splitmix64(seed ^ (uint64_t(j) * 0xD1342543DE82EF95ull));
            long long r = (long long)(__umul64hi(h, uint64_t(m)));                // uniform in [0, m)
            int64_t lo = 0, hi = N;                                               // utterance n: prefix[n] <= r < prefix[n+1]
            while (hi - lo > 1) {
                const int64_t mid = (lo + hi) >> 1;
                if (prefix[mid] <= r) lo = mid; else hi = mid;
            }
            const int64_t n = lo;
            long long q = r - prefix[n];                                          // q-th valid frame of utterance n
            int64_t t_found = -1;
            if (mask == nullptr) t_found = q;
            else {
                for (int64_t t0 = 0; t0 < T && t_found < 0; t0 += 32) {
                    const bool v = t0 + lane < T && mask[n * T + t0 + lane] != 0.f;
                    const unsigned b = __ballot_sync(0xffffffffu, v);
                    const int c = __popc(b);
                    if (q < c) t_found = t0 + __fns(b, 0, int(q) + 1);
                    else q -= c;
                }
            }
            if (t_found >= 0) src = n * int64_t(D) * T + t_found;
        }
        if (lane == 0) s_src = src;
    }
    __syncthreads();
    const long long src = s_src;
    const float jitter = m < K ? 0.01f * rsqrtf(float(D)) : 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float v = src >= 0 ? x[src + int64_t(d) * T] : 0.f;
        if (jitter != 0.f && src >= 0) {                                          // Box-Muller on two hashed uniforms
            const uint64_t h = splitmix64(seed ^ splitmix64(uint64_t(j) * uint64_t(D) + uint64_t(d) + 0x5851F42D4C957F2Dull));
            const float u1 = (float(uint32_t(h >> 40)) + 1.f) * (1.f / 16777217.f);
            const float u2 = float(uint32_t(h) >> 8) * (1.f / 16777216.f);
            v += jitter * sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
        }
        out[size_t(j) * D + d] = v;
    }
}

}  // namespace vq
