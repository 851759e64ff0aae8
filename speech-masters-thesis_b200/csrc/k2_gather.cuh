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
__global__ void __launch_bounds__(G_THREADS, VEC ? 4 : 6)      // the kernel is latency-bound: 6 resident CTAs (40 registers) beat 5
gather_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ mask,
              const float* __restrict__ k, const float* __restrict__ grad_xq, const float* __restrict__ grad_commit,
              int64_t N, int D, int64_t T, int K, int Ds,
              float* __restrict__ out, double* __restrict__ scalars, float* __restrict__ results,
              unsigned int total_blocks) {
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

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n = tile / tiles_per_utt, t0 = (tile % tiles_per_utt) * G_TT;
        const int tt = int(min(int64_t(G_TT), T - t0));
        __syncthreads();
        if (tid < G_TT) {
            int c = 0;
            float m = 0.f;
            if (tid < tt) {
                int64_t ci = idx[n * T + t0 + tid];
                c = int(min(max(ci, int64_t(0)), int64_t(K - 1)));
                m = mask ? mask[n * T + t0 + tid] : 1.f;
            }
            s_idx[tid] = c;
            s_mask[tid] = m;
            if (MODE == GM_FWD) msum_local += double(m);
        }
        __syncthreads();
        if (VEC) {
            constexpr int ES = G_TT + 4;
            // ---- gather: one warp per frame, coalesced reads of the codebook row, transposed into Es[depth][frame]
            for (int r = warp; r < tt; r += G_THREADS / 32) {
                const float* src = k + size_t(s_idx[r]) * D;
                for (int d = lane; d < D; d += 32) Es[d * ES + r] = __ldg(src + d);
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
            // ---- gather: one warp per row, coalesced 128-byte reads of the codebook row (L2-resident)
            for (int r = warp; r < tt; r += G_THREADS / 32) {
                const float* src = k + size_t(s_idx[r]) * D;
                float* dst = Es + size_t(r) * Ds;
                for (int d = lane; d < D; d += 32) dst[d] = __ldg(src + d);
            }
            __syncthreads();
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

}  // namespace vq
