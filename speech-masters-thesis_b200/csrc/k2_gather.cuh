// K2: codebook gather + straight-through output + commitment-loss reduction, NCT in / NCT out.
// Replaces bottleneck.py:143-145 (F.embedding), :194 (commit loss), :197 (straight-through),
// :118-124 (row-major -> NCT transpose) and the `* mask` of :201 in ONE pass over x:
//   algorithmic bytes per frame = 4D (read x) + 4D (write x_q) + 8 (idx) + 4 (mask); codebook rows come from L2.
// HBM-bound: every global access is a 128-byte coalesced segment along the frame axis; the gathered
// codebook rows are transposed through shared memory (odd row stride => conflict-free both ways).
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int G_TT = 64;          // frames per tile
constexpr int G_THREADS = 256;

enum GatherMode { GM_FWD = 0, GM_BWD = 1, GM_DECODE = 2 };

// Shared memory: Es[G_TT][Ds] floats with Ds odd, + idx/mask staging.
template <int MODE>
__global__ void __launch_bounds__(G_THREADS)
gather_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ mask,
              const float* __restrict__ k, const float* __restrict__ grad_xq, const float* __restrict__ grad_commit,
              int64_t N, int D, int64_t T, int K, int Ds,
              float* __restrict__ out, double* __restrict__ scalars, float* __restrict__ results,
              unsigned int total_blocks) {
    extern __shared__ __align__(16) float smem[];
    float* Es = smem;                                   // [G_TT][Ds]
    int* s_idx = reinterpret_cast<int*>(Es + size_t(G_TT) * Ds);   // [G_TT]
    float* s_mask = reinterpret_cast<float*>(s_idx + G_TT);        // [G_TT]
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
#pragma unroll 4
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
