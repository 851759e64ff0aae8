// K1 (exact path): FP32 register-tiled distance + argmin on the CUDA cores.
// Replaces bottleneck.py:92-100 (NCT flatten -- fused: tiles are read straight from NCT),
// :129-134 (distance + min) and the `fit` numerator of :140.
//
// Two uses: (1) any shape the tcgen05 kernel does not take (D > 512, unaligned T, ...);
// (2) LIST mode: the exact re-scan of the few rows the tcgen05 kernel flags as unsafe.
#pragma once
#include "vq_common.cuh"
#include "k1_prepare.cuh"

namespace vq {

constexpr int S_BM = 128;   // rows (frames) per tile
constexpr int S_BN = 64;    // codes per inner tile
constexpr int S_BK = 16;    // depth per step

// LIST = false: tile i covers frames [t0, t0+128) of utterance n (tiles never straddle utterances).
// LIST = true : tile i covers row_list[i*128 .. i*128+127]; the count lives in device memory.  The code range is split
//               over blockIdx.y so that a short list still fills the machine; the partial (distance, index) minima
//               meet in list_keys[] through a 64-bit atomicMin whose ordering is exactly "smaller distance, then lower
//               index" (torch.min's tie rule); the last block to finish writes the results.
template <bool LIST>
__global__ void __launch_bounds__(256)
assign_simt_kernel(const float* __restrict__ x, int64_t N, int D, int64_t T,
                   const float* __restrict__ k, const float* __restrict__ ee, int K,
                   int64_t* __restrict__ idx, float* __restrict__ min_d, double* __restrict__ scalars,
                   const int* __restrict__ row_list, AssignHeader* __restrict__ hdr,
                   unsigned long long* __restrict__ list_keys) {
    __shared__ __align__(16) float Xs[S_BK][S_BM];
    __shared__ __align__(16) float Es[S_BK][S_BN + 4];
    __shared__ long long row_off[S_BM];      // offset of x[n, 0, t] for each tile row, -1 when out of range
    __shared__ double red[32];

    // (LIST runs as a programmatic dependent of the tcgen05 kernel: nothing that kernel wrote may be read before this)
    if (LIST) asm volatile("griddepcontrol.wait;" ::: "memory");
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t tiles_per_utt = (T + S_BM - 1) / S_BM;
    const int64_t n_rows_list = LIST ? int64_t(*reinterpret_cast<volatile int*>(&hdr->unsafe_count)) : 0;
    const int64_t n_tiles = LIST ? (n_rows_list + S_BM - 1) / S_BM : N * tiles_per_utt;
    const bool vec_k = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(k) & 15) == 0);
    double tile_sum = 0.0;
    // codes [c_lo, c_hi) for this block (LIST mode: a 64-code-aligned slice per blockIdx.y)
    int c_lo = 0, c_hi = K;
    if (LIST) {
        const int per = ((K + int(gridDim.y) - 1) / int(gridDim.y) + S_BN - 1) / S_BN * S_BN;
        c_lo = min(K, int(blockIdx.y) * per);
        c_hi = min(K, c_lo + per);            // empty slices skip the tile loop but still take a ticket below
    }

    for (int64_t tile = blockIdx.x; tile < n_tiles && c_lo < c_hi; tile += gridDim.x) {
        __syncthreads();
        if (tid < S_BM) {
            long long off = -1;
            if (LIST) {
                int64_t j = tile * S_BM + tid;
                if (j < n_rows_list) {
                    int64_t r = row_list[j];
                    off = (r / T) * int64_t(D) * T + (r % T);
                }
            } else {
                int64_t n = tile / tiles_per_utt, t = (tile % tiles_per_utt) * S_BM + tid;
                if (t < T) off = n * int64_t(D) * T + t;
            }
            row_off[tid] = off;
        }
        __syncthreads();

        float xx[8];
        float bd[8];
        int bi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { xx[i] = 0.f; bd[i] = __int_as_float(0x7f800000); bi[i] = 0x7fffffff; }

        for (int c0 = c_lo; c0 < c_hi; c0 += S_BN) {
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

            for (int d0 = 0; d0 < D; d0 += S_BK) {
                // ---- stage X: 16 x 128 floats, coalesced along the frame axis
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int kk = (tid >> 7) + 2 * i, r = tid & 127;
                    long long off = row_off[r];
                    float v = 0.f;
                    if (off >= 0 && d0 + kk < D) v = x[off + int64_t(d0 + kk) * T];
                    Xs[kk][r] = v;
                }
                // ---- stage E: 64 codes x 16 depth, transposed into [depth][code]
                {
                    int c = tid >> 2, kq = (tid & 3) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0 + c < K) {
                        const float* src = k + size_t(c0 + c) * D + d0 + kq;
                        if (vec_k && d0 + kq + 3 < D) {
                            v = *reinterpret_cast<const float4*>(src);
                        } else {
                            if (d0 + kq + 0 < D) v.x = src[0];
                            if (d0 + kq + 1 < D) v.y = src[1];
                            if (d0 + kq + 2 < D) v.z = src[2];
                            if (d0 + kq + 3 < D) v.w = src[3];
                        }
                    }
                    Es[kq + 0][c] = v.x; Es[kq + 1][c] = v.y; Es[kq + 2][c] = v.z; Es[kq + 3][c] = v.w;
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < S_BK; ++kk) {
                    float4 a0 = *reinterpret_cast<const float4*>(&Xs[kk][ty * 8]);
                    float4 a1 = *reinterpret_cast<const float4*>(&Xs[kk][ty * 8 + 4]);
                    float4 b = *reinterpret_cast<const float4*>(&Es[kk][tx * 4]);
                    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
                    }
                    if (c0 == c_lo) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) xx[i] = fmaf(a[i], a[i], xx[i]);
                    }
                }
                __syncthreads();
            }
            // ---- fold this code tile into the running argmin (codes visited in increasing order)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int c = c0 + tx * 4 + j;
                if (c < c_hi) {
                    float e2 = ee[c];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float d = ref_distance(xx[i], acc[i][j], e2);
                        if (d < bd[i]) { bd[i] = d; bi[i] = c; }
                    }
                }
            }
        }
        // ---- reduce over the 16 lanes that share a row
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                float od = __shfl_xor_sync(0xffffffffu, bd[i], o);
                int oi = __shfl_xor_sync(0xffffffffu, bi[i], o);
                argmin_take(bd[i], bi[i], od, oi);
            }
        }
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int r = ty * 8 + i;
                if (row_off[r] >= 0) {
                    if (LIST) {
                        // orderable bits of the distance (monotone for all finite floats and +inf), then the index
                        const unsigned b = __float_as_uint(bd[i]);
                        const unsigned ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                        const unsigned long long key = (static_cast<unsigned long long>(ord) << 32) | unsigned(bi[i]);
                        atomicMin(&list_keys[tile * S_BM + r], key);
                    } else {
                        const int64_t row = (tile / tiles_per_utt) * T + (tile % tiles_per_utt) * S_BM + r;
                        idx[row] = bi[i] == 0x7fffffff ? 0 : bi[i];
                        if (min_d) min_d[row] = bd[i];
                        tile_sum += double(bd[i]);
                    }
                }
            }
        }
    }
    if (LIST) {
        if (n_rows_list == 0) return;         // common case (speech-like latents): nothing to do, nothing to re-arm
        // the last block to finish writes the results of the whole list (one thread per listed row) and re-arms the header
        __shared__ bool is_last;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned ticket = atomicAdd(&hdr->list_ticket, 1u);
            is_last = ticket == gridDim.x * gridDim.y - 1;
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        double sum = 0.0;
        for (int64_t j = tid; j < n_rows_list; j += blockDim.x) {
            const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(&list_keys[j]);
            const unsigned ord = unsigned(key >> 32);
            const unsigned b = (ord & 0x80000000u) ? (ord & 0x7fffffffu) : ~ord;
            const float d = __uint_as_float(b);
            const unsigned ci = unsigned(key & 0xffffffffu);
            const int row = row_list[j];
            idx[row] = ci == 0x7fffffffu ? 0 : int64_t(ci);
            if (min_d) min_d[row] = d;
            sum += double(d);
        }
        sum = block_sum(sum, red);
        if (tid == 0) {
            if (scalars && n_rows_list) {
                atomicAdd(&scalars[VQ_S_SUM_MIN_D], sum);
                atomicAdd(&scalars[VQ_S_UNSAFE_ROWS], double(n_rows_list));
            }
            hdr->unsafe_count = 0;            // ready for the next vq_assign on this workspace (no memset needed)
            hdr->list_ticket = 0;
        }
        return;
    }
    double s = block_sum(tile_sum, red);
    if (tid == 0 && scalars && s != 0.0) atomicAdd(&scalars[VQ_S_SUM_MIN_D], s);
}

}  // namespace vq
