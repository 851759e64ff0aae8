// K1 (exact path): FP32 register-tiled distance + argmin on the CUDA cores.
// Replaces bottleneck.py:92-100 (NCT flatten -- fused: tiles are read straight from NCT),
// :129-134 (distance + min) and the `fit` numerator of :140.
//
// Two kernels: (1) assign_simt_kernel, for any shape the tcgen05 kernel does not take (D > 512, unaligned T, ...);
// (2) assign_list_kernel, the exact re-scan of the few frames the tcgen05 kernel flags as unsafe.
#pragma once
#include "vq_common.cuh"
#include "k1_prepare.cuh"

namespace vq {

constexpr int S_BM = 128;   // rows (frames) per tile
constexpr int S_BN = 64;    // codes per inner tile
constexpr int S_BK = 16;    // depth per step

// Tile i covers frames [t0, t0+128) of utterance n (tiles never straddle utterances).
template <typename XT>
__global__ void __launch_bounds__(256)
assign_simt_kernel(const XT* __restrict__ x, int64_t N, int D, int64_t T,
                   const float* __restrict__ k, const float* __restrict__ ee, int K,
                   int64_t* __restrict__ idx, float* __restrict__ min_d, double* __restrict__ scalars) {
    __shared__ __align__(16) float Xs[S_BK][S_BM];
    __shared__ __align__(16) float Es[S_BK][S_BN + 4];
    __shared__ long long row_off[S_BM];      // offset of x[n, 0, t] for each tile row, -1 when out of range
    __shared__ double red[32];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t tiles_per_utt = (T + S_BM - 1) / S_BM;
    const int64_t n_tiles = N * tiles_per_utt;
    const bool vec_k = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(k) & 15) == 0);
    double tile_sum = 0.0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();
        if (tid < S_BM) {
            long long off = -1;
            int64_t n = tile / tiles_per_utt, t = (tile % tiles_per_utt) * S_BM + tid;
            if (t < T) off = n * int64_t(D) * T + t;
            row_off[tid] = off;
        }
        __syncthreads();

        float xx[8];
        float bd[8];
        int bi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { xx[i] = 0.f; bd[i] = __int_as_float(0x7f800000); bi[i] = 0x7fffffff; }

        for (int c0 = 0; c0 < K; c0 += S_BN) {
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
                    if (off >= 0 && d0 + kk < D) v = x_to_float(x[off + int64_t(d0 + kk) * T]);
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
                    if (c0 == 0) {
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
                if (c < K) {
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
                    const int64_t row = (tile / tiles_per_utt) * T + (tile % tiles_per_utt) * S_BM + r;
                    idx[row] = bi[i] == 0x7fffffff ? 0 : bi[i];
                    if (min_d) min_d[row] = bd[i];
                    tile_sum += double(bd[i]);
                }
            }
        }
    }
    double s = block_sum(tile_sum, red);
    if (tid == 0 && scalars && s != 0.0) atomicAdd(&scalars[VQ_S_SUM_MIN_D], s);
}

// ------------------------------------------------------------------------------------------------
// Exact re-scan of the frames the tcgen05 kernel could not prove safe (bottleneck.py:129-134 for those rows, FP32, lowest
// index on ties like torch.min).  The tcgen05 pass leaves, per listed frame, a 32-bit mask: bit (16 g + j) set means "the
// exact winner may be a code of scan group g (group of code tile nt = (nt >> own_shift) & 1) in residue chain j (code % 16 == j)"; every other code was
// proven out of reach by its FP16 score (see finish() in k1_assign_tc.cuh).  Typically two chains survive: K/16 codes
// instead of K -- the re-scan of 2 % of a batch went from 0.11 ms (all codes, 128-row tiles, 64-bit atomics) to ~0.01 ms.
//   one WARP per listed frame; the frame's D values live in registers (lane = depth quad), a candidate code row is one
//   coalesced 16-byte load per lane, eight candidates are reduced together with a transposing butterfly (9 shuffles).
// Launched as a programmatic dependent of the tcgen05 kernel; the frame count lives in device memory.
constexpr int L_WARPS = 8;      // (592 blocks of 8 warps: the usual empty-worklist launch costs ~1.3 us less than 2368 blocks of 4)

// One chain segment = the eight codes c = 128 nt + 16 b + res (b = 0..7) of residue chain `res` in code tile `nt`.
// dot8 leaves in every lane the partial dot products of its depth slice with the eight codes (loads first, then FMAs).
template <bool VEC, int MAXQ>
__device__ __forceinline__ void list_dot8(const float* __restrict__ k, int D, int K, int nt, int res, int lane,
                                          const float (&xr)[4 * MAXQ], float (&part)[8]) {
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int c = nt * 128 + 16 * b + res;
        float acc = 0.f;
        if (c < K) {
            const float* er = k + size_t(c) * D;
            if (VEC) {
#pragma unroll
                for (int q = 0; q < MAXQ; ++q)
                    if (4 * (lane + 32 * q) < D) {
                        const float4 e4 = __ldg(reinterpret_cast<const float4*>(er) + lane + 32 * q);
                        acc = fmaf(xr[4 * q + 0], e4.x, acc); acc = fmaf(xr[4 * q + 1], e4.y, acc);
                        acc = fmaf(xr[4 * q + 2], e4.z, acc); acc = fmaf(xr[4 * q + 3], e4.w, acc);
                    }
            } else {
#pragma unroll
                for (int i = 0; i < 4 * MAXQ; ++i)
                    if (lane + 32 * i < D) acc = fmaf(xr[i], __ldg(er + lane + 32 * i), acc);
            }
        }
        part[b] = acc;
    }
}
// transposing butterfly: afterwards every lane holds the full dot product of code b = 4 bit4 + 2 bit3 + bit2 of its lane id
__device__ __forceinline__ float list_reduce8(float (&part)[8], int lane) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const bool hi = (lane & 16) != 0;
        const float keep = hi ? part[b + 4] : part[b], give = hi ? part[b] : part[b + 4];
        part[b] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const bool hi = (lane & 8) != 0;
        const float keep = hi ? part[b + 2] : part[b], give = hi ? part[b] : part[b + 2];
        part[b] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
    }
    {
        const bool hi = (lane & 4) != 0;
        const float keep = hi ? part[1] : part[0], give = hi ? part[0] : part[1];
        part[0] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    }
    part[0] += __shfl_xor_sync(0xffffffffu, part[0], 2);
    part[0] += __shfl_xor_sync(0xffffffffu, part[0], 1);
    return part[0];
}
// walks the (group, chain, code tile) segments a mask selects, in increasing group / chain / tile order
// (own_shift: scan group of code tile nt = (nt >> own_shift) & 1, see plan_assign_tc)
// tiles: per group, the 24-bit map of its code tiles that can hold the winner (bit = (own index * scale) >> 16, as in the main kernel)
struct ListWalk {
    uint32_t mask;
    uint2 tiles;
    uint32_t scale;
    int n_code_tiles, g, res, nt, sh;
    __device__ __forceinline__ ListWalk(uint32_t m, uint2 t, uint32_t sc, int n, int own_shift)
        : mask(m), tiles(t), scale(sc), n_code_tiles(n), g(-1), res(0), nt(1 << 30), sh(own_shift) {}
    __device__ __forceinline__ bool tile_flagged() const {
        const uint32_t oi = sh ? ((uint32_t(nt) >> 2) << 1 | (uint32_t(nt) & 1u)) : (uint32_t(nt) >> 1);
        return (((g ? tiles.y : tiles.x) >> ((oi * scale) >> 16)) & 1u) != 0u;
    }
    __device__ __forceinline__ bool next() {
        for (;;) {
            if (!step()) return false;
            if (tile_flagged()) return true;
        }
    }
    __device__ __forceinline__ bool step() {
        nt += sh ? ((nt & 1) ? 3 : 1) : 2;                 // the group's next code tile
        while (nt >= n_code_tiles) {
            // next chain of this group, else the next group
            uint32_t mg = g >= 0 ? (mask >> (16 * g)) & 0xFFFFu & ~((2u << res) - 1u) : 0u;
            while (mg == 0u) {
                if (++g > 1) return false;
                mg = (mask >> (16 * g)) & 0xFFFFu;
                res = -1;
                if (mg) break;
            }
            res = __ffs(mg) - 1;
            nt = g << sh;                                      // the group's first code tile
        }
        return true;
    }
};

// The per-frame work is a chain of dependent L2 / DRAM round trips (frame id -> x -> code rows -> ||e||^2), so the kernel is
// latency-bound: two segments are evaluated per step (16 code rows and both norms in flight at once) and frames are spread
// one per warp over as many resident warps as the register budget allows (MAXQ = 1 for D <= 128).
template <bool VEC, int MAXQ, typename XT>
__global__ void __launch_bounds__(L_WARPS * 32)
assign_list_kernel(const XT* __restrict__ x, int64_t N, int D, int64_t T, const float* __restrict__ k,
                   const float* __restrict__ ee, int K, int n_code_tiles, int own_shift, int64_t* __restrict__ idx, float* __restrict__ min_d,
                   double* __restrict__ scalars, const int* __restrict__ row_list, const uint32_t* __restrict__ row_mask,
                   const uint2* __restrict__ row_tiles, uint32_t tile_scale, AssignHeader* __restrict__ hdr, unsigned int* hard_hint) {
    __shared__ double red[32];
    __shared__ bool is_last;
    asm volatile("griddepcontrol.wait;" ::: "memory");     // nothing the tcgen05 kernel wrote may be read before this
    const int n_list = *reinterpret_cast<volatile int*>(&hdr->unsafe_count);
    if (n_list == 0) {                                     // common case (speech-like latents): nothing to do, nothing to re-arm
        if (hard_hint && blockIdx.x == 0 && threadIdx.x == 0) *hard_hint = 0u;
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b_mine = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const float inf = __int_as_float(0x7f800000);
    double sum = 0.0;
    for (int j = blockIdx.x * L_WARPS + warp; j < n_list; j += gridDim.x * L_WARPS) {
        const int64_t row = row_list[j];
        const uint32_t mask = row_mask[j];
        const XT* src = x + (row / T) * int64_t(D) * T + (row % T);
        // this lane's slice of the frame: depths 4 (lane + 32 q) .. + 3   (VEC)   or   lane + 32 i   (scalar)
        float xr[4 * MAXQ];
        float xx = 0.f;
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = VEC ? 4 * (lane + 32 * q) + u : lane + 32 * (4 * q + u);
                const float v = d < D ? x_to_float(__ldg(src + int64_t(d) * T)) : 0.f;
                xr[4 * q + u] = v;
                xx = fmaf(v, v, xx);
            }
        xx = warp_sum(xx);
        float bd = inf;
        int bi = 0x7fffffff;
        ListWalk walk(mask, row_tiles[j], tile_scale, n_code_tiles, own_shift);
        while (walk.next()) {
            const int nt0 = walk.nt, res0 = walk.res;
            const bool two = walk.next();
            const int nt1 = walk.nt, res1 = walk.res;
            const int c0 = nt0 * 128 + 16 * b_mine + res0, c1 = nt1 * 128 + 16 * b_mine + res1;
            const float e0 = c0 < K ? __ldg(ee + c0) : inf;
            const float e1 = (two && c1 < K) ? __ldg(ee + c1) : inf;
            float p0[8], p1[8];
            list_dot8<VEC, MAXQ>(k, D, K, nt0, res0, lane, xr, p0);
            if (two) list_dot8<VEC, MAXQ>(k, D, K, nt1, res1, lane, xr, p1);
            const float d0 = list_reduce8(p0, lane);
            if (c0 < K) argmin_take(bd, bi, ref_distance(xx, d0, e0), c0);
            if (two) {
                const float d1 = list_reduce8(p1, lane);
                if (c1 < K) argmin_take(bd, bi, ref_distance(xx, d1, e1), c1);
            }
        }
        // combine the lanes (lanes sharing a code hold identical values)
#pragma unroll
        for (int o = 16; o >= 4; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            argmin_take(bd, bi, od, oi);
        }
        if (lane == 0) {
            idx[row] = bi == 0x7fffffff ? 0 : bi;
            if (min_d) min_d[row] = bd;
            sum += double(bd);
        }
    }
    // the last block to finish re-arms the header for the next vq_assign on this workspace (no memset needed)
    sum = block_sum(sum, red);
    if (threadIdx.x == 0) {
        if (scalars) {
            if (sum != 0.0) atomicAdd(&scalars[VQ_S_SUM_MIN_D], sum);
            if (blockIdx.x == 0) atomicAdd(&scalars[VQ_S_UNSAFE_ROWS], double(n_list));
        }
        __threadfence();
        const unsigned ticket = atomicAdd(&hdr->list_ticket, 1u);
        is_last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        hdr->unsafe_count = 0;
        hdr->list_ticket = 0;
        if (hard_hint) {                       // mapped host memory: which kernel variant the host launches next (see hard_hint())
            *hard_hint = (int64_t(n_list) * 128 > N * T) ? 1u : 0u;
            __threadfence_system();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Grouped (phoneme-conditioned) assignment, models/vqtts/bottleneck.py:38-58: frame j only competes among the l_bins codes
// of its aligned token tok[j]; K = n_vocab * l_bins (76 288 in the reference's TTS config).  The reference gathers a
// [frames, l_bins, D] copy of the codebook (frames * l_bins * D * 4 bytes!) and runs a bmm; here one warp per frame walks
// its token's code rows in place (they stay in L2), FP32, lowest index on ties.  Writes the relative index (what the
// reference returns), the absolute index (what dequantize / update_k use, :58) and optionally the winning distance.
template <bool VEC, int MAXQ>
__global__ void __launch_bounds__(L_WARPS * 32)
assign_grouped_kernel(const float* __restrict__ x, int64_t N, int D, int64_t T, const float* __restrict__ k,
                      const float* __restrict__ ee, int n_vocab, int l_bins, const int64_t* __restrict__ tok,
                      int64_t* __restrict__ q_rel, int64_t* __restrict__ q_abs, float* __restrict__ min_d,
                      double* __restrict__ scalars) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b_mine = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    const float inf = __int_as_float(0x7f800000);
    const int64_t rows = N * T;
    double sum = 0.0;
    for (int64_t row = int64_t(blockIdx.x) * L_WARPS + warp; row < rows; row += int64_t(gridDim.x) * L_WARPS) {
        const int64_t g = min(max(tok[row], int64_t(0)), int64_t(n_vocab - 1));
        const float* src = x + (row / T) * int64_t(D) * T + (row % T);
        const float* kg = k + size_t(g) * l_bins * D;
        const float* eg = ee + size_t(g) * l_bins;
        float xr[4 * MAXQ];
        float xx = 0.f;
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = VEC ? 4 * (lane + 32 * q) + u : lane + 32 * (4 * q + u);
                const float v = d < D ? __ldg(src + int64_t(d) * T) : 0.f;
                xr[4 * q + u] = v;
                xx = fmaf(v, v, xx);
            }
        xx = warp_sum(xx);
        float bd = inf;
        int bi = 0x7fffffff;
        for (int c0 = 0; c0 < l_bins; c0 += 16) {                 // two segments of eight consecutive codes per step
            const int ca = c0 + b_mine, cb = c0 + 8 + b_mine;
            const float ea = ca < l_bins ? __ldg(eg + ca) : inf, eb = cb < l_bins ? __ldg(eg + cb) : inf;
            float pa[8], pb[8];
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const int c = c0 + b;
                float acc = 0.f;
                if (c < l_bins) {
                    const float* er = kg + size_t(c) * D;
                    if (VEC) {
#pragma unroll
                        for (int q = 0; q < MAXQ; ++q)
                            if (4 * (lane + 32 * q) < D) {
                                const float4 e4 = __ldg(reinterpret_cast<const float4*>(er) + lane + 32 * q);
                                acc = fmaf(xr[4 * q + 0], e4.x, acc); acc = fmaf(xr[4 * q + 1], e4.y, acc);
                                acc = fmaf(xr[4 * q + 2], e4.z, acc); acc = fmaf(xr[4 * q + 3], e4.w, acc);
                            }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4 * MAXQ; ++i)
                            if (lane + 32 * i < D) acc = fmaf(xr[i], __ldg(er + lane + 32 * i), acc);
                    }
                }
                if (b < 8) pa[b] = acc; else pb[b - 8] = acc;
            }
            const float da = list_reduce8(pa, lane), db = list_reduce8(pb, lane);
            if (ca < l_bins) argmin_take(bd, bi, ref_distance(xx, da, ea), ca);
            if (cb < l_bins) argmin_take(bd, bi, ref_distance(xx, db, eb), cb);
        }
#pragma unroll
        for (int o = 16; o >= 4; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            argmin_take(bd, bi, od, oi);
        }
        if (lane == 0) {
            const int rel = bi == 0x7fffffff ? 0 : bi;
            q_rel[row] = rel;
            q_abs[row] = g * l_bins + rel;
            if (min_d) min_d[row] = bd;
            sum += double(bd);
        }
    }
    sum = block_sum(sum, red);
    if (threadIdx.x == 0 && scalars && sum != 0.0) atomicAdd(&scalars[VQ_S_SUM_MIN_D], sum);
}

}  // namespace vq
