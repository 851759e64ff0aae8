// K3: EMA codebook statistics and update.
// K3a (accumulate) replaces the dense one-hot build + [K,M]x[M,D] GEMM + row sums of bottleneck.py:64-68
//     with a scatter-add that never materialises the one-hot matrix:
//     algorithmic bytes per frame = 4D (x) + 8 (idx) + 4 (mask);  + 4K(D+1) per launch for the statistics.
// K3b (finalize) replaces bottleneck.py:78-89 (EMA lerp, usage threshold, dead-code revival, 4 metrics).
#pragma once
#include "vq_common.cuh"

namespace vq {

constexpr int E_TT = 128;        // frames per tile      (fallback kernel for emb_width > 128)
constexpr int E_DS = 32;         // depth slice staged at a time
constexpr int E_THREADS = 256;

// ---- sorted-run kernels.  Tiles of 64 frames x (up to 64 or 128) depths stream through a cp.async ring; one warp
// bitonic-sorts the NEXT tile's (code, frame) keys while the other 15 walk the current tile's rows in code order.  A run of
// equal codes is summed in registers (lane == depth) by the warp in whose row range it STARTS, so no two warps ever touch
// the same code in one tile:
//   SLAB = true  (K*64*4 B fits shared memory, e.g. K <= 512): the CTA owns a 64-deep depth slice and a private [K][64]
//                slab; a run is one non-atomic read-modify-write of the slab; the slab is flushed once per CTA.
//   SLAB = false (any K, D <= 128): a run is one coalesced FP32 reduction per 32-depth group straight to L2.
// Either way a row costs ~15 issued instructions (the first version spent ~80 per row and depth slice) and a hot code
// costs one update per tile instead of one per row.
constexpr int ER_TT = 64;                    // frames per tile
constexpr int ER_XS = ER_TT + 4;             // row stride (16-byte aligned rows)
#ifndef VQ_K3_WARPS
#define VQ_K3_WARPS 16
#endif
#ifndef VQ_K3_DW
#define VQ_K3_DW 64
#endif
#ifndef VQ_K3_STAGES
#define VQ_K3_STAGES 5
#endif
#ifndef VQ_K3_OCC
#define VQ_K3_OCC 1                          // CTAs per SM the launch aims for (build-time experiment knobs, see tools/experiment_k3.sh)
#endif
constexpr int ER_WARPS = VQ_K3_WARPS;        // the last warp sorts, the others accumulate
constexpr int ER_THREADS = ER_WARPS * 32;
constexpr int ER_PER = (ER_TT + ER_WARPS - 2) / (ER_WARPS - 1);   // sorted rows per accumulating warp (5)
constexpr int ER_LIST = 256;                 // valid tiles compacted per pass
constexpr int ER_DYN_SMEM_MAX = 227 * 1024 - 4096;   // the kernel also has ~2.1 KB of static shared memory (4 KB with the reserved 1 KB, rounded)

// flags[tile] = 1 when the 64-frame tile has at least one valid frame: padded tiles (35 % of an LJSpeech-like batch) are
// then never loaded.  One warp per tile.
__global__ void __launch_bounds__(256) ema_tile_flags_kernel(const float* __restrict__ mask, int64_t N, int64_t T, int tiles_per_utt,
                                                            unsigned char* __restrict__ flags) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    // the accumulate kernel may start zeroing its slab
    const int64_t tile = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tile >= N * tiles_per_utt) return;
    const int64_t n = tile / tiles_per_utt, t0 = (tile % tiles_per_utt) * ER_TT;
    bool any = false;
    for (int t = lane; t < ER_TT && t0 + t < T; t += 32) any |= mask[n * T + t0 + t] != 0.f;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flags[tile] = any ? 1 : 0;
}

template <bool SLAB> struct ErCfg;
template <> struct ErCfg<true>  { static constexpr int DW = VQ_K3_DW,  STAGES = VQ_K3_STAGES; };   // depth slice width, ring depth
template <> struct ErCfg<false> { static constexpr int DW = 128, STAGES = 6; };
template <bool SLAB> constexpr int er_stage_bytes() { return ErCfg<SLAB>::DW * ER_XS * 4 + ER_TT * 8 + ER_TT * 4; }
template <bool SLAB> inline size_t er_smem_bytes(int K) {
    return size_t(ErCfg<SLAB>::STAGES) * er_stage_bytes<SLAB>() + 3 * ER_TT * 4 + (SLAB ? (size_t(K) * ErCfg<true>::DW + K) * 4 : 0);
}

template <bool SLAB>
__global__ void __launch_bounds__(ER_THREADS, VQ_K3_OCC)
ema_accumulate_runs_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, const float* __restrict__ mask,
                           int64_t N, int D, int64_t T, int K, float* __restrict__ stats,
                           const unsigned char* __restrict__ tile_flags) {
    constexpr int DW = ErCfg<SLAB>::DW, STAGES = ErCfg<SLAB>::STAGES, NQ = DW / 32, STAGE_BYTES = er_stage_bytes<SLAB>();
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* stage0 = smem_raw;
    uint32_t* sorted = reinterpret_cast<uint32_t*>(smem_raw + size_t(STAGES) * STAGE_BYTES);   // [2][ER_TT] sorted keys (>= 0xFFFFFF00: no row) + [ER_TT] scratch
    float* slab = reinterpret_cast<float*>(sorted + 3 * ER_TT);                               // [K][DW]   (SLAB only)
    float* scnt = slab + size_t(K) * DW;                                                       // [K]       (SLAB, slice 0)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d0 = SLAB ? blockIdx.y * DW : 0;
    const int dn = min(DW, D - d0);
    const bool count_here = !SLAB || blockIdx.y == 0;
    const int tiles_per_utt = int((T + ER_TT - 1) / ER_TT);
    const int64_t n_tiles = N * tiles_per_utt;
    const bool vec16 = (T % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    float* sums = stats;
    float* counts = stats + size_t(K) * D;

    if (SLAB) {
        for (int i = tid; i < K * DW; i += ER_THREADS) slab[i] = 0.f;
        for (int i = tid; i < K; i += ER_THREADS) scnt[i] = 0.f;
    }

    // the loader's per-thread pattern is the same for every tile: precompute it
    constexpr int CHUNKS = DW * (ER_TT / 4) / ER_THREADS;           // 16-byte copies per thread per tile (2 or 4)
    int64_t g_off[CHUNKS];
    int s_off[CHUNKS], t_of[CHUNKS];
#pragma unroll
    for (int r = 0; r < CHUNKS; ++r) {
        const int i = tid + r * ER_THREADS, d = i >> 4, t = (i & 15) * 4;
        g_off[r] = int64_t(d0 + d) * T + t;
        s_off[r] = d * ER_XS + t;
        t_of[r] = d < dn ? t : ER_TT;                               // depths beyond D are never copied (nor read)
    }
    // launched as a programmatic dependent of ema_tile_flags_kernel: everything above overlapped it, the flags are read below
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // The CTA's tiles are first + k * step.  Only tiles with a valid frame enter the pipeline: warp 0 compacts them, up to
    // ER_LIST at a time, into (utterance, tile-in-utterance) lists in shared memory.  (Walking padded tiles through the
    // ring as empty copy groups left only ~26 KB of loads in flight per SM on a 35 %-padded batch: 1.5 TB/s.)
    __shared__ int s_ln[ER_LIST], s_lj[ER_LIST];
    __shared__ int s_nlist, s_scan;
    const int64_t first = blockIdx.x, step = gridDim.x;
    const int64_t n_mine = first < n_tiles ? (n_tiles - first + step - 1) / step : 0;      // tiles of this CTA
    if (tid == 0) s_scan = 0;
    auto build_list = [&]() {                       // warp 0 only; continues from candidate s_scan
        int count = 0;
        int64_t k = s_scan;
        while (count <= ER_LIST - 32 && k < n_mine) {
            const int64_t kk = k + lane;
            // LAST tiles first: in the training forward this kernel follows K1, which streamed x front to back, so the tail of x
            // is what the 126 MB L2 still holds (and K2, which follows, starts at the front -- what this kernel touched last)
            const int64_t tile = first + (n_mine - 1 - kk) * step;
            const bool ok = kk < n_mine && (tile_flags == nullptr || tile_flags[tile] != 0);
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int pos = count + __popc(m & ((1u << lane) - 1u));
                s_ln[pos] = int(tile / tiles_per_utt);
                s_lj[pos] = int(tile % tiles_per_utt);
            }
            count += __popc(m);
            k += 32;
        }
        if (lane == 0) { s_nlist = count; s_scan = int(min(k, n_mine)); }
    };

    // group g of the pipeline = x tile g  +  indices/mask of tile g + 1 (the sorter needs those one iteration early, and
    // keeping them out of their own tile's group lets STAGES - 2 tiles stay in flight instead of STAGES - 3)
    auto issue_im = [&](int g, int n_list) {                                       // indices + mask of list entry g
        if (g < n_list && tid < ER_TT) {
            float* Xs = reinterpret_cast<float*>(stage0 + size_t(g % STAGES) * STAGE_BYTES);
            int64_t* s_idx = reinterpret_cast<int64_t*>(Xs + DW * ER_XS);
            float* s_mask = reinterpret_cast<float*>(s_idx + ER_TT);
            const int64_t n = s_ln[g], t0 = int64_t(s_lj[g]) * ER_TT;
            if (t0 + tid < T) {
                cp_async8(s_idx + tid, idx + n * T + t0 + tid);
                if (mask) cp_async4(s_mask + tid, mask + n * T + t0 + tid);
                else s_mask[tid] = 1.f;
            } else {
                s_idx[tid] = -1;                   // frames beyond the utterance: no row (their x slots are never read)
                s_mask[tid] = 0.f;
            }
        }
    };
    auto issue = [&](int g, int n_list) {
        if (g < n_list) {
            float* Xs = reinterpret_cast<float*>(stage0 + size_t(g % STAGES) * STAGE_BYTES);
            const int64_t n = s_ln[g], t0 = int64_t(s_lj[g]) * ER_TT;
            const int tt = int(min(int64_t(ER_TT), T - t0));
            const float* src = x + n * int64_t(D) * T + t0;
            if (vec16) {
#pragma unroll
                for (int r = 0; r < CHUNKS; ++r)
                    if (t_of[r] < tt) cp_async16(Xs + s_off[r], src + g_off[r]);              // T % 4 == 0: never straddles tt
            } else {
                for (int i = tid; i < dn * ER_TT; i += ER_THREADS) {
                    const int d = i >> 6, t = i & (ER_TT - 1);
                    if (t < tt) cp_async4(Xs + d * ER_XS + t, src + int64_t(d0 + d) * T + t);
                }
            }
        }
        issue_im(g + 1, n_list);
        cp_async_commit();
    };

    // keys of a tile: (code << 8) | frame for valid rows, 0xFFFFFF00 | frame otherwise (all distinct); 64 keys = 2 per lane.
    // Sorted by RANK: every lane counts how many of the 64 keys are smaller than each of its two (16 broadcast 16-byte loads, 128
    // independent compare-adds) and stores its keys at those positions.  A bitonic network needs 21 dependent shuffle stages
    // (~1300 cycles of latency on this one warp, the longest chain of an iteration); the rank count has no chain at all.
    auto sort_tile = [&](bool exists, int st, uint32_t* out) {
        if (!exists) return;
        const float* Xs = reinterpret_cast<const float*>(stage0 + size_t(st) * STAGE_BYTES);
        const int64_t* s_idx = reinterpret_cast<const int64_t*>(Xs + DW * ER_XS);
        const float* s_mask = reinterpret_cast<const float*>(s_idx + ER_TT);
        uint32_t* tmp = sorted + 2 * ER_TT;
        uint32_t key[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int t = r * 32 + lane;
            const int64_t ci = s_idx[t];
            key[r] = (s_mask[t] != 0.f && ci >= 0) ? ((uint32_t(min(ci, int64_t(K - 1))) << 8) | uint32_t(t)) : (0xFFFFFF00u | uint32_t(t));
            tmp[t] = key[r];
        }
        __syncwarp();
#if defined(VQ_EXPERIMENT) && (VQ_EXPERIMENT & 2048)     /* timing experiment: no sorting (wrong results) */
        int rank0 = lane, rank1 = 32 + lane;
#else
        int rank0 = 0, rank1 = 0;
#pragma unroll
        for (int j4 = 0; j4 < ER_TT / 4; ++j4) {
            const uint4 q = reinterpret_cast<const uint4*>(tmp)[j4];
            rank0 += int(q.x < key[0]) + int(q.y < key[0]) + int(q.z < key[0]) + int(q.w < key[0]);
            rank1 += int(q.x < key[1]) + int(q.y < key[1]) + int(q.z < key[1]) + int(q.w < key[1]);
        }
#endif
        out[rank0] = key[0];
        out[rank1] = key[1];
        __syncwarp();                                  // (tmp is rewritten by this warp's next call)
    };

    for (;;) {
        __syncthreads();                               // previous list fully consumed (and s_scan initialised)
        if (warp == 0) build_list();
        __syncthreads();
        const int n_list = s_nlist;
        if (n_list == 0) break;
        issue_im(0, n_list);
        cp_async_commit();
#pragma unroll
        for (int g = 0; g < STAGES - 1; ++g) issue(g, n_list);
        cp_async_wait<STAGES - 1>();                   // indices / mask of tile 0 have landed
        __syncthreads();
        if (warp == ER_WARPS - 1) sort_tile(true, 0, sorted);

        for (int it = 0; it < n_list; ++it) {
            cp_async_wait<STAGES - 2>();               // x of tile it and indices / mask of tile it+1 have landed (this thread's copies)
            __syncthreads();                           // ... everyone's; sorted[it & 1] is complete; stage of tile it-1 is free
            issue(it + STAGES - 1, n_list);
            if (warp == ER_WARPS - 1) {
                sort_tile(it + 1 < n_list, (it + 1) % STAGES, sorted + ((it + 1) & 1) * ER_TT);
            } else {
                const float* Xs = reinterpret_cast<const float*>(stage0 + size_t(it % STAGES) * STAGE_BYTES);
                const uint32_t* keys = sorted + (it & 1) * ER_TT;
                const int lo = warp * ER_PER, hi = min(ER_TT, lo + ER_PER);
                int i = lo;
                const uint32_t kprev = (lo > 0 && lo < ER_TT) ? keys[lo - 1] : 0xFFFFFFFFu;
                // skip the tail of a run that started in an earlier warp's range
                while (i < hi && keys[i] < 0xFFFFFF00u && (keys[i] >> 8) == (kprev >> 8)) ++i;
                while (i < hi) {
                    uint32_t key = keys[i];
                    if (key >= 0xFFFFFF00u) break;                                // sorted: no more rows
                    const uint32_t code = key >> 8;
                    float a[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) a[q] = 0.f;
                    float cnt = 0.f;
                    do {                                                          // one run; may continue past hi
                        const float* col = Xs + (key & 255u);
#pragma unroll
                        for (int q = 0; q < NQ; ++q)
                            if (lane + 32 * q < dn) a[q] += col[(lane + 32 * q) * ER_XS];
                        cnt += 1.f;
                        ++i;
                        key = i < ER_TT ? keys[i] : 0xFFFFFFFFu;
                    } while (key < 0xFFFFFF00u && (key >> 8) == code);
#if defined(VQ_EXPERIMENT) && (VQ_EXPERIMENT & 256)      /* timing experiment: no updates */
                    if (a[0] + cnt == -12345.f) sums[0] = 1.f;
#else
                    if (SLAB) {
#pragma unroll
                        for (int q = 0; q < NQ; ++q) slab[size_t(code) * DW + lane + 32 * q] += a[q];
                        if (count_here && lane == 0) scnt[code] += cnt;
                    } else {
#pragma unroll
                        for (int q = 0; q < NQ; ++q)
                            if (lane + 32 * q < dn) atomicAdd(&sums[size_t(code) * D + lane + 32 * q], a[q]);
                        if (lane == 0) atomicAdd(&counts[code], cnt);
                    }
#endif
                }
            }
        }
        cp_async_wait<0>();                            // (only empty groups are left)
    }
    cp_async_wait<0>();
    if (SLAB) {
        __syncthreads();
        for (int i = tid; i < K * DW; i += ER_THREADS) {
            const int c = i / DW, d = i % DW;
            const float v = slab[i];
            if (d < dn && v != 0.f) atomicAdd(&sums[size_t(c) * D + d0 + d], v);
        }
        if (count_here)
            for (int c = tid; c < K; c += ER_THREADS)
                if (scnt[c] != 0.f) atomicAdd(&counts[c], scnt[c]);
    }
}

// (Tried and measured on B200, K=512, D=128, 256 utterances -- not kept: an "owner computes" variant without the sort
//  [TMA boxes + mbarrier ring, warp w owns code % 30 == w, one prep warp publishing per-owner frame masks]: 0.122 ms
//  against 0.115 ms for the kernel above.  With idle consumers the same ring streams the valid tiles in 0.068 ms, i.e.
//  2.8 TB/s: 64-frame tiles of a 64-deep slice are 256-byte pieces at a stride of T*4 bytes, which is what bounds both.)

// (Also tried and measured in round 2, not kept: two CTAs per SM on 32-deep slices [VQ_K3_DW=32 VQ_K3_STAGES=4 VQ_K3_OCC=2] so
//  that one CTA's per-tile barrier and sort hide behind the other's walk: 0.109 ms with 16 warps per CTA, 0.127 ms with 8,
//  against 0.096 ms -- twice the slices means every tile's keys are sorted and walked twice as often per byte.)

// (Also tried and measured, not kept: the whole shared memory as a six-stage TMA ring, two consumer warps per stage, groups
//  of equal codes found with match.any and flushed as coalesced 128-byte reductions straight to L2 instead of a slab:
//  0.139 ms -- 1.15 M vector reductions to 512 hot rows cost more than the shared-memory slab saves.)

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

// sum of the (already all-reduced) counts -> scalars[VQ_S_COUNT_TOTAL]; sum of the UPDATED cluster sizes
// mu k_elem + (1 - mu) counts -> scalars[VQ_S_ELEM_TOTAL] (the `n` of the optional Laplace smoothing)
__global__ void __launch_bounds__(256) ema_count_total_kernel(const float* __restrict__ counts, const float* __restrict__ k_elem,
                                                             float mu, float one_minus_mu, int K, double* scalars) {
    __shared__ double red[32];
    double s = 0.0, e = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < K; c += gridDim.x * blockDim.x) {
        s += double(counts[c]);
        e += double(__fadd_rn(__fmul_rn(mu, k_elem[c]), __fmul_rn(one_minus_mu, counts[c])));
    }
    s = block_sum(s, red);
    e = block_sum(e, red);
    if (threadIdx.x == 0) {
        if (s != 0.0) atomicAdd(&scalars[VQ_S_COUNT_TOTAL], s);
        if (e != 0.0) atomicAdd(&scalars[VQ_S_ELEM_TOTAL], e);
    }
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
    const double elem_total = scalars[VQ_S_ELEM_TOTAL];
    double ent = 0.0, used = 0.0, usage_n = 0.0, dk = 0.0;

    for (int c = blockIdx.x * warps_per_block + (threadIdx.x >> 5); c < K; c += gridDim.x * warps_per_block) {
        const float n_c = counts[c];
        // k_elem <- mu k_elem + (1 - mu) n_c                                   (bottleneck.py:80)
        float ke = __fadd_rn(__fmul_rn(mu, k_elem[c]), __fmul_rn(one_minus_mu, n_c));
        float ke_div = ke;
        if (laplace_eps > 0.f) {      // n = sum of the updated cluster sizes; (k_elem + eps) / (n + K eps) * n
            const float ntot = float(elem_total);
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
