// K1 (fast path): the distance contraction x.E^T on the 5th-generation tensor cores (tcgen05, sm_100a).
// Replaces bottleneck.py:92-100 (NCT flatten, fused away), :129-134 (distance + argmin) and the numerator of :140.
//
// One persistent CTA per SM, 16 warps, warp-specialised:
//   warp 0        TMA producer for x: FP32 [32 depth x 128 frames] boxes straight out of the NCT tensor
//   warp 1        MMA issuer (one lane): tcgen05.mma kind::f16, A (frames x depth, FP16) from TMEM, B (codes x depth,
//                 FP16, 128B-swizzled K-major) from shared memory, FP32 accumulators in TMEM (2 stages x 128 columns)
//   warp 2        TMA producer for the FP16 codebook image (resident in shared memory when it fits, else a ring)
//   warp 3        TMEM allocator
//   warps 4-7     front/back group: (front) FP32 smem tile -> FP16 pairs -> tcgen05.st into the TMEM A operand
//                 (thread == frame, so the NCT -> row-major transpose is free) while measuring ||x||^2 and the
//                 FP16 rounding residual ||x - fp16(x)||^2 per frame; (back) for the PREVIOUS tile: exact FP32
//                 re-scoring of the shortlisted code, the provable safety test, idx / min_d output
//   warps 8-15    scan groups: tcgen05.ld the accumulators (thread == frame; each group takes 64 of the 128 code
//                 columns), score s = x.e - ||e||^2/2, branch-free running (best, runner-up) with the code index packed
//                 into the low mantissa bits (1 LOP3 + 3 FMNMX per code)
//
// Exactness: FP16 operands only SHORTLIST.  A frame keeps the shortlisted code c1 iff its exact FP32 score beats the
// runner-up's approximate score by more than a rigorous bound on the FP16 error,
//     g1 > s2 + ||x - x16|| max||e16|| + ||x|| max||e - e16|| + slack,
// which proves c1 is the exact-arithmetic argmax; otherwise the frame goes to a worklist that the exact FP32 kernel
// (k1_assign_simt.cuh, LIST mode) re-scans.  So the output never depends on reduced-precision arithmetic.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "k1_prepare.cuh"
#include "vq_common.cuh"

namespace vq {
namespace tc {

constexpr int TM = 128;                      // frames per tile (UMMA M)
constexpr int TN = 128;                      // codes per accumulator tile (UMMA N)
constexpr int XCH = 32;                      // depth per x stage
constexpr int XS = 4;                        // x stages
constexpr int BKB = 64;                      // depth per codebook stage: 64 fp16 = 128 B = one swizzle row
constexpr int X_STAGE_BYTES = XCH * TM * 4;  // 16 KB
constexpr int B_STAGE_BYTES = TN * BKB * 2;  // 16 KB
constexpr int B_RING = 6;                    // streaming ring depth
constexpr int B_RESIDENT_MAX = 8;            // up to 8 stages (128 KB) stay resident
constexpr int THREADS = 512;
constexpr int ACC_COLS = 2 * TN;             // TMEM columns [0,256): two accumulator stages
constexpr uint32_t SPIN_LIMIT = 1u << 22;    // a lost barrier traps (after seconds) instead of hanging the GPU
constexpr uint32_t WAIT_HINT_NS = 2000;      // let the hardware park a waiting thread instead of spinning on issue slots

struct Params {
    const float* x; const float* k; const float* ee; const float* hn;
    AssignHeader* hdr; int* unsafe_rows;
    int64_t* idx; float* min_d; double* scalars; float* dbg;
    int N, D, Dp, K, Kp, T;
    int tiles_per_utt, n_tiles, n_nt, n_kb, n_xch, a_bufs, resident, b_stages, vec_k;
    uint32_t pack_mask;      // 0xFFFFFFC0, passed at run time so it lives in a register and the pack is ONE LOP3
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(WAIT_HINT_NS) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, FP16 operands, FP32 accumulate, M=128, N=128, K=16
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}

// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, rows of 64 fp16, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t b_desc_base(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr & 0x3FFFF) >> 4);           // start address   [0,14)
    d |= uint64_t(1) << 16;                              // leading byte offset (unused for swizzled K-major) [16,30)
    d |= uint64_t(1024 >> 4) << 32;                      // stride byte offset: 8 rows x 128 B        [32,46)
    d |= uint64_t(1) << 46;                              // descriptor version (Blackwell)
    d |= uint64_t(2) << 61;                              // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: FP32 accumulator, FP16 A and B, both K-major, N=128, M=128
constexpr uint32_t IDESC = (1u << 4) | (0u << 7) | (0u << 10) | (uint32_t(TN >> 3) << 17) | (uint32_t(TM >> 4) << 24);

struct Cand { float s1, s2; int c1; int pad; };          // per frame, per scan group: best, runner-up, best's code

struct Smem {           // control block placed after the data stages
    uint64_t x_full[XS], x_empty[XS];
    uint64_t b_full[B_RESIDENT_MAX], b_empty[B_RESIDENT_MAX];
    uint64_t a_full[2], a_empty[2], acc_full[2], acc_empty[2], cand_full[2], cand_empty[2];
    uint32_t tmem_base; uint32_t pad0;
    Cand cand[2][2][TM];
};

// RESCORE = true additionally re-scores the shortlisted code in exact FP32 (needed for min_d / sum(min_d) and for
// the audit output; it also halves the safety margin); RESCORE = false decides from the approximate scores alone.
template <bool RESCORE>
__global__ void __launch_bounds__(THREADS, 1)
assign_tc_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap b_map, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // round up to 1024 B with pointer arithmetic on the __shared__ array so every access stays an LDS/STS
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* xs_base = smem;                                        // XS x 16 KB
    uint8_t* bs_base = smem + XS * X_STAGE_BYTES;                   // b_stages x 16 KB (1024-aligned)
    Smem* ctl = reinterpret_cast<Smem*>(bs_base + size_t(p.b_stages) * B_STAGE_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < XS; ++i) { mbar_init(smem_u32(&ctl->x_full[i]), 1); mbar_init(smem_u32(&ctl->x_empty[i]), 128); }
        for (int i = 0; i < B_RESIDENT_MAX; ++i) { mbar_init(smem_u32(&ctl->b_full[i]), 1); mbar_init(smem_u32(&ctl->b_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&ctl->a_full[i]), 128);   mbar_init(smem_u32(&ctl->a_empty[i]), 1);
            mbar_init(smem_u32(&ctl->acc_full[i]), 1);   mbar_init(smem_u32(&ctl->acc_empty[i]), 256);
            mbar_init(smem_u32(&ctl->cand_full[i]), 256); mbar_init(smem_u32(&ctl->cand_empty[i]), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = ctl->tmem_base;
    const uint32_t a_col0 = ACC_COLS, a_stride = uint32_t(p.Dp) >> 1;   // A buffers: TMEM columns [256, 256 + a_bufs * Dp/2)

    const int first = blockIdx.x, step = gridDim.x;

    if (warp == 0) {
        // ============================================================ x producer
        if (lane == 0) {
            uint32_t q = 0;
            for (int tile = first; tile < p.n_tiles; tile += step) {
                const int n = tile / p.tiles_per_utt, t0 = (tile % p.tiles_per_utt) * TM;
                for (int ch = 0; ch < p.n_xch; ++ch, ++q) {
                    const uint32_t s = q % XS, ph = (q / XS) & 1;
                    mbar_wait(smem_u32(&ctl->x_empty[s]), ph ^ 1);
                    mbar_expect_tx(smem_u32(&ctl->x_full[s]), X_STAGE_BYTES);
                    tma_load_3d(smem_u32(xs_base + s * X_STAGE_BYTES), &x_map, smem_u32(&ctl->x_full[s]), t0, ch * XCH, n);
                }
            }
        }
    } else if (warp == 2) {
        // ============================================================ codebook producer
        if (lane == 0) {
            if (p.resident) {
                for (int nt = 0; nt < p.n_nt; ++nt)
                    for (int kb = 0; kb < p.n_kb; ++kb) {
                        const int s = nt * p.n_kb + kb;
                        mbar_expect_tx(smem_u32(&ctl->b_full[s]), B_STAGE_BYTES);
                        tma_load_2d(smem_u32(bs_base + size_t(s) * B_STAGE_BYTES), &b_map, smem_u32(&ctl->b_full[s]), kb * BKB, nt * TN);
                    }
            } else {
                uint32_t q = 0;
                for (int tile = first; tile < p.n_tiles; tile += step)
                    for (int nt = 0; nt < p.n_nt; ++nt)
                        for (int kb = 0; kb < p.n_kb; ++kb, ++q) {
                            const uint32_t s = q % p.b_stages, ph = (q / p.b_stages) & 1;
                            mbar_wait(smem_u32(&ctl->b_empty[s]), ph ^ 1);
                            mbar_expect_tx(smem_u32(&ctl->b_full[s]), B_STAGE_BYTES);
                            tma_load_2d(smem_u32(bs_base + size_t(s) * B_STAGE_BYTES), &b_map, smem_u32(&ctl->b_full[s]), kb * BKB, nt * TN);
                        }
            }
        }
    } else if (warp == 1) {
        // ============================================================ MMA issuer
        if (lane == 0) {
            uint32_t qa = 0, qb = 0, it = 0;
            for (int tile = first; tile < p.n_tiles; tile += step, ++it) {
                const uint32_t a = it % p.a_bufs, aph = (it / p.a_bufs) & 1;
                mbar_wait(smem_u32(&ctl->a_full[a]), aph);
                tc_fence_after();
                const uint32_t a_tmem = tmem + a_col0 + a * a_stride;
                for (int nt = 0; nt < p.n_nt; ++nt, ++qa) {
                    const uint32_t s = qa & 1, sph = (qa >> 1) & 1;
                    mbar_wait(smem_u32(&ctl->acc_empty[s]), sph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem + s * TN;
                    for (int kb = 0; kb < p.n_kb; ++kb) {
                        uint32_t bs;
                        if (p.resident) {
                            bs = nt * p.n_kb + kb;
                            if (it == 0) mbar_wait(smem_u32(&ctl->b_full[bs]), 0);
                        } else {
                            bs = qb % p.b_stages;
                            mbar_wait(smem_u32(&ctl->b_full[bs]), (qb / p.b_stages) & 1);
                        }
                        tc_fence_after();
                        const uint64_t bd = b_desc_base(smem_u32(bs_base + size_t(bs) * B_STAGE_BYTES));
#pragma unroll
                        for (int k4 = 0; k4 < BKB / 16; ++k4) {
                            // A: 16 fp16 along depth = 8 TMEM columns; B: +32 bytes inside the swizzle row
                            tc_mma_ts(d_tmem, a_tmem + uint32_t(kb * (BKB / 2) + k4 * 8), bd + uint64_t(k4 * 2), IDESC,
                                      (kb | k4) != 0 ? 1u : 0u);
                        }
                        if (!p.resident) { tc_commit(smem_u32(&ctl->b_empty[bs])); ++qb; }
                    }
                    tc_commit(smem_u32(&ctl->acc_full[s]));
                }
                tc_commit(smem_u32(&ctl->a_empty[a]));
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ============================================================ front/back group (thread == frame)
        const int wq = warp & 3, r = wq * 32 + lane;
        const uint32_t lane_base = uint32_t(wq * 32) << 16;
        const float e_norm_max = __uint_as_float(p.hdr->e_norm_max_bits);
        const float e_err_max = __uint_as_float(p.hdr->e_err_max_bits);
        uint32_t qx = 0, it = 0;
        float prev_xx = 0.f, prev_rr = 0.f;
        int prev_tile = -1;
        double sum_d = 0.0;

        auto rescore = [&](int tile, uint32_t itp, float xx, float rr) {
            const uint32_t cb = itp & 1, cph = (itp >> 1) & 1;
            mbar_wait(smem_u32(&ctl->cand_full[cb]), cph);
            const Cand ca = ctl->cand[cb][0][r], cc = ctl->cand[cb][1][r];
            mbar_arrive(smem_u32(&ctl->cand_empty[cb]));
            const int n = tile / p.tiles_per_utt, t = (tile % p.tiles_per_utt) * TM + r;
            if (t >= p.T) return;
            const bool a_wins = ca.s1 >= cc.s1;
            const int c1 = a_wins ? ca.c1 : cc.c1;
            const float s1 = fmaxf(ca.s1, cc.s1);
            // everything that is not c1 scored at most `bound` in FP16 arithmetic
            const float bound = fmaxf(fminf(ca.s1, cc.s1), fmaxf(ca.s2, cc.s2));
            const float xn = sqrtf(xx);
            const float acc_err = 1.2e-7f * float(p.Dp) * xn * e_norm_max;               // FP32 accumulation of D products (2^-23 D |x||e|)
            const float err = sqrtf(rr) * e_norm_max + xn * e_err_max + acc_err          // FP16 rounding of x and of E
                            + 1.6e-5f * (fabsf(bound) + fabsf(s1));                       // packed index bits, FP32 roundings of the scores
            const int64_t row = int64_t(n) * p.T + t;
            bool safe;
            float dot = 0.f;
            if (RESCORE) {
                const float* xr = p.x + (size_t(n) * p.D) * p.T + t;
                const float* er = p.k + size_t(c1) * p.D;
                int d = 0;
                if (p.vec_k) {
                    for (; d + 16 <= p.D; d += 16) {            // 16 independent loads in flight per thread
                        float xv[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) xv[u] = __ldg(xr + size_t(d + u) * p.T);
#pragma unroll
                        for (int u4 = 0; u4 < 4; ++u4) {
                            const float4 e4 = __ldg(reinterpret_cast<const float4*>(er + d + 4 * u4));
                            dot = fmaf(xv[4 * u4 + 0], e4.x, dot);
                            dot = fmaf(xv[4 * u4 + 1], e4.y, dot);
                            dot = fmaf(xv[4 * u4 + 2], e4.z, dot);
                            dot = fmaf(xv[4 * u4 + 3], e4.w, dot);
                        }
                    }
                }
                for (; d < p.D; ++d) dot = fmaf(__ldg(xr + size_t(d) * p.T), __ldg(er + d), dot);
                const float g1 = dot - p.hn[c1];
                // c1 is the exact argmax if its exact score clears every other code's approximate score + its error
                safe = g1 > bound + err + acc_err + 1.6e-5f * fabsf(g1);
                if (p.dbg) reinterpret_cast<float4*>(p.dbg)[row] = make_float4(s1, bound, g1, err);
            } else {
                // both approximate scores carry at most `err`
                safe = (s1 - bound) > 2.f * err;
            }
            if (safe) {
                p.idx[row] = c1;
                if (RESCORE) {
                    const float dist = ref_distance(xx, dot, p.ee[c1]);
                    if (p.min_d) p.min_d[row] = dist;
                    sum_d += double(dist);
                }
            } else {
                const int pos = atomicAdd(&p.hdr->unsafe_count, 1);
                p.unsafe_rows[pos] = int(row);
            }
        };

        for (int tile = first; tile < p.n_tiles; tile += step, ++it) {
            const uint32_t a = it % p.a_bufs, aph = (it / p.a_bufs) & 1;
            mbar_wait(smem_u32(&ctl->a_empty[a]), aph ^ 1);
            tc_fence_after();
            const uint32_t a_tmem = tmem + lane_base + a_col0 + a * a_stride;
            float xx = 0.f, rr = 0.f;
            for (int ch = 0; ch < p.n_xch; ++ch, ++qx) {
                const uint32_t s = qx % XS, ph = (qx / XS) & 1;
                mbar_wait(smem_u32(&ctl->x_full[s]), ph);
                const float* xs = reinterpret_cast<const float*>(xs_base + s * X_STAGE_BYTES) + r;
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float v0 = xs[(2 * j) * TM], v1 = xs[(2 * j + 1) * TM];
                    const __half2 h = __floats2half2_rn(v0, v1);              // low half = even depth, high half = odd depth
                    const float2 f = __half22float2(h);
                    const float r0 = v0 - f.x, r1 = v1 - f.y;
                    xx = fmaf(v0, v0, xx); xx = fmaf(v1, v1, xx);
                    rr = fmaf(r0, r0, rr); rr = fmaf(r1, r1, rr);
                    pk[j] = *reinterpret_cast<const uint32_t*>(&h);
                }
                mbar_arrive(smem_u32(&ctl->x_empty[s]));                      // the stage's data now lives in registers
                tc_st16(a_tmem + uint32_t(ch * (XCH / 2)), pk);
            }
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(smem_u32(&ctl->a_full[a]));
            if (prev_tile >= 0) rescore(prev_tile, it - 1, prev_xx, prev_rr);
            prev_tile = tile; prev_xx = xx; prev_rr = rr;
        }
        if (prev_tile >= 0) rescore(prev_tile, it - 1, prev_xx, prev_rr);
        sum_d = warp_sum(sum_d);
        if (lane == 0 && p.scalars && sum_d != 0.0) atomicAdd(&p.scalars[VQ_S_SUM_MIN_D], sum_d);
    } else if (warp >= 8) {
        // ============================================================ scan groups (thread == frame)
        const int wq = warp & 3, r = wq * 32 + lane, wg = (warp - 8) >> 2;
        const uint32_t lane_base = uint32_t(wq * 32) << 16;
        const float NEG_INF = __int_as_float(0xff800000);
        const uint32_t pack_mask = p.pack_mask;
        uint32_t qa = 0, it = 0;
        for (int tile = first; tile < p.n_tiles; tile += step, ++it) {
            float r1 = NEG_INF, r2 = NEG_INF;
            int rc1 = 0;
            for (int nt = 0; nt < p.n_nt; ++nt, ++qa) {
                const uint32_t s = qa & 1, sph = (qa >> 1) & 1;
                const int cbase = nt * TN + wg * 64;
                mbar_wait(smem_u32(&ctl->acc_full[s]), sph);
                tc_fence_after();
                uint32_t v0[32], v1[32];
                const uint32_t taddr = tmem + lane_base + s * TN + wg * 64;
                tc_ld32(taddr, v0);
                tc_ld32(taddr + 32, v1);
                tc_wait_ld();
                tc_fence_before();
                mbar_arrive(smem_u32(&ctl->acc_empty[s]));                    // accumulators are in registers: free the stage
                float t1 = NEG_INF, t2 = NEG_INF;
                const float4* hn4 = reinterpret_cast<const float4*>(p.hn + cbase);
#pragma unroll
                for (int j4 = 0; j4 < 16; ++j4) {
                    const float4 h = __ldg(hn4 + j4);
                    const float hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = j4 * 4 + jj;
                        const float acc = __uint_as_float(j < 32 ? v0[j & 31] : v1[j & 31]);
                        const float sc = acc - hh[jj];
                        uint32_t pkb;                                   // (score & ~63) | j : the code's column rides in the low mantissa bits
                        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(pkb) : "r"(__float_as_uint(sc)), "r"(pack_mask), "r"(uint32_t(j)));
                        const float pk = __uint_as_float(pkb);
                        const float lo = fminf(t1, pk);
                        t1 = fmaxf(t1, pk);
                        t2 = fmaxf(t2, lo);
                    }
                }
                // fold the tile-local pair into the running pair
                if (t1 > r1) {
                    r2 = fmaxf(r1, t2);
                    r1 = t1;
                    rc1 = cbase + int(__float_as_uint(t1) & 63u);
                } else {
                    r2 = fmaxf(r2, t1);
                }
            }
            const uint32_t cb = it & 1, cph = (it >> 1) & 1;
            mbar_wait(smem_u32(&ctl->cand_empty[cb]), cph ^ 1);
            Cand c; c.s1 = r1; c.s2 = r2; c.c1 = rc1; c.pad = 0;
            ctl->cand[cb][wg][r] = c;
            mbar_arrive(smem_u32(&ctl->cand_full[cb]));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

}  // namespace tc

// nullptr when the tcgen05 kernel takes this problem, else the reason it does not.
inline const char* tc_unsupported_reason(const float* x, int64_t N, int D, int64_t T, int K) {
    if (D > 512) return "emb_width > 512 (the FP16 A operand must fit 256 TMEM columns)";
    if (T % 4 != 0) return "T is not a multiple of 4 (TMA needs 16-byte global strides)";
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return "x is not 16-byte aligned";
    if (T >= (int64_t(1) << 31) || N >= (int64_t(1) << 31)) return "dimension too large for a tensor map";
    if (K > (1 << 24)) return "codebook too large";
    if (!tc::encode_tiled_fn()) return "cuTensorMapEncodeTiled is unavailable";
    return nullptr;
}

inline int launch_assign_tc(const float* x, int64_t N, int D, int64_t T, const float* k, int K, int64_t* idx, float* min_d,
                            double* scalars, const AssignWorkspace& w, cudaStream_t stream, float* dbg = nullptr) {
    using namespace tc;
    EncodeTiledFn encode = encode_tiled_fn();
    VQ_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is unavailable");
    Params p;
    p.x = x; p.k = k; p.ee = w.ee; p.hn = w.hn; p.hdr = w.hdr; p.unsafe_rows = w.unsafe_rows;
    p.idx = idx; p.min_d = min_d; p.scalars = scalars; p.dbg = dbg;
    p.N = int(N); p.D = D; p.Dp = w.Dp; p.K = K; p.Kp = w.Kp; p.T = int(T);
    p.tiles_per_utt = int((T + TM - 1) / TM);
    const int64_t n_tiles = N * p.tiles_per_utt;
    VQ_REQUIRE(n_tiles < (int64_t(1) << 31), "too many tiles");
    p.n_tiles = int(n_tiles);
    p.n_nt = w.Kp / TN; p.n_kb = w.Dp / BKB; p.n_xch = w.Dp / XCH;
    p.a_bufs = w.Dp <= 256 ? 2 : 1;
    p.resident = (p.n_nt * p.n_kb <= B_RESIDENT_MAX) ? 1 : 0;
    p.b_stages = p.resident ? p.n_nt * p.n_kb : B_RING;
    p.pack_mask = 0xFFFFFFC0u;
    p.vec_k = (D % 4 == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0) ? 1 : 0;

    CUtensorMap x_map, b_map;
    {
        cuuint64_t dims[3] = {cuuint64_t(T), cuuint64_t(D), cuuint64_t(N)};
        cuuint64_t strides[2] = {cuuint64_t(T) * 4, cuuint64_t(T) * cuuint64_t(D) * 4};
        cuuint32_t box[3] = {TM, XCH, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&x_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(x) failed with CUresult %s%lld", "", (long long)r);
    }
    {
        cuuint64_t dims[2] = {cuuint64_t(w.Dp), cuuint64_t(w.Kp)};
        cuuint64_t strides[1] = {cuuint64_t(w.Dp) * 2};
        cuuint32_t box[2] = {BKB, TN};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&b_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w.eb, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled(codebook) failed with CUresult %s%lld", "", (long long)r);
    }
    const size_t smem = 1024 + size_t(XS) * X_STAGE_BYTES + size_t(p.b_stages) * B_STAGE_BYTES + sizeof(Smem);
    VQ_REQUIRE(smem <= 227 * 1024, "shared memory budget exceeded");
    static bool configured = false;
    if (!configured) {
        VQ_CUDA_OK(cudaFuncSetAttribute(assign_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        VQ_CUDA_OK(cudaFuncSetAttribute(assign_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    const int grid = int(std::min<int64_t>(n_tiles, num_sms()));
    if (min_d || scalars || dbg) assign_tc_kernel<true><<<grid, THREADS, smem, stream>>>(x_map, b_map, p);
    else assign_tc_kernel<false><<<grid, THREADS, smem, stream>>>(x_map, b_map, p);
    VQ_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace vq
