// Codebook preparation for K1: ||e||^2 (FP32), the zero-padded FP16 image the tcgen05 kernel streams
// through shared memory, and the two norms that bound the FP16 shortlisting error.
// Replaces the `torch.sum(k_w**2, dim=0)` term of bottleneck.py:131-133 (computed once per call, not per row).
#pragma once
#include "vq_common.cuh"
#include <cuda_fp16.h>

namespace vq {

// Workspace header (device memory, first 256 bytes of the workspace).
struct AssignHeader {
    unsigned int e_norm_max_bits;   // max_c ||fp16(e_c)||            (float bits; non-negative => uint order == float order)
    unsigned int e_err_max_bits;    // max_c ||e_c - fp16(e_c)||
    int unsafe_count;               // rows queued for the exact fallback (reset by the fallback kernel's last block)
    unsigned int list_ticket;       // last-block ticket of the fallback kernel
    int pad[60];
};
static_assert(sizeof(AssignHeader) == 256, "header is 256 bytes");

struct AssignWorkspace {
    AssignHeader* hdr;
    float* ee;            // [Kp]  ||e||^2, +inf beyond K
    float* hn;            // [Kp]  ||e||^2 / 2, huge beyond K
    float* hn_off;        // [Kp]  ||e||^2 / 2 - B (key offset), written only for codebooks too large for shared memory
    __half* eb;           // [Kp][Dp] FP16 image, zero padded
    int* unsafe_rows;     // [N*T] frames the tcgen05 kernel could not prove safe
    uint32_t* unsafe_mask;           // [N*T] per listed frame: residue chains (column % 16) x scan group whose codes the exact re-scan visits
    uint2* unsafe_tiles;             // [N*T] ... and of those, which code tiles (24-bit map per scan group)
    int Kp, Dp;
    size_t bytes;
};

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

inline AssignWorkspace carve_workspace(void* base, int64_t rows, int K, int D) {
    AssignWorkspace w;
    w.Kp = round_up(K, 128);
    w.Dp = round_up(D, 64);
    char* p = static_cast<char*>(base);
    size_t off = 0;
    w.hdr = reinterpret_cast<AssignHeader*>(p + off);              off += 256;
    w.ee = reinterpret_cast<float*>(p + off);                      off += align256(size_t(w.Kp) * 4);
    w.hn = reinterpret_cast<float*>(p + off);                      off += align256(size_t(w.Kp) * 4);
    w.hn_off = reinterpret_cast<float*>(p + off);                  off += align256(size_t(w.Kp) * 4);
    w.eb = reinterpret_cast<__half*>(p + off);              off += align256(size_t(w.Kp) * w.Dp * 2);
    w.unsafe_rows = reinterpret_cast<int*>(p + off);               off += align256(size_t(rows) * 4);
    w.unsafe_mask = reinterpret_cast<uint32_t*>(p + off);          off += align256(size_t(rows) * 4);
    w.unsafe_tiles = reinterpret_cast<uint2*>(p + off);            off += align256(size_t(rows) * 8);
    w.bytes = off;
    return w;
}

// One warp per code row.
__global__ void __launch_bounds__(256) codebook_prepare_kernel(const float* __restrict__ k, int K, int D, int Kp, int Dp,
                                                              float* __restrict__ ee, float* __restrict__ hn,
                                                              __half* __restrict__ eb, AssignHeader* hdr) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= Kp) return;
    float s = 0.f, sb = 0.f, se = 0.f;
    for (int d = lane; d < Dp; d += 32) {
        float v = (warp < K && d < D) ? k[size_t(warp) * D + d] : 0.f;
        __half b = __float2half_rn(v);
        float vb = __half2float(b);
        if (eb) eb[size_t(warp) * Dp + d] = b;
        s = fmaf(v, v, s);
        sb = fmaf(vb, vb, sb);
        float r = v - vb;
        se = fmaf(r, r, se);
    }
    s = warp_sum(s); sb = warp_sum(sb); se = warp_sum(se);
    if (lane == 0) {
        if (warp < K) {
            ee[warp] = s;
            hn[warp] = 0.5f * s;
            atomicMax(&hdr->e_norm_max_bits, __float_as_uint(sqrtf(sb)));
            atomicMax(&hdr->e_err_max_bits, __float_as_uint(sqrtf(se)));
        } else {
            ee[warp] = __int_as_float(0x7f800000);   // +inf: never the minimum
            hn[warp] = 1.0e38f;                      // finite, so packed keys never become NaN
        }
    }
}

}  // namespace vq
