// Shared helpers for the vqb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/vqb200.h"

namespace vq {

// ------------------------------------------------------------------ error plumbing (thread-local)
inline char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int fail(const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
    snprintf(err_buf(), 512, fmt, a, b, c);
    return 1;
}
#define VQ_CUDA_OK(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            snprintf(vq::err_buf(), 512, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                     __FILE__, __LINE__);                                                    \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

#define VQ_REQUIRE(cond, msg)                                                                \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            snprintf(vq::err_buf(), 512, "vqb200: %s  [%s] (%s:%d)", msg, #cond, __FILE__, __LINE__); \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

inline int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember (kernel, device) pairs, not "done once per
// process", so a second GPU in the same process gets its opt-in too.
inline cudaError_t ensure_dynamic_smem_impl(const void* func, int bytes) {
    struct Entry { const void* func; unsigned long long devices; int bytes; };
    static Entry table[64];
    static int n_entries = 0;
    static std::mutex mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    Entry* hit = nullptr;
    for (int i = 0; i < n_entries; ++i)
        if (table[i].func == func && table[i].bytes == bytes) { hit = &table[i]; break; }
    if (hit && dev >= 0 && dev < 64 && ((hit->devices >> dev) & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    if (!hit && n_entries < 64) { table[n_entries] = Entry{func, 0ull, bytes}; hit = &table[n_entries++]; }
    if (hit && dev >= 0 && dev < 64) hit->devices |= 1ull << dev;
    return cudaSuccess;
}
template <class F>
inline cudaError_t ensure_dynamic_smem(F* func, int bytes) {
    return ensure_dynamic_smem_impl(reinterpret_cast<const void*>(func), bytes);
}

// latents come as FP32 or (bf16 autocast training / inference) as BF16; every kernel that reads them is templated on the type
__device__ __forceinline__ float x_to_float(float v) { return v; }
__device__ __forceinline__ float x_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of a double; result valid in thread 0.  `red` needs >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? red[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// The reference evaluates  (xx - 2*dot) + ee  left to right in FP32 (bottleneck.py:129-133);
// keep exactly those two roundings (2*dot is exact) and forbid FMA contraction.
__device__ __forceinline__ float ref_distance(float xx, float dot, float ee) {
    return __fadd_rn(__fsub_rn(xx, 2.0f * dot), ee);
}

// Lexicographic (distance, index) minimum: torch.min returns the lowest index on exact ties.
__device__ __forceinline__ void argmin_take(float& bd, int& bi, float d, int i) {
    if (d < bd || (d == bd && i < bi)) { bd = d; bi = i; }
}

__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v));
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// ---- Ampere-style asynchronous global -> shared copies (per-thread 4 / 8 / 16 bytes)
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(uint32_t(__cvta_generic_to_shared(dst))), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


}  // namespace vq
