// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, cta_group::1) on one SM, for several N, A sources and
// accumulator patterns.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t desc(uint32_t a) {
    return uint64_t((a & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(64) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 = TS same D, 1 = TS alternating D (2 tiles), 2 = SS same D, 3 = TS same D with 8 warps hammering tcgen05.ld
__global__ void __launch_bounds__(384, 1) bench(int n_cols, int mode, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stop = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = (1u << 4) | (uint32_t(n_cols >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    if (warp == 11) {
        const bool leader = elect_one();
        const uint64_t bd = desc(smem_u32(smem)), ad = desc(smem_u32(smem + 32768));
        long long t0 = clock64();
        if (mode == 4) {
            // the kernel's pattern: a batch of 9 MMAs, commit, wait for completion, next batch
            uint32_t ph = 0;
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if (leader) mma_ts(tmem, tmem + 256 + uint32_t((k & 7) * 8), bd + uint64_t((k & 3) * 2) + uint64_t(((k >> 2) & 1) * 1024), idesc, k != 0);
                if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(ph) : "memory");
                ph ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            long long t1 = clock64();
            if (leader && blockIdx.x == 0) { out[0] = (t1 - t0) * 8 / 9; out[1] = 0; }
            stop = 1;
        } else {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (leader) {
                    const uint32_t d = tmem + ((mode == 1) ? uint32_t((r & 1) * 256) : 0u);
                    const uint64_t b = bd + uint64_t((k & 3) * 2) + uint64_t((k >> 2) * 1024);
                    if (mode == 2) mma_ss(d, ad + uint64_t((k & 3) * 2), b, idesc, k != 0);
                    else mma_ts(d, tmem + 256 + uint32_t(k * 8), b, idesc, k != 0);
                }
            }
        }
        if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        long long t_issue = clock64();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t1 = clock64();
        if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue - t0; }
        stop = 1;
        }
    } else if (mode == 3 && warp < 8) {
        // accumulator readers: tcgen05.ld 32x32b.x32 in a loop
        uint32_t r[32];
        uint32_t sink = 0;
        const uint32_t taddr = tmem + (uint32_t((warp & 3) * 32) << 16) + uint32_t((warp >> 2) * 64);
        while (!stop) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                           "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                           "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            sink += r[0] ^ r[31];
        }
        if (sink == 0x12345 && blockIdx.x == 0) out[7] = sink;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 200;
    const char* names[] = {"TS same D", "TS alternating D", "SS same D", "TS same D + 8 warps of tcgen05.ld", "TS batches of 9 + commit + wait"};
    for (int grid : {148})
    for (int mode = 0; mode < 5; ++mode)
        for (int n : {64, 128, 256}) {
            if (mode == 1 && n == 256) continue;     // two 256-column tiles + A would not fit
            for (int trial = 0; trial < 2; ++trial) {
                out[0] = out[1] = 0;
                bench<<<grid, 384, 66 * 1024>>>(n, mode, reps, out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            printf("grid %3d %-36s N=%3d: %7.1f cycles/MMA to completion, %7.1f to issue  (ideal %d)\n", grid, names[mode], n,
                   double(out[0]) / (reps * 8), double(out[1]) / (reps * 8), n / 2);
        }
    return 0;
}
