// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 operands from shared memory, K = 16) by shape and cta_group,
// issued back to back by one thread (or by two threads into two accumulators) on every SM of the chip at once.
// It answers three design questions of the trunk kernels:
//   * does a half tile (M = 64 per CTA: cta_group::1 M = 64, or cta_group::2 M = 128) take half the time of a full one?
//     ("third tile pair with cta_group::2 M = 128")
//   * does N = 64 (output channels split between the CTAs of a pair) take half the time of N = 128, and how fast can one
//     thread issue such MMAs?  ("N-split pair kernel for small batches")
//   * what does a second issuing thread buy?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_shapes tools/micro/mma_shapes.cu ; run: build/mma_shapes
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                     "h"((uint16_t)1) : "memory");
}

constexpr int ROWS = 256;                    // rows per operand panel in shared memory (enough for every shape)
constexpr int PANEL = ROWS * 16;

// CG = cta_group, M / N = MMA shape (M over the whole cta_group), NISSUE = issuing threads (one warp each, own accumulator)
template <int CG, int M, int N, int NISSUE>
__global__ void __cluster_dims__(2, 1, 1) mma_kernel(int n_mma, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = smem_u32(smem);
    const uint32_t sA = base, sB = base + 2 * PANEL, bar = base + 4 * PANEL;
    uint32_t* holder = reinterpret_cast<uint32_t*>(smem + 4 * PANEL + 64);
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 4 * PANEL / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NISSUE; i++) mbar_init(bar + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(512u) : "memory");
        else asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(512u) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *holder;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const bool issuer_cta = (CG == 1) || (rank == 0);
    if (issuer_cta && warp < NISSUE && (threadIdx.x & 31) == 0) {
        const uint64_t a = make_desc(sA, PANEL, 128), b = make_desc(sB, PANEL, 128);
        const uint32_t d = tmem + (uint32_t)(warp * 256);
        long long t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < n_mma; i++) umma<CG>(d, a + (uint64_t)((i & 7) * 8), b, idesc);     // (A start row varies like a tap shift)
        long long t1 = clock64();
        commit<CG>(bar + 8 * warp);
        mbar_wait(bar + 8 * warp, 0);
        long long t2 = clock64();
        out[(blockIdx.x * 2 + warp) * 2 + 0] = t1 - t0;
        out[(blockIdx.x * 2 + warp) * 2 + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int CG, int M, int N, int NISSUE>
void run(const char* what, int grid) {
    const int n_mma = 4096;
    long long* d;
    cudaMalloc(&d, 148 * 4 * sizeof(long long));
    cudaMemset(d, 0, 148 * 4 * sizeof(long long));
    auto k = mma_kernel<CG, M, N, NISSUE>;
    const int smem = 4 * PANEL + 128;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; rep++) k<<<grid, 64, smem>>>(n_mma, d);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%-46s FAILED: %s\n", what, cudaGetErrorString(err)); exit(1); }
    long long h[148 * 4];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double issue = 0, total = 0, worst = 0;
    int n = 0;
    for (int b = 0; b < grid; b++)
        for (int w = 0; w < NISSUE; w++) {
            long long ti = h[(b * 2 + w) * 2], tt = h[(b * 2 + w) * 2 + 1];
            if (!tt) continue;
            issue += (double)ti; total += (double)tt; n++;
            if ((double)tt > worst) worst = (double)tt;
        }
    // cycles per MMA *of one issuing thread*; with NISSUE threads the tensor pipe retires NISSUE MMAs in that time
    printf("%-46s grid %3d: issue %6.1f  retire %6.1f (slowest %6.1f) cycles per MMA and thread; %5.2f MMA-cycles per 128x128x16 of useful tile per SM\n",
           what, grid, issue / n / n_mma, total / n / n_mma, worst / n_mma,
           (total / n / n_mma) / NISSUE / ((double)M / CG / 128.0 * (double)N / 128.0));
    cudaFree(d);
}

int main() {
    for (int grid : {2, 148}) {
        run<1, 128, 128, 1>("cta_group::1 M=128 N=128, 1 thread", grid);
        run<1, 128, 128, 2>("cta_group::1 M=128 N=128, 2 threads", grid);
        run<1, 64, 128, 1>("cta_group::1 M=64  N=128, 1 thread", grid);
        run<1, 128, 64, 1>("cta_group::1 M=128 N=64,  1 thread", grid);
        run<1, 128, 64, 2>("cta_group::1 M=128 N=64,  2 threads", grid);
        run<1, 128, 32, 1>("cta_group::1 M=128 N=32,  1 thread", grid);
        run<1, 128, 256, 1>("cta_group::1 M=128 N=256, 1 thread", grid);
        run<1, 128, 208, 1>("cta_group::1 M=128 N=208, 1 thread", grid);
        run<1, 128, 160, 1>("cta_group::1 M=128 N=160, 1 thread", grid);
        run<1, 128, 112, 1>("cta_group::1 M=128 N=112, 1 thread", grid);
        run<1, 128, 112, 2>("cta_group::1 M=128 N=112, 2 threads", grid);
        run<1, 128, 208, 2>("cta_group::1 M=128 N=208, 2 threads", grid);
        run<2, 256, 128, 1>("cta_group::2 M=256 N=128, 1 thread", grid);
        run<2, 256, 128, 2>("cta_group::2 M=256 N=128, 2 threads", grid);
        run<2, 128, 128, 1>("cta_group::2 M=128 N=128, 1 thread", grid);
        run<2, 256, 64, 1>("cta_group::2 M=256 N=64,  1 thread", grid);
        run<2, 256, 256, 1>("cta_group::2 M=256 N=256, 1 thread", grid);
    }
    return 0;
}
