// Micro-benchmark: how fast can every SM stream the SAME weight array from L2 into shared memory with cp.async.bulk,
// (a) unicast per CTA, (b) CTA pairs where each rank loads half of a stage and multicasts it to both CTAs?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/l2stream tools/micro/l2stream.cu ; run: build/l2stream
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int STAGE_BYTES = 16384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t a, uint32_t r) {
    uint32_t o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
    return o;
}
__device__ __forceinline__ void arrive_remote(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory");
}

template <int STAGES, bool MCAST>
__global__ void __cluster_dims__(2, 1, 1) stream_kernel(const uint8_t* w, int n_stage_src, int n_iter, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t base = smem_u32(smem);
    uint32_t bar_full = base + STAGES * STAGE_BYTES, bar_empty = bar_full + 8 * STAGES;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, MCAST ? 2 : 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    long long t0 = clock64();
    if (threadIdx.x == 0) {                 // producer
        for (int i = 0; i < n_iter; i++) {
            int st = i % STAGES;
            uint32_t par = (i / STAGES) & 1;
            mbar_wait(bar_empty + 8 * st, par ^ 1);
            const uint8_t* src = w + (size_t)(i % n_stage_src) * STAGE_BYTES;
            mbar_expect_tx(bar_full + 8 * st, STAGE_BYTES);
            if (MCAST) {
                uint32_t half = STAGE_BYTES / 2;
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                        base + st * STAGE_BYTES + rank * half),
                    "l"(src + rank * half), "r"(half), "r"(bar_full + 8 * st), "h"((uint16_t)3)
                    : "memory");
            } else {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 base + st * STAGE_BYTES),
                             "l"(src), "r"((uint32_t)STAGE_BYTES), "r"(bar_full + 8 * st)
                             : "memory");
            }
        }
    } else if (threadIdx.x == 32) {         // consumer: frees the stage at once (in both CTAs when multicasting)
        for (int i = 0; i < n_iter; i++) {
            int st = i % STAGES;
            uint32_t par = (i / STAGES) & 1;
            mbar_wait(bar_full + 8 * st, par);
            if (MCAST) {
                arrive_remote(map_to_rank(bar_empty + 8 * st, 0));
                arrive_remote(map_to_rank(bar_empty + 8 * st, 1));
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_empty + 8 * st) : "memory");
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int STAGES, bool MCAST>
void run(const uint8_t* w, int n_src, int n_iter, long long* d_cycles, int grid) {
    int smem = STAGES * STAGE_BYTES + 16 * STAGES;
    cudaFuncSetAttribute(stream_kernel<STAGES, MCAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; rep++) stream_kernel<STAGES, MCAST><<<grid, 64, smem>>>(w, n_src, n_iter, d_cycles);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
    long long h[160];
    cudaMemcpy(h, d_cycles, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0, sum = 0;
    for (int i = 0; i < grid; i++) { mx = h[i] > mx ? h[i] : mx; sum += h[i]; }
    double bytes = (double)n_iter * STAGE_BYTES;
    printf("stages %d %s grid %3d: %.1f B/clk/SM (slowest CTA), %.1f mean\n", STAGES, MCAST ? "multicast pairs" : "unicast        ",
           grid, bytes / mx, bytes / ((double)sum / grid));
}

int main() {
    int n_src = 9 * 1024 * 1024 / STAGE_BYTES;      // the trunk's weight array: 9.4 MB, L2-resident
    uint8_t* w;
    cudaMalloc(&w, (size_t)n_src * STAGE_BYTES);
    cudaMemset(w, 1, (size_t)n_src * STAGE_BYTES);
    long long* d_cycles;
    cudaMalloc(&d_cycles, 160 * sizeof(long long));
    int n_iter = 4 * n_src;
    for (int grid : {148, 74, 2}) {
        run<4, false>(w, n_src, n_iter, d_cycles, grid);
        run<8, false>(w, n_src, n_iter, d_cycles, grid);
        run<4, true>(w, n_src, n_iter, d_cycles, grid);
        run<8, true>(w, n_src, n_iter, d_cycles, grid);
    }
    return 0;
}
