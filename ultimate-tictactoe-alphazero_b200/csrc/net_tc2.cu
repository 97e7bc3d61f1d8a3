// net_tc2.cu -- __global__ wrappers and launchers of trunk_tc2_kernel (body: net_tc2_kernel.cuh)
#include "net_tc2_kernel.cuh"

namespace uttt {
namespace tc2 {

template <int LT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Cfg<LT>::THREADS, 1)
trunk_tc2_kernel(const __nv_bfloat16* __restrict__ wq,
                 const __nv_bfloat16* __restrict__ wq_in,
                 const __nv_bfloat16* __restrict__ wq_bias,
                 const __nv_bfloat16* __restrict__ planes,
                 const float* __restrict__ headw,
                 float* headfeat,
                 uint4* skip,
                 const int32_t* __restrict__ count,
                 int min_count, int max_count,
                 long long* dbg) {
    trunk_tc2_body<LT>(wq, wq_in, wq_bias, planes, headw, headfeat, skip, count, min_count, max_count, dbg);
}

}  // namespace tc2


cudaError_t trunk_tc2_init() {
    cudaError_t e = cudaFuncSetAttribute(tc2::trunk_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tc2::Cfg<2>::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc2::trunk_tc2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                tc2::Cfg<3>::SMEM_BYTES);
}

// largest batch the CTA-pair variants evaluate in one wave (7 positions per pair)
int trunk_tc2_capacity(int n_sm) { return (n_sm / 2) * tc2::Cfg<3>::MAX_P; }

int trunk_tc2_small_capacity(int n_sm) { return (n_sm / 2) * tc2::Cfg<2>::MAX_P; }

cudaError_t launch_trunk_tc2_small(const NetWeights& w, const __nv_bfloat16* planes, float* headfeat, const int32_t* count,
                                   int max_rows, float* skip, int n_sm, cudaStream_t s, long long* dbg) {
    int pairs = n_sm / 2;
    if (max_rows < pairs) pairs = max_rows < 1 ? 1 : max_rows;
    tc2::trunk_tc2_kernel<2><<<2 * pairs, tc2::Cfg<2>::THREADS, tc2::Cfg<2>::SMEM_BYTES, s>>>(
        w.res_w_bf16, w.conv_in_w_bf16, w.bias_blk, planes, w.head_w, headfeat, reinterpret_cast<uint4*>(skip), count, 0,
        trunk_tc2_small_capacity(n_sm), dbg);
    return cudaGetLastError();
}

// Two instantiations are enqueued; the queue length read on the device selects one:
//   batch <= 5 * pairs : 2 accumulator tiles per CTA (lowest latency, 8-stage weight ring)
//   batch <= 7 * pairs : 3 accumulator tiles per CTA
cudaError_t launch_trunk_tc2(const NetWeights& w, const __nv_bfloat16* planes, float* headfeat, const int32_t* count,
                             int max_rows, float* skip, int n_sm, cudaStream_t s, long long* dbg) {
    int pairs = n_sm / 2;
    if (max_rows < pairs) pairs = max_rows < 1 ? 1 : max_rows;
    const int cap2 = (n_sm / 2) * tc2::Cfg<2>::MAX_P, cap3 = (n_sm / 2) * tc2::Cfg<3>::MAX_P;
    tc2::trunk_tc2_kernel<2><<<2 * pairs, tc2::Cfg<2>::THREADS, tc2::Cfg<2>::SMEM_BYTES, s>>>(
        w.res_w_bf16, w.conv_in_w_bf16, w.bias_blk, planes, w.head_w, headfeat, reinterpret_cast<uint4*>(skip), count, 0,
        cap2, dbg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || max_rows <= cap2) return e;
    tc2::trunk_tc2_kernel<3><<<2 * pairs, tc2::Cfg<3>::THREADS, tc2::Cfg<3>::SMEM_BYTES, s>>>(
        w.res_w_bf16, w.conv_in_w_bf16, w.bias_blk, planes, w.head_w, headfeat, reinterpret_cast<uint4*>(skip), count, cap2,
        cap3, dbg);
    return cudaGetLastError();
}

}  // namespace uttt
