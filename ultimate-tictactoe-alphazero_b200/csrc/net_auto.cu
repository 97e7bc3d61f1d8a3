// net_auto.cu -- trunk_auto_kernel: ONE launch per evaluator round.  The queue length is only known on the device, so
// the engine used to enqueue both trunk kernels every round and the one whose range did not contain the batch size
// exited at once; an exit still costs a full kernel boundary (3.4 us under ncu, ~6 us between dependent launches).
// Here both bodies sit behind one device-side branch on *count: batches of up to one wave of 5-position groups run
// tc2::trunk_tc2_body<2> (cta_group::1 MMAs, next layer overlaps the epilogue), larger ones pp::trunk_pp_body<1>
// (two groups in flight, cta_group::2 MMAs).  Both use 19 warps, clusters of two CTAs and the same grid; shared memory is
// the larger of the two footprints.  Batches above 7 positions per pair still get a second launch (pp<2>).
#include "heads_fc.cuh"
#include "net_pp_kernel.cuh"
#include "net_tc2_kernel.cuh"

namespace uttt {

static_assert(tc2::Cfg<2>::THREADS == pp::THREADS, "one block size for both bodies");
constexpr int AUTO_SMEM = tc2::Cfg<2>::SMEM_BYTES > pp::Cfg<1>::SMEM_BYTES ? tc2::Cfg<2>::SMEM_BYTES : pp::Cfg<1>::SMEM_BYTES;
static_assert(HEADS_SMEM_BYTES <= AUTO_SMEM, "the fused heads reuse the trunk's shared memory");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pp::THREADS, 1)
trunk_auto_kernel(const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wq_in,
                  const __nv_bfloat16* __restrict__ wq_bias,                      // cta_group::1 packing (net_tc2)
                  const __nv_bfloat16* __restrict__ wq2, const __nv_bfloat16* __restrict__ wq2_in,
                  const __nv_bfloat16* __restrict__ wq2_bias,                     // per-CTA halves (net_pp)
                  const __nv_bfloat16* __restrict__ planes, const float* __restrict__ headw, float* headfeat, uint4* skip,
                  const int32_t* __restrict__ count, int small_cap, int max_count, long long* dbg,
                  HeadsFC fc, float* policy, float* value /* null: the heads' FC layers are a separate kernel */) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const bool stamp = dbg && blockIdx.x == 0 && threadIdx.x == 0;          // diagnostics: phases of CTA 0
    if (stamp) dbg[200] = clock64();
    const int n_pos = *count;
    const bool small = n_pos <= small_cap;
    if (small)
        tc2::trunk_tc2_body<2>(wq, wq_in, wq_bias, planes, headw, headfeat, skip, count, 0, small_cap, dbg);
    else
        pp::trunk_pp_body<1>(wq2, wq2_in, wq2_bias, planes, headw, headfeat, skip, count, small_cap, max_count, dbg);
    if (stamp) dbg[201] = clock64();
    if (policy == nullptr || n_pos > max_count) return;
    // Fused heads (the host passes policy / value only if no batch can exceed one group per pair): the pair's head
    // features were written to global memory by both CTAs' last epilogues and ordered by the cluster barrier that ends
    // the body (release / acquire); they are read back with ld.global.cg.  The two CTAs take alternate positions of
    // the pair (at most 4 each); the body's shared memory is free now.
    const int n_pairs = (int)gridDim.x >> 1, pair = (int)blockIdx.x >> 1;
    const int P = small ? tc2::group_positions<2>(n_pos, n_pairs) : pp::pair_positions<1>(n_pos, n_pairs);
    const int first = pair * P, last = min(n_pos, first + P);
    const int row0 = first + (int)tcx::cluster_rank();
    const int np = row0 < last ? (last - row0 + 1) >> 1 : 0;
    static_assert((pp::Cfg<1>::MAX_PA + pp::Cfg<1>::MAX_PB + 1) / 2 <= HEADS_P, "one heads call per CTA");
    if (np == 0) return;
    heads_fc_block(fc, headfeat, row0, 2, np, policy, value, 1, reinterpret_cast<float*>(smem), stamp ? dbg + 203 : nullptr);
    if (stamp) dbg[202] = clock64();
}

cudaError_t trunk_auto_init() {
    return cudaFuncSetAttribute(trunk_auto_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AUTO_SMEM);
}

// batches of 1 .. 7 positions per CTA pair in one launch; larger ones are left to launch_trunk_pp_large.
// policy / value non-null: the heads' FC layers run in the kernel's tail (only valid if max_rows <= trunk_pp_cap1)
cudaError_t launch_trunk_auto(const NetWeights& w, const __nv_bfloat16* planes, float* headfeat, const int32_t* count, int max_rows,
                              float* skip, int n_sm, cudaStream_t s, long long* dbg, float* policy, float* value) {
    int pairs = n_sm / 2;
    if (max_rows < pairs) pairs = max_rows < 1 ? 1 : max_rows;
    const int small_cap = (n_sm / 2) * tc2::Cfg<2>::MAX_P;
    const int cap1 = (n_sm / 2) * (pp::Cfg<1>::MAX_PA + pp::Cfg<1>::MAX_PB);
    if (policy && max_rows > cap1) return cudaErrorInvalidValue;
    trunk_auto_kernel<<<2 * pairs, pp::THREADS, AUTO_SMEM, s>>>(w.res_w_bf16, w.conv_in_w_bf16, w.bias_blk, w.res_w_2sm, w.conv_in_w_2sm,
                                                                w.bias_blk_2sm, planes, w.head_w, headfeat,
                                                                reinterpret_cast<uint4*>(skip), count, small_cap, cap1, dbg,
                                                                heads_fc_of(w), policy, value);
    return cudaGetLastError();
}

}  // namespace uttt
