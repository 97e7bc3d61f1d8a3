// net_auto.cu -- trunk_auto_kernel: ONE launch per evaluator round.  The queue length is only known on the device, so
// the engine used to enqueue both trunk kernels every round and the one whose range did not contain the batch size
// exited at once; an exit still costs a full kernel boundary (3.4 us under ncu, ~6 us between dependent launches).
// Here the bodies sit behind one device-side branch on *count: batches of up to one wave of 5-position groups run
// tc2::trunk_tc2_body<2> (next layer overlaps the epilogue) -- with cta_group::2 MMAs while a CTA owns a single accumulator
// tile (up to 2 positions per pair: 128 -> 147 us per forward, the tile's cta_group::1 MMAs saturate shared memory and the
// weight stream slows them down; with two tiles per CTA the pair form measures the same as cta_group::1, which stays) --,
// larger ones pp::trunk_pp_body<1> (two groups in flight, cta_group::2 MMAs).  All use 19 warps, clusters of two CTAs and
// the same grid; shared memory is the largest footprint.  Batches above 7 positions per pair still get a second launch (pp<2>).
#include "heads_fc.cuh"
#include "net_pp_kernel.cuh"
#include "net_tc2_kernel.cuh"

namespace uttt {

static_assert(tc2::Cfg<2>::THREADS == pp::THREADS, "one block size for both bodies");
constexpr int max3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }
constexpr int AUTO_SMEM = max3(tc2::Cfg<2>::SMEM_BYTES, tc2::Cfg<2, false, true>::SMEM_BYTES, pp::Cfg<1>::SMEM_BYTES);
static_assert(HEADS_SMEM_BYTES <= AUTO_SMEM, "the fused heads reuse the trunk's shared memory");

// Slot mode.  With an atomic counter the leaves of a round land in the evaluator queue in arrival order, which differs
// from run to run; that is harmless as long as every row is computed by the same arithmetic, but the one-tile group of
// trunk_pp_body<1> accumulates in another order (split K), so WHICH leaves fall into it must not depend on arrival order
// if a self-play run is to be reproducible.  In slot mode a leaf stays in the slot of its tree (planes[slot], policy[slot],
// value[slot]) and flags[slot] says whether the slot holds one this round; row i of the batch is the i-th flagged slot.
// Every CTA derives the batch size and the slots of its own <= 7 positions from the same scan (warp 0, one L2 round trip).
constexpr int SLOT_CHUNKS = 17;                                 // 544 slots >= one group per CTA pair on 148 SMs (518)
__device__ __forceinline__ void scan_slots(const uint8_t* __restrict__ flags, int n_slots, int n_pairs, int pair, int small_cap,
                                           int* s_npos, int* s_src) {
    const int lane = threadIdx.x & 31;
    uint32_t mine = 0;                                          // bit j: slot lane + 32 j holds a leaf
#pragma unroll
    for (int j = 0; j < SLOT_CHUNKS; j++) {
        const int idx = lane + 32 * j;
        if (idx < n_slots && flags[idx]) mine |= 1u << j;
    }
    uint32_t bal[SLOT_CHUNKS];
    int n_pos = 0;
#pragma unroll
    for (int j = 0; j < SLOT_CHUNKS; j++) {
        bal[j] = __ballot_sync(0xFFFFFFFFu, (mine >> j) & 1u);
        n_pos += __popc(bal[j]);
    }
    const int P = n_pos <= small_cap ? tc2::group_positions<2>(n_pos, n_pairs) : pp::pair_positions<1>(n_pos, n_pairs);
    const int first = pair * P, last = min(n_pos, first + P);
    int run = 0;
#pragma unroll
    for (int j = 0; j < SLOT_CHUNKS; j++) {
        if ((bal[j] >> lane) & 1u) {
            const int rank = run + __popc(bal[j] & ((1u << lane) - 1u));
            if (rank >= first && rank < last) s_src[rank - first] = lane + 32 * j;
        }
        run += __popc(bal[j]);
    }
    if (lane == 0) *s_npos = n_pos;
}

constexpr int AUTO_SMEM_TOTAL = AUTO_SMEM + 64;                 // + the slot table of the pair and the batch size

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pp::THREADS, 1)
trunk_auto_kernel(const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wq_in,
                  const __nv_bfloat16* __restrict__ wq_bias,                      // cta_group::1 packing (net_tc2)
                  const __nv_bfloat16* __restrict__ wq2, const __nv_bfloat16* __restrict__ wq2_in,
                  const __nv_bfloat16* __restrict__ wq2_bias,                     // per-CTA halves, 18-block stages (net_pp<1>)
                  const __nv_bfloat16* __restrict__ planes, const float* __restrict__ headw, float* headfeat, uint4* skip,
                  const int32_t* __restrict__ count, int small_cap, int max_count, long long* dbg,
                  HeadsFC fc, float* policy, float* value /* null: the heads' FC layers are a separate kernel */,
                  const uint8_t* __restrict__ slot_flags, int n_slots /* slot mode (needs the fused heads), else null */,
                  const __nv_bfloat16* __restrict__ wq8, const __nv_bfloat16* __restrict__ wq8_in /* per-CTA halves in 8-block
                  stages (with wq2_bias): */, int pair_cap /* batches of up to pair_cap positions run cta_group::2 MMAs */) {
    extern __shared__ __align__(1024) uint8_t smem[];
    pdl_trigger();          // the next round's tree kernel may be scheduled (it waits for this grid's policy / value rows)
    pdl_wait();             // this round's tree kernel has finished: slot flags / queue length and planes are visible
    const bool stamp = dbg && blockIdx.x == 0 && threadIdx.x == 0;          // diagnostics: phases of CTA 0
    if (stamp) dbg[200] = clock64();
    const int n_pairs = (int)gridDim.x >> 1, pair = (int)blockIdx.x >> 1;
    int* s_src = reinterpret_cast<int*>(smem + AUTO_SMEM);                  // [8] slots of this pair's positions
    int* s_npos = s_src + 8;
    int n_pos;
    if (slot_flags) {
        if (threadIdx.x < 32) scan_slots(slot_flags, n_slots, n_pairs, pair, small_cap, s_npos, s_src);
        __syncthreads();
        n_pos = *s_npos;
    } else {
        n_pos = *count;
    }
    const int* src = slot_flags ? s_src : nullptr;
    const bool small = n_pos <= small_cap;
    if (n_pos <= pair_cap)
        tc2::trunk_tc2_body<2, false, true>(wq8, wq8_in, wq2_bias, planes, headw, headfeat, skip, count, 0, small_cap, dbg, n_pos, src);
    else if (small)
        tc2::trunk_tc2_body<2>(wq, wq_in, wq_bias, planes, headw, headfeat, skip, count, 0, small_cap, dbg, n_pos, src);
    else
        pp::trunk_pp_body<1>(wq2, wq2_in, wq2_bias, planes, headw, headfeat, skip, count, small_cap, max_count, dbg, n_pos, src);
    if (stamp) dbg[201] = clock64();
    if (policy == nullptr || n_pos > max_count) return;
    // Fused heads (the host passes policy / value only if no batch can exceed one group per pair): the pair's head
    // features were written to global memory by both CTAs' last epilogues and ordered by the cluster barrier that ends
    // the body (release / acquire); they are read back with ld.global.cg.  The two CTAs take alternate positions of
    // the pair (at most 4 each); the body's shared memory is free now.
    const int P = small ? tc2::group_positions<2>(n_pos, n_pairs) : pp::pair_positions<1>(n_pos, n_pairs);
    const int first = pair * P, last = min(n_pos, first + P);
    const int rank = (int)tcx::cluster_rank();
    const int row0 = first + rank;
    const int np = row0 < last ? (last - row0 + 1) >> 1 : 0;
    static_assert((pp::Cfg<1>::MAX_PA + pp::Cfg<1>::MAX_PB + 1) / 2 <= HEADS_P, "one heads call per CTA");
    if (np == 0) return;
    heads_fc_block(fc, headfeat, row0, 2, np, policy, value, 1, reinterpret_cast<float*>(smem), stamp ? dbg + 203 : nullptr,
                   src ? src + rank : nullptr);
    if (stamp) dbg[202] = clock64();
}

cudaError_t trunk_auto_init() {
    return cudaFuncSetAttribute(trunk_auto_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AUTO_SMEM_TOTAL);
}

// batches of 1 .. 7 positions per CTA pair in one launch; larger ones are left to launch_trunk_pp_large.
// policy / value non-null: the heads' FC layers run in the kernel's tail (only valid if max_rows <= trunk_pp_cap1).
// slot_flags non-null: slot mode (see scan_slots; needs the fused heads and n_slots <= 32 * SLOT_CHUNKS)
cudaError_t launch_trunk_auto(const NetWeights& w, const __nv_bfloat16* planes, float* headfeat, const int32_t* count, int max_rows,
                              float* skip, int n_sm, cudaStream_t s, long long* dbg, float* policy, float* value,
                              const uint8_t* slot_flags, int n_slots) {
    int pairs = n_sm / 2;
    if (max_rows < pairs) pairs = max_rows < 1 ? 1 : max_rows;
    const int small_cap = (n_sm / 2) * tc2::Cfg<2>::MAX_P;
    const int cap1 = (n_sm / 2) * (pp::Cfg<1>::MAX_PA + pp::Cfg<1>::MAX_PB);
    static const bool one_tile_pair = !(getenv("UTTT_TC2_PAIR") && atoi(getenv("UTTT_TC2_PAIR")) == 0);   // (0: comparison runs)
    if (policy && max_rows > cap1) return cudaErrorInvalidValue;
    if (slot_flags && (!policy || n_slots > 32 * SLOT_CHUNKS || n_slots > max_rows)) return cudaErrorInvalidValue;
    return launch_pdl(trunk_auto_kernel, dim3(2 * pairs), dim3(pp::THREADS), AUTO_SMEM_TOTAL, s, w.res_w_bf16, w.conv_in_w_bf16,
                      w.bias_blk, w.res_w_2sm18, w.conv_in_w_2sm18, w.bias_blk_2sm, planes, w.head_w, headfeat,
                      reinterpret_cast<uint4*>(skip), count, small_cap, cap1, dbg, heads_fc_of(w), policy, value, slot_flags, n_slots,
                      w.res_w_2sm, w.conv_in_w_2sm, one_tile_pair ? 2 * (n_sm / 2) : 0);
}

}  // namespace uttt
