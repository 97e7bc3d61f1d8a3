// net_tc.cu -- the residual trunk of DualNetwork (dual_network.py:28-45,97-99: 16 blocks x 2
// 3x3 convolutions, 128 -> 128 channels, BatchNorm + ReLU + skip) as ONE persistent sm_100a kernel:
// tcgen05.mma implicit GEMM with the activations resident in shared memory for all 32 layers.
//
// Mapping (per CTA, one CTA per SM, persistent over groups of 5 positions):
//   * GEMM rows: each position occupies 100 rows of a padded row space, index = 10*r + c
//     (r,c in 0..8; column 9 and the 10-row gap after row 8 are zero padding), so the 3x3 tap
//     (dy,dx) is a constant row shift of 10*dy + dx and the zero rows supply the halo.
//     5 positions = 500 rows = 4 MMA tiles of M = 128 (81 % of the issued rows are real cells).
//   * A operand (activations, bf16): shared memory, canonical K-major NO-swizzle layout stored as
//     16 channel panels [panel = ci/8][row][8 ci]: a core matrix (8 rows x 16 B) is 128 contiguous
//     bytes for ANY starting row, so a tap shift is just a different descriptor start address
//     (SBO = 128 B between 8-row groups, LBO = panel stride between the two K halves).
//   * B operand (weights, bf16, BN scale folded): pre-packed in HBM in exactly the shared-memory
//     image [layer][tap][ci/8][co][ci%8]; streamed by one elected thread with bulk async copies
//     (cp.async.bulk -> UBLKCP, mbarrier complete_tx) through a 5-stage ring of 16 KiB stages (4 K-blocks).
//     Each stage feeds 16 MMAs (4 tiles x 4 K-steps), i.e. weights are re-used across the 4 tiles.
//   * D accumulators: 4 x (128 lanes x 128 fp32 columns) = all 512 TMEM columns.
//   * Epilogue (16 warps, one TMEM lane quarter of one tile each): tcgen05.ld -> +shift (+skip)
//     -> ReLU -> zero the padding rows -> bf16 -> back into the SAME shared-memory buffer in place
//     (all MMAs of the layer have completed).  The skip connection (fp16 panels) lives in HBM/L2, written
//     and re-read by the same thread; only conv_input's output and the last block's output touch
//     HBM otherwise.
//
// Warp roles: warps 0-15 epilogue, warp 16 weight producer, warps 17-20 MMA issuers (one elected lane each,
// one accumulator tile each: a single issuing thread cannot keep the tensor pipe busy with K=16 MMAs of
// 64 cycles because every issue costs ~80 cycles of descriptor/uniform-register traffic); warp 17 owns TMEM.
#include "tc_common.cuh"

namespace uttt {

constexpr int TC_P = 5;                       // positions per group
constexpr int TC_POS_ROWS = 100;              // padded rows per position
constexpr int TC_TILES = 4;
constexpr int TC_M = 128 * TC_TILES;          // 512 GEMM rows per group
constexpr int TC_LEAD = 11;                   // zero rows before row 0 (largest negative shift)
constexpr int TC_AROWS = 536;                 // >= TC_LEAD + 512 + 11
constexpr int TC_PANEL_BYTES = TC_AROWS * 16; // one 8-channel panel
constexpr int TC_A_BYTES = 18 * TC_PANEL_BYTES;   // 16 channel panels + the constant panel pair of the bias MMA
constexpr int TC_STAGE_BYTES = 16384;         // half a tap: 64 ci x 128 co bf16
constexpr int TC_STAGES = 4;
constexpr int TC_STAGES_PER_LAYER = 18;
constexpr int TC_IN_STAGES = 3;                // conv_input: 9 taps x (K=16: 3 real channels) in 3 stages of 4 taps
constexpr int TC_BIAS_BYTES = 4096;             // [2 panels][128 co][8]: BN shift as bf16 hi + lo in k = 0, 1
constexpr int TC_GROUP_STAGES = (TC_IN_STAGES + 1) + NET_LAYERS * (TC_STAGES_PER_LAYER + 1);   // each layer starts with its bias block
constexpr int TC_GROUP_LAYERS = NET_LAYERS + 1;  // conv_input runs as layer -1 through the same pipeline
constexpr int TC_BAR_OFF = TC_A_BYTES + TC_STAGES * TC_STAGE_BYTES;
constexpr int TC_SMEM_BYTES = TC_BAR_OFF + 256;
constexpr int TC_ISSUERS = 4;                  // one MMA-issuing warp per accumulator tile
constexpr int TC_THREADS = (17 + TC_ISSUERS) * 32;
constexpr int TC_EPI_WARPS = 16;

// instruction descriptor (kind::f16): D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major A and B, N=128, M=128
constexpr uint32_t TC_IDESC = tcx::IDESC_M128_N128_BF16;

using namespace tcx;

__global__ void __launch_bounds__(TC_THREADS, 1)
trunk_tc_kernel(const __nv_bfloat16* __restrict__ wq,   // [32][72 K-blocks][2][128][8] bf16
                const __nv_bfloat16* __restrict__ wq_in,// conv_input: [12 taps (9 used)][2][128][8] bf16
                const __nv_bfloat16* __restrict__ wq_bias,   // [33][2][128][8] bf16: per layer the BN shift as a K=16 B block
                const __nv_bfloat16* __restrict__ planes,   // network input [rows][3][81] bf16
                const float* __restrict__ headw,        // [3][128] head 1x1 convs (BN scale folded) + [384..386] shifts
                float* headfeat,                        // out: [rows][243] = relu(policy conv)[2][81], relu(value conv)[81]
                float* resid,                           // [gridDim][16 panels][512 rows][8] fp16 skip connection (L2-resident)
                const int32_t* __restrict__ count,
                int min_count,                          // batches up to this size are handled by trunk_tc2_kernel
                long long* dbg) {                       // optional [32][4] clock64 timeline of CTA 0 (diagnostics)
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_pos = *count;
    if (n_pos <= min_count) return;
    if (dbg && blockIdx.x == 0 && threadIdx.x == 0)          // diagnostics: histogram of evaluator batch sizes (16 per bucket)
        atomicAdd(reinterpret_cast<unsigned long long*>(dbg) + 128 + min(n_pos >> 4, 63), 1ull);
    // positions per group: as few as keeps every group in the first wave (latency matters when the batch is
    // small), at most 5; 4 positions would need the same 4 tiles as 5, so it is never chosen.
    int P = (n_pos + (int)gridDim.x - 1) / (int)gridDim.x;
    P = P < 1 ? 1 : (P >= 4 ? TC_P : P);
    const int tiles = (P * TC_POS_ROWS + 127) / 128;
    const int n_groups = (n_pos + P - 1) / P;
    if ((int)blockIdx.x >= n_groups) return;

    uint8_t* sA = smem;
    const uint32_t sA_u = smem_u32(sA);
    const uint32_t sB_u = sA_u + TC_A_BYTES;
    const uint32_t bar_u = sA_u + TC_BAR_OFF;
    // barriers: full[5] @0, empty[5] @40, accum_full[4] @80, act_ready[4] @112; tmem base holder @144
    // Per-tile barriers let the four accumulator tiles drift apart: while one tile is in its epilogue the
    // other tiles' MMAs keep the tensor pipe busy.  A tile only synchronises with its row neighbours, because
    // its MMAs read 11 halo rows of the tile before and after it:
    //   act_ready[t]  <- epilogue warps (t,0..3), (t-1,3), (t+1,0)      (the rows tile t's MMAs read are written)
    //   accum_full[t] <- issuer t's tcgen05.commit                        (tile t's MMAs of this layer retired)
    //   epilogue warp (t,q) may overwrite its rows once accum_full[t], and accum_full[t-1] (q==0) or
    //   accum_full[t+1] (q==3), have completed: nobody reads the old values any more.
    const uint32_t bar_full = bar_u, bar_empty = bar_u + 8 * TC_STAGES, bar_accum = bar_u + 16 * TC_STAGES,
                   bar_act = bar_u + 16 * TC_STAGES + 32;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + TC_BAR_OFF + 16 * TC_STAGES + 64);

    if (threadIdx.x == 0) {
        for (int i = 0; i < TC_STAGES; i++) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, tiles); }
        for (int t = 0; t < TC_TILES; t++) {
            mbar_init(bar_accum + 8 * t, 1);
            mbar_init(bar_act + 8 * t, 4 + (t > 0 ? 1 : 0) + (t < tiles - 1 ? 1 : 0));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 17) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the whole activation buffer once: lead/tail margins and padding rows stay zero forever
    for (int i = threadIdx.x; i < TC_A_BYTES / 16; i += TC_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    // constant panel 16, every row = (1, 1, 0, ..., 0): the A operand of the per-layer bias MMA (see net_tc2.cu)
    for (int i = threadIdx.x; i < TC_AROWS; i += TC_THREADS)
        reinterpret_cast<uint4*>(sA + (size_t)16 * TC_PANEL_BYTES)[i] = make_uint4(0x3F803F80u, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    int iter = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x, iter++) {
        if (warp < TC_EPI_WARPS) {
            // ================= epilogue warps: one GEMM row (TMEM lane) per thread =================
            const int tile = warp >> 2, quarter = warp & 3;
            if (tile >= tiles) continue;
            const int m = tile * 128 + quarter * 32 + lane;
            const int pos = m / TC_POS_ROWS, idx = m - pos * TC_POS_ROWS;
            const int r = idx / 10, c = idx - 10 * r;
            const int gpos = g * P + pos;
            const bool valid = (pos < P) && (r < 9) && (c < 9) && (gpos < n_pos);
            float* hrow = headfeat + (size_t)gpos * 243 + (size_t)(r * 9 + c);
            uint4* rrow = reinterpret_cast<uint4*>(resid) + (size_t)blockIdx.x * (16 * TC_M) + (size_t)m;
            uint8_t* srow = sA + (size_t)(TC_LEAD + m) * 16;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tile * 128);

            // prologue: the three input planes of this row go into channel panel 0 (channels 3..15 are zero): the
            // A operand of conv_input, which runs as "layer -1" on the tensor pipe with K = 16 per tap
            {
                uint4 pk = make_uint4(0, 0, 0, 0);
                if (valid) {
                    const __nv_bfloat16* px = planes + (size_t)gpos * 243 + (size_t)(r * 9 + c);
                    uint32_t x0 = (uint32_t)__bfloat16_as_ushort(px[0]), x1 = (uint32_t)__bfloat16_as_ushort(px[81]),
                             x2 = (uint32_t)__bfloat16_as_ushort(px[162]);
                    pk = make_uint4(x0 | (x1 << 16), x2, 0u, 0u);
                }
                *reinterpret_cast<uint4*>(srow) = pk;
                *reinterpret_cast<uint4*>(srow + TC_PANEL_BYTES) = make_uint4(0, 0, 0, 0);
            }
            const bool nb_lo = (quarter == 0) && (tile > 0);            // rows also read by tile-1's MMAs
            const bool nb_hi = (quarter == 3) && (tile < tiles - 1);    // rows also read by tile+1's MMAs
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_act + 8 * tile);
                if (nb_lo) mbar_arrive(bar_act + 8 * (tile - 1));
                if (nb_hi) mbar_arrive(bar_act + 8 * (tile + 1));
            }

            // the last layer (heads' 1x1 convs instead of a write-back) is a separate instantiation of the body so that
            // its extra live registers do not burden the 32 common layers
            int layer = -1;
            auto epilogue_layer = [&](auto last_tag) {
                constexpr bool last = decltype(last_tag)::value;
                const uint32_t lpar = (uint32_t)((iter * TC_GROUP_LAYERS + layer + 1) & 1);
                const bool second = (layer >= 0) && (layer & 1) != 0;   // conv2 of a block: add the skip connection
                const bool keep = second || (layer < 0);                // output feeds the next block: keep it as skip
                // skip connection (fp16 panels in L2): the first 8 of the 16 panels are fetched while the MMAs still
                // run, the rest two chunk pairs ahead of their use (register budget: 672 threads x 96)
                const uint4 zero4 = make_uint4(0, 0, 0, 0);
                uint4 sk[8];
#pragma unroll
                for (int j = 0; j < 8; j++) sk[j] = (second && valid) ? rrow[(size_t)j * TC_M] : zero4;
                mbar_wait(bar_accum + 8 * tile, lpar, 128);
                if (nb_lo) mbar_wait(bar_accum + 8 * (tile - 1), lpar, 64);
                if (nb_hi) mbar_wait(bar_accum + 8 * (tile + 1), lpar, 64);
                if (dbg && blockIdx.x == 0 && iter == 0 && threadIdx.x == 0 && layer >= 0) dbg[layer * 4 + 2] = clock64();
                tc_fence_after();
                // 8 chunks of 16 accumulator columns, TMEM loads double-buffered against the math / stores
                float va[16], vb[16];
                float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f;               // last layer: the heads' 1x1 convolutions of this row
                tmem_ld16(taddr, va);
#pragma unroll
                for (int ch = 0; ch < 8; ch++) {
                    float* v = (ch & 1) ? vb : va;
                    tmem_ld_wait();
                    if (ch < 7) tmem_ld16(taddr + (uint32_t)((ch + 1) * 16), (ch & 1) ? va : vb);
                    f16x8_add2(sk[(2 * ch) & 7], v);
                    f16x8_add2(sk[(2 * ch + 1) & 7], v + 8);
                    if (ch < 4 && second && valid) {             // refill the two registers just consumed: panels +8
                        sk[(2 * ch) & 7] = rrow[(size_t)(2 * ch + 8) * TC_M];
                        sk[(2 * ch + 1) & 7] = rrow[(size_t)(2 * ch + 9) * TC_M];
                    }
                    if constexpr (last) {
                        // policy_conv / value_conv (1x1, dual_network.py:102,111): three dot products over this row
                        const float* hw = headw + ch * 16;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            float x = fmaxf(v[j], 0.0f);
                            h0 = fmaf(x, __ldg(hw + j), h0);
                            h1 = fmaf(x, __ldg(hw + 128 + j), h1);
                            h2 = fmaf(x, __ldg(hw + 256 + j), h2);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 2; j++) {
                            const float* w8 = v + 8 * j;
                            uint4 pk = valid ? make_uint4(relu_bf16x2(w8[0], w8[1]), relu_bf16x2(w8[2], w8[3]),
                                                          relu_bf16x2(w8[4], w8[5]), relu_bf16x2(w8[6], w8[7]))
                                             : zero4;                    // padding rows stay zero
                            *reinterpret_cast<uint4*>(srow + (size_t)(ch * 2 + j) * TC_PANEL_BYTES) = pk;
                            if (keep && valid)
                                rrow[(size_t)(ch * 2 + j) * TC_M] = make_uint4(relu_f16x2(w8[0], w8[1]), relu_f16x2(w8[2], w8[3]),
                                                                               relu_f16x2(w8[4], w8[5]), relu_f16x2(w8[6], w8[7]));
                        }
                    }
                }
                if constexpr (!last) {
                    fence_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bar_act + 8 * tile);
                        if (nb_lo) mbar_arrive(bar_act + 8 * (tile - 1));
                        if (nb_hi) mbar_arrive(bar_act + 8 * (tile + 1));
                    }
                } else if (valid) {
                    hrow[0] = fmaxf(h0 + __ldg(headw + 384), 0.0f);
                    hrow[81] = fmaxf(h1 + __ldg(headw + 385), 0.0f);
                    hrow[162] = fmaxf(h2 + __ldg(headw + 386), 0.0f);
                }
                if (dbg && blockIdx.x == 0 && iter == 0 && threadIdx.x == 0 && layer >= 0) dbg[layer * 4 + 3] = clock64();
            };
#pragma unroll 1
            for (layer = -1; layer < NET_LAYERS - 1; layer++) epilogue_layer(std::false_type{});
            epilogue_layer(std::true_type{});          // layer == NET_LAYERS - 1
            tc_fence_before();
        } else if (warp == 16) {
            // ================= weight producer =================
#pragma unroll 1
            int gn = iter * TC_GROUP_STAGES;
#pragma unroll 1
            for (int layer = -1; layer < NET_LAYERS; layer++) {
                const int n_st = 1 + ((layer < 0) ? TC_IN_STAGES : TC_STAGES_PER_LAYER);
#pragma unroll 1
                for (int st = 0; st < n_st; st++, gn++) {
                    const int stage = gn % TC_STAGES;
                    const uint32_t par = (uint32_t)((gn / TC_STAGES) & 1);
                    mbar_wait(bar_empty + 8 * stage, par ^ 1u);
                    if (lane == 0) {
                        const __nv_bfloat16* src;
                        uint32_t bytes = TC_STAGE_BYTES;
                        if (st == 0) { src = wq_bias + (size_t)(layer + 1) * (TC_BIAS_BYTES / 2); bytes = TC_BIAS_BYTES; }
                        else if (layer < 0) src = wq_in + (size_t)(st - 1) * (TC_STAGE_BYTES / 2);
                        else src = wq + (size_t)(layer * TC_STAGES_PER_LAYER + st - 1) * (TC_STAGE_BYTES / 2);
                        mbar_expect_tx(bar_full + 8 * stage, bytes);
                        bulk_g2s(sB_u + stage * TC_STAGE_BYTES, src, bytes, bar_full + 8 * stage);
                    }
                    __syncwarp();
                }
            }
        } else {
            // ================= MMA issuers: warp 17+t drives accumulator tile t =================
            const int tile = warp - 17;
            if (tile >= tiles) continue;
            const bool leader = elect_one();
            const uint32_t tmem_d = tmem_base + (uint32_t)(tile * 128);
            const uint32_t a_tile = sA_u + (uint32_t)(TC_LEAD + tile * 128) * 16u;
#pragma unroll 1
            for (int layer = -1; layer < NET_LAYERS; layer++) {
                mbar_wait(bar_act + 8 * tile, (uint32_t)((iter * TC_GROUP_LAYERS + layer + 1) & 1), 32);
                tc_fence_after();
                if (dbg && blockIdx.x == 0 && iter == 0 && tile == 0 && leader && layer >= 0) dbg[layer * 4 + 0] = clock64();
                const int n_st = 1 + ((layer < 0) ? TC_IN_STAGES : TC_STAGES_PER_LAYER);
                const int st0 = iter * TC_GROUP_STAGES +
                                ((layer < 0) ? 0 : (TC_IN_STAGES + 1) + layer * (TC_STAGES_PER_LAYER + 1));
#pragma unroll 1
                for (int s = 0; s < n_st; s++) {
                    const int gn = st0 + s;
                    const int stage = gn % TC_STAGES;
                    const uint32_t par = (uint32_t)((gn / TC_STAGES) & 1);
                    mbar_wait(bar_full + 8 * stage, par);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t b0 = sB_u + (uint32_t)stage * TC_STAGE_BYTES;
                        if (s == 0) {
                            // accumulator := BN shift (constant panel x bias block); starts the layer's accumulation
                            umma_bf16(tmem_d, make_desc(a_tile + 16u * TC_PANEL_BYTES, TC_PANEL_BYTES, 128),
                                      make_desc(b0, 2048, 128), TC_IDESC, 0u);
                        } else if (layer < 0) {
                            // conv_input: block j of stage s is tap 4(s-1)+j, K = 16 (channel panels 0,1)
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int tap = 4 * (s - 1) + j;
                                if (tap < 9) {
                                    const int shift = (tap / 3 - 1) * 10 + (tap % 3 - 1);
                                    umma_bf16(tmem_d, make_desc(a_tile + (uint32_t)(shift * 16), TC_PANEL_BYTES, 128),
                                              make_desc(b0 + (uint32_t)j * 4096u, 2048, 128), TC_IDESC, 1u);
                                }
                            }
                        } else {
                            // stage s holds K-blocks 4(s-1) .. 4(s-1)+3 in the order of tcx::kblock_of
#pragma unroll
                            for (int ks = 0; ks < 4; ks++) {
                                int q, tap, unit;
                                kblock_decode(4 * (s - 1) + ks, q, tap, unit);
                                const uint32_t a0 = a_tile + (uint32_t)(tap_shift(tap) * 16) + (uint32_t)(2 * unit) * TC_PANEL_BYTES;
                                umma_bf16(tmem_d, make_desc(a0, TC_PANEL_BYTES, 128),
                                          make_desc(b0 + (uint32_t)ks * 4096u, 2048, 128), TC_IDESC, 1u);
                            }
                        }
                        umma_commit(bar_empty + 8 * stage);          // frees the weight stage when the MMAs retire
                        if (s == n_st - 1) {
                            umma_commit(bar_accum + 8 * tile);       // this tile's accumulator is complete
                            if (dbg && blockIdx.x == 0 && iter == 0 && tile == 0 && layer >= 0) dbg[layer * 4 + 1] = clock64();
                        }
                    }
                    __syncwarp();
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 17) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

int trunk_tc_smem_bytes() { return TC_SMEM_BYTES; }

cudaError_t trunk_tc_init() {
    return cudaFuncSetAttribute(trunk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
}

cudaError_t launch_trunk_tc(const NetWeights& w, const __nv_bfloat16* planes, float* headfeat, const int32_t* count,
                            int max_rows, float* resid, int n_sm, cudaStream_t s, long long* dbg, int min_count) {
    int grid = max_rows < n_sm ? max_rows : n_sm;
    if (grid < 1) grid = 1;
    trunk_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(w.res_w_bf16, w.conv_in_w_bf16, w.bias_blk, planes, w.head_w, headfeat, resid,
                                                            count, min_count, dbg);
    return cudaGetLastError();
}

}  // namespace uttt
