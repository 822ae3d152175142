// tcgen05 / TMA / TMEM GEMM engine on sm_100a with split-bf16 operands, two front ends:
//   * nn.Linear          C[M,N] = act(A[M,K] W[N,K]^T * scale + shift + res_pre) + res_post     (rp_linear_tc)
//   * nn.Conv2d (NHWC)   implicit GEMM: the A tile of every (filter tap, 64-channel block) K step is ONE
//                        5-D TMA box of the activation tensor [plane][image][H][W][C]; zero padding and
//                        image borders are TMA out-of-bounds zero fill, stride-2 is the map's element
//                        stride.  Eval-mode BatchNorm is the scale/shift epilogue            (rp_conv2d_tc)
//
// Operands are bf16 "planes": a float32 value x is carried as P bf16 numbers x0 = bf16(x),
// x1 = bf16(x - x0) (P = 2 keeps 16 mantissa bits).  The tensor core then evaluates
//   P = 1 :  a0 b0                         (plain bf16, BASELINE.json configs[3] throughput mode)
//   P = 2 :  a0 b0 + a0 b1 + a1 b0         ("bf16x3": fp32-class products, fp32 accumulation in TMEM;
//                                           measured end to end: holds the 1e-4 parity bar)
// into ONE fp32 TMEM accumulator.  Reference call sites: vision_transformer.py:323,331; mlp.py:21-24;
// src/model.py:127-134; extractor.py:51-65.
//
// Kernel anatomy (persistent, one CTA per SM, 320 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor (128B swizzle) into a STAGES-deep shared-memory ring
//              guarded by full/empty mbarriers;
//   warp 1     MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=16) reading K-major
//              SWIZZLE_128B smem descriptors; tcgen05.commit releases ring slots and publishes the
//              accumulator; also owns TMEM alloc/dealloc (512 columns = 2 accumulator stages x 256);
//   warps 2-9  epilogue: tcgen05.ld (32 lanes x 32 columns per warp and step), transpose through a
//              warp-private shared-memory patch so that scale / shift / activation / residual / plane
//              split run on coalesced 128-byte row segments (fp32 and/or re-split bf16 plane outputs).
// The two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// Convolution tiles are R whole output rows of one image (R*Wo <= 128 pixels) so that the pixels of a
// tile are one rectangular TMA box for every filter tap; unused accumulator rows are never stored.
#include <cstdlib>
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int A_TILE = BM * BK * 2;   // 16 KiB
constexpr int EPI_WARPS = 16;                      // four per TMEM lane quarter, each owning a quarter of the columns
constexpr int NTHREADS = 32 * (2 + EPI_WARPS);     // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr int TMEM_COLS = 512, ACC_STRIDE = 256;
constexpr int STG_LD = 16;                         // staging patch: 32 rows x 16 floats, 16-byte chunks XOR-swizzled
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
constexpr int SMEM_BUDGET = 200 * 1024;

template <int P, int BN>
struct Cfg {
    static constexpr int B_TILE = BN * BK * 2;
    static constexpr int STAGE_BYTES = P * (A_TILE + B_TILE);
    static constexpr int STAGES_RAW = (SMEM_BUDGET - STG_BYTES) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
    static constexpr int SMEM = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(STAGES >= 2, "pipeline too shallow");
};

struct Geom {
    int M, N, K;          // GEMM view: output rows (pixels), output columns, reduction length
    int ksplit, kb_per_split;   // nn.Linear only: split the reduction over `ksplit` tiles writing float32 partials (0: off)
    // convolution only
    int Ho, Wo, R, tiles_per_img, C, KW, stride, pad, cblocks, taps;
};

struct EpiParams {
    const float* scale;      // [N] or null
    const float* shift;      // [N] or null (nn.Linear bias / folded BN shift)
    const float* res_pre;    // [M,N] or null: added before the activation
    const float* res_post;   // added after the activation, row index modulo res_post_rows (0: same rows)
    int res_post_rows;
    float* out_f32;          // [M,N] or null
    __nv_bfloat16* out_planes;   // [P_out][M][N] or null
    int p_out;
    int act;
};

using tc::gelu_fast;

__device__ __forceinline__ float act_fn(float v, int act) {
    if (act == RP_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    if (act == RP_ACT_RELU) return fmaxf(v, 0.0f);
    return v;
}

// FOLD: long reductions (K > 1600: the two 5x5 convolutions that produce the tokens) are cut into segments of at most
// FOLD_KB K blocks.  tcgen05 accumulates in fp32 with TRUNCATION, so one TMEM accumulator that lives for n MMAs carries
// a bias of ~n 2^-25 (measured 3.5e-9 K relative: 1.6e-5 at K = 4800, the largest single term of the front end's token
// error).  Every segment gets a fresh accumulator (the two TMEM stages alternate); the epilogue warps add the finished
// segments in fp32 registers (round to nearest) while the next segment's MMAs run, and run the usual epilogue on the sum.
// Measured (64 pairs): segments of 12 K blocks on every K > 768 layer cost 0.14 ms per step (the unrolled, spilling
// epilogue of the FOLD variant is on the critical path of the short-K 3x3 layers) for a pose error against the float64
// reference of 3.7e-5 rad / 2.9e-5 instead of 4.8e-5 / 4.9e-5; segments of 25 on the 5x5 layers only keep most of the
// gain (their chains were the 1.6e-5 / 1.0e-5 terms) at a fraction of the cost.
constexpr int FOLD_KB = 25;

template <int P, int BN, bool CONV, bool FOLD = false>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, EpiParams ep, Geom g) {
    using C = Cfg<P, BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a SHARED pointer (LDS / STS)
    float* staging = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + STG_BYTES);
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + C::STAGES;          // [STAGES]
    uint64_t* tfull = bars + 2 * C::STAGES;      // [2]
    uint64_t* tempty = bars + 2 * C::STAGES + 2; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = g.M, N = g.N;
    const int tiles_n = CONV ? 1 : (N + BN - 1) / BN;
    const int tiles_m = CONV ? (g.M / (g.Ho * g.Wo)) * g.tiles_per_img : (M + BM - 1) / BM;
    const int nsplit = (!CONV && g.ksplit > 1) ? g.ksplit : 1;
    const int ntiles = tiles_m * tiles_n * nsplit;
    const int ksteps_all = CONV ? g.taps * g.cblocks : (g.K + BK - 1) / BK;
    const int nseg = FOLD ? (ksteps_all + FOLD_KB - 1) / FOLD_KB : 1;       // accumulator segments per tile
    const int seg_len = (ksteps_all + nseg - 1) / nseg;
    // split-K (weight-bandwidth-bound skinny GEMMs: the 26880 -> 512 regressor layer): tile = (split, m, n), every
    // split reduces kb_per_split K blocks into its own float32 partial [split][M][N]; short accumulation chains
    auto k_range = [&](int tile, int& ks0, int& ks1) {
        const int split = tile / (tiles_m * tiles_n);
        ks0 = nsplit > 1 ? split * g.kb_per_split : 0;
        ks1 = nsplit > 1 ? min(ksteps_all, ks0 + g.kb_per_split) : ksteps_all;
    };
    // bytes one stage receives from TMA: full boxes, zero-filled where out of bounds
    const uint32_t stage_tx = CONV ? (uint32_t)(P * (g.R * g.Wo * 128 + C::B_TILE)) : (uint32_t)C::STAGE_BYTES;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        for (int i = 0; i < C::STAGES; ++i) {
            tc::mbar_init(&full[i], 1);
            tc::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&tfull[i], 1);
            tc::mbar_init(&tempty[i], EPI_WARPS);     // one elected arrive per epilogue warp (512 arrives on one mbarrier
                                                      // serialise in the shared-memory atomic unit)
        }
        tc::fence_barrier_init();
    }
    rp::pdl_launch_dependents();                  // the next kernel may start its prologue (common.cuh)
    if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    rp::pdl_wait();                               // the previous kernel has completed: its outputs are visible
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto a_tile = [&](int stage, int p) { return smem + stage * C::STAGE_BYTES + p * A_TILE; };
    auto b_tile = [&](int stage, int p) { return smem + stage * C::STAGE_BYTES + P * A_TILE + p * C::B_TILE; };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (convergent warp)
        int stage = 0, phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int tmn = tile % (tiles_m * tiles_n);
            const int tm = tmn / tiles_n, n0 = (tmn % tiles_n) * BN;
            const int img = CONV ? tm / g.tiles_per_img : 0;
            const int oy0 = CONV ? (tm % g.tiles_per_img) * g.R : 0;
            int ks0, ks1;
            k_range(tile, ks0, ks1);
            for (int ks = ks0; ks < ks1; ++ks) {
                tc::mbar_wait(&empty[stage], phase ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(&full[stage], stage_tx);
                    if (CONV) {
                        const int tap = ks / g.cblocks, cb = ks - tap * g.cblocks;
                        const int ky = tap / g.KW, kx = tap - ky * g.KW;
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            tc::tma_load_5d(a_tile(stage, p), &tmA, &full[stage], cb * 64, kx - g.pad,
                                            oy0 * g.stride + ky - g.pad, img, p);
                            tc::tma_load_3d(b_tile(stage, p), &tmB, &full[stage], tap * g.C + cb * 64, n0, p);
                        }
                    } else {
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            tc::tma_load_3d(a_tile(stage, p), &tmA, &full[stage], ks * BK, tm * BM, p);
                            tc::tma_load_3d(b_tile(stage, p), &tmB, &full[stage], ks * BK, n0, p);
                        }
                    }
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (convergent warp)
        constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
        int stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int kt0, kt1;
            k_range(tile, kt0, kt1);
          for (int seg = 0; seg < nseg; ++seg) {
            const int ks0 = FOLD ? kt0 + seg * seg_len : kt0, ks1 = FOLD ? min(kt1, ks0 + seg_len) : kt1;
            tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
            tc::tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
            for (int ks = ks0; ks < ks1; ++ks) {
                tc::mbar_wait(&full[stage], phase);
                tc::tcgen05_fence_after();
                // descriptors of the stage's tiles; a K step of 16 bf16 = 32 B inside the 128-byte swizzle row = +2
                const uint64_t a0 = tc::make_kmajor_sw128_desc(tc::smem_u32(a_tile(stage, 0)));
                const uint64_t b0 = tc::make_kmajor_sw128_desc(tc::smem_u32(b_tile(stage, 0)));
                const uint64_t a1 = tc::make_kmajor_sw128_desc(tc::smem_u32(a_tile(stage, P - 1)));
                const uint64_t b1 = tc::make_kmajor_sw128_desc(tc::smem_u32(b_tile(stage, P - 1)));
                if (tc::elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        uint32_t accum = (ks > ks0 || k > 0) ? 1u : 0u;
                        // smallest terms first; all land in the same fp32 accumulator
                        if (P == 2) {
                            tc::umma_bf16(d_tmem, a1 + 2 * k, b0 + 2 * k, idesc, accum);
                            tc::umma_bf16(d_tmem, a0 + 2 * k, b1 + 2 * k, idesc, 1u);
                            accum = 1u;
                        }
                        tc::umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, accum);
                    }
                    tc::umma_commit(&empty[stage]);           // frees the smem slot when these MMAs retire
                    if (ks + 1 == ks1) tc::umma_commit(&tfull[acc]);   // accumulator complete -> epilogue
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..17)
        // TMEM hands every thread one ROW; global memory wants a warp to touch whole row segments.  Each warp
        // owns a 32-row x (BN/4)-column slab of the tile and walks it in 16-column chunks: tcgen05.ld (x16),
        // transpose through a private 2 KiB shared-memory patch (16-byte chunks XOR-swizzled by row pair:
        // conflict free both ways), then the epilogue arithmetic on the coalesced side -- 4 lanes x 16 B per row,
        // 8 rows per instruction.  ncu showed the 8-warp version issue bound at ~0.25 IPC per scheduler with the
        // per-column constants and residual loads exposed: 16 warps double the latency hiding, the constants are
        // requested before the TMEM load and all residual loads of a chunk are in flight before the first store.
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int part = (warp - 2) >> 2;                // column quarter owned by this warp
        float* stg = staging + (warp - 2) * 32 * STG_LD;
        int acc = 0, acc_phase = 0;
        const bool vec4 = (N % 4) == 0;
        const int rr = lane >> 2, cq = lane & 3;         // read side: row inside an 8-row group, 16-byte chunk
        constexpr int NCHUNK = BN / 16;                  // 16-column chunks per tile (4, 8 or 12)
        constexpr int CH_PER_PART = NCHUNK / 4;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int tmn = tile % (tiles_m * tiles_n);
            const int tm = tmn / tiles_n, n0 = (tmn % tiles_n) * BN;
            float* const out_f32 = ep.out_f32 ? ep.out_f32 + (size_t)(tile / (tiles_m * tiles_n)) * M * N : nullptr;   // split-K partial
            int row_base, valid_rows;                    // global row of tile row 0; rows of the tile that exist
            if (CONV) {
                const int img = tm / g.tiles_per_img, oy0 = (tm % g.tiles_per_img) * g.R;
                row_base = (img * g.Ho + oy0) * g.Wo;
                valid_rows = min(g.R, g.Ho - oy0) * g.Wo;
            } else {
                row_base = tm * BM;
                valid_rows = min(BM, M - row_base);
            }
            // finished accumulator segments are summed in registers (FOLD): this thread's row, its CH_PER_PART chunks
            float facc[FOLD ? CH_PER_PART * 16 : 1];
            if constexpr (FOLD) {
                for (int seg = 0; seg + 1 < nseg; ++seg) {
                    tc::mbar_wait(&tfull[acc], acc_phase);
                    tc::tcgen05_fence_after();
                    const uint32_t t_seg = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
#pragma unroll
                    for (int ci = 0; ci < CH_PER_PART; ++ci) {
                        uint32_t r[16];
                        tc::tmem_ld_32x32b_x16(t_seg + (part * CH_PER_PART + ci) * 16, r);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            facc[ci * 16 + i] = seg == 0 ? __uint_as_float(r[i]) : facc[ci * 16 + i] + __uint_as_float(r[i]);
                    }
                    tc::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&tempty[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
            tc::mbar_wait(&tfull[acc], acc_phase);
            tc::tcgen05_fence_after();
            const uint32_t t_row = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
            constexpr int CI_UNROLL = FOLD ? CH_PER_PART : 1;    // facc[] needs compile-time indices; otherwise keep the loop rolled
#pragma unroll(CI_UNROLL)
            for (int ci = 0; ci < CH_PER_PART; ++ci) {
                const int c0 = (part * CH_PER_PART + ci) * 16;
                if (n0 + c0 >= N) break;                 // warp-uniform
                const int col = n0 + c0 + cq * 4;        // the lane's 4 columns: the same for all row groups
                const bool colv = vec4 && col < N;
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (colv && ep.scale) sc = __ldg(reinterpret_cast<const float4*>(ep.scale + col));
                if (colv && ep.shift) sh = __ldg(reinterpret_cast<const float4*>(ep.shift + col));
                float4 rpre[4], rpost[4];
                if (vec4 && !FOLD) {          // FOLD: the partial sums occupy these registers; residuals are read at use
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int rt = q * 32 + it * 8 + rr;
                        const bool ok = colv && rt < valid_rows;
                        const int row = row_base + rt;
                        rpre[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                        rpost[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok && ep.res_pre) rpre[it] = __ldg(reinterpret_cast<const float4*>(ep.res_pre + (size_t)row * N + col));
                        if (ok && ep.res_post) {
                            const int rq = ep.res_post_rows > 0 ? row % ep.res_post_rows : row;
                            rpost[it] = __ldg(reinterpret_cast<const float4*>(ep.res_post + (size_t)rq * N + col));
                        }
                    }
                }
                uint32_t r[16];
                tc::tmem_ld_32x32b_x16(t_row + c0, r);
                tc::tmem_ld_wait();
                if constexpr (FOLD) {
                    if (nseg > 1) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(facc[ci * 16 + i] + __uint_as_float(r[i]));
                    }
                }
                __syncwarp();                            // the previous chunk's readers are done with the patch
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(stg + lane * STG_LD + ((j ^ ((lane >> 1) & 3)) << 2)) =
                        make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
                if (vec4) {
#pragma unroll
                    for (int it = 0; it < 4; ++it) {
                        const int rl = it * 8 + rr;                  // row inside the warp's 32-row slab
                        const int rt = q * 32 + rl;
                        if (!colv || rt >= valid_rows) continue;
                        const size_t o = (size_t)(row_base + rt) * N + col;
                        const float4 a = *reinterpret_cast<const float4*>(stg + rl * STG_LD + ((cq ^ ((rl >> 1) & 3)) << 2));
                        if constexpr (FOLD) {
                            rpre[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                            rpost[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (ep.res_pre) rpre[it] = __ldg(reinterpret_cast<const float4*>(ep.res_pre + o));
                            if (ep.res_post) {
                                const int rq = ep.res_post_rows > 0 ? (row_base + rt) % ep.res_post_rows : (row_base + rt);
                                rpost[it] = __ldg(reinterpret_cast<const float4*>(ep.res_post + (size_t)rq * N + col));
                            }
                        }
                        float v0 = fmaf(a.x, sc.x, sh.x) + rpre[it].x, v1 = fmaf(a.y, sc.y, sh.y) + rpre[it].y;
                        float v2 = fmaf(a.z, sc.z, sh.z) + rpre[it].z, v3 = fmaf(a.w, sc.w, sh.w) + rpre[it].w;
                        if (ep.act == RP_ACT_GELU) {
                            v0 = gelu_fast(v0); v1 = gelu_fast(v1); v2 = gelu_fast(v2); v3 = gelu_fast(v3);
                        } else if (ep.act == RP_ACT_RELU) {
                            v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
                        }
                        v0 += rpost[it].x; v1 += rpost[it].y; v2 += rpost[it].z; v3 += rpost[it].w;
                        if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = make_float4(v0, v1, v2, v3);
                        if (ep.out_planes) {
                            for (int p = 0; p < ep.p_out; ++p) {
                                __nv_bfloat162 h01 = __floats2bfloat162_rn(v0, v1), h23 = __floats2bfloat162_rn(v2, v3);
                                uint2 w;
                                w.x = *reinterpret_cast<uint32_t*>(&h01);
                                w.y = *reinterpret_cast<uint32_t*>(&h23);
                                *reinterpret_cast<uint2*>(ep.out_planes + (size_t)p * M * N + o) = w;
                                v0 -= __uint_as_float(w.x << 16); v1 -= __uint_as_float(w.x & 0xffff0000u);   // residue
                                v2 -= __uint_as_float(w.y << 16); v3 -= __uint_as_float(w.y & 0xffff0000u);
                            }
                        }
                    }
                } else {
#pragma unroll 1
                    for (int it = 0; it < 4; ++it) {
                        const int rl = it * 8 + rr;
                        const int rt = q * 32 + rl;
                        const int colx = n0 + c0 + cq * 4;
                        if (rt >= valid_rows || colx >= N) continue;
                        const int row = row_base + rt;
                        const size_t o = (size_t)row * N + colx;
                        const int nvalid = min(4, N - colx);
                        for (int j = 0; j < nvalid; ++j) {
                            float x = stg[rl * STG_LD + ((cq ^ ((rl >> 1) & 3)) << 2) + j];
                            if (ep.scale) x *= ep.scale[colx + j];
                            if (ep.shift) x += ep.shift[colx + j];
                            if (ep.res_pre) x += ep.res_pre[o + j];
                            x = act_fn(x, ep.act);
                            if (ep.res_post) {
                                const int rq = ep.res_post_rows > 0 ? row % ep.res_post_rows : row;
                                x += ep.res_post[(size_t)rq * N + colx + j];
                            }
                            if (out_f32) out_f32[o + j] = x;
                            if (ep.out_planes)
                                for (int p = 0; p < ep.p_out; ++p) {
                                    __nv_bfloat16 h = __float2bfloat16_rn(x);
                                    ep.out_planes[(size_t)p * M * N + o + j] = h;
                                    x -= __bfloat162float(h);
                                }
                        }
                    }
                }
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------ plane producers
// x -> P bf16 planes (x0 = bf16(x), x1 = bf16(x - x0), ...), 8 elements per thread, 16-byte stores
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           long long n, int P) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    float v[8];
    if (i + 8 <= n) {
        float4 a = *reinterpret_cast<const float4*>(x + i), b = *reinterpret_cast<const float4*>(x + i + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
        for (int j = 0; j < 8; ++j) v[j] = (i + j < n) ? x[i + j] : 0.f;
    }
    for (int p = 0; p < P; ++p) {
        __nv_bfloat16 h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            h[j] = __float2bfloat16_rn(v[j]);
            v[j] -= __bfloat162float(h[j]);
        }
        __nv_bfloat16* dst = out + (size_t)p * n + i;
        if (i + 8 <= n) {
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
        } else {
            for (int j = 0; j < 8 && i + j < n; ++j) dst[j] = h[j];
        }
    }
}

// x [R][C] float32 -> bf16 planes of the TRANSPOSE, [P][C][R]: the weight-gradient GEMM dW = dY^T X of the training path
// contracts over the rows of both factors, so it runs on the K-major engine above with both operands transposed
// (train_path.py).  64 x 32 tile through shared memory: 128-byte global reads and writes (two rows per thread on the
// store side, packed bf16x2).
__global__ void __launch_bounds__(256) transpose_split_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                                     int R, int C, int P) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    __shared__ float t[32][65];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        t[tx][ty + 8 * k] = (r < R && c < C) ? x[(size_t)r * C + c] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int cl = ty + 8 * m, c = c0 + cl, r = r0 + 2 * tx;
        if (c >= C || r >= R) continue;
        float v0 = t[cl][2 * tx], v1 = t[cl][2 * tx + 1];
        for (int p = 0; p < P; ++p) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
            const uint32_t w = *reinterpret_cast<uint32_t*>(&h);
            __nv_bfloat16* dst = out + ((size_t)p * C + c) * R + r;
            if (r + 1 < R) *reinterpret_cast<uint32_t*>(dst) = w;
            else dst[0] = __ushort_as_bfloat16((unsigned short)(w & 0xffffu));
            v0 -= __uint_as_float(w << 16);
            v1 -= __uint_as_float(w & 0xffff0000u);
        }
    }
}

// LayerNorm (eps 1e-6, vision_transformer.py:396) writing bf16 planes directly (A operand of the next GEMM).
// One warp per row; cols % 8 == 0 and cols <= 256: lane l owns columns [8l, 8l+8) -- two 16-byte loads and one
// 16-byte store per plane (the first version stored 2 bytes per lane: 3.4 TB/s; HBM-bound work wants full sectors).
__global__ void __launch_bounds__(256) layernorm_planes_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                               const float* __restrict__ b, __nv_bfloat16* __restrict__ out,
                                                               int rows, int cols, float eps, int P) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const int c0 = lane * 8;
    const bool act = c0 < cols;
    float v[8];
    float s = 0.f;
    if (act) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(x + (size_t)warp * cols + c0));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(x + (size_t)warp * cols + c0 + 4));
        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    const float mean = rp::warp_sum(s) / (float)cols;
    float q = 0.f;
    if (act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q += d * d; }
    }
    const float rstd = 1.0f / sqrtf(rp::warp_sum(q) / (float)cols + eps);
    if (!act) return;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + c0)), g1 = __ldg(reinterpret_cast<const float4*>(g + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c0)), b1 = __ldg(reinterpret_cast<const float4*>(b + c0 + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * gg[i] + bb[i];
    for (int p = 0; p < P; ++p) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
            v[2 * i] -= __uint_as_float(w[i] << 16);
            v[2 * i + 1] -= __uint_as_float(w[i] & 0xffff0000u);
        }
        *reinterpret_cast<uint4*>(out + ((size_t)p * rows + warp) * cols + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// nn.MaxPool2d(3,2,1) on NHWC float32 -> float32 and/or bf16 planes (input of the first tensor-core conv)
__global__ void __launch_bounds__(256) maxpool_planes_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                             __nv_bfloat16* __restrict__ yp, int P, int n_img, int H, int W,
                                                             int C4, int Ho, int Wo) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    long long total = (long long)n_img * Ho * Wo * C4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C4);
        long long pix = idx / C4;
        int ox = (int)(pix % Wo), oy = (int)((pix / Wo) % Ho), n = (int)(pix / ((long long)Wo * Ho));
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            int iy = oy * 2 - 1 + dy;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                int ix = ox * 2 - 1 + dx;
                if (ix < 0 || ix >= W) continue;
                float4 v = x[(((long long)n * H + iy) * W + ix) * C4 + c];
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        if (y) y[idx] = m;
        if (yp) {
            float v[4] = {m.x, m.y, m.z, m.w};
            for (int p = 0; p < P; ++p) {
                __nv_bfloat16 h[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    h[j] = __float2bfloat16_rn(v[j]);
                    v[j] -= __bfloat162float(h[j]);
                }
                uint2 w;
                w.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
                w.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
                *reinterpret_cast<uint2*>(yp + ((size_t)p * total + idx) * 4) = w;
            }
        }
    }
}

template <int P, int BN, bool CONV, bool FOLD = false>
int launch_gemm_tc(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiParams& ep, const Geom& g, int ntiles,
                   int device, cudaStream_t st, const char* what) {
    using C = Cfg<P, BN>;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<P, BN, CONV, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            rp::set_error("%s: cudaFuncSetAttribute(%d): %s", what, C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    int grid = ntiles < rp::num_sms(device) ? ntiles : rp::num_sms(device);
    rp::launch(gemm_tc_kernel<P, BN, CONV, FOLD>, dim3(grid), dim3(NTHREADS), (size_t)(C::SMEM), st, tmA, tmB, ep, g);
    return rp::finish_launch(what);
}

}  // namespace

extern "C" int rp_split_planes_bf16(const float* x, void* planes, int64_t n, int P, int device, void* stream) {
    RP_REQUIRE(x && planes && n > 0 && (P == 1 || P == 2 || P == 3), RP_EINVAL, "rp_split_planes: bad argument");
    RP_REQUIRE(rp::aligned16(x) && rp::aligned16(planes) && (n % 8) == 0, RP_EALIGN,
               "rp_split_planes: 16-byte alignment and n %% 8 == 0 required");
    RP_GUARD(device);
    long long threads = (n + 7) / 8;
    rp::launch(split_planes_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        x, static_cast<__nv_bfloat16*>(planes), n, P);
    return rp::finish_launch("rp_split_planes");
}

extern "C" int rp_transpose_split_planes_bf16(const float* x, void* planes, int R, int C, int P, int device, void* stream) {
    RP_REQUIRE(x && planes && R > 0 && C > 0 && (P == 1 || P == 2), RP_EINVAL, "rp_transpose_split_planes: bad argument");
    RP_REQUIRE((R % 2) == 0 && (reinterpret_cast<uintptr_t>(planes) & 3) == 0, RP_EALIGN,
               "rp_transpose_split_planes: R must be even and planes 4-byte aligned");
    RP_GUARD(device);
    dim3 grid((R + 63) / 64, (C + 31) / 32);
    RP_REQUIRE(grid.y <= 65535, RP_EINVAL, "rp_transpose_split_planes: too many columns");
    rp::launch(transpose_split_planes_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, x, static_cast<__nv_bfloat16*>(planes), R, C, P);
    return rp::finish_launch("rp_transpose_split_planes");
}

extern "C" int rp_layernorm_planes_bf16(const float* x, const float* gamma, const float* beta, void* planes, int rows,
                                        int cols, float eps, int P, int device, void* stream) {
    RP_REQUIRE(x && gamma && beta && planes && rows > 0 && cols > 0 && cols <= 256 && (cols % 8) == 0 && (P == 1 || P == 2), RP_EINVAL,
               "rp_layernorm_planes: bad argument (cols <= 256, cols %% 8 == 0)");
    RP_REQUIRE(rp::aligned16(x) && rp::aligned16(gamma) && rp::aligned16(beta) && rp::aligned16(planes), RP_EALIGN,
               "rp_layernorm_planes: 16-byte alignment");
    RP_GUARD(device);
    rp::launch(layernorm_planes_kernel, dim3((rows + 7) / 8), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        x, gamma, beta, static_cast<__nv_bfloat16*>(planes), rows, cols, eps, P);
    return rp::finish_launch("rp_layernorm_planes");
}

extern "C" int rp_maxpool3x3s2_planes(const float* x, float* y_f32, void* y_planes, int P, int n_img, int H, int W, int C,
                                      int device, void* stream) {
    RP_REQUIRE(x && (y_f32 || y_planes) && n_img > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0 && (P == 1 || P == 2),
               RP_EINVAL, "rp_maxpool3x3s2_planes: bad argument");
    RP_REQUIRE(rp::aligned16(x) && rp::aligned16(y_f32) && rp::aligned16(y_planes), RP_EALIGN, "rp_maxpool3x3s2_planes: alignment");
    RP_GUARD(device);
    int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    long long total = (long long)n_img * Ho * Wo * (C / 4);
    long long blocks = (total + 255) / 256;
    long long cap = (long long)rp::num_sms(device) * 16;
    rp::launch(maxpool_planes_kernel, dim3((unsigned)(blocks < cap ? blocks : cap)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y_f32), static_cast<__nv_bfloat16*>(y_planes), P,
        n_img, H, W, C / 4, Ho, Wo);
    return rp::finish_launch("rp_maxpool3x3s2_planes");
}

// out[m][n] = act(bias[n] + sum over splits of part[s][m][n]), fixed summation order (deterministic)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias,
                                                            float* __restrict__ out, long long mn, int N, int nsplit, int act) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= mn) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s2 = 0;
    for (; s2 + 8 <= nsplit; s2 += 8) {                    // eight partial loads in flight, added in split order
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(part + (size_t)(s2 + u) * mn + i));
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; s2 < nsplit; ++s2) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(part + (size_t)s2 * mn + i));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + (i % N)));
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    acc.x = act_fn(acc.x, act); acc.y = act_fn(acc.y, act); acc.z = act_fn(acc.z, act); acc.w = act_fn(acc.w, act);
    *reinterpret_cast<float4*>(out + i) = acc;
}

extern "C" size_t rp_linear_tc_splitk_workspace_bytes(int M, int N, int K, int* ksplit_out) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    const int kblocks = (K + BK - 1) / BK;
    const int tiles_mn = ((M + BM - 1) / BM) * ((N + 191) / 192);
    int want = (148 + tiles_mn - 1) / tiles_mn;            // about one tile per SM
    if (want > kblocks) want = kblocks;
    if (want < 1) want = 1;
    const int kbps = (kblocks + want - 1) / want;
    const int ksplit = (kblocks + kbps - 1) / kbps;
    if (ksplit_out) *ksplit_out = ksplit;
    return (size_t)ksplit * M * N * sizeof(float);
}

// Split-K nn.Linear for skinny problems whose time is the weight stream (pose_regressor.0: [B,26880] x [512,26880]^T,
// src/model.py:91-98,189): every tile reduces a slice of K into a float32 partial, a second kernel adds the partials
// in a fixed order and applies bias + activation.  Accumulation chains stay short (K / ksplit), which also keeps the
// truncation bias of the tensor-memory accumulator away from a 26 880-long reduction.
extern "C" int rp_linear_tc_splitk(const void* A_planes, const void* W_planes, const float* bias, float* out_f32, int M, int N,
                                   int K, int P, int act, void* workspace, size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(A_planes && W_planes && out_f32 && workspace, RP_EINVAL, "rp_linear_tc_splitk: null pointer");
    RP_REQUIRE(M > 0 && N > 0 && (N % 4) == 0 && K > 0 && (K % 8) == 0, RP_EINVAL, "rp_linear_tc_splitk: bad shape M=%d N=%d K=%d", M, N, K);
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_linear_tc_splitk: P must be 1 or 2");
    RP_REQUIRE(act >= RP_ACT_NONE && act <= RP_ACT_RELU, RP_EINVAL, "rp_linear_tc_splitk: bad act %d", act);
    int ksplit = 1;
    const size_t need = rp_linear_tc_splitk_workspace_bytes(M, N, K, &ksplit);
    RP_REQUIRE(workspace_bytes >= need, RP_EWORKSPACE, "rp_linear_tc_splitk: workspace %zu < %zu bytes", workspace_bytes, need);
    RP_REQUIRE(rp::aligned16(A_planes) && rp::aligned16(W_planes) && rp::aligned16(workspace) && rp::aligned16(out_f32) &&
                   rp::aligned16(bias), RP_EALIGN, "rp_linear_tc_splitk: 16-byte alignment");
    RP_GUARD(device);
    constexpr int BN = 192;
    CUtensorMap tmA, tmB;
    int rc = tc::make_planes_tmap(&tmA, A_planes, P, M, K, BM);
    if (rc) return rc;
    rc = tc::make_planes_tmap(&tmB, W_planes, P, N, K, BN);
    if (rc) return rc;
    EpiParams ep{nullptr, nullptr, nullptr, nullptr, 0, static_cast<float*>(workspace), nullptr, 0, RP_ACT_NONE};
    Geom g{};
    g.M = M; g.N = N; g.K = K;
    const int kblocks = (K + BK - 1) / BK;
    g.ksplit = ksplit;
    g.kb_per_split = (kblocks + ksplit - 1) / ksplit;
    const int ntiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * ksplit;
    cudaStream_t st = (cudaStream_t)stream;
    rc = (P == 1) ? launch_gemm_tc<1, BN, false>(tmA, tmB, ep, g, ntiles, device, st, "rp_linear_tc_splitk")
                  : launch_gemm_tc<2, BN, false>(tmA, tmB, ep, g, ntiles, device, st, "rp_linear_tc_splitk");
    if (rc) return rc;
    const long long mn = (long long)M * N;
    rp::launch(splitk_reduce_kernel, dim3((unsigned)((mn / 4 + 255) / 256)), dim3(256), (size_t)(0), st, static_cast<const float*>(workspace), bias, out_f32, mn, N,
                                                                        ksplit, act);
    return rp::finish_launch("rp_linear_tc_splitk(reduce)");
}

extern "C" int rp_linear_tc(const void* A_planes, const void* W_planes, const float* bias, const float* residual,
                            float* out_f32, void* out_planes, int M, int N, int K, int P, int P_out, int act, int device,
                            void* stream) {
    RP_REQUIRE(A_planes && W_planes && (out_f32 || out_planes), RP_EINVAL, "rp_linear_tc: null pointer");
    RP_REQUIRE(M > 0 && N > 0 && K > 0 && (K % 8) == 0, RP_EINVAL, "rp_linear_tc: bad shape M=%d N=%d K=%d (K%%8==0)", M, N, K);
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_linear_tc: P must be 1 (bf16) or 2 (bf16x3)");
    RP_REQUIRE(!out_planes || (P_out >= 1 && P_out <= 2), RP_EINVAL, "rp_linear_tc: bad P_out");
    RP_REQUIRE(act >= RP_ACT_NONE && act <= RP_ACT_RELU, RP_EINVAL, "rp_linear_tc: bad act %d", act);
    RP_REQUIRE(rp::aligned16(A_planes) && rp::aligned16(W_planes), RP_EALIGN, "rp_linear_tc: operands must be 16-byte aligned");
    RP_GUARD(device);
    constexpr int BN = 192;
    CUtensorMap tmA, tmB;
    int rc = tc::make_planes_tmap(&tmA, A_planes, P, M, K, BM);
    if (rc) return rc;
    rc = tc::make_planes_tmap(&tmB, W_planes, P, N, K, BN);
    if (rc) return rc;
    EpiParams ep{nullptr, bias, nullptr, residual, 0, out_f32, static_cast<__nv_bfloat16*>(out_planes), P_out, act};
    Geom g{};
    g.M = M; g.N = N; g.K = K;
    int ntiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    cudaStream_t st = (cudaStream_t)stream;
    if (P == 1) return launch_gemm_tc<1, BN, false>(tmA, tmB, ep, g, ntiles, device, st, "rp_linear_tc");
    return launch_gemm_tc<2, BN, false>(tmA, tmB, ep, g, ntiles, device, st, "rp_linear_tc");
}

extern "C" int rp_conv2d_tc(const void* x_planes, const void* w_planes, const float* scale, const float* shift,
                            const float* res_pre, const float* res_post, int res_post_rows, float* out_f32,
                            void* out_planes, int n_img, int H, int W, int C, int O, int KH, int KW, int stride, int pad,
                            int P, int P_out, int act, int device, void* stream) {
    RP_REQUIRE(x_planes && w_planes && (out_f32 || out_planes), RP_EINVAL, "rp_conv2d_tc: null pointer");
    RP_REQUIRE(n_img > 0 && H > 0 && W > 0 && C > 0 && (C % 64) == 0 && KH > 0 && KW > 0 && (stride == 1 || stride == 2) && pad >= 0,
               RP_EINVAL, "rp_conv2d_tc: bad shape n=%d H=%d W=%d C=%d (C%%64==0) k=%dx%d s=%d p=%d", n_img, H, W, C, KH, KW, stride, pad);
    RP_REQUIRE(O == 64 || O == 128 || O == 192, RP_EINVAL, "rp_conv2d_tc: O must be 64, 128 or 192 (got %d)", O);
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_conv2d_tc: P must be 1 or 2");
    RP_REQUIRE(!out_planes || (P_out >= 1 && P_out <= 2), RP_EINVAL, "rp_conv2d_tc: bad P_out");
    RP_REQUIRE(act >= RP_ACT_NONE && act <= RP_ACT_RELU, RP_EINVAL, "rp_conv2d_tc: bad act %d", act);
    RP_REQUIRE(rp::aligned16(x_planes) && rp::aligned16(w_planes), RP_EALIGN, "rp_conv2d_tc: 16-byte alignment");
    Geom g{};
    g.Ho = (H + 2 * pad - KH) / stride + 1;
    g.Wo = (W + 2 * pad - KW) / stride + 1;
    RP_REQUIRE(g.Ho > 0 && g.Wo > 0 && g.Wo <= 128 && g.Wo * stride <= 256, RP_EINVAL, "rp_conv2d_tc: unsupported output width %d", g.Wo);
    g.R = 128 / g.Wo;
    if (g.R > g.Ho) g.R = g.Ho;
    if (g.R * stride > 256) g.R = 256 / stride;
    g.tiles_per_img = (g.Ho + g.R - 1) / g.R;
    g.C = C; g.KW = KW; g.stride = stride; g.pad = pad; g.cblocks = C / 64; g.taps = KH * KW;
    long long M = (long long)n_img * g.Ho * g.Wo;
    RP_REQUIRE(M < (1ll << 31), RP_EINVAL, "rp_conv2d_tc: too many output pixels");
    g.M = (int)M; g.N = O; g.K = KH * KW * C;
    RP_GUARD(device);
    // activation map: [plane][image][H][W][C], box = [1][1][R*stride][Wo*stride][64] traversed with the conv stride
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    RP_REQUIRE(fn != nullptr, RP_EINVAL, "rp_conv2d_tc: cuTensorMapEncodeTiled entry point unavailable");
    CUtensorMap tmA, tmB;
    {
        cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img, (cuuint64_t)P};
        cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                              (cuuint64_t)n_img * H * W * C * 2};
        cuuint32_t box[5] = {64, (cuuint32_t)(g.Wo * stride), (cuuint32_t)(g.R * stride), 1, 1};
        cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1, 1};
        CUresult r = fn(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x_planes), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RP_REQUIRE(r == CUDA_SUCCESS, RP_EINVAL, "rp_conv2d_tc: activation tensor map failed (CUresult %d)", (int)r);
    }
    int rc = tc::make_planes_tmap(&tmB, w_planes, P, O, g.K, O);
    if (rc) return rc;
    EpiParams ep{scale, shift, res_pre, res_post, res_post_rows, out_f32, static_cast<__nv_bfloat16*>(out_planes), P_out, act};
    int ntiles = n_img * g.tiles_per_img;
    cudaStream_t st = (cudaStream_t)stream;
    const char* what = "rp_conv2d_tc";
    // reductions longer than FOLD_KB K blocks (K > 768) sum their accumulator segments in registers (see gemm_tc_kernel);
    // RELPOSE_CONV_FOLD=0 keeps the single-chain accumulation for A/B measurements
    static const bool fold_enabled = !(getenv("RELPOSE_CONV_FOLD") && getenv("RELPOSE_CONV_FOLD")[0] == '0');
    const bool fold = fold_enabled && g.taps * g.cblocks > FOLD_KB;
#define RP_CONV_DISPATCH(PP, NN)                                                                              \
    return fold ? launch_gemm_tc<PP, NN, true, true>(tmA, tmB, ep, g, ntiles, device, st, what)               \
                : launch_gemm_tc<PP, NN, true, false>(tmA, tmB, ep, g, ntiles, device, st, what)
    if (P == 1) {
        if (O == 64) RP_CONV_DISPATCH(1, 64);
        if (O == 128) RP_CONV_DISPATCH(1, 128);
        RP_CONV_DISPATCH(1, 192);
    }
    if (O == 64) RP_CONV_DISPATCH(2, 64);
    if (O == 128) RP_CONV_DISPATCH(2, 128);
    RP_CONV_DISPATCH(2, 192);
#undef RP_CONV_DISPATCH
}
