// Fused transformer MLP half-block on sm_100a tensor cores:
//     out = x + fc2( GELU( fc1( LayerNorm(x) ) ) )          x, out float32 [M,192]; hidden width 768
// i.e. `x = x + self.mlp(self.norm2(x))` of Block.forward (vision_transformer.py:352-353, mlp.py:20-26) and
// `out = f + mlp(norm2(f))` of CrossBlock.forward (vision_transformer.py:295-296) in ONE launch.  The unfused
// sequence (layernorm_planes -> linear_tc+GELU -> linear_tc+residual) wrote the [M,768] hidden activation to HBM
// as bf16 planes and read it back (2 x 226 MB per layer at 64 pairs) plus a LayerNorm round trip; here the hidden
// activation never leaves the SM: fc1 accumulates a 64-column chunk in tensor memory, the epilogue warps apply
// bias + exact-erf GELU and hand the chunk back THROUGH TENSOR MEMORY (tcgen05.st, two bf16 per column) as the
// K-major A operand of fc2 (tcgen05.mma with A in TMEM), which accumulates the [128,192] output tile in tensor
// memory over the twelve chunks.
//
// Operands are split-bf16 planes as in gemm_tc.cu (P = 1 bf16, P = 2 "bf16x3" = fp32-class products).
//
// One persistent CTA per SM, 768 threads, one 128-row tile at a time:
//   warp 0      TMA producer of the fc1 weights: "units" [P][64 hidden rows][64 K] bf16 (128-byte swizzle), three
//               per chunk, into the fc1 ring.
//   warps 1, 2  fc1 issuers (even / odd chunks): 36 tcgen05.mma M128 x N64 x K16 per chunk (bf16x3), up to three
//               chunks ahead of the GELU (three fc1 accumulators in TMEM).
//   warp 3      fc2 issuer: 12 tcgen05.mma M128 x N192 x K16 per chunk, A = GELU chunk in TMEM, B = the chunk's
//               [192 x 64] slice of W2, which this warp also loads (its ring slot is free exactly when its own
//               MMAs retire).  Three issuer warps on three schedulers because a single issuer was the bottleneck.
//   warps 4-19  LayerNorm of the tile's rows straight into the swizzled A-operand planes (no HBM round trip; the
//               next tile's LayerNorm runs after the current tile's last GELU chunk), per-chunk GELU
//               (TMEM -> registers -> TMEM); out != x only: output epilogue (acc2 + bias + residual, transposed
//               through shared memory for coalesced float4 stores).
//   warps 20-23 in place (out == x) only: output warpgroup, x += acc2 + b2 as vector reductions at the memory side.
// Registers (setmaxnreg, per warpgroup): control 40, GELU 96, output 56 = the 768 x 80 the CTA is launched with.
// Tensor memory (512 columns): acc1[3] = 0..191, acc2 = 192..383, GELU chunk buffers [2][P][32] = 384..511.
//
// Measured history at 64 pairs, bf16x3 (profiles/r01_mlp_fused_history.md): unfused 214 (+27 LayerNorm) us ->
// 161 (first fused version, one issuer, fc1 one chunk ahead) -> 150 (fc1 two ahead) -> 131 (three issuers,
// N = 192 fc2) -> 116 (GELU chunk through TMEM, double buffered) -> 114 us.
//
// Where a tile's ~48 k cycles go (tools/probes/mlp_trace_probe.cu, clock64 at every barrier of one CTA;
// profiles/r02_mlp_trace_*.txt): twelve chunks at 3.0-3.3 k (tensor bound: 36 x 48 + 12 x 96 = 2 880 cycles of
// tcgen05.mma per chunk; the N = 64 fc1 products read 6 KB of shared memory per 32-cycle instruction, 128 B/clk ->
// 48 cycles) + ~9 k at the tile boundary, where the single 96 KB A-operand buffer serialises [last fc1 product
// retires] -> LayerNorm of the next rows (1.3-2.3 k to issue the loads, 5.6 k to reduce / normalise / split / store: 16
// warps in lockstep, ~570 dependent-ish instructions per thread) -> first fc1 chunk (1.7 k) with only three fc2
// products (3.4 k) queued on the tensor pipe.  Tried against that boundary in round 2, all measured, none kept:
// start offsets per CTA group (the phases are not bandwidth collisions between CTAs: +delay, no gain); TMA reduce-add of
// 2 KB boxes from the GELU warps (bulk stores of small boxes cost ~220 cycles each: tma_store_probe.cu, 9 B/clk);
// LayerNorm of the next tile by the output warpgroup into an L2-resident planes ring fetched by TMA
// (profiles/r02_mlp_ln_ring_experiment.patch: one warp per scheduler executes its ~200-instruction dependent chain per
// row pair at ~10 cycles per instruction, 30 k cycles per tile against the 16 warps' 7 k; the accumulator drain through
// one 16 KB box buffer took 11 k and stalled fc2).  What is left is a second A-operand buffer, i.e. 96 KB of shared
// memory this layout does not have.
#include <cstdlib>
#include "ln_rows.cuh"
#include "rows_ln_epilogue.cuh"
#include "tc_common.cuh"

namespace {

constexpr int D = 192, HID = 768, CH = 64, NCH = HID / CH;   // embed width, hidden width, hidden chunk
constexpr int BM = 128;
constexpr int KB = D / 64;                                   // K blocks of fc1 (3)
constexpr int NT = D / 64;                                   // 64-column thirds of the fc2 output (3)
constexpr int EPI_WARPS = 16;
constexpr int CTRL_WARPS = 4;                                 // TMA producer, two fc1 issuers, fc2 issuer
constexpr int OUT_WARPS = 4;                                  // in-place output warpgroup (one warp per TMEM lane quarter)
constexpr int NTHREADS = 32 * (CTRL_WARPS + EPI_WARPS + OUT_WARPS);
constexpr int TILE16K = BM * 64 * 2;                         // one [128 x 64] bf16 operand tile
constexpr int UNIT1 = 64 * 64 * 2;                           // one plane of a weight unit (8 KiB)
constexpr int NA1 = 3, LEAD = 2;                              // fc1 accumulators in TMEM; fc1 runs LEAD chunks ahead of fc2
constexpr int ACC1_COL = 0, ACC2_COL = NA1 * CH, H_COL = ACC2_COL + D, TMEM_COLS = 512;   // 192 + 192 + 128
constexpr int STG_LD = 16;

template <int P>
struct MCfg {
    static constexpr int NH = 2;                             // hidden-chunk buffers (A operand of fc2) in TENSOR memory
    static constexpr int H_STRIDE = P * (CH / 2);            // columns per buffer: 32 per plane (two bf16 per column)
    static constexpr int G1 = (P == 1) ? 2 : 1;              // fc1 weight ring: groups of three units (one chunk each)
    static constexpr int G2 = (P == 1) ? 2 : 1;              // fc2 weight ring: groups of three units (one chunk each)
    static constexpr int NU = 3 * (G1 + G2);                 // units in shared memory: fc1 ring first, then fc2 ring
    static constexpr int UNIT = P * UNIT1;
    static constexpr int OFF_XN = 0;                         // [P][KB] tiles of 16 KiB
    static constexpr int OFF_H = OFF_XN + P * KB * TILE16K;  // staging patches of the output epilogue (32 KiB)
    static constexpr int OFF_W = OFF_H + EPI_WARPS * 32 * STG_LD * 4;
    static constexpr int OFF_BAR = OFF_W + NU * UNIT;
    static constexpr int SMEM = OFF_BAR + 512 + 1024 /*align slack*/;
    static_assert(H_COL + NH * H_STRIDE <= TMEM_COLS, "tensor memory budget");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct MlpParams {
    const float* x;        // [M,192] input = residual
    const float* gamma;    // LayerNorm weight / bias [192] (AIN = false)
    const float* beta;
    const float* b1;       // [768]
    const float* b2;       // [192]
    float* out;            // [M,192]
    int M;
    float eps;
    // optional second output: LayerNorm(out) with the NEXT layer's norm weights as bf16 planes [P][M][192]
    __nv_bfloat16* ln_planes;
    const float* gamma2;
    const float* beta2;
    float eps2;
    int inplace;           // out == x: the output epilogue is a TMA reduce-add of acc + b2 into x (tmOut)
};

// Timeline probe (tools/probes/mlp_trace_probe.cu defines RP_MLP_TRACE and includes this file): lane 0 of every warp of
// one CTA appends (clock64 << 8 | tag) to a per-warp list.  Compiles to nothing in the library.
#ifdef RP_MLP_TRACE
constexpr int TRACE_EV = 2048;
__device__ long long g_mlp_trace[NTHREADS / 32][TRACE_EV];
#define MLP_TR(tag)                                                                                         \
    do {                                                                                                    \
        if (blockIdx.x == RP_MLP_TRACE_CTA && lane == 0 && trn < TRACE_EV) g_mlp_trace[warp][trn++] = (clock64() << 8) | (tag); \
    } while (0)
#else
#define MLP_TR(tag) do { } while (0)
#endif

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// AIN: the LayerNorm output (fc1's A operand) arrives as bf16 planes [P][M][192] by TMA -- written by the producer of x
// (rows_ln_epilogue.cuh) -- instead of being computed here.  ncu of the AIN = false kernel: the in-kernel LayerNorm held
// 30 % of the sixteen epilogue warps' time (lockstep loads + reductions while the GELU pipeline and, behind it, the
// tensor pipe stood still at every row-tile boundary).
template <int P, bool AIN>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_fused_tc_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                    const __grid_constant__ CUtensorMap tmXN, MlpParams prm) {
    using C = MCfg<P>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a SHARED pointer (LDS / STS)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* wfull = bars;                      // [NU]
    uint64_t* wempty = bars + C::NU;             // [NU]
    uint64_t* xn_full = bars + 2 * C::NU;
    uint64_t* xn_free = xn_full + 1;             // AIN: both fc1 issuers' last products of the tile have retired
    uint64_t* acc1_full = xn_full + 2;           // [NA1]
    uint64_t* acc1_empty = xn_full + 5;          // [NA1]
    uint64_t* h_full = xn_full + 8;              // [2]
    uint64_t* h_empty = xn_full + 10;            // [2]
    uint64_t* acc2_full = xn_full + 12;
    uint64_t* acc2_empty = xn_full + 13;
    uint64_t* turn = xn_full + 14;               // [2] hand-off between the two fc1 issuers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xn_full + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef RP_MLP_TRACE
    int trn = 0;
#endif
    const int M = prm.M;
    const int ntiles = (M + BM - 1) / BM;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmW1);
        tc::prefetch_tmap(&tmW2);
        if (AIN) tc::prefetch_tmap(&tmXN);
        for (int i = 0; i < C::NU; ++i) {
            tc::mbar_init(&wfull[i], 1);
            tc::mbar_init(&wempty[i], 1);
        }
        tc::mbar_init(xn_full, AIN ? 1 : EPI_WARPS); // one elected arrive per epilogue warp (512 arrives on one
                                                    // mbarrier serialise in the shared-memory atomic unit) | TMA bytes
        tc::mbar_init(xn_free, 2);
        for (int i = 0; i < NA1; ++i) {
            tc::mbar_init(&acc1_full[i], 1);
            tc::mbar_init(&acc1_empty[i], EPI_WARPS);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&h_full[i], EPI_WARPS);
            tc::mbar_init(&h_empty[i], 1);
        }
        tc::mbar_init(acc2_full, 1);
        tc::mbar_init(acc2_empty, prm.inplace ? OUT_WARPS : EPI_WARPS);
        tc::mbar_init(&turn[0], 1);
        tc::mbar_init(&turn[1], 1);
        tc::fence_barrier_init();
    }
    rp::pdl_launch_dependents();                  // the next kernel may start its prologue (common.cuh)
    if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    rp::pdl_wait();                               // the previous kernel has completed: its outputs are visible
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto xn_tile = [&](int p, int kb) { return smem + C::OFF_XN + (p * KB + kb) * TILE16K; };
    // ring units are grouped in threes, planes outermost inside a group: the three units of an fc2 chunk
    // (output rows 0..63, 64..127, 128..191) then form one contiguous [192 x 64] B tile per plane
    auto w_unit = [&](int u, int p) { return smem + C::OFF_W + (u / 3) * (3 * C::UNIT) + p * (3 * UNIT1) + (u % 3) * UNIT1; };

    // Every mbarrier of the weight rings has exactly ONE consumer: a parity wait is only unambiguous for a waiter
    // that observes every phase in order.  The fc1 ring (units 0 .. 3 G1-1) is consumed by the two fc1 issuers,
    // the fc2 ring (the rest) by the fc2 issuer.  Where both fc1 issuers alternate on the same units (G1 odd),
    // the `turn` hand-off below makes the later one wait until the earlier one has observed ITS fill.
    // Register budget: the CTA owns 768 x 80 registers; the control (40) and output (56) warpgroups give back exactly what
    // the four GELU warpgroups take (96): 128 x 40 + 128 x 56 + 512 x 96 = 768 x 80.  setmaxnreg is per warpgroup
    // (warps 0-3 | 4-19 | 20-23) and an increase can only draw on registers released inside the CTA.
    if (warp < CTRL_WARPS) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;" ::: "memory");
    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (convergent warp)
        // fc1 weights only: the fc2 issuer feeds its own ring, so a full fc2 ring (waiting for a GELU chunk) never
        // holds back the fc1 weights of the next chunks / the next tile behind it in a common load order
        uint32_t g1 = 0;                         // fc1 chunks loaded so far by this CTA
        auto load_unit = [&](int u, uint32_t fill, const CUtensorMap* tm, int c0, int c1) {
            tc::mbar_wait(&wempty[u], (fill & 1) ^ 1);
            if (tc::elect_one_sync()) {
                tc::mbar_expect_tx(&wfull[u], (uint32_t)C::UNIT);
#pragma unroll
                for (int p = 0; p < P; ++p) tc::tma_load_3d(w_unit(u, p), tm, &wfull[u], c0, c1, p);
            }
            __syncwarp();
        };
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int s = 0; s < NCH; ++s, ++g1) {                                        // W1[s*64.., kb*64..]
                const int u0 = 3 * (int)(g1 % C::G1);
                for (int kb = 0; kb < KB; ++kb) load_unit(u0 + kb, g1 / C::G1, &tmW1, kb * 64, s * CH);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------------------------ fc1 issuers (convergent warps)
        // Two warps share the fc1 chunks (even / odd): ncu showed ONE issuer warp spending 80 % of its time just
        // executing its own instruction stream (barrier polls, descriptor set-up, R2UR moves: ~90 instructions per
        // 12-MMA unit at ~11 cycles each next to four busy GELU warps) with the tensor pipe 65 % idle.  The three
        // issuers sit on different schedulers; MMAs of different issuers touch different accumulators.
        constexpr uint32_t idesc = tc::make_idesc_bf16(BM, 64);
        const uint32_t sel = (uint32_t)(warp - 1);
        uint32_t it = 0, n = 0;                 // tiles seen / chunks issued by this warp
        uint64_t dxn0[KB], dxn1[KB];
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            dxn0[kb] = tc::make_kmajor_sw128_desc(tc::smem_u32(xn_tile(0, kb)));
            dxn1[kb] = tc::make_kmajor_sw128_desc(tc::smem_u32(xn_tile(P - 1, kb)));
        }
        // AIN: the second issuer also fetches the tiles' LayerNorm planes -- six [128 x 64] boxes straight into the
        // swizzled A-operand tiles (rows beyond M: TMA zero fill, never stored).  It issues the last chunk of a tile, so
        // it is idle exactly while those products retire, and the weight producer keeps running ahead undisturbed.
        auto load_xn = [&](int tile) {
            if (tc::elect_one_sync()) {
                tc::mbar_expect_tx(xn_full, (uint32_t)(P * KB * TILE16K));
#pragma unroll
                for (int p = 0; p < P; ++p)
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) tc::tma_load_3d(xn_tile(p, kb), &tmXN, xn_full, kb * 64, tile * BM, p);
            }
            __syncwarp();
        };
        if (AIN && sel == 1 && (int)blockIdx.x < ntiles) load_xn(blockIdx.x);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            tc::mbar_wait(xn_full, it & 1);
            tc::tcgen05_fence_after();
            for (int s = (int)sel; s < NCH; s += 2, ++n) {
                // ---- fc1 of chunk s: acc1[b] = LN(x) . W1[s*64 .. s*64+64, :]^T       (NCH is even: g = 2n + sel)
                const uint32_t g = it * NCH + (uint32_t)s;
                const uint32_t b = g % NA1;
                const int u0 = 3 * (int)(g % C::G1);
                const uint32_t fpar = (g / C::G1) & 1;
                // the other issuer has observed the previous chunk's fill (see the ring comment above)
                if (sel == 0) { if (n > 0) tc::mbar_wait(&turn[0], (n - 1) & 1); }
                else tc::mbar_wait(&turn[1], n & 1);
                MLP_TR(1);
                tc::mbar_wait(&acc1_empty[b], ((g / NA1) & 1) ^ 1);
                tc::tcgen05_fence_after();
                MLP_TR(2);
                const uint32_t d = tmem_base + ACC1_COL + b * CH;
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    const int u = u0 + kb;
                    tc::mbar_wait(&wfull[u], fpar);
                    tc::tcgen05_fence_after();
                    const uint64_t w0 = tc::make_kmajor_sw128_desc(tc::smem_u32(w_unit(u, 0)));
                    const uint64_t w1 = tc::make_kmajor_sw128_desc(tc::smem_u32(w_unit(u, P - 1)));
                    if (tc::elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
                            if (P == 2) {            // smallest terms first (truncating fp32 accumulation)
                                tc::umma_bf16(d, dxn1[kb] + 2 * k, w0 + 2 * k, idesc, accum);
                                tc::umma_bf16(d, dxn0[kb] + 2 * k, w1 + 2 * k, idesc, 1u);
                                accum = 1u;
                            }
                            tc::umma_bf16(d, dxn0[kb] + 2 * k, w0 + 2 * k, idesc, accum);
                        }
                        tc::umma_commit(&wempty[u]);
                        if (kb == KB - 1) {
                            tc::umma_commit(&acc1_full[b]);
                            if (AIN && s + 2 >= NCH) tc::umma_commit(xn_free);   // this issuer's last chunk of the tile
                            tc::mbar_arrive(&turn[sel ^ 1]);
                        }
                    }
                    __syncwarp();
                }
                MLP_TR(3);
            }
            if (AIN && sel == 1 && tile + (int)gridDim.x < ntiles) {
                tc::mbar_wait(xn_free, it & 1);              // both issuers' last fc1 products of this tile have retired
                load_xn(tile + (int)gridDim.x);
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ fc2 issuer (convergent warp)
        // acc2[128 x 192] += GELU chunk [128 x 64] . W2[:, chunk]^T : the chunk's three ring units are one
        // contiguous [192 rows x 64 K] tile per plane, so every K step is ONE N = 192 instruction.
        // The warp also loads its own weights: a ring slot can only be refilled once the MMAs reading it have
        // retired, which this warp is the first to know.
        constexpr uint32_t idesc = tc::make_idesc_bf16(BM, D);
        uint32_t c2 = 0, it = 0;                // fc2 chunks / tiles issued so far by this CTA
        const uint32_t total = (uint32_t)((ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * NCH;
        auto load_chunk = [&](uint32_t c) {     // W2[:, (c % 12)*64 ..] -> ring group c % G2 (three 64-row units)
            const int u0 = 3 * C::G1 + 3 * (int)(c % C::G2);
            if (tc::elect_one_sync()) {
#pragma unroll
                for (int nn = 0; nn < NT; ++nn) {
                    tc::mbar_expect_tx(&wfull[u0 + nn], (uint32_t)C::UNIT);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        tc::tma_load_3d(w_unit(u0 + nn, p), &tmW2, &wfull[u0 + nn], (int)(c % NCH) * CH, nn * 64, p);
                }
            }
            __syncwarp();
        };
        for (uint32_t c = 0; c < (uint32_t)C::G2 && c < total; ++c) load_chunk(c);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            for (int j = 0; j < NCH; ++j, ++c2) {
                const uint32_t hb = c2 % C::NH;
                const int u = 3 * C::G1 + 3 * (int)(c2 % C::G2);
                const uint32_t fpar = (c2 / C::G2) & 1;
                MLP_TR(4);
                tc::mbar_wait(&h_full[hb], (c2 / C::NH) & 1);
                MLP_TR(5);
                if (j == 0) tc::mbar_wait(acc2_empty, (it & 1) ^ 1);
#pragma unroll
                for (int nn = 0; nn < NT; ++nn) tc::mbar_wait(&wfull[u + nn], fpar);
                tc::tcgen05_fence_after();
                const uint32_t d = tmem_base + ACC2_COL;
                const uint32_t ah0 = tmem_base + H_COL + hb * C::H_STRIDE;              // GELU chunk, plane 0 / plane 1
                const uint32_t ah1 = ah0 + (P - 1) * (CH / 2);
                const uint64_t w0 = tc::make_kmajor_sw128_desc(tc::smem_u32(w_unit(u, 0)));
                const uint64_t w1 = tc::make_kmajor_sw128_desc(tc::smem_u32(w_unit(u, P - 1)));
                if (tc::elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t accum = (j > 0 || k > 0) ? 1u : 0u;
                        if (P == 2) {
                            tc::umma_bf16_ts(d, ah1 + 8 * k, w0 + 2 * k, idesc, accum);
                            tc::umma_bf16_ts(d, ah0 + 8 * k, w1 + 2 * k, idesc, 1u);
                            accum = 1u;
                        }
                        tc::umma_bf16_ts(d, ah0 + 8 * k, w0 + 2 * k, idesc, accum);
                    }
#pragma unroll
                    for (int nn = 0; nn < NT; ++nn) tc::umma_commit(&wempty[u + nn]);
                    tc::umma_commit(&h_empty[hb]);
                    if (j == NCH - 1) tc::umma_commit(acc2_full);
                }
                __syncwarp();
                MLP_TR(6);
                if (c2 + C::G2 < total) {       // refill this slot with the chunk G2 ahead as soon as it is free
#pragma unroll
                    for (int nn = 0; nn < NT; ++nn) tc::mbar_wait(&wempty[u + nn], fpar);
                    load_chunk(c2 + C::G2);
                }
                MLP_TR(7);
            }
        }
    } else if (warp < CTRL_WARPS + EPI_WARPS) {
        // ------------------------------------------------------------------ LayerNorm / GELU / output warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 96;" ::: "memory");
        const int ew = warp - CTRL_WARPS;
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int part = ew >> 2;                        // which 16 of a chunk's 64 hidden columns / 48 of the 192 outputs
        const int r = q * 32 + lane;                     // the thread's row in TMEM-side work
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        float* stg = reinterpret_cast<float*>(smem + C::OFF_H) + ew * 32 * STG_LD;
        const int rr = lane >> 2, cq = lane & 3;         // coalesced side of the final epilogue
        uint32_t b1 = 0, ph1 = 0, c2 = 0, it = 0;

        // LayerNorm (eps 1e-6, vision_transformer.py:396) of the 8 rows this warp owns, written as the bf16 planes
        // of the fc1 A operand: three K-major SWIZZLE_128B tiles [128 rows x 64 columns] per plane.
        // LayerNorm of the 8 rows this warp owns, straight into the swizzled A-operand planes (ln_rows.cuh): two phases,
        // so that the rows' global-memory latency can hide behind a wait
        float vln[4][12];
        auto ln_load = [&](int tile) { lnrows::load8(prm.x, M, tile * BM, ew, lane, vln); };
        auto ln_finish = [&]() {
            lnrows::finish8<P>(vln, prm.gamma, prm.beta, prm.eps, smem + C::OFF_XN, ew, lane);
            tc::fence_proxy_async_smem();       // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(xn_full);
        };

        // ---- output of a finished tile: acc2 + b2 + x, transposed through warp-private staging patches for coalesced
        // float4 stores.  DEFERRED: it runs inside the NEXT tile's chunk loop (before its third GELU chunk), not right
        // after the last GELU chunk.  The in-order tensor pipe has the next tile's first fc1 products queued ahead of
        // this tile's last two fc2 products, so waiting for acc2_full at the tile boundary parked all sixteen warps
        // for ~5 k cycles per tile with the GELU pipeline empty behind them (ncu: the output epilogue held 21 % of the
        // epilogue warps' samples for 1.3 k instructions); two chunks later the accumulator is long complete and the
        // tensor pipe still has two fc1 chunks queued while these warps are busy here.
        auto out_epilogue = [&](int tile, uint32_t itp) {
            const int valid_rows = min(BM, M - tile * BM);
            float* staging = reinterpret_cast<float*>(smem + C::OFF_H);
            auto wait_acc = [&]() {
                tc::mbar_wait(acc2_full, itp & 1);
                tc::tcgen05_fence_after();
                MLP_TR(20);
            };
            auto release = [&]() {                           // acc2 is in registers: fc2 of the following tile may overwrite it
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(acc2_empty);
                MLP_TR(21);
            };
            if (prm.ln_planes)
                rowsln::epilogue<P>(t_lane + ACC2_COL, staging, q, part, lane, prm.b2, prm.x, prm.out, prm.ln_planes, prm.gamma2,
                                    prm.beta2, prm.eps2, tile * BM, valid_rows, (size_t)M * D, wait_acc, release);
            else
                rowsln::epilogue<0>(t_lane + ACC2_COL, staging, q, part, lane, prm.b2, prm.x, prm.out, nullptr, nullptr, nullptr,
                                    0.f, tile * BM, valid_rows, 0, wait_acc, release);
        };

        int prev_tile = -1;
        if (!AIN && (int)blockIdx.x < ntiles) { ln_load(blockIdx.x); ln_finish(); }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            // ---- twelve hidden chunks: acc1 -> + b1 -> GELU -> bf16 planes (A operand of fc2)
#pragma unroll 1
            for (int j = 0; j < NCH; ++j, ++c2) {
                const uint32_t b = b1;
                if (j == 2 && prev_tile >= 0 && !prm.inplace) { MLP_TR(17); out_epilogue(prev_tile, it - 1); MLP_TR(18); }
                if (j == 2) {
                    // pull the next tile's rows towards L2 now; its LayerNorm runs before this tile's last two chunks
                    const int nrow = (tile + (int)gridDim.x) * BM + ew * 8 + (lane >> 2);
                    if (nrow < M && (lane & 3) < 3)      // 8 rows x 768 B = 48 lines of 128 B: 24 lanes x 2
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.x + (size_t)nrow * D + (lane & 3) * 64));
                    if (nrow < M && (lane & 3) < 3)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(prm.x + (size_t)nrow * D + (lane & 3) * 64 + 32));
                }
                // the next tile's rows are requested before the last GELU chunk's wait (their latency hides behind it);
                // they are normalised after that chunk -- see below
                const bool ln_next = !AIN && j == NCH - 1 && tile + (int)gridDim.x < ntiles;
                float bias[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(prm.b1 + j * CH + part * 16 + 4 * i));
                    bias[4 * i] = t.x; bias[4 * i + 1] = t.y; bias[4 * i + 2] = t.z; bias[4 * i + 3] = t.w;
                }
                MLP_TR(10);
                tc::mbar_wait(&acc1_full[b], ph1);
                MLP_TR(11);
                if (++b1 == NA1) { b1 = 0; ph1 ^= 1; }
                tc::tcgen05_fence_after();
                uint32_t a[16];
                tc::tmem_ld_32x32b_x16(t_lane + ACC1_COL + b * CH + part * 16, a);
                tc::tmem_ld_wait();
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc1_empty[b]);  // fc1 of chunk j+2 may overwrite acc1[b]
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = tc::gelu_fast(__uint_as_float(a[i]) + bias[i]);
                const uint32_t hb = c2 % C::NH;
                MLP_TR(12);
                tc::mbar_wait(&h_empty[hb], ((c2 / C::NH) & 1) ^ 1);
                tc::tcgen05_fence_after();
                MLP_TR(13);
                // the chunk goes back to TENSOR memory as the K-major A operand of fc2 (two bf16 per column): no
                // shared-memory traffic, no generic->async proxy fence, and room for two buffers
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        w[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                        if (p + 1 < P) {
                            v[2 * i] -= __uint_as_float(w[i] << 16);
                            v[2 * i + 1] -= __uint_as_float(w[i] & 0xffff0000u);
                        }
                    }
                    tc::tmem_st_32x32b_x8(t_lane + H_COL + hb * C::H_STRIDE + p * (CH / 2) + part * 8, w);
                }
                tc::tmem_st_wait();
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&h_full[hb]);
                MLP_TR(14);
                if (ln_next) {
                    // LayerNorm of the NEXT tile right after the last GELU chunk: every fc1 product of this tile has
                    // been observed complete (chunk by chunk), so the LayerNorm planes are dead; while these warps are
                    // busy here the tensor pipe still has this tile's last three fc2 products queued (one more than
                    // when the LayerNorm ran before the last two chunks).
                    ln_load(tile + (int)gridDim.x);
                    MLP_TR(15);
                    ln_finish();
                    MLP_TR(16);
                }
            }
            prev_tile = tile;
        }
        if (prev_tile >= 0 && !prm.inplace) out_epilogue(prev_tile, it - 1);
    } else {
        // ------------------------------------------------------------------ output warpgroup (in place: out == x)
        // x += acc2 + b2 with vector reductions performed at the memory side: no residual loads, nothing for the GELU
        // warps to do at all.  The timeline probe (tools/probes/mlp_trace_probe.cu) showed the sixteen GELU warps spending
        // 6-9 k of the ~52 k cycles of a tile in the output epilogue (the SM's 32 B/clk store path: 98 KB per tile) with
        // the GELU pipeline, and behind it the fc2 issuer waiting for acc2, parked; these four warps drain the
        // accumulator while the GELU warps normalise the next tile's rows and release it long before fc2 needs it.
        // Thread = one row (TMEM lane), six passes of 32 columns; a rounding-exact match of the load / add / store
        // epilogue: (acc + b2) + x is one fp32 addition either way.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;" ::: "memory");
        if (prm.inplace) {
            const int q = warp & 3;
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int row = tile * BM + q * 32 + lane;
                float* dst = prm.out + (size_t)row * D;
                tc::mbar_wait(acc2_full, it & 1);
                tc::tcgen05_fence_after();
                MLP_TR(20);
#pragma unroll 1
                for (int pass = 0; pass < 6; ++pass) {
                    uint32_t a[2][16];
#pragma unroll
                    for (int g = 0; g < 2; ++g) tc::tmem_ld_32x32b_x16(t_lane + ACC2_COL + pass * 32 + g * 16, a[g]);
                    tc::tmem_ld_wait();
                    if (pass == 5) {                              // acc2 is in registers: fc2 of the following tile may overwrite it
                        tc::tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(acc2_empty);
                        MLP_TR(21);
                    }
                    if (row < M) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 sh = __ldg(reinterpret_cast<const float4*>(prm.b2 + pass * 32 + 4 * i));
                            const uint32_t* av = &a[i >> 2][4 * (i & 3)];
                            tc::red_add_v4_f32(dst + pass * 32 + 4 * i, __uint_as_float(av[0]) + sh.x, __uint_as_float(av[1]) + sh.y,
                                               __uint_as_float(av[2]) + sh.z, __uint_as_float(av[3]) + sh.w);
                        }
                    }
                }
                MLP_TR(18);
            }
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int P, bool AIN>
int launch_mlp(const CUtensorMap& tmW1, const CUtensorMap& tmW2, const CUtensorMap& tmXN, const MlpParams& prm, int device,
               cudaStream_t st) {
    using C = MCfg<P>;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fused_tc_kernel<P, AIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            rp::set_error("rp_mlp_tc: cudaFuncSetAttribute(%d): %s", C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    const int ntiles = (prm.M + BM - 1) / BM;
    const int grid = ntiles < rp::num_sms(device) ? ntiles : rp::num_sms(device);
    rp::launch(mlp_fused_tc_kernel<P, AIN>, dim3(grid), dim3(NTHREADS), (size_t)(C::SMEM), st, tmW1, tmW2, tmXN, prm);
    return rp::finish_launch("rp_mlp_tc");
}

}  // namespace

// xn_planes != null: LayerNorm(x) is supplied as bf16 planes [P][M][192] (ln_gamma / ln_beta unused).
// out_ln_planes != null: also writes LayerNorm(out; ln2_gamma, ln2_beta, eps2) as bf16 planes [P][M][192].
extern "C" int rp_mlp_tc_ex(const float* x, const void* xn_planes, const float* ln_gamma, const float* ln_beta, float eps,
                            const void* W1_planes, const float* b1, const void* W2_planes, const float* b2, float* out,
                            void* out_ln_planes, const float* ln2_gamma, const float* ln2_beta, float eps2, int M, int dim,
                            int hidden, int P, int device, void* stream) {
    RP_REQUIRE(x && (xn_planes || (ln_gamma && ln_beta)) && W1_planes && b1 && W2_planes && b2 && out && M > 0, RP_EINVAL,
               "rp_mlp_tc: null pointer or M <= 0");
    RP_REQUIRE(!out_ln_planes || (ln2_gamma && ln2_beta), RP_EINVAL, "rp_mlp_tc: LayerNorm planes requested without weights");
    RP_REQUIRE(dim == D && hidden == HID, RP_EINVAL, "rp_mlp_tc: built for dim=192, hidden=768 (got %d, %d)", dim, hidden);
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_mlp_tc: P must be 1 (bf16) or 2 (bf16x3)");
    RP_REQUIRE(rp::aligned16(x) && rp::aligned16(out) && rp::aligned16(W1_planes) && rp::aligned16(W2_planes) &&
                   rp::aligned16(b1) && rp::aligned16(b2) && rp::aligned16(ln_gamma) && rp::aligned16(ln_beta) &&
                   rp::aligned16(xn_planes) && rp::aligned16(out_ln_planes) && rp::aligned16(ln2_gamma) && rp::aligned16(ln2_beta),
               RP_EALIGN, "rp_mlp_tc: 16-byte alignment");
    RP_GUARD(device);
    CUtensorMap tmW1, tmW2, tmXN;
    int rc = tc::make_planes_tmap(&tmW1, W1_planes, P, HID, D, 64);      // [P][768][192], box 64 rows x 64 K
    if (rc) return rc;
    rc = tc::make_planes_tmap(&tmW2, W2_planes, P, D, HID, 64);          // [P][192][768]
    if (rc) return rc;
    tmXN = tmW1;
    if (xn_planes) {
        rc = tc::make_planes_tmap(&tmXN, xn_planes, P, M, D, BM);        // [P][M][192], box 128 rows x 64 K
        if (rc) return rc;
    }
    // out == x: the residual update happens at the memory side (the output warpgroup's vector reductions); RELPOSE_MLP_INPLACE=0
    // keeps the load / add / store epilogue on the GELU warps for A/B runs.  Not combined with the LayerNorm-planes output
    // (that one needs the sum in registers).
    static const bool inplace_ok = [] { const char* e = getenv("RELPOSE_MLP_INPLACE"); return !(e && e[0] == '0'); }();
    const int inplace = (out == x && !out_ln_planes && inplace_ok) ? 1 : 0;
    MlpParams prm{x, ln_gamma, ln_beta, b1, b2, out, M, eps, static_cast<__nv_bfloat16*>(out_ln_planes), ln2_gamma, ln2_beta, eps2, inplace};
    cudaStream_t st = (cudaStream_t)stream;
    if (xn_planes) {
        if (P == 1) return launch_mlp<1, true>(tmW1, tmW2, tmXN, prm, device, st);
        return launch_mlp<2, true>(tmW1, tmW2, tmXN, prm, device, st);
    }
    if (P == 1) return launch_mlp<1, false>(tmW1, tmW2, tmXN, prm, device, st);
    return launch_mlp<2, false>(tmW1, tmW2, tmXN, prm, device, st);
}

extern "C" int rp_mlp_tc(const float* x, const float* ln_gamma, const float* ln_beta, float eps, const void* W1_planes,
                         const float* b1, const void* W2_planes, const float* b2, float* out, int M, int dim, int hidden,
                         int P, int device, void* stream) {
    return rp_mlp_tc_ex(x, nullptr, ln_gamma, ln_beta, eps, W1_planes, b1, W2_planes, b2, out, nullptr, nullptr, nullptr, 0.f, M,
                        dim, hidden, P, device, stream);
}
