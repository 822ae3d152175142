// LayerNorm fused into the QKV projection on sm_100a tensor cores:
//     out = LayerNorm(x) W^T + b            x float32 [M,192], W [N,192] (N = 576 for qkv), out as bf16 planes
// i.e. `self.qkv(self.norm1(x))` of Block.forward / Attention.forward (vision_transformer.py:350,323) and of
// CrossBlock.forward / CrossAttention.forward (vision_transformer.py:288-289,191-194) in ONE launch.  The unfused
// pair (layernorm_planes -> linear_tc) wrote the normalised activations to HBM as bf16 planes and read them back
// (2 x 57 MB per layer at 64 pairs, one extra launch); here the LayerNorm output goes straight into the swizzled
// shared-memory A operand and stays there for all N / 192 column tiles of the row tile.
//
// Operands are split-bf16 planes as in gemm_tc.cu (P = 1 bf16, P = 2 "bf16x3").  One persistent CTA per SM,
// 576 threads, one 128-row tile at a time:
//   warp 0      TMA producer: weight units [P][192 rows][64 K] (128-byte swizzle) into a ring
//   warp 1      MMA issuer: per column tile 3 K blocks x 4 K steps x (3 | 1) tcgen05.mma M128 x N192 x K16 into one of
//               two TMEM accumulators (the epilogue of column tile i overlaps the MMAs of tile i+1)
//   warps 2-17  LayerNorm of the tile's rows into the A-operand planes; epilogue.  Planes-only outputs (the QKV case)
//               take the TMA-store epilogue: a thread keeps its TMEM row, adds the bias, splits the planes and writes
//               them into the swizzled box layout of a [128 x 64] tile, one thread issues a bulk tensor store per
//               64-column round (85 us instead of 105 us with the generic epilogue: tcgen05.ld -> transpose through
//               a warp-private shared-memory patch -> float32 and/or planes with per-lane stores).
//               The next row tile's LayerNorm runs as soon as the last column tile's MMAs have retired, before that
//               tile's epilogue, so the tensor pipe does not drain at row-tile boundaries.
#include <stdlib.h>

#include "ln_rows.cuh"
#include "rows_ln_epilogue.cuh"
#include "tc_common.cuh"

namespace {

constexpr int D = 192, BM = 128, BN = 192, KB = D / 64;
constexpr int EPI_WARPS = 16;
constexpr int NTHREADS = 32 * (2 + EPI_WARPS);
constexpr int TILE16K = BM * 64 * 2;                  // one [128 x 64] bf16 A tile
constexpr int BTILE = BN * 64 * 2;                    // one plane of a weight unit: [192 x 64] bf16 = 24 KiB
constexpr int TMEM_COLS = 512, ACC_STRIDE = 256;
constexpr int STG_LD = 16;
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;

template <int P>
struct LCfg {
    static constexpr int NS = (P == 1) ? 4 : 2;       // weight ring depth
    static constexpr int UNIT = P * BTILE;
    static constexpr int OFF_XN = 0;                  // [P][KB] tiles of 16 KiB
    static constexpr int OFF_STG = OFF_XN + P * KB * TILE16K;
    static constexpr int OFF_W = OFF_STG + STG_BYTES;
    static constexpr int OFF_BAR = OFF_W + NS * UNIT;
    static constexpr int SMEM = OFF_BAR + 256 + 1024 /*align slack*/;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct LnLinParams {
    const float* x;        // [M,192] (AIN = false: LayerNorm input)
    const float* gamma;
    const float* beta;
    const float* bias;     // [N] or null
    float* out_f32;        // [M,N] or null
    __nv_bfloat16* out_planes;   // [P_out][M][N] or null
    int p_out;
    int M, N;
    float eps;
    // EPI == 3 (N == 192): out_f32 = acc + bias + residual, and optionally LayerNorm(out_f32; gamma2, beta2) as planes
    const float* residual; // [M,192]
    __nv_bfloat16* ln_planes;    // [P][M][192] or null
    const float* gamma2;
    const float* beta2;
    float eps2;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// TMAOUT: planes-only output through TMA stores.  The generic epilogue (transpose through shared memory, per-lane
// 8-byte global stores with 64-bit address arithmetic, run-time plane loop) executes ~20 instructions per output
// element in long dependent chains (ncu: 0.24 IPC per scheduler, 51 k cycles per row tile against 10 k cycles of
// tensor work).  Here a thread keeps its TMEM row: bias, plane split, two 16-byte shared-memory stores per plane
// into the 128-byte-swizzled box layout of a [128 x 64] bf16 tile, and ONE thread hands the tile to the TMA
// (cp.async.bulk.tensor store; rows / columns outside the tensor are clipped by the hardware).
// AIN: the A operand arrives as bf16 planes [P][M][192] by TMA (six [128 x 64] boxes per row tile) instead of being
// LayerNorm-ed here: the producer of x wrote them (rows_ln_epilogue.cuh).  EPI == 3 (N == 192, one column tile): the
// residual-stream epilogue of rows_ln_epilogue.cuh -- out_f32 = acc + bias + residual and the next LayerNorm's planes.
template <int P, int EPI, bool AIN>
__global__ void __launch_bounds__(NTHREADS, 1)
ln_linear_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ CUtensorMap tmXN, LnLinParams prm) {
    using C = LCfg<P>;
    constexpr bool TMAOUT = (EPI == 1);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a SHARED pointer (LDS / STS)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* wfull = bars;                      // [NS]
    uint64_t* wempty = bars + C::NS;             // [NS]
    uint64_t* xn_full = bars + 2 * C::NS;
    uint64_t* tfull = xn_full + 1;               // [2]
    uint64_t* tempty = xn_full + 3;              // [2]
    uint64_t* xn_free = xn_full + 5;             // AIN: the row tile's last MMAs have retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xn_full + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = prm.M, N = prm.N;
    const int ntiles = (M + BM - 1) / BM;
    const int NT = (N + BN - 1) / BN;            // column tiles per row tile

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmW);
        if (TMAOUT) tc::prefetch_tmap(&tmOut);
        if (AIN) tc::prefetch_tmap(&tmXN);
        for (int i = 0; i < C::NS; ++i) {
            tc::mbar_init(&wfull[i], 1);
            tc::mbar_init(&wempty[i], 1);
        }
        tc::mbar_init(xn_full, AIN ? 1 : EPI_WARPS);
        tc::mbar_init(xn_free, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&tfull[i], 1);
            tc::mbar_init(&tempty[i], EPI_WARPS);
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

    auto xn_tile = [&](int p, int kb) { return smem + C::OFF_XN + (p * KB + kb) * TILE16K; };
    auto w_unit = [&](int s, int p) { return smem + C::OFF_W + s * C::UNIT + p * BTILE; };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (convergent warp)
        int s = 0, ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int nt = 0; nt < NT; ++nt) {
                for (int kb = 0; kb < KB; ++kb) {
                    tc::mbar_wait(&wempty[s], ph ^ 1);
                    if (tc::elect_one_sync()) {
                        tc::mbar_expect_tx(&wfull[s], (uint32_t)C::UNIT);
#pragma unroll
                        for (int p = 0; p < P; ++p) tc::tma_load_3d(w_unit(s, p), &tmW, &wfull[s], kb * 64, nt * BN, p);
                    }
                    __syncwarp();
                    if (++s == C::NS) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (convergent warp)
        constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
        int s = 0, ph = 0;
        uint32_t c = 0, it = 0;                 // column tiles / row tiles issued so far by this CTA
        uint64_t dxn0[KB], dxn1[KB];
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
            dxn0[kb] = tc::make_kmajor_sw128_desc(tc::smem_u32(xn_tile(0, kb)));
            dxn1[kb] = tc::make_kmajor_sw128_desc(tc::smem_u32(xn_tile(P - 1, kb)));
        }
        // AIN: this warp also fetches the row tiles' A planes (six [128 x 64] boxes into the swizzled A tiles; rows beyond
        // M are TMA zero fill): it is idle exactly while the tile's last products retire, and the weight producer keeps
        // running ahead undisturbed.
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
        if (AIN && (int)blockIdx.x < ntiles) load_xn(blockIdx.x);
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            tc::mbar_wait(xn_full, it & 1);
            tc::tcgen05_fence_after();
            for (int nt = 0; nt < NT; ++nt, ++c) {
                const uint32_t acc = c & 1;
                tc::mbar_wait(&tempty[acc], ((c >> 1) & 1) ^ 1);
                tc::tcgen05_fence_after();
                const uint32_t d = tmem_base + acc * ACC_STRIDE;
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    tc::mbar_wait(&wfull[s], ph);
                    tc::tcgen05_fence_after();
                    const uint64_t w0 = tc::make_kmajor_sw128_desc(tc::smem_u32(w_unit(s, 0)));
                    const uint64_t w1 = tc::make_kmajor_sw128_desc(tc::smem_u32(w_unit(s, P - 1)));
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
                        tc::umma_commit(&wempty[s]);
                        if (kb == KB - 1) tc::umma_commit(&tfull[acc]);
                        if (AIN && kb == KB - 1 && nt == NT - 1) tc::umma_commit(xn_free);
                    }
                    __syncwarp();
                    if (++s == C::NS) { s = 0; ph ^= 1; }
                }
            }
            if (AIN && tile + (int)gridDim.x < ntiles) {
                tc::mbar_wait(xn_free, it & 1);              // the tile's last products have retired
                load_xn(tile + (int)gridDim.x);
            }
        }
    } else {
        // ------------------------------------------------------------------ LayerNorm / epilogue warps
        const int ew = warp - 2;
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int part = ew >> 2;                        // 48-column slab of the 192-column tile
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        float* stg = reinterpret_cast<float*>(smem + C::OFF_STG) + ew * 32 * STG_LD;
        const int rr = lane >> 2, cq = lane & 3;
        const bool vec4 = (N % 4) == 0;
        uint32_t c = 0;

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

        if (!AIN && (int)blockIdx.x < ntiles) { ln_load(blockIdx.x); ln_finish(); }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int row_base = tile * BM;
            const int valid_rows = min(BM, M - row_base);
            for (int nt = 0; nt < NT; ++nt, ++c) {
                const uint32_t acc = c & 1;
                const int n0 = nt * BN;
                // every MMA of this row tile has retired once its last column tile is complete: the LayerNorm planes
                // are dead, write the next row tile's before draining this accumulator (its rows were requested
                // before the wait)
                const bool ln_next = (!AIN && nt == NT - 1 && tile + (int)gridDim.x < ntiles);
                if (ln_next) ln_load(tile + (int)gridDim.x);
                tc::mbar_wait(&tfull[acc], (c >> 1) & 1);
                tc::tcgen05_fence_after();
                if (ln_next) ln_finish();
                const uint32_t t_row = t_lane + acc * ACC_STRIDE;
                if constexpr (EPI == 3) {
                    float* staging = reinterpret_cast<float*>(smem + C::OFF_STG);
                    auto wait_acc = [&]() {};                        // already complete (tfull above)
                    auto release = [&]() {
                        tc::tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&tempty[acc]);
                    };
                    if (prm.ln_planes)
                        rowsln::epilogue<P>(t_row, staging, q, part, lane, prm.bias, prm.residual, prm.out_f32, prm.ln_planes,
                                            prm.gamma2, prm.beta2, prm.eps2, row_base, valid_rows, (size_t)M * D, wait_acc, release);
                    else
                        rowsln::epilogue<0>(t_row, staging, q, part, lane, prm.bias, prm.residual, prm.out_f32, nullptr, nullptr,
                                            nullptr, 0.f, row_base, valid_rows, 0, wait_acc, release);
                    continue;                                        // tempty released inside
                } else if constexpr (TMAOUT) {
                    uint8_t* stg8 = smem + C::OFF_STG;               // [P][128 rows][128 B], 128-byte swizzle
                    const int r = q * 32 + lane;
                    const uint32_t row_off = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128;
                    const uint32_t sw = (uint32_t)(r & 7);
                    // one store pipeline per TMEM lane quarter: its four warps (the 64-column round split four ways) own
                    // rows 32 q .. 32 q + 31 = a [32 x 128 B] slice of each plane's staging tile, synchronise among
                    // themselves (128-thread named barriers; they sit on one scheduler) and issue their own bulk stores,
                    // so a quarter waiting for its previous store's read does not hold the other three
                    const bool leader = (part == 0 && lane == 0);
                    const int qbar = 1 + q;
#pragma unroll 1
                    for (int rd = 0; rd < BN / 64; ++rd) {             // 64-column rounds of the 192-column tile
                        const int c0 = n0 + rd * 64;
                        if (c0 >= N) break;                          // CTA-uniform
                        uint32_t a[16];
                        tc::tmem_ld_32x32b_x16(t_row + rd * 64 + part * 16, a);
                        float bia[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int cb = c0 + part * 16 + 4 * i;
                            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (prm.bias && cb < N) t4 = __ldg(reinterpret_cast<const float4*>(prm.bias + cb));   // N % 8 == 0
                            bia[4 * i] = t4.x; bia[4 * i + 1] = t4.y; bia[4 * i + 2] = t4.z; bia[4 * i + 3] = t4.w;
                        }
                        tc::tmem_ld_wait();
                        float v[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(a[i]) + bia[i];
                        // the previous round's TMA store has finished reading the staging tile
                        if (leader) tc::tma_store_wait_read();
                        asm volatile("bar.sync %0, 128;" ::"r"(qbar) : "memory");
#pragma unroll
                        for (int p = 0; p < P; ++p) {
#pragma unroll
                            for (int hc = 0; hc < 2; ++hc) {
                                uint4 w;
                                w.x = pack_bf16x2(v[8 * hc + 0], v[8 * hc + 1]);
                                w.y = pack_bf16x2(v[8 * hc + 2], v[8 * hc + 3]);
                                w.z = pack_bf16x2(v[8 * hc + 4], v[8 * hc + 5]);
                                w.w = pack_bf16x2(v[8 * hc + 6], v[8 * hc + 7]);
                                const uint32_t cchunk = (uint32_t)(part * 2 + hc);
                                *reinterpret_cast<uint4*>(stg8 + p * TILE16K + row_off + ((cchunk ^ sw) << 4)) = w;
                                if (p + 1 < P) {
                                    v[8 * hc + 0] -= __uint_as_float(w.x << 16); v[8 * hc + 1] -= __uint_as_float(w.x & 0xffff0000u);
                                    v[8 * hc + 2] -= __uint_as_float(w.y << 16); v[8 * hc + 3] -= __uint_as_float(w.y & 0xffff0000u);
                                    v[8 * hc + 4] -= __uint_as_float(w.z << 16); v[8 * hc + 5] -= __uint_as_float(w.z & 0xffff0000u);
                                    v[8 * hc + 6] -= __uint_as_float(w.w << 16); v[8 * hc + 7] -= __uint_as_float(w.w & 0xffff0000u);
                                }
                            }
                        }
                        tc::fence_proxy_async_smem();                // generic-proxy writes -> visible to the TMA
                        asm volatile("bar.sync %0, 128;" ::"r"(qbar) : "memory");
                        if (leader) {
#pragma unroll
                            for (int p = 0; p < P; ++p)
                                tc::tma_store_3d(&tmOut, stg8 + p * TILE16K + q * 4096, c0, row_base + 32 * q, p);
                            tc::tma_store_commit();
                        }
                    }
                } else {
#pragma unroll 1
                for (int ci = 0; ci < 3; ++ci) {
                    const int c0 = part * 48 + ci * 16;
                    if (n0 + c0 >= N) break;                 // warp-uniform
                    const int col = n0 + c0 + cq * 4;
                    const bool colv = vec4 && col < N;
                    float4 sh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (colv && prm.bias) sh = __ldg(reinterpret_cast<const float4*>(prm.bias + col));
                    uint32_t r[16];
                    tc::tmem_ld_32x32b_x16(t_row + c0, r);
                    tc::tmem_ld_wait();
                    __syncwarp();                            // the previous chunk's readers are done with the patch
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(stg + lane * STG_LD + ((j ^ ((lane >> 1) & 3)) << 2)) =
                            make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int rl = t * 8 + rr;
                        const int rt = q * 32 + rl;
                        if (rt >= valid_rows) continue;
                        const float4 a = *reinterpret_cast<const float4*>(stg + rl * STG_LD + ((cq ^ ((rl >> 1) & 3)) << 2));
                        float vv[4] = {a.x + sh.x, a.y + sh.y, a.z + sh.z, a.w + sh.w};
                        const size_t o = (size_t)(row_base + rt) * N + col;
                        if (colv) {
                            if (prm.out_f32) *reinterpret_cast<float4*>(prm.out_f32 + o) = make_float4(vv[0], vv[1], vv[2], vv[3]);
                            if (prm.out_planes) {
                                for (int p = 0; p < prm.p_out; ++p) {
                                    uint2 w;
                                    w.x = pack_bf16x2(vv[0], vv[1]);
                                    w.y = pack_bf16x2(vv[2], vv[3]);
                                    *reinterpret_cast<uint2*>(prm.out_planes + (size_t)p * M * N + o) = w;
                                    vv[0] -= __uint_as_float(w.x << 16); vv[1] -= __uint_as_float(w.x & 0xffff0000u);
                                    vv[2] -= __uint_as_float(w.y << 16); vv[3] -= __uint_as_float(w.y & 0xffff0000u);
                                }
                            }
                        } else {
                            const int colx = n0 + c0 + cq * 4;
                            for (int j = 0; j < 4 && colx + j < N; ++j) {
                                float xj = (&a.x)[j] + (prm.bias ? prm.bias[colx + j] : 0.f);
                                const size_t oj = (size_t)(row_base + rt) * N + colx + j;
                                if (prm.out_f32) prm.out_f32[oj] = xj;
                                if (prm.out_planes)
                                    for (int p = 0; p < prm.p_out; ++p) {
                                        const __nv_bfloat16 h = __float2bfloat16_rn(xj);
                                        prm.out_planes[(size_t)p * M * N + oj] = h;
                                        xj -= __bfloat162float(h);
                                    }
                            }
                        }
                    }
                }
                }
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[acc]);
            }
        }
    }

    if (TMAOUT && warp >= 2 && warp < 6 && lane == 0) tc::tma_store_wait_all();   // the quarter leaders' bulk stores are complete
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int P, int EPI, bool AIN>
int launch_ln_linear(const CUtensorMap& tmW, const CUtensorMap& tmOut, const CUtensorMap& tmXN, const LnLinParams& prm, int device,
                     cudaStream_t st) {
    using C = LCfg<P>;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(ln_linear_tc_kernel<P, EPI, AIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            rp::set_error("rp_ln_linear_tc: cudaFuncSetAttribute(%d): %s", C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    const int ntiles = (prm.M + BM - 1) / BM;
    const int grid = ntiles < rp::num_sms(device) ? ntiles : rp::num_sms(device);
    rp::launch(ln_linear_tc_kernel<P, EPI, AIN>, dim3(grid), dim3(NTHREADS), (size_t)(C::SMEM), st, tmW, tmOut, tmXN, prm);
    return rp::finish_launch("rp_ln_linear_tc");
}

template <int P, bool AIN>
int dispatch_ln_linear(const void* xn_planes, const void* W_planes, LnLinParams& prm, int device, cudaStream_t st) {
    const int M = prm.M, N = prm.N;
    CUtensorMap tmW, tmXN, tmOut;
    int rc = tc::make_planes_tmap(&tmW, W_planes, P, N, D, BN);           // [P][N][192], box 192 rows x 64 K
    if (rc) return rc;
    tmXN = tmW; tmOut = tmW;
    if (AIN) {
        rc = tc::make_planes_tmap(&tmXN, xn_planes, P, M, D, BM);        // [P][M][192], box 128 rows x 64 K
        if (rc) return rc;
    }
    if (prm.residual) return launch_ln_linear<P, 3, AIN>(tmW, tmOut, tmXN, prm, device, st);
    // planes-only output with as many planes as the operands and 16-byte row pitch: TMA-store epilogue
    if (!prm.out_f32 && prm.out_planes && prm.p_out == P && (N % 8) == 0) {
        rc = tc::make_planes_tmap(&tmOut, prm.out_planes, P, M, N, 32);  // [P][M][N], box 32 rows x 64 columns (one lane quarter)
        if (rc) return rc;
        return launch_ln_linear<P, 1, AIN>(tmW, tmOut, tmXN, prm, device, st);
    }
    return launch_ln_linear<P, 0, AIN>(tmW, tmOut, tmXN, prm, device, st);
}

}  // namespace

// Generalised entry point.  A operand: xn_planes != null -> bf16 planes [P][M][192] supplied by the producer of x (the
// LayerNorm arguments are unused), else LayerNorm(x) computed in the kernel.  residual != null (N must be 192):
// out_f32 = A W^T + bias + residual, and out_ln_planes != null additionally receives LayerNorm(out_f32; ln2_*) as bf16
// planes [P][M][192] -- the attention projection with the Block's norm2 folded in (vision_transformer.py:351-352).
extern "C" int rp_ln_linear_tc_ex(const float* x, const void* xn_planes, const float* ln_gamma, const float* ln_beta, float eps,
                                  const void* W_planes, const float* bias, const float* residual, float* out_f32, void* out_planes,
                                  void* out_ln_planes, const float* ln2_gamma, const float* ln2_beta, float eps2, int M, int N,
                                  int K, int P, int P_out, int device, void* stream) {
    RP_REQUIRE((xn_planes || (x && ln_gamma && ln_beta)) && W_planes && (out_f32 || out_planes) && M > 0 && N > 0, RP_EINVAL,
               "rp_ln_linear_tc: null pointer or empty shape");
    RP_REQUIRE(K == D, RP_EINVAL, "rp_ln_linear_tc: built for K = 192 (got %d)", K);
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_ln_linear_tc: P must be 1 (bf16) or 2 (bf16x3)");
    RP_REQUIRE(!out_planes || (P_out >= 1 && P_out <= 2), RP_EINVAL, "rp_ln_linear_tc: bad P_out");
    RP_REQUIRE(!residual || (N == D && out_f32 && !out_planes && bias), RP_EINVAL,
               "rp_ln_linear_tc: the residual epilogue needs N = 192, a bias, a float32 output and no plane output");
    RP_REQUIRE(!out_ln_planes || (residual && ln2_gamma && ln2_beta), RP_EINVAL,
               "rp_ln_linear_tc: LayerNorm planes need the residual epilogue and the norm weights");
    RP_REQUIRE(rp::aligned16(x) && rp::aligned16(xn_planes) && rp::aligned16(W_planes) && rp::aligned16(ln_gamma) &&
                   rp::aligned16(ln_beta) && rp::aligned16(bias) && rp::aligned16(out_f32) && rp::aligned16(out_planes) &&
                   rp::aligned16(residual) && rp::aligned16(out_ln_planes) && rp::aligned16(ln2_gamma) && rp::aligned16(ln2_beta),
               RP_EALIGN, "rp_ln_linear_tc: 16-byte alignment");
    RP_GUARD(device);
    LnLinParams prm{x, ln_gamma, ln_beta, bias, out_f32, static_cast<__nv_bfloat16*>(out_planes), P_out, M, N, eps,
                    residual, static_cast<__nv_bfloat16*>(out_ln_planes), ln2_gamma, ln2_beta, eps2};
    cudaStream_t st = (cudaStream_t)stream;
    if (xn_planes) {
        if (P == 1) return dispatch_ln_linear<1, true>(xn_planes, W_planes, prm, device, st);
        return dispatch_ln_linear<2, true>(xn_planes, W_planes, prm, device, st);
    }
    if (P == 1) return dispatch_ln_linear<1, false>(xn_planes, W_planes, prm, device, st);
    return dispatch_ln_linear<2, false>(xn_planes, W_planes, prm, device, st);
}

extern "C" int rp_ln_linear_tc(const float* x, const float* ln_gamma, const float* ln_beta, float eps, const void* W_planes,
                               const float* bias, float* out_f32, void* out_planes, int M, int N, int K, int P, int P_out,
                               int device, void* stream) {
    RP_REQUIRE(x && ln_gamma && ln_beta, RP_EINVAL, "rp_ln_linear_tc: null pointer");
    return rp_ln_linear_tc_ex(x, nullptr, ln_gamma, ln_beta, eps, W_planes, bias, nullptr, out_f32, out_planes, nullptr, nullptr,
                              nullptr, 0.f, M, N, K, P, P_out, device, stream);
}
