// Fused self-attention (A5: vision_transformer.py:321-329) on tcgen05 tensor cores, flash style: nothing of
// size 576x576 ever leaves the SM.  softmax(q k^T * 0.125) v for every (image, head), N = 576 tokens, d = 64.
//
// Operands are the bf16 "planes" the QKV GEMM epilogue wrote (gemm_tc.cu): P = 1 plain bf16, P = 2 split bf16
// (x = x0 + x1; every product is evaluated as a0 b0 + a0 b1 + a1 b0 into one fp32 TMEM accumulator, fp32 class).
//
// Work item = (image, head, 128-query tile); 576 = 4.5 tiles, the 5th tile is half empty (TMA zero fill, rows
// never stored; the softmax warps of its two empty lane quarters keep the barrier protocol and skip the arithmetic).  Keys/values stream in 6 blocks of 96.  Persistent CTAs (one or two per SM), 384 threads:
//   warp 0      TMA producer: Q tile (once per item), K and V blocks through two independent 2-stage rings
//   warp 1      S issuer + TMEM owner:  S_j = Q K_j^T   M=128 N=96 K=64, A,B K-major SWIZZLE_128B -> TMEM S[j&1];
//               runs up to two key blocks ahead of the softmax (s_free mbarriers)
//   warp 2      O issuer:  O_j = P_j V_j   M=128 N=64 K=96, A = P_j read from TENSOR memory (K-major, two bf16 per
//               column), B = V_j as loaded by TMA ([key][d] rows = MN-major SWIZZLE_128B) -> TMEM O
//   warps 4-11  softmax: NSPLIT = 2 threads per query row (TMEM lane).  tcgen05.ld S_j, online max / sum,
//               p = 2^((s-m) c), P_j re-split into bf16 planes and written back to tensor memory (tcgen05.st),
//               mbarrier arrive.  The un-normalised output is carried in registers:
//               o = o * alpha + O_j (tcgen05.ld of the per-block product), so TMEM is never rescaled.
// S is double buffered in TMEM, O and P are single buffers: P_j may only be overwritten / O_j only be replaced
// after the softmax threads have seen pv_done(j-1), and the O issuer starts PV_j only after p_ready(j), which
// each softmax warp signals after reading O_{j-1}.
// Measured at 64 pairs, bf16x3: 140 us (one issuer warp, P through shared memory) -> 127 us; four threads per
// row (NSPLIT = 4, 93 registers) measured 132 us, i.e. the softmax chain is not short of warps.
//
// CPS = 2 (two CTAs per SM): the chain S_j -> softmax -> P_j -> PV_j -> fold of ONE query tile is latency bound (the
// ncu capture of the CPS = 1 kernel shows the tensor pipe 41 % active and the softmax warps at ~0.3 IPC), so two
// co-resident CTAs -- two independent tiles per SM -- hide each other's gaps.  To fit twice: the V ring drops to one
// stage (V_{j+1} can only be used after softmax j+1 anyway), tensor memory drops to 256 columns by writing P_j over
// S_j (the softmax threads have S_j in registers by then; the buffer is released by the PV product's commit instead
// of by the softmax threads), and registers are capped at 88 per thread.
#include "tc_common.cuh"

namespace {

constexpr int NTOK = RP_NTOK, HD = RP_HDIM, EMB = RP_EMBED, HEADS = RP_HEADS;
constexpr int BM = 128, BKV = 96, NBLK = NTOK / BKV, QTILES = (NTOK + BM - 1) / BM;   // 6 key blocks, 5 query tiles
constexpr int Q_TILE = BM * 128;          // bytes of one [128 x 64] bf16 tile
constexpr int KV_TILE = BKV * 128;        // bytes of one [96 x 64] bf16 tile
constexpr int P_SUB = BM * 128;           // P_j is [128 x 96] = one full and one half-used 64-wide K-major sub-tile
constexpr int K_STAGES = 2;
constexpr int ATT_CTRL = 4;              // warp 0 TMA, warp 1 S = QK^T issuer, warp 2 O = PV issuer, warp 3 idle (fills the warpgroup: setmaxnreg)
constexpr int NSPLIT = 2;                // softmax threads per query row (4 NSPLIT warps: NSPLIT per TMEM lane quarter)
constexpr int ATT_THREADS = 32 * (ATT_CTRL + 4 * NSPLIT);
constexpr int HB = BKV / NSPLIT, HO = HD / NSPLIT;  // key columns / output columns per softmax thread
static_assert((NSPLIT == 2 || NSPLIT == 4) && HB % 8 == 0 && HO % 16 == 0, "softmax split");
// CPS = 1: S[0] 0..95, S[1] 96..191, O 192..255, P planes 256..351 (two bf16 per column); 512 columns allocated
// CPS = 2: S[0] 0..95, S[1] 96..191, O 192..255; P_j planes are written over S[j & 1]; 256 columns allocated
constexpr int S_COL = 0, O_COL = 2 * BKV, P_COL = O_COL + HD, P_PLANE = BKV / 2;
static_assert(NTOK % BKV == 0 && BKV % 16 == 0 && (NBLK % 2) == 0, "key blocking");

template <int P, int CPS>
struct ACfg {
    static constexpr int V_STAGES = CPS == 2 ? 1 : 2;
    static constexpr int TMEM_COLS = CPS == 2 ? 256 : 512;
    static constexpr int Q_BYTES = P * Q_TILE;
    static constexpr int KV_BYTES = P * KV_TILE;
    static constexpr int OFF_K = Q_BYTES;
    static constexpr int OFF_V = OFF_K + K_STAGES * KV_BYTES;
    static constexpr int OFF_XCH = OFF_V + V_STAGES * KV_BYTES;   // float [2 parity][NSPLIT][128 rows]: row-max exchange
    static constexpr int OFF_BAR = OFF_XCH + 2 * NSPLIT * BM * 4;
    static constexpr int SMEM = OFF_BAR + 256 + 1024 /*align slack*/;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int P, int CPS>
__global__ void __launch_bounds__(ATT_THREADS, CPS)
self_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                         float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_planes, int p_out, int n_img,
                         float scale_log2, int kv_xor, float* __restrict__ lse_out, int skip_dead) {
    using C = ACfg<P, CPS>;
    constexpr int V_STAGES = C::V_STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a SHARED pointer (LDS / STS)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* q_free = bars + 1;
    uint64_t* k_full = bars + 2;     // [2]
    uint64_t* k_free = bars + 4;     // [2]
    uint64_t* v_full = bars + 6;     // [2]
    uint64_t* v_free = bars + 8;     // [2]
    uint64_t* s_full = bars + 10;    // [2]
    uint64_t* p_ready = bars + 12;
    uint64_t* pv_done = bars + 13;
    uint64_t* s_free = bars + 14;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = n_img * HEADS * QTILES;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmQ);
        tc::prefetch_tmap(&tmKV);
        tc::mbar_init(q_full, 1);
        tc::mbar_init(q_free, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&k_full[i], 1);
            tc::mbar_init(&k_free[i], 1);
            tc::mbar_init(&v_full[i], 1);
            tc::mbar_init(&v_free[i], 1);
            tc::mbar_init(&s_full[i], 1);
            // CPS = 1: one elected arrive per softmax warp (S is in registers); CPS = 2: the commit of the PV product
            // that read P_j out of the same columns
            tc::mbar_init(&s_free[i], CPS == 2 ? 1 : 4 * NSPLIT);
        }
        tc::mbar_init(p_ready, 4 * NSPLIT);
        tc::mbar_init(pv_done, 1);
        tc::fence_barrier_init();
    }
    rp::pdl_launch_dependents();                  // the next kernel may start its prologue (common.cuh)
    if (warp == 1) tc::tmem_alloc(tmem_slot, C::TMEM_COLS);
    rp::pdl_wait();                               // the previous kernel has completed: its outputs are visible
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto q_tile = [&](int p) { return smem + p * Q_TILE; };
    auto k_tile = [&](int st, int p) { return smem + C::OFF_K + st * C::KV_BYTES + p * KV_TILE; };
    auto v_tile = [&](int st, int p) { return smem + C::OFF_V + st * C::KV_BYTES + p * KV_TILE; };

    // CPS = 2: 2 x 384 threads share the 64 K registers of the SM = 80 each at launch; the control warpgroup hands most
    // of its share to the two softmax warpgroups (48 S values + 32 O values + temporaries per thread live at once).
    // The pool is per CTA: 128 x 32 + 256 x 104 = 30 720 = 384 x 80 -- a larger request would wait forever.
    // (the instruction sits at the head of the warpgroup-uniform branch that dominates the role code, so ptxas
    // allocates each side with its own budget)
    if (warp < ATT_CTRL) {
    if constexpr (CPS == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    if (warp == 0) {
        // ---------------------------------------------------------------------------- TMA producer (convergent warp)
        int ks = 0, kph = 0, vs = 0, vph = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int qt = tile % QTILES, h = (tile / QTILES) % HEADS, img = tile / (QTILES * HEADS);
            tc::mbar_wait(q_free, (it & 1) ^ 1);
            if (tc::elect_one_sync()) {
                tc::mbar_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
                for (int p = 0; p < P; ++p) tc::tma_load_4d(q_tile(p), &tmQ, q_full, h * HD, qt * BM, img, p);
            }
            __syncwarp();
            for (int j = 0; j < NBLK; ++j) {
                tc::mbar_wait(&k_free[ks], kph ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(&k_full[ks], C::KV_BYTES);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        tc::tma_load_4d(k_tile(ks, p), &tmKV, &k_full[ks], EMB + h * HD, j * BKV, img ^ kv_xor, p);
                }
                __syncwarp();
                if (++ks == K_STAGES) { ks = 0; kph ^= 1; }
                tc::mbar_wait(&v_free[vs], vph ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(&v_full[vs], C::KV_BYTES);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        tc::tma_load_4d(v_tile(vs, p), &tmKV, &v_full[vs], 2 * EMB + h * HD, j * BKV, img ^ kv_xor, p);
                }
                __syncwarp();
                if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------------------- S = Q K^T issuer (convergent warp)
        // Two issuer warps on two schedulers: a single issuer spent most of its time executing its own instruction
        // stream (barrier polls, descriptor set-up, ~250 instructions per key block) while the tensor pipe idled.
        // This one runs up to two key blocks ahead of the softmax (S is double buffered in tensor memory).
        constexpr uint32_t idesc_s = tc::make_idesc_bf16(BM, BKV);
        int ks = 0, kph = 0, it = 0;
        uint32_t g = 0;     // global key-block counter of this CTA
        const uint64_t dq0 = tc::make_kmajor_sw128_desc(tc::smem_u32(q_tile(0)));
        const uint64_t dq1 = tc::make_kmajor_sw128_desc(tc::smem_u32(q_tile(P - 1)));
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            tc::mbar_wait(q_full, it & 1);
            for (int j = 0; j < NBLK; ++j, ++g) {
                tc::mbar_wait(&s_free[g & 1], ((g >> 1) & 1) ^ 1);      // softmax has read S_{g-2} out of this buffer
                tc::mbar_wait(&k_full[ks], kph);
                tc::tcgen05_fence_after();
                const uint32_t d = tmem_base + S_COL + (g & 1) * BKV;
                const uint64_t dk0 = tc::make_kmajor_sw128_desc(tc::smem_u32(k_tile(ks, 0)));
                const uint64_t dk1 = tc::make_kmajor_sw128_desc(tc::smem_u32(k_tile(ks, P - 1)));
                if (tc::elect_one_sync()) {
                    // tcgen05 accumulates with truncation (bias ~ chain length x 2^-25): the small correction terms of
                    // ALL K steps go first, the main terms last, so the full-magnitude chain is K/16 long, not 3K/16
                    uint32_t accum = 0u;
                    if (P == 2) {
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) {          // 16 bf16 = 32 B inside the swizzle row = +2
                            tc::umma_bf16(d, dq1 + 2 * k, dk0 + 2 * k, idesc_s, accum);
                            tc::umma_bf16(d, dq0 + 2 * k, dk1 + 2 * k, idesc_s, 1u);
                            accum = 1u;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        tc::umma_bf16(d, dq0 + 2 * k, dk0 + 2 * k, idesc_s, accum);
                        accum = 1u;
                    }
                    tc::umma_commit(&k_free[ks]);
                    tc::umma_commit(&s_full[g & 1]);
                    if (j + 1 == NBLK) tc::umma_commit(q_free);       // last S of this item is in flight
                }
                __syncwarp();
                if (++ks == K_STAGES) { ks = 0; kph ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ---------------------------------------------------------------------------- O = P V issuer (convergent warp)
        // A = P_j read from TENSOR memory (K-major, two bf16 per column, written by the softmax threads with
        // tcgen05.st), B = V_j as loaded by TMA ([key][d] rows = MN-major SWIZZLE_128B).
        constexpr uint32_t idesc_o = tc::make_idesc_bf16(BM, HD) | tc::IDESC_B_MN;
        int vs = 0, vph = 0;
        uint32_t g = 0;
        const uint32_t d = tmem_base + O_COL;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int j = 0; j < NBLK; ++j, ++g) {
                tc::mbar_wait(p_ready, g & 1);
                tc::mbar_wait(&v_full[vs], vph);
                tc::tcgen05_fence_after();
                const uint32_t pbase = tmem_base + (CPS == 2 ? S_COL + (g & 1) * BKV : P_COL);
                const uint32_t ap0 = pbase, ap1 = pbase + (P - 1) * P_PLANE;
                const uint64_t dv0 = tc::make_mnmajor_sw128_desc(tc::smem_u32(v_tile(vs, 0)), 0);
                const uint64_t dv1 = tc::make_mnmajor_sw128_desc(tc::smem_u32(v_tile(vs, P - 1)), 0);
                if (tc::elect_one_sync()) {
                    uint32_t accum = 0u;
                    if (P == 2) {
#pragma unroll
                        for (int kk = 0; kk < BKV / 16; ++kk) {
                            const uint32_t b_off = (kk * 16 * 128) >> 4;                       // MN-major: 16 keys = 16 rows
                            tc::umma_bf16_ts(d, ap1 + 8 * kk, dv0 + b_off, idesc_o, accum);    // 16 keys = 8 columns
                            tc::umma_bf16_ts(d, ap0 + 8 * kk, dv1 + b_off, idesc_o, 1u);
                            accum = 1u;
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < BKV / 16; ++kk) {
                        const uint32_t b_off = (kk * 16 * 128) >> 4;
                        tc::umma_bf16_ts(d, ap0 + 8 * kk, dv0 + b_off, idesc_o, accum);
                        accum = 1u;
                    }
                    tc::umma_commit(&v_free[vs]);
                    if (CPS == 2) tc::umma_commit(&s_free[g & 1]);      // P_j (= the S buffer) has been consumed
                    tc::umma_commit(pv_done);
                }
                __syncwarp();
                if (++vs == V_STAGES) { vs = 0; vph ^= 1; }
            }
        }
    }
    } else {
        if constexpr (CPS == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
        // ---------------------------------------------------------------------------- softmax warps
        // NSPLIT threads per query row: warps w, w+4, .. share a TMEM lane quarter, thread `hsel` owns key columns
        // [HB hsel, HB hsel + HB) of every block and output columns [HO hsel, HO hsel + HO).  They agree on the
        // running maximum through a double-buffered shared-memory slot and a named barrier per block; the
        // partial sums are only combined at the end.  ncu history: one thread per row = one softmax warp per
        // scheduler at 0.22 IPC; two per row (143 registers) still left the schedulers two warps each to hide
        // the MUFU / TMEM latencies of a serial chain; four per row halves the registers and doubles the warps.
        const int quarter = warp & 3;                      // TMEM lane quarter this warp may touch
        const int hsel = (warp - ATT_CTRL) >> 2;           // which slice of the columns
        const int r = quarter * 32 + lane;                 // query row inside the tile
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        float* xch = reinterpret_cast<float*>(smem + C::OFF_XCH);
        const int bar_id = 1 + quarter;
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int qt = tile % QTILES, h = (tile / QTILES) % HEADS, img = tile / (QTILES * HEADS);
            if (skip_dead && qt * BM + quarter * 32 >= NTOK) {
                // Rows past the last token (lane quarters 2 and 3 of the fifth query tile: zero-filled Q, nothing is ever
                // stored): keep the barrier protocol in lock step with the live warps -- same waits, same arrivals -- and
                // skip the arithmetic; the issue slots go to the live quarters and to the SM's other CTA.  Both warps of
                // a quarter take this branch together, so the row exchanges (bar.sync) are skipped by both.
                for (int j = 0; j < NBLK; ++j, ++g) {
                    tc::mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                    if constexpr (CPS == 1) {
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&s_free[g & 1]);
                    }
                    if (j > 0) tc::mbar_wait(pv_done, (g - 1) & 1);   // P_j may not be announced before PV_{j-1} retired
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(p_ready);
                }
                tc::mbar_wait(pv_done, (g - 1) & 1);
                continue;
            }
            float m = -INFINITY, l = 0.f, alpha_prev = 1.f;
            float o[HO];
#pragma unroll
            for (int i = 0; i < HO; ++i) o[i] = 0.f;
            auto fold_o = [&]() {                           // o = o * alpha_prev + O_j (this thread's columns)
                // 16 columns at a time: the exponentials of the next block (HB registers) are live across this call
#pragma unroll
                for (int c = 0; c < HO; c += 16) {
                    uint32_t t[16];
                    tc::tmem_ld_32x32b_x16(t_lane + O_COL + hsel * HO + c, t);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[c + i] = fmaf(o[c + i], alpha_prev, __uint_as_float(t[i]));
                }
            };
            for (int j = 0; j < NBLK; ++j, ++g) {
                tc::mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                tc::tcgen05_fence_after();
                uint32_t s[HB];
                {
                    const uint32_t t_s = t_lane + S_COL + (g & 1) * BKV + hsel * HB;
                    if constexpr (HB == 48) {
                        tc::tmem_ld_32x32b_x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
                        tc::tmem_ld_32x32b_x16(t_s + 32, *reinterpret_cast<uint32_t(*)[16]>(&s[32]));
                    } else {
                        tc::tmem_ld_32x32b_x16(t_s, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
                        tc::tmem_ld_32x32b_x8(t_s + 16, *reinterpret_cast<uint32_t(*)[8]>(&s[16]));
                    }
                    tc::tmem_ld_wait();
                }
                if constexpr (CPS == 1) {
                    tc::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&s_free[g & 1]);     // S_{g+2} may overwrite this buffer
                }
                float bmax = __uint_as_float(s[0]);
#pragma unroll
                for (int i = 1; i < HB; ++i) bmax = fmaxf(bmax, __uint_as_float(s[i]));
                float* slot = xch + (g & 1) * NSPLIT * BM;
                slot[hsel * BM + r] = bmax;
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSPLIT) : "memory");
#pragma unroll
                for (int o2 = 1; o2 < NSPLIT; ++o2) bmax = fmaxf(bmax, slot[((hsel + o2) % NSPLIT) * BM + r]);
                const float m_new = fmaxf(m, bmax);
                const float alpha = tc::fast_exp2((m - m_new) * scale_log2);      // 0 on the first block
                const float ms = m_new * scale_log2;
                float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
                for (int i = 0; i < HB; i += 2) {
                    const float p0 = tc::fast_exp2(fmaf(__uint_as_float(s[i]), scale_log2, -ms));
                    const float p1 = tc::fast_exp2(fmaf(__uint_as_float(s[i + 1]), scale_log2, -ms));
                    sum0 += p0; sum1 += p1;
                    s[i] = __float_as_uint(p0); s[i + 1] = __float_as_uint(p1);
                }
                l = l * alpha + (sum0 + sum1);
                m = m_new;
                if (j > 0) {
                    // O_{j-1} is complete: fold it into the register accumulator; P may now be overwritten
                    tc::mbar_wait(pv_done, (g - 1) & 1);
                    tc::tcgen05_fence_after();
                    fold_o();
                }
                alpha_prev = alpha;
                // P_j -> bf16 planes in TENSOR memory (K-major A operand of the PV product: key 2c, 2c+1 in column c)
#pragma unroll
                for (int p = 0; p < P; ++p) {
#pragma unroll
                    uint32_t w[HB / 2];
#pragma unroll
                    for (int i = 0; i < HB / 2; ++i) {
                        const float v0 = __uint_as_float(s[2 * i]), v1 = __uint_as_float(s[2 * i + 1]);
                        w[i] = pack_bf16x2(v0, v1);
                        if (p + 1 < P) {
                            s[2 * i] = __float_as_uint(v0 - __uint_as_float(w[i] << 16));
                            s[2 * i + 1] = __float_as_uint(v1 - __uint_as_float(w[i] & 0xffff0000u));
                        }
                    }
                    // CPS = 2: over S_j.  Every thread of the row loaded its S columns before the row-max barrier above.
                    const uint32_t t_p = t_lane + (CPS == 2 ? S_COL + (g & 1) * BKV : P_COL) + p * P_PLANE + hsel * (HB / 2);
#pragma unroll
                    for (int ci = 0; ci < HB / 16; ++ci)
                        tc::tmem_st_32x32b_x8(t_p + ci * 8, *reinterpret_cast<uint32_t(*)[8]>(&w[ci * 8]));
                    if constexpr ((HB / 2) % 8 == 4)
                        tc::tmem_st_32x32b_x4(t_p + (HB / 16) * 8, *reinterpret_cast<uint32_t(*)[4]>(&w[(HB / 16) * 8]));
                }
                tc::tmem_st_wait();
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(p_ready);
            }
            // last product of the item
            tc::mbar_wait(pv_done, (g - 1) & 1);
            tc::tcgen05_fence_after();
            fold_o();
            // combine the two partial sums of the row (same slot discipline as the maxima: one more barrier)
            float* slot = xch + (g & 1) * NSPLIT * BM;
            slot[hsel * BM + r] = l;
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSPLIT) : "memory");
            float lsum = 0.f;                                   // same order in every thread of the row
#pragma unroll
            for (int o2 = 0; o2 < NSPLIT; ++o2) lsum += slot[o2 * BM + r];
            const float inv = 1.0f / lsum;
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NSPLIT) : "memory");      // slot parity g&1 is reused by the next tile's block 0
            const int row = qt * BM + r;
            // training: log2-sum-exp of the scaled row, L = m c + log2(sum), so that P = 2^(s c - L) can be recomputed
            // by the backward kernel (attention_bwd_tc.cu) instead of being stored
            if (lse_out && hsel == 0 && row < NTOK)
                lse_out[((size_t)img * HEADS + h) * NTOK + row] = fmaf(m, scale_log2, log2f(lsum));
            if (row < NTOK) {
                // out[n, row, h*64 + d]  ((attn @ v).transpose(1,2).reshape(B,N,C), vision_transformer.py:329)
                const size_t o_idx = ((size_t)img * NTOK + row) * EMB + h * HD + hsel * HO;
#pragma unroll
                for (int i = 0; i < HO; ++i) o[i] *= inv;
                if (out_f32) {
#pragma unroll
                    for (int i = 0; i < HO; i += 4)
                        *reinterpret_cast<float4*>(out_f32 + o_idx + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
                }
                if (out_planes) {
                    const size_t plane = (size_t)n_img * NTOK * EMB;
                    for (int p = 0; p < p_out; ++p) {
#pragma unroll
                        for (int i = 0; i < HO; i += 8) {
                            uint4 w;
                            w.x = pack_bf16x2(o[i], o[i + 1]);
                            w.y = pack_bf16x2(o[i + 2], o[i + 3]);
                            w.z = pack_bf16x2(o[i + 4], o[i + 5]);
                            w.w = pack_bf16x2(o[i + 6], o[i + 7]);
                            *reinterpret_cast<uint4*>(out_planes + p * plane + o_idx + i) = w;
                            o[i] -= __uint_as_float(w.x << 16); o[i + 1] -= __uint_as_float(w.x & 0xffff0000u);
                            o[i + 2] -= __uint_as_float(w.y << 16); o[i + 3] -= __uint_as_float(w.y & 0xffff0000u);
                            o[i + 4] -= __uint_as_float(w.z << 16); o[i + 5] -= __uint_as_float(w.z & 0xffff0000u);
                            o[i + 6] -= __uint_as_float(w.w << 16); o[i + 7] -= __uint_as_float(w.w & 0xffff0000u);
                        }
                    }
                }
            }
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

// bf16 [P][n_img][576][576] qkv planes, box = [1][1][box_rows][64 columns], 128-byte swizzle, OOB rows -> 0
int make_qkv_tmap(CUtensorMap* out, const void* base, int P, int n_img, int box_rows) {
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    if (!fn) {
        rp::set_error("cuTensorMapEncodeTiled entry point unavailable");
        return RP_EINVAL;
    }
    const cuuint64_t ld = 3 * EMB;
    cuuint64_t gdim[4] = {ld, (cuuint64_t)NTOK, (cuuint64_t)n_img, (cuuint64_t)P};
    cuuint64_t gstr[3] = {ld * 2, ld * NTOK * 2, ld * NTOK * 2 * (cuuint64_t)n_img};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rp::set_error("qkv tensor map failed (CUresult %d) n_img=%d box_rows=%d", (int)r, n_img, box_rows);
        return RP_EINVAL;
    }
    return RP_OK;
}

// RELPOSE_ATT_CPS=1 selects the one-CTA-per-SM variant (A/B measurements); default: two CTAs per SM
int attention_cps() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RELPOSE_ATT_CPS");
        v = (e && e[0] == '1') ? 1 : 2;
    }
    return v;
}

template <int P, int CPS>
int launch_attention_cps(const void* qkv_planes, float* out_f32, void* out_planes, int p_out, int n_img, int device,
                         cudaStream_t st, int kv_xor, float* lse_out) {
    using C = ACfg<P, CPS>;
    CUtensorMap tmQ, tmKV;
    int rc = make_qkv_tmap(&tmQ, qkv_planes, P, n_img, BM);
    if (rc) return rc;
    rc = make_qkv_tmap(&tmKV, qkv_planes, P, n_img, BKV);
    if (rc) return rc;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(self_attention_tc_kernel<P, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            rp::set_error("rp_self_attention_tc: cudaFuncSetAttribute(%d): %s", C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    const int ntiles = n_img * HEADS * QTILES;
    const int slots = CPS * rp::num_sms(device);
    const int grid = ntiles < slots ? ntiles : slots;
    const float scale_log2 = 0.125f * 1.4426950408889634f;     // head_dim^-0.5 * log2(e)
    // RELPOSE_ATT_SKIP_DEAD=0: the softmax warps of the rows past the last token do the full arithmetic (A/B measurements)
    static const int skip_dead = [] { const char* e = getenv("RELPOSE_ATT_SKIP_DEAD"); return (e && e[0] == '0') ? 0 : 1; }();
    rp::launch(self_attention_tc_kernel<P, CPS>, dim3(grid), dim3(ATT_THREADS), (size_t)(C::SMEM), st, tmQ, tmKV, out_f32, static_cast<__nv_bfloat16*>(out_planes),
                                                                       p_out, n_img, scale_log2, kv_xor, lse_out, skip_dead);
    return rp::finish_launch("rp_self_attention_tc");
}

template <int P>
int launch_attention(const void* qkv_planes, float* out_f32, void* out_planes, int p_out, int n_img, int device,
                     cudaStream_t st, int kv_xor = 0, float* lse_out = nullptr) {
    if (attention_cps() == 1) return launch_attention_cps<P, 1>(qkv_planes, out_f32, out_planes, p_out, n_img, device, st, kv_xor, lse_out);
    return launch_attention_cps<P, 2>(qkv_planes, out_f32, out_planes, p_out, n_img, device, st, kv_xor, lse_out);
}

}  // namespace

extern "C" int rp_self_attention_tc(const void* qkv_planes, float* out_f32, void* out_planes, int n_img, int P, int P_out,
                                    int device, void* stream) {
    RP_REQUIRE(qkv_planes && (out_f32 || out_planes) && n_img > 0, RP_EINVAL, "rp_self_attention_tc: bad argument");
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_self_attention_tc: P must be 1 (bf16) or 2 (bf16x3)");
    RP_REQUIRE(!out_planes || (P_out >= 1 && P_out <= 2), RP_EINVAL, "rp_self_attention_tc: bad P_out");
    RP_REQUIRE(rp::aligned16(qkv_planes) && rp::aligned16(out_f32) && rp::aligned16(out_planes), RP_EALIGN,
               "rp_self_attention_tc: 16-byte alignment");
    RP_GUARD(device);
    if (P == 1) return launch_attention<1>(qkv_planes, out_f32, out_planes, P_out, n_img, device, (cudaStream_t)stream);
    return launch_attention<2>(qkv_planes, out_f32, out_planes, P_out, n_img, device, (cudaStream_t)stream);
}

// Training forward (A11): the same kernel, which also writes the log2-sum-exp of every scaled score row,
// lse [n_img][3][576] -- all the backward kernel (rp_attention_bwd_tc) needs to recompute the probabilities.
extern "C" int rp_self_attention_tc_lse(const void* qkv_planes, float* out_f32, float* lse, int n_img, int P, int device,
                                        void* stream) {
    RP_REQUIRE(qkv_planes && out_f32 && lse && n_img > 0, RP_EINVAL, "rp_self_attention_tc_lse: bad argument");
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_self_attention_tc_lse: P must be 1 (bf16) or 2 (bf16x3)");
    RP_REQUIRE(rp::aligned16(qkv_planes) && rp::aligned16(out_f32), RP_EALIGN, "rp_self_attention_tc_lse: 16-byte alignment");
    RP_GUARD(device);
    if (P == 1) return launch_attention<1>(qkv_planes, out_f32, nullptr, 0, n_img, device, (cudaStream_t)stream, 0, lse);
    return launch_attention<2>(qkv_planes, out_f32, nullptr, 0, n_img, device, (cudaStream_t)stream, 0, lse);
}

// Plain cross attention between the two views of each pair (--noess ablation, vision_transformer.py:239-253):
// image n's queries attend to the keys/values of image n^1.  Same kernel; only the TMA coordinates of K and V change.
extern "C" int rp_cross_attention_tc(const void* qkv_planes, float* out_f32, void* out_planes, int n_img, int P, int P_out,
                                     int device, void* stream) {
    RP_REQUIRE(qkv_planes && (out_f32 || out_planes) && n_img > 0 && n_img % 2 == 0, RP_EINVAL,
               "rp_cross_attention_tc: n_img must be a positive even number");
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_cross_attention_tc: P must be 1 (bf16) or 2 (bf16x3)");
    RP_REQUIRE(!out_planes || (P_out >= 1 && P_out <= 2), RP_EINVAL, "rp_cross_attention_tc: bad P_out");
    RP_REQUIRE(rp::aligned16(qkv_planes) && rp::aligned16(out_f32) && rp::aligned16(out_planes), RP_EALIGN,
               "rp_cross_attention_tc: 16-byte alignment");
    RP_GUARD(device);
    if (P == 1) return launch_attention<1>(qkv_planes, out_f32, out_planes, P_out, n_img, device, (cudaStream_t)stream, 1);
    return launch_attention<2>(qkv_planes, out_f32, out_planes, P_out, n_img, device, (cudaStream_t)stream, 1);
}
