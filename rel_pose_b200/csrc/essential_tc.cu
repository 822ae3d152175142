// Essential Matrix Module core (A7: vision_transformer.py:198-223) on tcgen05 tensor cores.
//
//   dir 0:  S1 = q2 k1^T / 8,  A1 = softmax(S1,-1) * softmax(S1,-2),  F1 = V1'^T A1 V1',  V1' = [v1 | pos]
//   dir 1:  S2 = q1 k2^T / 8,  A2 likewise,                            F2 = V2'^T A2 V2'
// per pair b (views 2b, 2b+1) and head h; F is 70 x 70.  The dual softmax needs the row AND the column
// log-sum-exp of the same S, so the module is two passes over S, both with S recomputed on the tensor cores
// and never written to HBM (the reference materialises [B,3,576,576] eight times):
//
//   em_stats_tc_kernel   rows x columns of S (which = 0) and of S^T (which = 1): S tile = R C_j^T by tcgen05
//                        (M128 x N96 x K64), one softmax thread per row keeps a running (max, sum);
//                        lse2[row] = max * c + log2(sum),  c = 0.125 * log2(e).
//   em_accum_tc_kernel   one CTA per (pair, dir, head), looping over the 5 row tiles i and 6 key blocks j:
//                          S_ij (tcgen05) -> A_ij = 2^(2 c s - lse2_r[i] - lse2_c[j]) (softmax threads, written as
//                          bf16 planes into swizzled shared memory) -> T_i += A_ij [v_j | pos_j] (tcgen05,
//                          N = 64 with V_j as loaded by TMA = MN-major operand, N = 16 with pos^T K-major);
//                        after the 6 blocks T_i (TMEM) is re-split to planes in shared memory and
//                          F += [v_i | pos_i]^T T_i   (tcgen05, K = the 128 rows of the tile)
//                        accumulates in TMEM across the 5 tiles; F is read once at the end: no partial
//                        matrices, no atomics, bit-reproducible.
// Operands are the bf16 planes of the QKV GEMM (P = 1 bf16, P = 2 split bf16: a0 b0 + a0 b1 + a1 b0).
// Operand rows that only pad an MMA shape (M = 128 for the 6 positional rows, N = 16 for 6 columns) are NOT
// materialised: the descriptors run over neighbouring shared memory, whose (finite) contents only reach
// accumulator rows / columns that are never read.
#include <cstdlib>
#include "tc_common.cuh"

namespace {

constexpr int NTOK = RP_NTOK, HD = RP_HDIM, EMB = RP_EMBED, HEADS = RP_HEADS, EMW = RP_EMW, NPOS = RP_NPOS;
constexpr int BM = 128, BKV = 96, NBLK = NTOK / BKV, RTILES = (NTOK + BM - 1) / BM;   // 6 key blocks, 5 row tiles
constexpr int R_TILE = BM * 128;          // [128 x 64] bf16
constexpr int C_TILE = BKV * 128;         // [96 x 64] bf16
constexpr int POS_SUB = 1024;             // [8 x 64] bf16: one 8-row group of a K-major SWIZZLE_128B tile
constexpr int P_SUB = BM * 128;
constexpr int EM_THREADS = 192;
constexpr float SCALE_LOG2 = 0.125f * 1.4426950408889634f;

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ============================================================================================ stats pass
template <int P>
struct SCfg {
    static constexpr int R_BYTES = P * R_TILE;
    static constexpr int C_BYTES = P * C_TILE;
    static constexpr int OFF_C = R_BYTES;
    static constexpr int OFF_BAR = OFF_C + 2 * C_BYTES;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
};

// Two CTAs per SM (82 KB of shared memory and 256 tensor-memory columns each): the one-thread-per-row running
// (max, sum) chain of one CTA fills the gaps of the other.
template <int P>
__global__ void __launch_bounds__(EM_THREADS, 2)
em_stats_tc_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC,
                   float* __restrict__ lse2, int B, int skip_dead) {
    using C = SCfg<P>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a SHARED pointer (LDS / STS)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* r_full = bars + 0;
    uint64_t* r_free = bars + 1;
    uint64_t* c_full = bars + 2;    // [2]
    uint64_t* c_free = bars + 4;    // [2]
    uint64_t* s_full = bars + 6;    // [2]
    uint64_t* s_free = bars + 8;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = B * 2 * 2 * HEADS * RTILES;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmR);
        tc::prefetch_tmap(&tmC);
        tc::mbar_init(r_full, 1);
        tc::mbar_init(r_free, 1);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&c_full[i], 1);
            tc::mbar_init(&c_free[i], 1);
            tc::mbar_init(&s_full[i], 1);
            tc::mbar_init(&s_free[i], 128);
        }
        tc::fence_barrier_init();
    }
    rp::pdl_launch_dependents();                  // the next kernel may start its prologue (common.cuh)
    if (warp == 1) tc::tmem_alloc(tmem_slot, 256);
    rp::pdl_wait();                               // the previous kernel has completed: its outputs are visible
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto r_tile = [&](int p) { return smem + p * R_TILE; };
    auto c_tile = [&](int st, int p) { return smem + C::OFF_C + st * C::C_BYTES + p * C_TILE; };
    // item -> (tile, head, which, dir, pair); which = 0: rows = queries, columns = keys; 1: swapped
    auto decode = [&](int item, int& tile, int& h, int& which, int& dir, int& b) {
        tile = item % RTILES; item /= RTILES;
        h = item % HEADS; item /= HEADS;
        which = item & 1; dir = (item >> 1) & 1; b = item >> 2;
    };

    if (warp == 0) {
        // TMA producer (convergent warp, one elected lane issues)
        int cs = 0, cph = 0, it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            int tile, h, which, dir, b;
            decode(item, tile, h, which, dir, b);
            const int q_img = 2 * b + (1 - dir), kv_img = 2 * b + dir;
            const int r_img = which == 0 ? q_img : kv_img, r_col = (which == 0 ? 0 : EMB) + h * HD;
            const int c_img = which == 0 ? kv_img : q_img, c_col = (which == 0 ? EMB : 0) + h * HD;
            tc::mbar_wait(r_free, (it & 1) ^ 1);
            if (tc::elect_one_sync()) {
                tc::mbar_expect_tx(r_full, C::R_BYTES);
#pragma unroll
                for (int p = 0; p < P; ++p) tc::tma_load_4d(r_tile(p), &tmR, r_full, r_col, tile * BM, r_img, p);
            }
            __syncwarp();
            for (int j = 0; j < NBLK; ++j) {
                tc::mbar_wait(&c_free[cs], cph ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(&c_full[cs], C::C_BYTES);
#pragma unroll
                    for (int p = 0; p < P; ++p) tc::tma_load_4d(c_tile(cs, p), &tmC, &c_full[cs], c_col, j * BKV, c_img, p);
                }
                __syncwarp();
                if (++cs == 2) { cs = 0; cph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // MMA issuer (convergent warp)
        constexpr uint32_t idesc_s = tc::make_idesc_bf16(BM, BKV);
        int cs = 0, cph = 0, it = 0;
        uint32_t g = 0;
        const uint64_t dr0 = tc::make_kmajor_sw128_desc(tc::smem_u32(r_tile(0)));
        const uint64_t dr1 = tc::make_kmajor_sw128_desc(tc::smem_u32(r_tile(P - 1)));
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            tc::mbar_wait(r_full, it & 1);
            for (int j = 0; j < NBLK; ++j, ++g) {
                tc::mbar_wait(&s_free[g & 1], ((g >> 1) & 1) ^ 1);
                tc::mbar_wait(&c_full[cs], cph);
                tc::tcgen05_fence_after();
                const uint32_t d = tmem_base + (g & 1) * BKV;
                const uint64_t dc0 = tc::make_kmajor_sw128_desc(tc::smem_u32(c_tile(cs, 0)));
                const uint64_t dc1 = tc::make_kmajor_sw128_desc(tc::smem_u32(c_tile(cs, P - 1)));
                if (tc::elect_one_sync()) {
                    uint32_t accum = 0u;
                    if (P == 2) {                          // small correction terms first (see em_accum_tc_kernel)
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) {
                            tc::umma_bf16(d, dr1 + 2 * k, dc0 + 2 * k, idesc_s, accum);
                            tc::umma_bf16(d, dr0 + 2 * k, dc1 + 2 * k, idesc_s, 1u);
                            accum = 1u;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        tc::umma_bf16(d, dr0 + 2 * k, dc0 + 2 * k, idesc_s, accum);
                        accum = 1u;
                    }
                    tc::umma_commit(&c_free[cs]);
                    tc::umma_commit(&s_full[g & 1]);
                    if (j + 1 == NBLK) tc::umma_commit(r_free);
                }
                __syncwarp();
                if (++cs == 2) { cs = 0; cph ^= 1; }
            }
        }
    } else {
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint32_t g = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            int tile, h, which, dir, b;
            decode(item, tile, h, which, dir, b);
            float m = -INFINITY, l = 0.f;
            const bool dead = skip_dead && tile * BM + quarter * 32 >= NTOK;   // lane quarter entirely past the last token: nothing to compute or store
            for (int j = 0; j < NBLK; ++j, ++g) {
                tc::mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                if (dead) {                                       // keep the protocol (every thread arrives), skip the arithmetic
                    tc::mbar_arrive(&s_free[g & 1]);
                    continue;
                }
                tc::tcgen05_fence_after();
                uint32_t s[BKV];
                {
                    uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
                    uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
                    uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
                    const uint32_t t_s = t_lane + (g & 1) * BKV;
                    tc::tmem_ld_32x32b_x32(t_s, s0);
                    tc::tmem_ld_32x32b_x32(t_s + 32, s1);
                    tc::tmem_ld_32x32b_x32(t_s + 64, s2);
                    tc::tmem_ld_wait();
                }
                tc::tcgen05_fence_before();
                tc::mbar_arrive(&s_free[g & 1]);          // S buffer is in registers now
                float bmax = __uint_as_float(s[0]);
#pragma unroll
                for (int i = 1; i < BKV; ++i) bmax = fmaxf(bmax, __uint_as_float(s[i]));
                const float m_new = fmaxf(m, bmax);
                const float ms = m_new * SCALE_LOG2;
                float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
                for (int i = 0; i < BKV; i += 2) {
                    sum0 += tc::fast_exp2(fmaf(__uint_as_float(s[i]), SCALE_LOG2, -ms));
                    sum1 += tc::fast_exp2(fmaf(__uint_as_float(s[i + 1]), SCALE_LOG2, -ms));
                }
                l = l * tc::fast_exp2((m - m_new) * SCALE_LOG2) + (sum0 + sum1);
                m = m_new;
            }
            const int row = tile * BM + r;
            if (row < NTOK)
                lse2[((((size_t)b * 2 + dir) * 2 + which) * HEADS + h) * NTOK + row] = fmaf(m, SCALE_LOG2, log2f(l));
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, 256);
    }
}

// ============================================================================================ accumulate pass
constexpr int RING = 3;                                // unified K/V ring: loads and uses follow the same order
// TMEM columns: S double buffer, then one region that holds the per-block product T_j = A_ij [v_j | pos_j]
// (T_v 64 + T_pos 16 columns) and, after T has been folded into registers, the per-tile products
// F1 = v_i^T [T_v | T_pos] (80 columns) and G = [T_v | T_pos]^T pos_i (16 columns).
constexpr int T_S = 0, T_X = 2 * BKV, T_TV = T_X, T_TP = T_X + 64, T_F1 = T_X, T_G = T_X + 80;   // <= 288

template <int P>
struct ECfg {
    static constexpr int OFF_POSI = 0;                                   // [P][2][1 KiB]  pos^T of the row tile
    static constexpr int OFF_POSJ = OFF_POSI + P * 2 * POS_SUB;          // [RING][P][2][1 KiB]
    static constexpr int OFF_Q = OFF_POSJ + RING * P * 2 * POS_SUB;      // [P][16 KiB]
    static constexpr int OFF_VI = OFF_Q + P * R_TILE;                    // [P][16 KiB]
    static constexpr int OFF_RING = OFF_VI + P * R_TILE;                 // [RING][P][12 KiB]
    static constexpr int OFF_A = OFF_RING + RING * P * C_TILE;           // [P][32 KiB]   A planes, later T planes
    static constexpr int OFF_CL = OFF_A + P * 2 * P_SUB;                 // 576 floats
    static constexpr int OFF_BAR = OFF_CL + NTOK * 4;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int SLOT_BYTES = P * C_TILE;
    static constexpr int POSJ_BYTES = P * 2 * POS_SUB;
};

template <int P>
__global__ void __launch_bounds__(EM_THREADS, 1)
em_accum_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                   const __grid_constant__ CUtensorMap tmPos, const float* __restrict__ lse2, float* __restrict__ bil,
                   int B, int width, int em_flags) {
    using C = ECfg<P>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a SHARED pointer (LDS / STS)
    float* cl = reinterpret_cast<float*>(smem + C::OFF_CL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* q_free = bars + 1;
    uint64_t* vi_full = bars + 2;
    uint64_t* vi_free = bars + 3;
    uint64_t* ring_full = bars + 4;     // [3]
    uint64_t* ring_free = bars + 7;     // [3]
    uint64_t* s_full = bars + 10;       // [2]
    uint64_t* p_ready = bars + 12;
    uint64_t* pv_done = bars + 13;
    uint64_t* t_ready = bars + 14;
    uint64_t* f_done = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = B * 2 * HEADS;
    const bool has_pos = width > HD;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmQ);
        tc::prefetch_tmap(&tmKV);
        tc::prefetch_tmap(&tmPos);
        tc::mbar_init(q_full, 1);
        tc::mbar_init(q_free, 1);
        tc::mbar_init(vi_full, 1);
        tc::mbar_init(vi_free, 1);
        for (int i = 0; i < RING; ++i) {
            tc::mbar_init(&ring_full[i], 1);
            tc::mbar_init(&ring_free[i], 1);
        }
        tc::mbar_init(&s_full[0], 1);
        tc::mbar_init(&s_full[1], 1);
        tc::mbar_init(p_ready, 128);
        tc::mbar_init(pv_done, 1);
        tc::mbar_init(t_ready, 128);
        tc::mbar_init(f_done, 1);
        tc::fence_barrier_init();
    }
    rp::pdl_launch_dependents();                  // the next kernel may start its prologue (common.cuh)
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    rp::pdl_wait();                               // the previous kernel has completed: its outputs are visible
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto posi_tile = [&](int p) { return smem + C::OFF_POSI + p * 2 * POS_SUB; };
    auto posj_tile = [&](int st, int p) { return smem + C::OFF_POSJ + st * C::POSJ_BYTES + p * 2 * POS_SUB; };
    auto q_tile = [&](int p) { return smem + C::OFF_Q + p * R_TILE; };
    auto vi_tile = [&](int p) { return smem + C::OFF_VI + p * R_TILE; };
    auto ring_tile = [&](int st, int p) { return smem + C::OFF_RING + st * C::SLOT_BYTES + p * C_TILE; };
    // A_ij planes: two 64-wide K-major sub-tiles.  After the last block of a row tile the same memory holds
    // T_i as an MN-major operand of two 64-column atoms: atom 0 = T_v, atom 1 = [T_pos (8) | stale A].
    auto a_tile = [&](int p) { return smem + C::OFF_A + p * 2 * P_SUB; };

    if (warp == 0) {
        // ---------------------------------------------------------------------------- TMA producer (convergent warp)
        int rs = 0, rph = 0;
        uint32_t tt = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int h = item % HEADS, dir = (item / HEADS) & 1, b = item / (2 * HEADS);
            const int q_img = 2 * b + (1 - dir), kv_img = 2 * b + dir;
            auto load_k = [&](int j) {
                tc::mbar_wait(&ring_free[rs], rph ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(&ring_full[rs], C::SLOT_BYTES);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        tc::tma_load_4d(ring_tile(rs, p), &tmKV, &ring_full[rs], EMB + h * HD, j * BKV, kv_img, p);
                }
                __syncwarp();
                if (++rs == RING) { rs = 0; rph ^= 1; }
            };
            auto load_v = [&](int j) {
                tc::mbar_wait(&ring_free[rs], rph ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(&ring_full[rs], C::SLOT_BYTES + (has_pos ? C::POSJ_BYTES : 0));
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        tc::tma_load_4d(ring_tile(rs, p), &tmKV, &ring_full[rs], 2 * EMB + h * HD, j * BKV, kv_img, p);
                        if (has_pos) {
                            tc::tma_load_4d(posj_tile(rs, p), &tmPos, &ring_full[rs], j * BKV, 0, b, p);
                            tc::tma_load_4d(posj_tile(rs, p) + POS_SUB, &tmPos, &ring_full[rs], j * BKV + 64, 0, b, p);
                        }
                    }
                }
                __syncwarp();
                if (++rs == RING) { rs = 0; rph ^= 1; }
            };
            for (int tile = 0; tile < RTILES; ++tile, ++tt) {
                tc::mbar_wait(q_free, (tt & 1) ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(q_full, P * R_TILE);
#pragma unroll
                    for (int p = 0; p < P; ++p) tc::tma_load_4d(q_tile(p), &tmQ, q_full, h * HD, tile * BM, q_img, p);
                }
                __syncwarp();
                // same order as the issuer consumes: K0, then (K_{j+1}), V_j
                load_k(0);
                for (int j = 0; j < NBLK; ++j) {
                    if (j + 1 < NBLK) load_k(j + 1);
                    load_v(j);
                    if (j == 0) {
                        // left factor of F for this row tile (needed only after the 6 blocks)
                        tc::mbar_wait(vi_free, (tt & 1) ^ 1);
                        if (tc::elect_one_sync()) {
                            tc::mbar_expect_tx(vi_full, P * R_TILE + (has_pos ? P * 2 * POS_SUB : 0));
#pragma unroll
                            for (int p = 0; p < P; ++p) {
                                // --cross_features (:219-220): the left factor is the OTHER view's [v | pos]
                                tc::tma_load_4d(vi_tile(p), &tmQ, vi_full, 2 * EMB + h * HD, tile * BM,
                                                (em_flags & RP_EM_CROSS_FEATURES) ? q_img : kv_img, p);
                                if (has_pos) {
                                    tc::tma_load_4d(posi_tile(p), &tmPos, vi_full, tile * BM, 0, b, p);
                                    tc::tma_load_4d(posi_tile(p) + POS_SUB, &tmPos, vi_full, tile * BM + 64, 0, b, p);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------------------- MMA issuer (convergent warp)
        // tcgen05 accumulates in fp32 with truncation: a chain of n accumulations carries a bias of ~n 2^-25.
        // Every product group therefore issues its small correction terms (a1 b0, a0 b1) for ALL K steps first
        // and the main terms (a0 b0) last, and no TMEM accumulator lives longer than one key block / row tile:
        // T_j and F_t are folded in fp32 (round to nearest) by the softmax threads.
        // All 32 lanes run the control flow; one elected lane issues (see tc::elect_one_sync).
        constexpr uint32_t idesc_s = tc::make_idesc_bf16(BM, BKV);
        constexpr uint32_t idesc_tv = tc::make_idesc_bf16(BM, HD) | tc::IDESC_B_MN;
        constexpr uint32_t idesc_tp = tc::make_idesc_bf16(BM, 16);
        constexpr uint32_t idesc_fvv = tc::make_idesc_bf16(BM, HD) | tc::IDESC_A_MN | tc::IDESC_B_MN;
        constexpr uint32_t idesc_fvp = tc::make_idesc_bf16(BM, 16) | tc::IDESC_A_MN | tc::IDESC_B_MN;
        constexpr uint32_t idesc_g = tc::make_idesc_bf16(BM, 16) | tc::IDESC_A_MN;
        int rs = 0, rph = 0;
        uint32_t g = 0, tt = 0;
        auto kdesc = [&](const uint8_t* base, uint32_t off) { return tc::make_kmajor_sw128_desc(tc::smem_u32(base) + off); };
        auto mdesc = [&](const uint8_t* base, uint32_t off, uint32_t lbo) {
            return tc::make_mnmajor_sw128_desc(tc::smem_u32(base) + off, lbo);
        };
        const uint64_t dq0 = kdesc(q_tile(0), 0), dq1 = kdesc(q_tile(P - 1), 0);
        const uint64_t da0 = kdesc(a_tile(0), 0), da1 = kdesc(a_tile(P - 1), 0);          // A planes as K-major A operand
        const uint64_t dvi0 = mdesc(vi_tile(0), 0, 0), dvi1 = mdesc(vi_tile(P - 1), 0, 0);
        const uint64_t dt0 = mdesc(a_tile(0), 0, 0), dt1 = mdesc(a_tile(P - 1), 0, 0);    // T_v: MN-major B operand
        const uint64_t dtp0 = mdesc(a_tile(0) + P_SUB, 0, 0), dtp1 = mdesc(a_tile(P - 1) + P_SUB, 0, 0);   // T_pos atom
        const uint64_t dta0 = mdesc(a_tile(0), 0, P_SUB), dta1 = mdesc(a_tile(P - 1), 0, P_SUB);          // [T_v|T_pos] as A
        const uint64_t dpi0 = kdesc(posi_tile(0), 0), dpi1 = kdesc(posi_tile(P - 1), 0);
        auto issue_s = [&](uint32_t gb) {
            tc::mbar_wait(&ring_full[rs], rph);
            tc::tcgen05_fence_after();
            const uint32_t d = tmem_base + T_S + (gb & 1) * BKV;
            const uint64_t dk0 = kdesc(ring_tile(rs, 0), 0), dk1 = kdesc(ring_tile(rs, P - 1), 0);
            if (tc::elect_one_sync()) {
                uint32_t acc = 0;
                if (P == 2) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        tc::umma_bf16(d, dq1 + 2 * k, dk0 + 2 * k, idesc_s, acc);
                        tc::umma_bf16(d, dq0 + 2 * k, dk1 + 2 * k, idesc_s, 1u);
                        acc = 1u;
                    }
                }
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    tc::umma_bf16(d, dq0 + 2 * k, dk0 + 2 * k, idesc_s, acc);
                    acc = 1u;
                }
                tc::umma_commit(&ring_free[rs]);
                tc::umma_commit(&s_full[gb & 1]);
            }
            __syncwarp();
            if (++rs == RING) { rs = 0; rph ^= 1; }
        };
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            for (int tile = 0; tile < RTILES; ++tile, ++tt) {
                tc::mbar_wait(q_full, tt & 1);
                tc::tcgen05_fence_after();
                issue_s(g);
                for (int j = 0; j < NBLK; ++j, ++g) {
                    if (j + 1 < NBLK) {
                        issue_s(g + 1);
                        if (j + 2 == NBLK) {
                            if (tc::elect_one_sync()) tc::umma_commit(q_free);
                            __syncwarp();
                        }
                    }
                    tc::mbar_wait(p_ready, g & 1);
                    tc::mbar_wait(&ring_full[rs], rph);
                    tc::tcgen05_fence_after();
                    const uint64_t dv0 = mdesc(ring_tile(rs, 0), 0, 0), dv1 = mdesc(ring_tile(rs, P - 1), 0, 0);
                    const uint64_t dpj0 = kdesc(posj_tile(rs, 0), 0), dpj1 = kdesc(posj_tile(rs, P - 1), 0);
                    if (tc::elect_one_sync()) {
                        // T_j = A_ij [v_j | pos_j]  (fresh accumulator every block)
#pragma unroll
                        for (int pass = (P == 2 ? 0 : 1); pass < 2; ++pass) {
#pragma unroll
                            for (int kk = 0; kk < BKV / 16; ++kk) {
                                const uint32_t a_off = ((kk >> 2) * P_SUB + (kk & 3) * 32) >> 4;     // K-major: 16 keys = 32 B
                                const uint32_t v_off = (kk * 16 * 128) >> 4;                         // MN-major: 16 keys = 16 rows
                                const uint32_t p_off = ((kk >> 2) * POS_SUB + (kk & 3) * 32) >> 4;
                                const uint32_t first = (pass == (P == 2 ? 0 : 1) && kk == 0) ? 0u : 1u;
                                if (pass == 0) {
                                    tc::umma_bf16(tmem_base + T_TV, da1 + a_off, dv0 + v_off, idesc_tv, first);
                                    tc::umma_bf16(tmem_base + T_TV, da0 + a_off, dv1 + v_off, idesc_tv, 1u);
                                    if (has_pos) {
                                        tc::umma_bf16(tmem_base + T_TP, da1 + a_off, dpj0 + p_off, idesc_tp, first);
                                        tc::umma_bf16(tmem_base + T_TP, da0 + a_off, dpj1 + p_off, idesc_tp, 1u);
                                    }
                                } else {
                                    tc::umma_bf16(tmem_base + T_TV, da0 + a_off, dv0 + v_off, idesc_tv, first);
                                    if (has_pos) tc::umma_bf16(tmem_base + T_TP, da0 + a_off, dpj0 + p_off, idesc_tp, first);
                                }
                            }
                        }
                        tc::umma_commit(&ring_free[rs]);
                        tc::umma_commit(pv_done);
                    }
                    __syncwarp();
                    if (++rs == RING) { rs = 0; rph ^= 1; }
                }
                // F_t = v_i^T [T_v | T_pos]  and  G_t = [T_v | T_pos]^T pos_i   (K = the 128 rows of the tile)
                tc::mbar_wait(t_ready, tt & 1);
                tc::mbar_wait(vi_full, tt & 1);
                tc::tcgen05_fence_after();
                if (tc::elect_one_sync()) {
#pragma unroll
                    for (int pass = (P == 2 ? 0 : 1); pass < 2; ++pass) {
#pragma unroll
                        for (int kk = 0; kk < BM / 16; ++kk) {
                            const uint32_t mn_off = (kk * 16 * 128) >> 4;                            // 16 rows of an MN-major tile
                            const uint32_t k_off = ((kk >> 2) * POS_SUB + (kk & 3) * 32) >> 4;       // 16 K elements, K-major
                            const uint32_t first = (pass == (P == 2 ? 0 : 1) && kk == 0) ? 0u : 1u;
                            const int na = pass == 0 ? 2 : 1;
#pragma unroll
                            for (int t = 0; t < na; ++t) {
                                // pass 0: (a1, b0) then (a0, b1); pass 1: (a0, b0)
                                const bool a_hi = pass == 0 && t == 0, b_hi = pass == 0 && t == 1;
                                const uint32_t acc = (t == 0) ? first : 1u;
                                tc::umma_bf16(tmem_base + T_F1, (a_hi ? dvi1 : dvi0) + mn_off, (b_hi ? dt1 : dt0) + mn_off, idesc_fvv, acc);
                                if (has_pos) {
                                    tc::umma_bf16(tmem_base + T_F1 + 64, (a_hi ? dvi1 : dvi0) + mn_off, (b_hi ? dtp1 : dtp0) + mn_off,
                                                  idesc_fvp, acc);
                                    tc::umma_bf16(tmem_base + T_G, (a_hi ? dta1 : dta0) + mn_off, (b_hi ? dpi1 : dpi0) + k_off, idesc_g, acc);
                                }
                            }
                        }
                    }
                    tc::umma_commit(vi_free);
                    tc::umma_commit(f_done);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------------------- softmax warps
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int st = threadIdx.x - 64;                   // 0..127 among the softmax threads
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128;
        const uint32_t sw = (uint32_t)(r & 7);
        uint32_t g = 0, tt = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int h = item % HEADS, dir = (item / HEADS) & 1, b = item / (2 * HEADS);
            const float* lse_r = lse2 + ((((size_t)b * 2 + dir) * 2 + 0) * HEADS + h) * NTOK;
            const float* lse_c = lse2 + ((((size_t)b * 2 + dir) * 2 + 1) * HEADS + h) * NTOK;
            float* dst = bil + (((size_t)b * 2 + dir) * HEADS + h) * (size_t)(width * width);
            // F_t (TMEM) of row tile `tile_done` is folded into the output: fixed order, one owner thread per
            // element, so the result is bit-reproducible.  Rows 0..63: F1 lanes; rows 64..69: G columns.
            auto fold_f = [&](int tile_done) {
                if (quarter < 2) {
                    float* row = dst + r * width;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t t[32];
                        tc::tmem_ld_32x32b_x32(t_lane + T_F1 + half * 32, t);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                            float* q4 = row + half * 32 + i;
                            if (tile_done > 0) { acc.x = q4[0]; acc.y = q4[1]; acc.z = q4[2]; acc.w = q4[3]; }
                            q4[0] = acc.x + __uint_as_float(t[i]); q4[1] = acc.y + __uint_as_float(t[i + 1]);
                            q4[2] = acc.z + __uint_as_float(t[i + 2]); q4[3] = acc.w + __uint_as_float(t[i + 3]);
                        }
                    }
                }
                if (has_pos && quarter < 3) {
                    uint32_t t[16];
                    tc::tmem_ld_32x32b_x16(t_lane + (quarter < 2 ? T_F1 + 64 : T_G), t);   // warp-uniform address
                    uint32_t gq[16];
                    tc::tmem_ld_32x32b_x16(t_lane + T_G, gq);
                    tc::tmem_ld_wait();
                    if (quarter < 2) {                              // F[a][64+u] = F1[a][64+u]
#pragma unroll
                        for (int u = 0; u < NPOS; ++u) {
                            float* q1 = dst + r * width + HD + u;
                            *q1 = (tile_done > 0 ? *q1 : 0.f) + __uint_as_float(t[u]);
                        }
                    }
                    if (r < EMW) {                                  // F[64+u][c] = G[c][u],  c = r < 70
#pragma unroll
                        for (int u = 0; u < NPOS; ++u) {
                            float* q1 = dst + (HD + u) * width + r;
                            *q1 = (tile_done > 0 ? *q1 : 0.f) + __uint_as_float(gq[u]);
                        }
                    }
                }
            };
            asm volatile("bar.sync 1, 128;" ::: "memory");          // everyone is done with the previous item's cl[]
            // --use_single_softmax (:201-203): A = softmax(S,-1) = 2^(c s - lse2_r): no column term, exponent counted once
            const bool single = (em_flags & RP_EM_SINGLE_SOFTMAX) != 0;
            const float emul = single ? SCALE_LOG2 : 2.f * SCALE_LOG2;
            for (int c = st; c < NTOK; c += 128) cl[c] = single ? 0.f : lse_c[c];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int tile = 0; tile < RTILES; ++tile, ++tt) {
                const int row = tile * BM + r;
                const float rl = row < NTOK ? lse_r[row] : INFINITY;   // rows past the end contribute exactly 0
                float tacc[72];
#pragma unroll
                for (int i = 0; i < 72; ++i) tacc[i] = 0.f;
                auto fold_t = [&]() {                               // tacc += T_j (the block product that just retired)
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t t[32];
                        tc::tmem_ld_32x32b_x32(t_lane + T_TV + half * 32, t);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) tacc[half * 32 + i] += __uint_as_float(t[i]);
                    }
                    if (has_pos) {
                        uint32_t t[16];
                        tc::tmem_ld_32x32b_x16(t_lane + T_TP, t);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 8; ++i) tacc[64 + i] += __uint_as_float(t[i]);
                    }
                };
                for (int j = 0; j < NBLK; ++j, ++g) {
                    tc::mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                    tc::tcgen05_fence_after();
                    uint32_t s[BKV];
                    {
                        uint32_t(&s0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[0]);
                        uint32_t(&s1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[32]);
                        uint32_t(&s2)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
                        const uint32_t t_s = t_lane + T_S + (g & 1) * BKV;
                        tc::tmem_ld_32x32b_x32(t_s, s0);
                        tc::tmem_ld_32x32b_x32(t_s + 32, s1);
                        tc::tmem_ld_32x32b_x32(t_s + 64, s2);
                        tc::tmem_ld_wait();
                    }
                    // A = softmax(S,-1) * softmax(S,-2) = 2^(2 c s - lse2_r - lse2_c)     (:205-206)
                    const float4* cl4 = reinterpret_cast<const float4*>(cl + j * BKV);
#pragma unroll
                    for (int i = 0; i < BKV; i += 4) {
                        const float4 c4 = cl4[i >> 2];
                        s[i] = __float_as_uint(tc::fast_exp2(fmaf(__uint_as_float(s[i]), emul, -(rl + c4.x))));
                        s[i + 1] = __float_as_uint(tc::fast_exp2(fmaf(__uint_as_float(s[i + 1]), emul, -(rl + c4.y))));
                        s[i + 2] = __float_as_uint(tc::fast_exp2(fmaf(__uint_as_float(s[i + 2]), emul, -(rl + c4.z))));
                        s[i + 3] = __float_as_uint(tc::fast_exp2(fmaf(__uint_as_float(s[i + 3]), emul, -(rl + c4.w))));
                    }
                    // The A buffer and the T / F accumulator region are free once the previous product has retired:
                    // inside a tile that is PV_{j-1} (fold its T_j), at a tile start the F update of the previous
                    // tile (fold its F_t; it also read the T planes that alias the A buffer).
                    if (j > 0) {
                        tc::mbar_wait(pv_done, (g - 1) & 1);
                        tc::tcgen05_fence_after();
                        fold_t();
                    } else if (tile > 0) {
                        tc::mbar_wait(f_done, (tt - 1) & 1);
                        tc::tcgen05_fence_after();
                        fold_f(tile - 1);
                    } else if (tt > 0) {
                        tc::mbar_wait(f_done, (tt - 1) & 1);     // previous item: already folded, ordering only
                    }
#pragma unroll
                    for (int c = 0; c < BKV / 8; ++c) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(s[c * 8 + i]);
                        const uint32_t off = (uint32_t)(c >> 3) * P_SUB + row_off + ((((uint32_t)c & 7) ^ sw) << 4);
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            uint4 w;
                            w.x = pack2(v[0], v[1]); w.y = pack2(v[2], v[3]); w.z = pack2(v[4], v[5]); w.w = pack2(v[6], v[7]);
                            *reinterpret_cast<uint4*>(a_tile(p) + off) = w;
                            if (p + 1 < P) {
                                v[0] -= __uint_as_float(w.x << 16); v[1] -= __uint_as_float(w.x & 0xffff0000u);
                                v[2] -= __uint_as_float(w.y << 16); v[3] -= __uint_as_float(w.y & 0xffff0000u);
                                v[4] -= __uint_as_float(w.z << 16); v[5] -= __uint_as_float(w.z & 0xffff0000u);
                                v[6] -= __uint_as_float(w.w << 16); v[7] -= __uint_as_float(w.w & 0xffff0000u);
                            }
                        }
                    }
                    tc::fence_proxy_async_smem();
                    tc::tcgen05_fence_before();
                    tc::mbar_arrive(p_ready);
                }
                // last block product of the tile, then T_i (registers) -> bf16 planes in shared memory, laid out as
                // the MN-major operand [row r][64 columns] per atom: rows of 128 swizzled bytes like the A planes
                tc::mbar_wait(pv_done, (g - 1) & 1);
                tc::tcgen05_fence_after();
                fold_t();
#pragma unroll
                for (int c = 0; c < 9; ++c) {
                    if (c == 8 && !has_pos) break;
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = tacc[c * 8 + i];
                    // chunks 0..7: T_v in atom 0; chunk 8: T_pos = chunk 0 of atom 1
                    const uint32_t off = (c < 8 ? 0u : (uint32_t)P_SUB) + row_off + ((((uint32_t)c & 7) ^ sw) << 4);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        uint4 w;
                        w.x = pack2(v[0], v[1]); w.y = pack2(v[2], v[3]); w.z = pack2(v[4], v[5]); w.w = pack2(v[6], v[7]);
                        *reinterpret_cast<uint4*>(a_tile(p) + off) = w;
                        if (p + 1 < P) {
                            v[0] -= __uint_as_float(w.x << 16); v[1] -= __uint_as_float(w.x & 0xffff0000u);
                            v[2] -= __uint_as_float(w.y << 16); v[3] -= __uint_as_float(w.y & 0xffff0000u);
                            v[4] -= __uint_as_float(w.z << 16); v[5] -= __uint_as_float(w.z & 0xffff0000u);
                            v[6] -= __uint_as_float(w.w << 16); v[7] -= __uint_as_float(w.w & 0xffff0000u);
                        }
                    }
                }
                tc::fence_proxy_async_smem();
                tc::tcgen05_fence_before();
                tc::mbar_arrive(t_ready);
            }
            // last row tile of this (pair, dir, head)
            tc::mbar_wait(f_done, (tt - 1) & 1);
            tc::tcgen05_fence_after();
            fold_f(RTILES - 1);
            tc::tcgen05_fence_before();
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ============================================================================================ accumulate pass, version 2
// Same mathematics and operand layouts as em_accum_tc_kernel above; what changed is where the time went (ncu, round 1:
// tensor pipe 20.8 % active, one softmax warp per scheduler at ~0.2 IPC, the issuer idle 37 % of the time waiting for
// A_ij, the softmax warps idle 24 % waiting for the product that frees the single A buffer):
//   * FOUR softmax threads per row (16 warps): 24 score columns and 16 + 2 product columns each instead of 96 / 72;
//   * A_ij goes to TENSOR memory (tcgen05.st over S_ij, like P in attention_tc.cu) and is the TS-mode A operand of
//     T_j = A_ij [v_j | pos_j]: no shared-memory round trip, no generic->async proxy fence, and A_(j+1) (other S buffer)
//     can be written while T_j is still running -- the softmax and the product now overlap;
//   * a work unit is ONE row tile of a (pair, direction, head): 1920 units at 64 pairs instead of 384, so the grid
//     fills 148 SMs evenly (12.97 waves instead of 2.59); the per-tile forms go to a workspace and are added in tile
//     order by em_reduce_tiles_kernel (bit-reproducible, the same order the old kernel used).
constexpr int EM2_CTRL = 4, EM2_SPLIT = 4, EM2_THREADS = 32 * (EM2_CTRL + 4 * EM2_SPLIT);     // 640
constexpr int HB2 = BKV / EM2_SPLIT, HT2 = HD / EM2_SPLIT;                                     // 24 score / 16 T_v columns per thread
constexpr int A_PLANE = BKV / 2;                                                               // 48 columns per bf16 plane of A_ij

constexpr int EM_INTERNAL_NO_SKIP_DEAD = 1 << 30;   // launcher-only bit of em_flags (RELPOSE_EM_SKIP_DEAD=0, A/B measurements)

template <int P>
__global__ void __launch_bounds__(EM2_THREADS, 1)
em_accum2_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                    const __grid_constant__ CUtensorMap tmPos, const float* __restrict__ lse2, float* __restrict__ part,
                    int B, int width, int em_flags) {
    using C = ECfg<P>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a SHARED pointer (LDS / STS)
    float* cl = reinterpret_cast<float*>(smem + C::OFF_CL);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* q_free = bars + 1;
    uint64_t* vi_full = bars + 2;
    uint64_t* vi_free = bars + 3;
    uint64_t* ring_full = bars + 4;     // [3]
    uint64_t* ring_free = bars + 7;     // [3]
    uint64_t* s_full = bars + 10;       // [2]
    uint64_t* p_ready = bars + 12;
    uint64_t* pv_done = bars + 13;
    uint64_t* t_ready = bars + 14;
    uint64_t* f_done = bars + 15;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nunits = B * 2 * HEADS * RTILES;
    const bool has_pos = width > HD;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmQ);
        tc::prefetch_tmap(&tmKV);
        tc::prefetch_tmap(&tmPos);
        tc::mbar_init(q_full, 1);
        tc::mbar_init(q_free, 1);
        tc::mbar_init(vi_full, 1);
        tc::mbar_init(vi_free, 1);
        for (int i = 0; i < RING; ++i) {
            tc::mbar_init(&ring_full[i], 1);
            tc::mbar_init(&ring_free[i], 1);
        }
        tc::mbar_init(&s_full[0], 1);
        tc::mbar_init(&s_full[1], 1);
        tc::mbar_init(p_ready, 4 * EM2_SPLIT);      // one elected arrive per softmax warp
        tc::mbar_init(pv_done, 1);
        tc::mbar_init(t_ready, 4 * EM2_SPLIT);
        tc::mbar_init(f_done, 1);
        tc::fence_barrier_init();
    }
    rp::pdl_launch_dependents();                  // the next kernel may start its prologue (common.cuh)
    if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
    rp::pdl_wait();                               // the previous kernel has completed: its outputs are visible
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto posi_tile = [&](int p) { return smem + C::OFF_POSI + p * 2 * POS_SUB; };
    auto posj_tile = [&](int st, int p) { return smem + C::OFF_POSJ + st * C::POSJ_BYTES + p * 2 * POS_SUB; };
    auto q_tile = [&](int p) { return smem + C::OFF_Q + p * R_TILE; };
    auto vi_tile = [&](int p) { return smem + C::OFF_VI + p * R_TILE; };
    auto ring_tile = [&](int st, int p) { return smem + C::OFF_RING + st * C::SLOT_BYTES + p * C_TILE; };
    // T_i planes (MN-major operand of the F / G products): atom 0 = T_v, atom 1 = [T_pos (8) | never read]
    auto t_tile = [&](int p) { return smem + C::OFF_A + p * 2 * P_SUB; };
    auto decode = [&](int unit, int& tile, int& h, int& dir, int& b) {
        tile = unit % RTILES;
        const int item = unit / RTILES;
        h = item % HEADS; dir = (item / HEADS) & 1; b = item / (2 * HEADS);
    };

    if (warp == 0) {
        // ---------------------------------------------------------------------------- TMA producer (convergent warp)
        int rs = 0, rph = 0;
        uint32_t tt = 0;
        for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x, ++tt) {
            int tile, h, dir, b;
            decode(unit, tile, h, dir, b);
            const int q_img = 2 * b + (1 - dir), kv_img = 2 * b + dir;
            auto load_k = [&](int j) {
                tc::mbar_wait(&ring_free[rs], rph ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(&ring_full[rs], C::SLOT_BYTES);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        tc::tma_load_4d(ring_tile(rs, p), &tmKV, &ring_full[rs], EMB + h * HD, j * BKV, kv_img, p);
                }
                __syncwarp();
                if (++rs == RING) { rs = 0; rph ^= 1; }
            };
            auto load_v = [&](int j) {
                tc::mbar_wait(&ring_free[rs], rph ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(&ring_full[rs], C::SLOT_BYTES + (has_pos ? C::POSJ_BYTES : 0));
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        tc::tma_load_4d(ring_tile(rs, p), &tmKV, &ring_full[rs], 2 * EMB + h * HD, j * BKV, kv_img, p);
                        if (has_pos) {
                            tc::tma_load_4d(posj_tile(rs, p), &tmPos, &ring_full[rs], j * BKV, 0, b, p);
                            tc::tma_load_4d(posj_tile(rs, p) + POS_SUB, &tmPos, &ring_full[rs], j * BKV + 64, 0, b, p);
                        }
                    }
                }
                __syncwarp();
                if (++rs == RING) { rs = 0; rph ^= 1; }
            };
            tc::mbar_wait(q_free, (tt & 1) ^ 1);
            if (tc::elect_one_sync()) {
                tc::mbar_expect_tx(q_full, P * R_TILE);
#pragma unroll
                for (int p = 0; p < P; ++p) tc::tma_load_4d(q_tile(p), &tmQ, q_full, h * HD, tile * BM, q_img, p);
            }
            __syncwarp();
            load_k(0);                  // same order as the issuer consumes: K0, then (K_{j+1}), V_j
            for (int j = 0; j < NBLK; ++j) {
                if (j + 1 < NBLK) load_k(j + 1);
                load_v(j);
                if (j == 0) {
                    tc::mbar_wait(vi_free, (tt & 1) ^ 1);
                    if (tc::elect_one_sync()) {
                        tc::mbar_expect_tx(vi_full, P * R_TILE + (has_pos ? P * 2 * POS_SUB : 0));
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            tc::tma_load_4d(vi_tile(p), &tmQ, vi_full, 2 * EMB + h * HD, tile * BM,
                                            (em_flags & RP_EM_CROSS_FEATURES) ? q_img : kv_img, p);
                            if (has_pos) {
                                tc::tma_load_4d(posi_tile(p), &tmPos, vi_full, tile * BM, 0, b, p);
                                tc::tma_load_4d(posi_tile(p) + POS_SUB, &tmPos, vi_full, tile * BM + 64, 0, b, p);
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------------------- MMA issuer (convergent warp)
        // Issue order per unit: S_0, then for every block S_(j+1) before T_j.  tcgen05.mma of one thread execute in issue
        // order, so S_(j+2) (which overwrites the buffer A_j was read from) is behind T_j in the pipe by construction.
        constexpr uint32_t idesc_s = tc::make_idesc_bf16(BM, BKV);
        constexpr uint32_t idesc_tv = tc::make_idesc_bf16(BM, HD) | tc::IDESC_B_MN;
        constexpr uint32_t idesc_tp = tc::make_idesc_bf16(BM, 16);
        constexpr uint32_t idesc_fvv = tc::make_idesc_bf16(BM, HD) | tc::IDESC_A_MN | tc::IDESC_B_MN;
        constexpr uint32_t idesc_fvp = tc::make_idesc_bf16(BM, 16) | tc::IDESC_A_MN | tc::IDESC_B_MN;
        constexpr uint32_t idesc_g = tc::make_idesc_bf16(BM, 16) | tc::IDESC_A_MN;
        int rs = 0, rph = 0;
        uint32_t g = 0, tt = 0;
        auto kdesc = [&](const uint8_t* base, uint32_t off) { return tc::make_kmajor_sw128_desc(tc::smem_u32(base) + off); };
        auto mdesc = [&](const uint8_t* base, uint32_t off, uint32_t lbo) {
            return tc::make_mnmajor_sw128_desc(tc::smem_u32(base) + off, lbo);
        };
        const uint64_t dq0 = kdesc(q_tile(0), 0), dq1 = kdesc(q_tile(P - 1), 0);
        const uint64_t dvi0 = mdesc(vi_tile(0), 0, 0), dvi1 = mdesc(vi_tile(P - 1), 0, 0);
        const uint64_t dt0 = mdesc(t_tile(0), 0, 0), dt1 = mdesc(t_tile(P - 1), 0, 0);
        const uint64_t dtp0 = mdesc(t_tile(0) + P_SUB, 0, 0), dtp1 = mdesc(t_tile(P - 1) + P_SUB, 0, 0);
        const uint64_t dta0 = mdesc(t_tile(0), 0, P_SUB), dta1 = mdesc(t_tile(P - 1), 0, P_SUB);
        const uint64_t dpi0 = kdesc(posi_tile(0), 0), dpi1 = kdesc(posi_tile(P - 1), 0);
        auto issue_s = [&](uint32_t gb) {
            tc::mbar_wait(&ring_full[rs], rph);
            tc::tcgen05_fence_after();
            const uint32_t d = tmem_base + T_S + (gb & 1) * BKV;
            const uint64_t dk0 = kdesc(ring_tile(rs, 0), 0), dk1 = kdesc(ring_tile(rs, P - 1), 0);
            if (tc::elect_one_sync()) {
                uint32_t acc = 0;
                if (P == 2) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k) {
                        tc::umma_bf16(d, dq1 + 2 * k, dk0 + 2 * k, idesc_s, acc);
                        tc::umma_bf16(d, dq0 + 2 * k, dk1 + 2 * k, idesc_s, 1u);
                        acc = 1u;
                    }
                }
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) {
                    tc::umma_bf16(d, dq0 + 2 * k, dk0 + 2 * k, idesc_s, acc);
                    acc = 1u;
                }
                tc::umma_commit(&ring_free[rs]);
                tc::umma_commit(&s_full[gb & 1]);
            }
            __syncwarp();
            if (++rs == RING) { rs = 0; rph ^= 1; }
        };
        bool s0_issued = false;      // S_0 of this unit was already issued behind the previous unit's last block
        for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x, ++tt) {
            if (!s0_issued) {
                tc::mbar_wait(q_full, tt & 1);
                tc::tcgen05_fence_after();
                issue_s(g);
            }
            for (int j = 0; j < NBLK; ++j, ++g) {
                if (j + 1 < NBLK) {
                    issue_s(g + 1);
                    if (j + 2 == NBLK) {
                        if (tc::elect_one_sync()) tc::umma_commit(q_free);
                        __syncwarp();
                    }
                }
                tc::mbar_wait(p_ready, g & 1);
                tc::mbar_wait(&ring_full[rs], rph);
                tc::tcgen05_fence_after();
                const uint64_t dv0 = mdesc(ring_tile(rs, 0), 0, 0), dv1 = mdesc(ring_tile(rs, P - 1), 0, 0);
                const uint64_t dpj0 = kdesc(posj_tile(rs, 0), 0), dpj1 = kdesc(posj_tile(rs, P - 1), 0);
                const uint32_t a0 = tmem_base + T_S + (g & 1) * BKV, a1 = a0 + (P - 1) * A_PLANE;   // A_ij planes over S_ij
                if (tc::elect_one_sync()) {
                    // T_j = A_ij [v_j | pos_j]  (fresh accumulator every block; A from tensor memory)
#pragma unroll
                    for (int pass = (P == 2 ? 0 : 1); pass < 2; ++pass) {
#pragma unroll
                        for (int kk = 0; kk < BKV / 16; ++kk) {
                            const uint32_t v_off = (kk * 16 * 128) >> 4;                         // MN-major: 16 keys = 16 rows
                            const uint32_t p_off = ((kk >> 2) * POS_SUB + (kk & 3) * 32) >> 4;   // K-major pos^T: 16 keys = 32 B
                            const uint32_t first = (pass == (P == 2 ? 0 : 1) && kk == 0) ? 0u : 1u;
                            if (pass == 0) {
                                tc::umma_bf16_ts(tmem_base + T_TV, a1 + 8 * kk, dv0 + v_off, idesc_tv, first);
                                tc::umma_bf16_ts(tmem_base + T_TV, a0 + 8 * kk, dv1 + v_off, idesc_tv, 1u);
                                if (has_pos) {
                                    tc::umma_bf16_ts(tmem_base + T_TP, a1 + 8 * kk, dpj0 + p_off, idesc_tp, first);
                                    tc::umma_bf16_ts(tmem_base + T_TP, a0 + 8 * kk, dpj1 + p_off, idesc_tp, 1u);
                                }
                            } else {
                                tc::umma_bf16_ts(tmem_base + T_TV, a0 + 8 * kk, dv0 + v_off, idesc_tv, first);
                                if (has_pos) tc::umma_bf16_ts(tmem_base + T_TP, a0 + 8 * kk, dpj0 + p_off, idesc_tp, first);
                            }
                        }
                    }
                    tc::umma_commit(&ring_free[rs]);
                    tc::umma_commit(pv_done);
                }
                __syncwarp();
                if (++rs == RING) { rs = 0; rph ^= 1; }
            }
            // The next unit's first score block goes into the pipe BEFORE this unit's F step: its softmax (exponentials,
            // A_0 into tensor memory) then overlaps the 72 small F / G products instead of waiting behind them.  S_0 of the
            // next unit writes the S buffer whose A block was read by this unit's T_4 (issued above: in-order pipe).
            s0_issued = false;
            if (unit + (int)gridDim.x < nunits) {
                tc::mbar_wait(q_full, (tt + 1) & 1);
                tc::tcgen05_fence_after();
                issue_s(g);
                s0_issued = true;
            }
            // F_t = v_i^T [T_v | T_pos]  and  G_t = [T_v | T_pos]^T pos_i   (K = the 128 rows of the tile)
            tc::mbar_wait(t_ready, tt & 1);
            tc::mbar_wait(vi_full, tt & 1);
            tc::tcgen05_fence_after();
            if (tc::elect_one_sync()) {
#pragma unroll
                for (int pass = (P == 2 ? 0 : 1); pass < 2; ++pass) {
#pragma unroll
                    for (int kk = 0; kk < BM / 16; ++kk) {
                        const uint32_t mn_off = (kk * 16 * 128) >> 4;
                        const uint32_t k_off = ((kk >> 2) * POS_SUB + (kk & 3) * 32) >> 4;
                        const uint32_t first = (pass == (P == 2 ? 0 : 1) && kk == 0) ? 0u : 1u;
                        const int na = pass == 0 ? 2 : 1;
#pragma unroll
                        for (int t = 0; t < na; ++t) {
                            const bool a_hi = pass == 0 && t == 0, b_hi = pass == 0 && t == 1;
                            const uint32_t acc = (t == 0) ? first : 1u;
                            tc::umma_bf16(tmem_base + T_F1, (a_hi ? dvi1 : dvi0) + mn_off, (b_hi ? dt1 : dt0) + mn_off, idesc_fvv, acc);
                            if (has_pos) {
                                tc::umma_bf16(tmem_base + T_F1 + 64, (a_hi ? dvi1 : dvi0) + mn_off, (b_hi ? dtp1 : dtp0) + mn_off,
                                              idesc_fvp, acc);
                                tc::umma_bf16(tmem_base + T_G, (a_hi ? dta1 : dta0) + mn_off, (b_hi ? dpi1 : dpi0) + k_off, idesc_g, acc);
                            }
                        }
                    }
                }
                tc::umma_commit(vi_free);
                tc::umma_commit(f_done);
            }
            __syncwarp();
        }
    } else if (warp >= EM2_CTRL) {
        // ---------------------------------------------------------------------------- softmax warps (4 threads per row)
        const int quarter = warp & 3;                              // TMEM lane quarter
        const int hsel = (warp - EM2_CTRL) >> 2;                   // column slice of this thread
        const int r = quarter * 32 + lane;
        const int st = threadIdx.x - 32 * EM2_CTRL;                // 0..511 among the softmax threads
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128;
        const uint32_t sw = (uint32_t)(r & 7);
        const bool single = (em_flags & RP_EM_SINGLE_SOFTMAX) != 0;
        const float emul = single ? SCALE_LOG2 : 2.f * SCALE_LOG2;
        uint32_t g = 0, tt = 0;
        float* prev_dst = nullptr;
        // F_t of the unit that just finished: rows 0..63 = F1 lanes (this thread: 16 of the 64 v columns, thread 0 also
        // the 6 positional columns), rows 64..69 = G columns (thread 1 of rows < 70)
        auto fold_f = [&](float* dst) {
            if (quarter < 2) {
                uint32_t t[16];
                tc::tmem_ld_32x32b_x16(t_lane + T_F1 + hsel * HT2, t);
                tc::tmem_ld_wait();
                float* row = dst + r * width + hsel * HT2;
#pragma unroll
                for (int i = 0; i < 16; i += 2) *reinterpret_cast<float2*>(row + i) = make_float2(__uint_as_float(t[i]), __uint_as_float(t[i + 1]));
            }
            if (has_pos && hsel == 0 && quarter < 2) {
                uint32_t t[16];
                tc::tmem_ld_32x32b_x16(t_lane + T_F1 + 64, t);
                tc::tmem_ld_wait();
#pragma unroll
                for (int u = 0; u < NPOS; ++u) dst[r * width + HD + u] = __uint_as_float(t[u]);
            }
            if (has_pos && hsel == 1 && quarter < 3) {
                uint32_t t[16];
                tc::tmem_ld_32x32b_x16(t_lane + T_G, t);
                tc::tmem_ld_wait();
                if (r < EMW) {
#pragma unroll
                    for (int u = 0; u < NPOS; ++u) dst[(HD + u) * width + r] = __uint_as_float(t[u]);
                }
            }
        };
        for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x, ++tt) {
            int tile, h, dir, b;
            decode(unit, tile, h, dir, b);
            const float* lse_r = lse2 + ((((size_t)b * 2 + dir) * 2 + 0) * HEADS + h) * NTOK;
            const float* lse_c = lse2 + ((((size_t)b * 2 + dir) * 2 + 1) * HEADS + h) * NTOK;
            float* dst = part + (size_t)unit * (size_t)(width * width);
            asm volatile("bar.sync 1, 512;" ::: "memory");          // everyone is done with the previous unit's cl[]
            for (int c = st; c < NTOK; c += 4 * EM2_SPLIT * 32) cl[c] = single ? 0.f : lse_c[c];
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const int row = tile * BM + r;
            const float rl = row < NTOK ? lse_r[row] : INFINITY;    // rows past the end contribute exactly 0
            // Lane quarters that lie entirely past the last token (quarters 2 and 3 of the fifth row tile): Q is zero-filled,
            // so S_ij = 0 exactly, and the zero bits already sitting in the S buffer ARE the A_ij planes these rows would
            // write (2^-inf = 0).  Such a warp keeps every wait / arrival of the protocol and skips the loads, exponentials,
            // plane split and T_j folds (its T rows stay 0); all four warps of a quarter agree, so its bar.sync is skipped too.
            const bool dead = !(em_flags & EM_INTERNAL_NO_SKIP_DEAD) && tile * BM + quarter * 32 >= NTOK;
            float tacc[HT2], tpos[8];
#pragma unroll
            for (int i = 0; i < HT2; ++i) tacc[i] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) tpos[i] = 0.f;
            auto fold_t = [&]() {                                   // tacc += T_j (the block product that just retired)
                uint32_t t[16];
                tc::tmem_ld_32x32b_x16(t_lane + T_TV + hsel * HT2, t);
                if (has_pos && hsel == 0) {
                    uint32_t tp[8];
                    tc::tmem_ld_32x32b_x8(t_lane + T_TP, tp);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) tpos[i] += __uint_as_float(tp[i]);
                } else {
                    tc::tmem_ld_wait();
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) tacc[i] += __uint_as_float(t[i]);
            };
            for (int j = 0; j < NBLK; ++j, ++g) {
                tc::mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                tc::tcgen05_fence_after();
                if (!dead) {
                uint32_t s[HB2];
                {
                    const uint32_t t_s = t_lane + T_S + (g & 1) * BKV + hsel * HB2;
                    tc::tmem_ld_32x32b_x16(t_s, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
                    tc::tmem_ld_32x32b_x8(t_s + 16, *reinterpret_cast<uint32_t(*)[8]>(&s[16]));
                    tc::tmem_ld_wait();
                }
                // A = softmax(S,-1) * softmax(S,-2) = 2^(2 c s - lse2_r - lse2_c)     (:205-206)
                const float4* cl4 = reinterpret_cast<const float4*>(cl + j * BKV + hsel * HB2);
#pragma unroll
                for (int i = 0; i < HB2; i += 4) {
                    const float4 c4 = cl4[i >> 2];
                    s[i] = __float_as_uint(tc::fast_exp2(fmaf(__uint_as_float(s[i]), emul, -(rl + c4.x))));
                    s[i + 1] = __float_as_uint(tc::fast_exp2(fmaf(__uint_as_float(s[i + 1]), emul, -(rl + c4.y))));
                    s[i + 2] = __float_as_uint(tc::fast_exp2(fmaf(__uint_as_float(s[i + 2]), emul, -(rl + c4.z))));
                    s[i + 3] = __float_as_uint(tc::fast_exp2(fmaf(__uint_as_float(s[i + 3]), emul, -(rl + c4.w))));
                }
                // all four threads of the row hold their S_ij columns in registers: A_ij may now overwrite S_ij
                asm volatile("bar.sync %0, 128;" ::"r"(2 + quarter) : "memory");
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    uint32_t w[HB2 / 2];
#pragma unroll
                    for (int i = 0; i < HB2 / 2; ++i) {
                        const float v0 = __uint_as_float(s[2 * i]), v1 = __uint_as_float(s[2 * i + 1]);
                        w[i] = pack2(v0, v1);
                        if (p + 1 < P) {
                            s[2 * i] = __float_as_uint(v0 - __uint_as_float(w[i] << 16));
                            s[2 * i + 1] = __float_as_uint(v1 - __uint_as_float(w[i] & 0xffff0000u));
                        }
                    }
                    const uint32_t t_a = t_lane + T_S + (g & 1) * BKV + p * A_PLANE + hsel * (HB2 / 2);
                    tc::tmem_st_32x32b_x8(t_a, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
                    tc::tmem_st_32x32b_x4(t_a + 8, *reinterpret_cast<uint32_t(*)[4]>(&w[8]));
                }
                }   // !dead
                // the T / F accumulator region must be drained before T_j may be issued: fold the previous block's
                // product (inside a unit) or the previous unit's F_t
                if (j > 0) {
                    tc::mbar_wait(pv_done, (g - 1) & 1);
                    tc::tcgen05_fence_after();
                    if (!dead) fold_t();
                } else if (tt > 0) {
                    tc::mbar_wait(f_done, (tt - 1) & 1);
                    tc::tcgen05_fence_after();
                    fold_f(prev_dst);
                }
                tc::tmem_st_wait();
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(p_ready);
            }
            // last block product of the unit, then T_i (registers) -> bf16 planes in shared memory (MN-major operand)
            tc::mbar_wait(pv_done, (g - 1) & 1);
            tc::tcgen05_fence_after();
            if (!dead) fold_t();
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                if (cc == 2 && !(has_pos && hsel == 0)) break;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = cc < 2 ? tacc[cc * 8 + i] : tpos[i];
                // T_v chunk (2 hsel + cc) of atom 0; T_pos = chunk 0 of atom 1
                const uint32_t c = cc < 2 ? (uint32_t)(2 * hsel + cc) : 0u;
                const uint32_t off = (cc < 2 ? 0u : (uint32_t)P_SUB) + row_off + ((c ^ sw) << 4);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    uint4 w;
                    w.x = pack2(v[0], v[1]); w.y = pack2(v[2], v[3]); w.z = pack2(v[4], v[5]); w.w = pack2(v[6], v[7]);
                    *reinterpret_cast<uint4*>(t_tile(p) + off) = w;
                    if (p + 1 < P) {
                        v[0] -= __uint_as_float(w.x << 16); v[1] -= __uint_as_float(w.x & 0xffff0000u);
                        v[2] -= __uint_as_float(w.y << 16); v[3] -= __uint_as_float(w.y & 0xffff0000u);
                        v[4] -= __uint_as_float(w.z << 16); v[5] -= __uint_as_float(w.z & 0xffff0000u);
                        v[6] -= __uint_as_float(w.w << 16); v[7] -= __uint_as_float(w.w & 0xffff0000u);
                    }
                }
            }
            tc::fence_proxy_async_smem();
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(t_ready);
            prev_dst = dst;
        }
        if (tt > 0) {                                              // F_t of this CTA's last unit
            tc::mbar_wait(f_done, (tt - 1) & 1);
            tc::tcgen05_fence_after();
            fold_f(prev_dst);
            tc::tcgen05_fence_before();
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// bil[item][e] = sum over the item's RTILES row tiles of part[item * RTILES + t][e], in tile order (deterministic)
__global__ void __launch_bounds__(256) em_reduce_tiles_kernel(const float* __restrict__ part, float* __restrict__ bil,
                                                              long long total, int ww) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long item = idx / ww;
        const int e = (int)(idx - item * ww);
        const float* p = part + item * RTILES * (long long)ww + e;
        float acc = p[0];
#pragma unroll
        for (int t = 1; t < RTILES; ++t) acc += p[(long long)t * ww];
        bil[idx] = acc;
    }
}

// pos [B][576][6] float32 -> pos^T bf16 planes [P][B][8][576] (rows 6, 7 zero): K-major operand of the EM products
__global__ void __launch_bounds__(256) pos_planes_kernel(const float* __restrict__ pos, __nv_bfloat16* __restrict__ out,
                                                         int B, int P) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    const int total = B * 8 * NTOK;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int tok = idx % NTOK, u = (idx / NTOK) & 7, b = idx / (8 * NTOK);
        float v = u < NPOS ? pos[((size_t)b * NTOK + tok) * NPOS + u] : 0.f;
        for (int p = 0; p < P; ++p) {
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            out[(size_t)p * total + idx] = h;
            v -= __bfloat162float(h);
        }
    }
}

int make_map4(CUtensorMap* out, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t d3, cuuint32_t b0,
              cuuint32_t b1, const char* what) {
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    if (!fn) {
        rp::set_error("cuTensorMapEncodeTiled entry point unavailable");
        return RP_EINVAL;
    }
    cuuint64_t gdim[4] = {d0, d1, d2, d3};
    cuuint64_t gstr[3] = {d0 * 2, d0 * d1 * 2, d0 * d1 * d2 * 2};
    cuuint32_t box[4] = {b0, b1, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rp::set_error("%s: tensor map failed (CUresult %d)", what, (int)r);
        return RP_EINVAL;
    }
    return RP_OK;
}

// RELPOSE_EM_V1=1 keeps the round-1 accumulate kernel (A/B measurements)
bool em_use_v1() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RELPOSE_EM_V1");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

template <int P>
int launch_essential(const void* qkv_planes, const float* pos, float* bil, int B, int flags, float* lse2, void* pos_planes,
                     float* part, int device, cudaStream_t st) {
    const int width = pos ? EMW : HD;
    CUtensorMap tmR, tmC, tmPos;
    int rc = make_map4(&tmR, qkv_planes, 3 * EMB, NTOK, 2 * (cuuint64_t)B, P, 64, BM, "rp_essential_tc(rows)");
    if (rc) return rc;
    rc = make_map4(&tmC, qkv_planes, 3 * EMB, NTOK, 2 * (cuuint64_t)B, P, 64, BKV, "rp_essential_tc(cols)");
    if (rc) return rc;
    if (pos) {
        rp::launch(pos_planes_kernel, dim3((B * 8 * NTOK + 255) / 256), dim3(256), (size_t)(0), st, pos, static_cast<__nv_bfloat16*>(pos_planes), B, P);
        rc = rp::finish_launch("rp_essential_tc(pos planes)");
        if (rc) return rc;
    }
    // without positional encodings the map still has to be valid (it is never used by the kernel)
    rc = make_map4(&tmPos, pos ? pos_planes : qkv_planes, pos ? NTOK : 3 * EMB, pos ? 8 : NTOK, pos ? B : 2 * (cuuint64_t)B, P, 64, 8,
                   "rp_essential_tc(pos)");
    if (rc) return rc;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(em_stats_tc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCfg<P>::SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(em_accum_tc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, ECfg<P>::SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(em_accum2_tc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, ECfg<P>::SMEM);
        if (e != cudaSuccess) {
            rp::set_error("rp_essential_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    const int sms = rp::num_sms(device);
    const int n_stats = B * 2 * 2 * HEADS * RTILES, n_acc = B * 2 * HEADS;
    // RELPOSE_EM_SKIP_DEAD=0: the softmax warps of lane quarters past the last token do the full arithmetic (A/B measurements)
    static const int no_skip = [] { const char* e = getenv("RELPOSE_EM_SKIP_DEAD"); return (e && e[0] == '0') ? EM_INTERNAL_NO_SKIP_DEAD : 0; }();
    rp::launch(em_stats_tc_kernel<P>, dim3(n_stats < 2 * sms ? n_stats : 2 * sms), dim3(EM_THREADS), (size_t)(SCfg<P>::SMEM), st, tmR, tmC, lse2, B, no_skip ? 0 : 1);
    rc = rp::finish_launch("rp_essential_tc(stats)");
    if (rc) return rc;
    if (em_use_v1()) {
        rp::launch(em_accum_tc_kernel<P>, dim3(n_acc < sms ? n_acc : sms), dim3(EM_THREADS), (size_t)(ECfg<P>::SMEM), st, tmR, tmC, tmPos, lse2, bil, B, width, flags);
        return rp::finish_launch("rp_essential_tc(accum)");
    }
    const int n_units = n_acc * RTILES;
    rp::launch(em_accum2_tc_kernel<P>, dim3(n_units < sms ? n_units : sms), dim3(EM2_THREADS), (size_t)(ECfg<P>::SMEM), st, tmR, tmC, tmPos, lse2, part, B, width, flags | no_skip);
    rc = rp::finish_launch("rp_essential_tc(accum)");
    if (rc) return rc;
    const long long total = (long long)n_acc * width * width;
    long long blocks = (total + 255) / 256;
    rp::launch(em_reduce_tiles_kernel, dim3((unsigned)(blocks < 4096 ? blocks : 4096)), dim3(256), (size_t)(0), st, part, bil, total, width * width);
    return rp::finish_launch("rp_essential_tc(reduce)");
}

}  // namespace

extern "C" size_t rp_essential_tc_workspace_bytes(int B, int P) {
    if (B <= 0 || P <= 0) return 0;
    size_t lse = (size_t)B * 2 * 2 * HEADS * NTOK * sizeof(float);
    size_t posp = (size_t)P * B * 8 * NTOK * 2;
    size_t part = (size_t)B * 2 * HEADS * RTILES * EMW * EMW * sizeof(float);      // per-row-tile bilinear forms
    return lse + posp + part;
}

extern "C" int rp_essential_ex_tc(const void* qkv_planes, const float* pos, float* bil, int B, int P, int flags, void* workspace,
                                  size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(qkv_planes && bil && B > 0 && (flags & ~(RP_EM_SINGLE_SOFTMAX | RP_EM_CROSS_FEATURES)) == 0, RP_EINVAL,
               "rp_essential_tc: bad argument");
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_essential_tc: P must be 1 (bf16) or 2 (bf16x3)");
    RP_REQUIRE(rp::aligned16(qkv_planes), RP_EALIGN, "rp_essential_tc: qkv planes must be 16-byte aligned");
    RP_REQUIRE(workspace && workspace_bytes >= rp_essential_tc_workspace_bytes(B, P), RP_EWORKSPACE,
               "rp_essential_tc: workspace %zu < %zu bytes", workspace_bytes, rp_essential_tc_workspace_bytes(B, P));
    RP_REQUIRE(rp::aligned16(workspace), RP_EALIGN, "rp_essential_tc: workspace must be 16-byte aligned");
    RP_GUARD(device);
    float* lse2 = static_cast<float*>(workspace);
    char* posp = static_cast<char*>(workspace) + (size_t)B * 2 * 2 * HEADS * NTOK * sizeof(float);
    float* part = reinterpret_cast<float*>(posp + (size_t)P * B * 8 * NTOK * 2);
    if (P == 1) return launch_essential<1>(qkv_planes, pos, bil, B, flags, lse2, posp, part, device, (cudaStream_t)stream);
    return launch_essential<2>(qkv_planes, pos, bil, B, flags, lse2, posp, part, device, (cudaStream_t)stream);
}

extern "C" int rp_essential_tc(const void* qkv_planes, const float* pos, float* bil, int B, int P, void* workspace,
                               size_t workspace_bytes, int device, void* stream) {
    return rp_essential_ex_tc(qkv_planes, pos, bil, B, P, 0, workspace, workspace_bytes, device, stream);
}
