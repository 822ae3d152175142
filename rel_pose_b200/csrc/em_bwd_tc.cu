// Backward of the Essential Matrix Module core (A11: the gradient of vision_transformer.py:198-223) on tcgen05 tensor
// cores, flash style: the dual-softmax matrix A = softmax(S,-1) .* softmax(S,-2) is recomputed tile by tile from q, k and
// the row / column log-sum-exp vectors of the forward pass (rp_essential_tc keeps them in its workspace) -- no 576 x 576
// tensor is written, where the materialised path (train_path.EssentialMaterialisedFn) stores seven of them per step.
//
// Per (pair b, direction d, head h) = "mat":   q = q of image 2b+1-d,  k, v = k, v of image 2b+d,  V' = [v | pos_b] (70 wide)
//   S = q k^T,  R = 2^(S c - lr_i),  C = 2^(S c - lc_j),  A = R C,  T = A V',  F = V'^T T          (forward)
//   dT = V' dF,  dA_ij = <dT_i, V'_j> = <dTv_i, v_j> + <dTp_i, pos_j>      (64 columns on the tensor cores + 6 on the FMA pipe)
//   dr_i = sum_j A dA,  dc_j = sum_i A dA,  dS = 0.125 (2 A dA - R dr_i - C dc_j)
//   dq = dS k,  dk = dS^T q,  dv = (T dF^T)[:, :64] + (A^T dT)[:, :64]                  (the positional columns get no gradient)
// Four launches:
//   em_bwd_prep_kernel   dT = V' dF (small SIMT product), written as the bf16 planes of dTv and the per-token "packs"
//                        packQ[i] = {lr_i, dr_i, dTp_i[6]},  packK[j] = {lc_j, dc_j, pos_j[6]}  (8 floats per token)
//   em_bwd_tc_kernel<0>  pass A.  Row items (rows = queries): S = q k_b^T, X = dTv v_b^T per 64-key block, A and dA in the
//                        compute threads, dr_i, Tp_i = sum_j A pos_j in registers, Tv += A v_b (A planes through tensor
//                        memory).  Column items (rows = keys): the transposed problem, dc_j and dv2 += A^T dTv_b.
//   em_bwd_mid_kernel    dv = Tv dF_vv^T + Tp dF_vp^T + dv2  -> the v columns of d_qkv
//   em_bwd_tc_kernel<1>  pass B.  Row items: dS -> dq += dS k_b; column items: dS^T -> dk += dS^T q_b.
// Rows and columns are symmetric: with "row pack" and "column pack" (the 8 floats above) A = 2^(2 S c - L_row - L_col),
// dA = X + <row6, col6>, dS = 0.125 (2 A dA - 2^(S c - L_row) d_row - 2^(S c - L_col) d_col) hold for both kinds of item;
// only the tensors the operands and packs come from differ.  The column pack of a block travels with the block's operand
// tiles (one 2 KB cp.async.bulk on the same mbarrier).
// Kernel skeleton, tensor-memory budget (256 columns, two CTAs per SM) and barrier protocol: attention_bwd_tc.cu.
#include "tc_common.cuh"

namespace {

constexpr int NTOK = RP_NTOK, HD = RP_HDIM, EMB = RP_EMBED, HEADS = RP_HEADS, EMW = RP_EMW, NPOS = RP_NPOS;
constexpr int P = 2;
constexpr int BM = 128, BN = 64, NBLK = NTOK / BN, TILES = (NTOK + BM - 1) / BM;
constexpr int TILE_BYTES = BM * 128, BLK_BYTES = BN * 128;
constexpr int PACK = 8;                                  // floats per token pack
constexpr int COL_BYTES = BN * PACK * 4;                 // column packs of one block
constexpr int OFF_TA = 0, OFF_TB = OFF_TA + P * TILE_BYTES, OFF_BA = OFF_TB + P * TILE_BYTES, OFF_BB = OFF_BA + P * BLK_BYTES;
constexpr int OFF_COL = OFF_BB + P * BLK_BYTES, OFF_XCH = OFF_COL + COL_BYTES, OFF_BAR = OFF_XCH + 2 * BM * PACK * 4;
constexpr int SMEM = OFF_BAR + 128 + 1024;
constexpr int CTRL_WARPS = 4, COMPUTE_WARPS = 8, THREADS = 32 * (CTRL_WARPS + COMPUTE_WARPS);
constexpr int S_COL = 0, X_COL = 64, ACC_COL = 128, TMEM_COLS = 256;
constexpr int PLANE_COLS = BN / 2;
static_assert(EMW == HD + NPOS && NPOS == 6 && OFF_BAR % 8 == 0 && 2 * SMEM <= 227 * 1024, "layout");

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void store_planes(uint32_t t_dst, uint32_t (&v)[32]) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float v0 = __uint_as_float(v[2 * i]), v1 = __uint_as_float(v[2 * i + 1]);
            w[i] = pack_bf16x2(v0, v1);
            if (p + 1 < P) {
                v[2 * i] = __float_as_uint(v0 - __uint_as_float(w[i] << 16));
                v[2 * i + 1] = __float_as_uint(v1 - __uint_as_float(w[i] & 0xffff0000u));
            }
        }
        tc::tmem_st_32x32b_x8(t_dst + p * PLANE_COLS, *reinterpret_cast<uint32_t(*)[8]>(&w[0]));
        tc::tmem_st_32x32b_x8(t_dst + p * PLANE_COLS + 8, *reinterpret_cast<uint32_t(*)[8]>(&w[8]));
    }
}

// global -> shared bulk copy that completes on an mbarrier (bytes % 16 == 0, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}

struct EmOut {
    float* Tv;      // [mats][576][64]   pass A, row items
    float* Tp;      // [mats][576][8]    pass A, row items (6 used)
    float* dv2;     // [mats][576][64]   pass A, column items
    float* dqkv;    // [2B][576][576]    pass B
};

template <int PASS>
__global__ void __launch_bounds__(THREADS, 2)
em_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                 const __grid_constant__ CUtensorMap tmDT128, const __grid_constant__ CUtensorMap tmDT64,
                 float* __restrict__ packQ, float* __restrict__ packK, EmOut out, int n_mats, float scale_log2, float scale) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* tile_full = bars + 0;
    uint64_t* tile_free = bars + 1;
    uint64_t* blk_full = bars + 2;
    uint64_t* blk_free = bars + 3;
    uint64_t* s_full = bars + 4;
    uint64_t* p_ready = bars + 5;
    uint64_t* acc_full = bars + 6;
    uint64_t* acc_free = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const float* sCol = reinterpret_cast<const float*>(smem + OFF_COL);
    float* xch = reinterpret_cast<float*>(smem + OFF_XCH);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = 2 * n_mats * TILES;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmQKV128);
        tc::prefetch_tmap(&tmQKV64);
        tc::prefetch_tmap(&tmDT128);
        tc::prefetch_tmap(&tmDT64);
        tc::mbar_init(tile_full, 1);
        tc::mbar_init(tile_free, 1);
        tc::mbar_init(blk_full, 1);
        tc::mbar_init(blk_free, 1);
        tc::mbar_init(s_full, 1);
        tc::mbar_init(p_ready, COMPUTE_WARPS);
        tc::mbar_init(acc_full, 1);
        tc::mbar_init(acc_free, COMPUTE_WARPS);
        tc::fence_barrier_init();
    }
    rp::pdl_launch_dependents();
    if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    rp::pdl_wait();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // item -> (kind, mat, row tile); kind 1 = column item (rows = keys).  mat = (b * 2 + d) * 3 + h
    auto decode = [&](int item, int& kv, int& mat, int& h, int& img_q, int& img_k, int& tile) {
        kv = item & 1;
        const int r = item >> 1;
        tile = r % TILES;
        mat = r / TILES;
        h = mat % HEADS;
        const int d = (mat / HEADS) & 1, b = mat / (2 * HEADS);
        img_q = 2 * b + 1 - d;
        img_k = 2 * b + d;
    };

    if (warp < CTRL_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
        if (warp == 0) {
            // ------------------------------------------------------------------------ TMA producer (convergent warp)
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
                int kv, mat, h, img_q, img_k, tile;
                decode(item, kv, mat, h, img_q, img_k, tile);
                // row item:    tiles q (image q), dTv (mat);   blocks k, v (image k);    column packs = packK
                // column item: tiles k, v (image k);           blocks q (image q), dTv;  column packs = packQ
                const CUtensorMap* mTB = kv ? &tmQKV128 : &tmDT128;
                const CUtensorMap* mBB = kv ? &tmDT64 : &tmQKV64;
                const int cTA = kv ? EMB + h * HD : h * HD, iTA = kv ? img_k : img_q;
                const int cTB = kv ? 2 * EMB + h * HD : 0, iTB = kv ? img_k : mat;
                const int cBA = kv ? h * HD : EMB + h * HD, iBA = kv ? img_q : img_k;
                const int cBB = kv ? 0 : 2 * EMB + h * HD, iBB = kv ? mat : img_k;
                const float* cpack = (kv ? packQ : packK) + (size_t)mat * NTOK * PACK;
                tc::mbar_wait(tile_free, (it & 1) ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(tile_full, 2 * P * TILE_BYTES);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        tc::tma_load_4d(smem + OFF_TA + p * TILE_BYTES, &tmQKV128, tile_full, cTA, tile * BM, iTA, p);
                        tc::tma_load_4d(smem + OFF_TB + p * TILE_BYTES, mTB, tile_full, cTB, tile * BM, iTB, p);
                    }
                }
                __syncwarp();
                for (int b = 0; b < NBLK; ++b, ++g) {
                    tc::mbar_wait(blk_free, (g & 1) ^ 1);
                    if (tc::elect_one_sync()) {
                        tc::mbar_expect_tx(blk_full, 2 * P * BLK_BYTES + COL_BYTES);
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            tc::tma_load_4d(smem + OFF_BA + p * BLK_BYTES, &tmQKV64, blk_full, cBA, b * BN, iBA, p);
                            tc::tma_load_4d(smem + OFF_BB + p * BLK_BYTES, mBB, blk_full, cBB, b * BN, iBB, p);
                        }
                        bulk_load(smem + OFF_COL, cpack + (size_t)b * BN * PACK, COL_BYTES, blk_full);
                    }
                    __syncwarp();
                }
            }
        } else if (warp == 1) {
            // ------------------------------------------------------------------------ MMA issuer (convergent warp)
            constexpr uint32_t idesc_s = tc::make_idesc_bf16(BM, BN);
            constexpr uint32_t idesc_p = tc::make_idesc_bf16(BM, HD) | tc::IDESC_B_MN;
            const uint64_t ta0 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_TA));
            const uint64_t ta1 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_TA + TILE_BYTES));
            const uint64_t tb0 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_TB));
            const uint64_t tb1 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_TB + TILE_BYTES));
            const uint64_t ba0 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_BA));
            const uint64_t ba1 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_BA + BLK_BYTES));
            const uint64_t bb0 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_BB));
            const uint64_t bb1 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_BB + BLK_BYTES));
            // the product's B operand, the same block tile read MN-major: pass A = block B (v / dTv), pass B = block A (k / q)
            const int off_m = PASS == 0 ? OFF_BB : OFF_BA;
            const uint64_t m0 = tc::make_mnmajor_sw128_desc(tc::smem_u32(smem + off_m), 0);
            const uint64_t m1 = tc::make_mnmajor_sw128_desc(tc::smem_u32(smem + off_m + BLK_BYTES), 0);
            const uint32_t dS = tmem_base + S_COL, dX = tmem_base + X_COL, dAcc = tmem_base + ACC_COL;
            const uint32_t aT = PASS == 0 ? dS : dX;                 // planes of A (pass A) over S, of dS (pass B) over X
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
                tc::mbar_wait(tile_full, it & 1);
                for (int b = 0; b < NBLK; ++b, ++g) {
                    tc::mbar_wait(blk_full, g & 1);
                    tc::tcgen05_fence_after();
                    if (tc::elect_one_sync()) {
                        uint32_t accum = 0u;
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) {
                            tc::umma_bf16(dS, ta1 + 2 * k, ba0 + 2 * k, idesc_s, accum);
                            tc::umma_bf16(dS, ta0 + 2 * k, ba1 + 2 * k, idesc_s, 1u);
                            accum = 1u;
                        }
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) tc::umma_bf16(dS, ta0 + 2 * k, ba0 + 2 * k, idesc_s, 1u);
                        accum = 0u;
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) {
                            tc::umma_bf16(dX, tb1 + 2 * k, bb0 + 2 * k, idesc_s, accum);
                            tc::umma_bf16(dX, tb0 + 2 * k, bb1 + 2 * k, idesc_s, 1u);
                            accum = 1u;
                        }
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) tc::umma_bf16(dX, tb0 + 2 * k, bb0 + 2 * k, idesc_s, 1u);
                        tc::umma_commit(s_full);
                    }
                    __syncwarp();
                    tc::mbar_wait(p_ready, g & 1);
                    if (b == 0) tc::mbar_wait(acc_free, (it & 1) ^ 1);
                    tc::tcgen05_fence_after();
                    if (tc::elect_one_sync()) {
                        uint32_t accum = b == 0 ? 0u : 1u;
#pragma unroll
                        for (int kk = 0; kk < BN / 16; ++kk) {
                            const uint32_t b_off = (kk * 16 * 128) >> 4;
                            tc::umma_bf16_ts(dAcc, aT + PLANE_COLS + 8 * kk, m0 + b_off, idesc_p, accum);
                            tc::umma_bf16_ts(dAcc, aT + 8 * kk, m1 + b_off, idesc_p, 1u);
                            accum = 1u;
                        }
#pragma unroll
                        for (int kk = 0; kk < BN / 16; ++kk)
                            tc::umma_bf16_ts(dAcc, aT + 8 * kk, m0 + ((kk * 16 * 128) >> 4), idesc_p, 1u);
                        tc::umma_commit(blk_free);
                        if (b + 1 == NBLK) {
                            tc::umma_commit(acc_full);
                            tc::umma_commit(tile_free);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
        // ---------------------------------------------------------------------------- compute warps
        const int quarter = warp & 3;
        const int hsel = (warp - CTRL_WARPS) >> 2;
        const int r = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int bar_id = 1 + quarter;
        uint32_t g = 0;
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            int kv, mat, h, img_q, img_k, tile;
            decode(item, kv, mat, h, img_q, img_k, tile);
            const int row = tile * BM + r;
            float* rpack = (kv ? packK : packQ) + ((size_t)mat * NTOK + (row < NTOK ? row : 0)) * PACK;
            float Lrow = 0.f, drow = 0.f, r6[NPOS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (row < NTOK) {
                const float4 p0 = *reinterpret_cast<const float4*>(rpack), p1 = *reinterpret_cast<const float4*>(rpack + 4);
                Lrow = p0.x; drow = p0.y;
                r6[0] = p0.z; r6[1] = p0.w; r6[2] = p1.x; r6[3] = p1.y; r6[4] = p1.z; r6[5] = p1.w;
            }
            float rowsum = 0.f, tp[NPOS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int b = 0; b < NBLK; ++b, ++g) {
                tc::mbar_wait(s_full, g & 1);
                tc::tcgen05_fence_after();
                uint32_t s[32], x[32];
                tc::tmem_ld_32x32b_x32(t_lane + S_COL + hsel * 32, s);
                tc::tmem_ld_32x32b_x32(t_lane + X_COL + hsel * 32, x);
                tc::tmem_ld_wait();
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                const float4* cp = reinterpret_cast<const float4*>(sCol + hsel * 32 * PACK);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float4 c0 = cp[2 * i], c1 = cp[2 * i + 1];             // {L, d, c6[0], c6[1]}, {c6[2..5]}
                    const float sv = __uint_as_float(s[i]);
                    float dA = __uint_as_float(x[i]);
                    dA = fmaf(r6[0], c0.z, dA); dA = fmaf(r6[1], c0.w, dA); dA = fmaf(r6[2], c1.x, dA);
                    dA = fmaf(r6[3], c1.y, dA); dA = fmaf(r6[4], c1.z, dA); dA = fmaf(r6[5], c1.w, dA);
                    if (PASS == 0) {
                        const float a = tc::fast_exp2(fmaf(sv, 2.0f * scale_log2, -(Lrow + c0.x)));
                        rowsum = fmaf(a, dA, rowsum);
                        if (!kv) {                                               // Tp_i = sum_j A_ij pos_j
                            tp[0] = fmaf(a, c0.z, tp[0]); tp[1] = fmaf(a, c0.w, tp[1]); tp[2] = fmaf(a, c1.x, tp[2]);
                            tp[3] = fmaf(a, c1.y, tp[3]); tp[4] = fmaf(a, c1.z, tp[4]); tp[5] = fmaf(a, c1.w, tp[5]);
                        }
                        s[i] = __float_as_uint(a);
                    } else {
                        const float R = tc::fast_exp2(fmaf(sv, scale_log2, -Lrow));
                        const float C = tc::fast_exp2(fmaf(sv, scale_log2, -c0.x));
                        const float a = R * C;
                        x[i] = __float_as_uint(scale * (2.0f * a * dA - R * drow - C * c0.y));
                    }
                }
                if (PASS == 0) store_planes(t_lane + S_COL + hsel * (PLANE_COLS / 2), s);
                else store_planes(t_lane + X_COL + hsel * (PLANE_COLS / 2), x);
                tc::tmem_st_wait();
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(p_ready);
            }
            if (PASS == 0) {
                // the two threads of a row combine their partial sums: d_row into the row pack, Tp (row items)
                float* mine = xch + ((size_t)hsel * BM + r) * PACK;
                mine[0] = rowsum;
#pragma unroll
                for (int c = 0; c < NPOS; ++c) mine[1 + c] = tp[c];
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                if (hsel == 0 && row < NTOK) {
                    const float* other = xch + ((size_t)BM + r) * PACK;
                    rpack[1] = rowsum + other[0];
                    if (!kv) {
                        float* tpo = out.Tp + ((size_t)mat * NTOK + row) * PACK;
#pragma unroll
                        for (int c = 0; c < NPOS; ++c) tpo[c] = tp[c] + other[1 + c];
                    }
                }
            }
            // ------------------------------------------------------------------------ epilogue: accumulator -> output rows
            tc::mbar_wait(acc_full, it & 1);
            tc::tcgen05_fence_after();
            {
                uint32_t a[32];
                tc::tmem_ld_32x32b_x32(t_lane + ACC_COL + hsel * 32, a);
                tc::tmem_ld_wait();
                if (row < NTOK) {
                    float* dst;
                    if (PASS == 0) dst = (kv ? out.dv2 : out.Tv) + ((size_t)mat * NTOK + row) * HD + hsel * 32;
                    else dst = out.dqkv + ((size_t)(kv ? img_k : img_q) * NTOK + row) * (3 * EMB) + (kv ? EMB : 0) + h * HD + hsel * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(dst + i) = make_float4(__uint_as_float(a[i]), __uint_as_float(a[i + 1]),
                                                                           __uint_as_float(a[i + 2]), __uint_as_float(a[i + 3]));
                }
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(acc_free);
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// dT = V' dF for 64 token rows of one mat; writes the bf16 planes of dTv, packQ = {lr, 0, dTp}, packK = {lc, 0, pos}
__global__ void __launch_bounds__(256)
em_bwd_prep_kernel(const __nv_bfloat16* __restrict__ qkv_planes, const float* __restrict__ pos, const float* __restrict__ lse2,
                   const float* __restrict__ dF, __nv_bfloat16* __restrict__ dT_planes, float* __restrict__ packQ,
                   float* __restrict__ packK, int n_img, int n_mats) {
    __shared__ float sF[EMW][EMW + 1];
    __shared__ float sV[64][EMW + 1];
    const int mat = blockIdx.x, row0 = blockIdx.y * 64;
    const int h = mat % HEADS, d = (mat / HEADS) & 1, b = mat / (2 * HEADS);
    const int img_k = 2 * b + d;
    for (int i = threadIdx.x; i < EMW * EMW; i += 256) sF[i / EMW][i % EMW] = dF[(size_t)mat * EMW * EMW + i];
    const size_t qplane = (size_t)n_img * NTOK * 3 * EMB;
    for (int i = threadIdx.x; i < 64 * EMW; i += 256) {
        const int r = i / EMW, c = i % EMW, row = row0 + r;
        float v;
        if (c < HD) {
            const size_t idx = ((size_t)img_k * NTOK + row) * (3 * EMB) + 2 * EMB + h * HD + c;
            v = __bfloat162float(qkv_planes[idx]) + __bfloat162float(qkv_planes[qplane + idx]);
        } else {
            v = pos[((size_t)b * NTOK + row) * NPOS + (c - HD)];
        }
        sV[r][c] = v;
    }
    __syncthreads();
    const size_t tplane = (size_t)n_mats * NTOK * HD;
    for (int i = threadIdx.x; i < 64 * (EMW / 2); i += 256) {
        const int r = i / (EMW / 2), c = 2 * (i % (EMW / 2)), row = row0 + r;
        float t0 = 0.f, t1 = 0.f;
#pragma unroll 10
        for (int a = 0; a < EMW; ++a) {
            const float va = sV[r][a];
            t0 = fmaf(va, sF[a][c], t0);
            t1 = fmaf(va, sF[a][c + 1], t1);
        }
        if (c < HD) {
            const size_t idx = ((size_t)mat * NTOK + row) * HD + c;
            const uint32_t w0 = pack_bf16x2(t0, t1);
            const uint32_t w1 = pack_bf16x2(t0 - __uint_as_float(w0 << 16), t1 - __uint_as_float(w0 & 0xffff0000u));
            *reinterpret_cast<uint32_t*>(dT_planes + idx) = w0;
            *reinterpret_cast<uint32_t*>(dT_planes + tplane + idx) = w1;
        } else {
            float* q = packQ + ((size_t)mat * NTOK + row) * PACK + 2 + (c - HD);
            q[0] = t0;
            q[1] = t1;
        }
    }
    if (threadIdx.x < 64) {
        const int row = row0 + threadIdx.x;
        const int bd = mat / HEADS;                                           // b * 2 + d
        const float* lr = lse2 + (((size_t)bd * 2 + 0) * HEADS + h) * NTOK;
        const float* lc = lse2 + (((size_t)bd * 2 + 1) * HEADS + h) * NTOK;
        float* q = packQ + ((size_t)mat * NTOK + row) * PACK;
        float* k = packK + ((size_t)mat * NTOK + row) * PACK;
        q[0] = lr[row];
        q[1] = 0.f;
        k[0] = lc[row];
        k[1] = 0.f;
#pragma unroll
        for (int c = 0; c < NPOS; ++c) k[2 + c] = sV[threadIdx.x][HD + c];
    }
}

// dv = (T dF^T)[:, :64] + dv2 for 64 token rows of one mat -> the v columns of image 2b+d in d_qkv
__global__ void __launch_bounds__(256)
em_bwd_mid_kernel(const float* __restrict__ Tv, const float* __restrict__ Tp, const float* __restrict__ dv2,
                  const float* __restrict__ dF, float* __restrict__ dqkv) {
    __shared__ float sF[HD][EMW + 1];
    __shared__ float sT[64][EMW + 1];
    const int mat = blockIdx.x, row0 = blockIdx.y * 64;
    const int h = mat % HEADS, d = (mat / HEADS) & 1, b = mat / (2 * HEADS);
    const int img_k = 2 * b + d;
    for (int i = threadIdx.x; i < HD * EMW; i += 256) sF[i / EMW][i % EMW] = dF[(size_t)mat * EMW * EMW + i];   // rows a < 64
    for (int i = threadIdx.x; i < 64 * EMW; i += 256) {
        const int r = i / EMW, c = i % EMW;
        const size_t tok = (size_t)mat * NTOK + row0 + r;
        sT[r][c] = c < HD ? Tv[tok * HD + c] : Tp[tok * PACK + (c - HD)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * HD; i += 256) {
        const int r = i / HD, a = i % HD;
        float acc = 0.f;
#pragma unroll 10
        for (int c = 0; c < EMW; ++c) acc = fmaf(sT[r][c], sF[a][c], acc);
        const size_t tok = (size_t)mat * NTOK + row0 + r;
        dqkv[((size_t)img_k * NTOK + row0 + r) * (3 * EMB) + 2 * EMB + h * HD + a] = acc + dv2[tok * HD + a];
    }
}

int make_tok_tmap(CUtensorMap* out, const void* base, int ld, int n_img, int box_rows) {
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    if (!fn) {
        rp::set_error("cuTensorMapEncodeTiled entry point unavailable");
        return RP_EINVAL;
    }
    cuuint64_t gdim[4] = {(cuuint64_t)ld, (cuuint64_t)NTOK, (cuuint64_t)n_img, (cuuint64_t)P};
    cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * NTOK * 2, (cuuint64_t)ld * NTOK * 2 * (cuuint64_t)n_img};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rp::set_error("rp_em_bwd_tc: tensor map failed (CUresult %d) ld=%d n=%d box_rows=%d", (int)r, ld, n_img, box_rows);
        return RP_EINVAL;
    }
    return RP_OK;
}

struct Ws {
    size_t dT, packQ, packK, Tv, Tp, dv2, total;
    explicit Ws(int B) {
        const size_t mats = (size_t)B * 2 * HEADS, tok = mats * NTOK;
        dT = 0;
        packQ = dT + (size_t)P * tok * HD * 2;
        packK = packQ + tok * PACK * 4;
        Tv = packK + tok * PACK * 4;
        Tp = Tv + tok * HD * 4;
        dv2 = Tp + tok * PACK * 4;
        total = dv2 + tok * HD * 4;
    }
};

}  // namespace

extern "C" size_t rp_em_bwd_tc_workspace_bytes(int B) { return B > 0 ? Ws(B).total : 0; }

extern "C" int rp_em_bwd_tc(const void* qkv_planes, const float* pos, const float* lse2, const float* d_bil, float* d_qkv, int B,
                            void* workspace, size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(qkv_planes && pos && lse2 && d_bil && d_qkv && workspace && B > 0, RP_EINVAL, "rp_em_bwd_tc: bad argument");
    const Ws w(B);
    RP_REQUIRE(workspace_bytes >= w.total, RP_EWORKSPACE, "rp_em_bwd_tc: workspace %zu < %zu bytes", workspace_bytes, w.total);
    RP_REQUIRE(rp::aligned16(qkv_planes) && rp::aligned16(d_qkv) && rp::aligned16(workspace), RP_EALIGN, "rp_em_bwd_tc: 16-byte alignment");
    RP_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int n_img = 2 * B, mats = B * 2 * HEADS;
    char* base = static_cast<char*>(workspace);
    __nv_bfloat16* dT = reinterpret_cast<__nv_bfloat16*>(base + w.dT);
    float* packQ = reinterpret_cast<float*>(base + w.packQ);
    float* packK = reinterpret_cast<float*>(base + w.packK);
    EmOut out{reinterpret_cast<float*>(base + w.Tv), reinterpret_cast<float*>(base + w.Tp), reinterpret_cast<float*>(base + w.dv2), d_qkv};

    em_bwd_prep_kernel<<<dim3(mats, NTOK / 64), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(qkv_planes), pos, lse2, d_bil, dT, packQ,
                                                              packK, n_img, mats);
    int rc = rp::finish_launch("rp_em_bwd_tc (prep)");
    if (rc) return rc;

    CUtensorMap tmQKV128, tmQKV64, tmDT128, tmDT64;
    if ((rc = make_tok_tmap(&tmQKV128, qkv_planes, 3 * EMB, n_img, BM))) return rc;
    if ((rc = make_tok_tmap(&tmQKV64, qkv_planes, 3 * EMB, n_img, BN))) return rc;
    if ((rc = make_tok_tmap(&tmDT128, dT, HD, mats, BM))) return rc;
    if ((rc = make_tok_tmap(&tmDT64, dT, HD, mats, BN))) return rc;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(em_bwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(em_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) {
            rp::set_error("rp_em_bwd_tc: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    const int nitems = 2 * mats * TILES;
    const int slots = 2 * rp::num_sms(device);
    const int grid = nitems < slots ? nitems : slots;
    const float scale = 0.125f;
    // the prep kernel was launched without the programmatic attribute: the pass-A grid waits for it in pdl_wait()
    rp::launch(em_bwd_tc_kernel<0>, dim3(grid), dim3(THREADS), (size_t)SMEM, st, tmQKV128, tmQKV64, tmDT128, tmDT64, packQ, packK, out,
               mats, scale * 1.4426950408889634f, scale);
    if ((rc = rp::finish_launch("rp_em_bwd_tc (pass A)"))) return rc;
    em_bwd_mid_kernel<<<dim3(mats, NTOK / 64), 256, 0, st>>>(out.Tv, out.Tp, out.dv2, d_bil, d_qkv);
    if ((rc = rp::finish_launch("rp_em_bwd_tc (mid)"))) return rc;
    rp::launch(em_bwd_tc_kernel<1>, dim3(grid), dim3(THREADS), (size_t)SMEM, st, tmQKV128, tmQKV64, tmDT128, tmDT64, packQ, packK, out,
               mats, scale * 1.4426950408889634f, scale);
    return rp::finish_launch("rp_em_bwd_tc (pass B)");
}
