// Backward of the fused self-attention (A11: the gradient of vision_transformer.py:321-329) on tcgen05 tensor cores,
// flash style: like the forward kernel (attention_tc.cu) nothing of size 576x576 is written to memory -- the
// probabilities are RECOMPUTED from q, k and the log-sum-exp vector the forward kernel saved.
//
//   S = q k^T,  P = 2^(S c - L)  (c = 0.125 log2 e, L_i = log2-sum-exp of row i),  O = P v
//   D_i  = <dO_i, O_i>                              (attn_bwd_prep_kernel, together with the bf16 planes of dO)
//   dP   = dO v^T,   dS = 0.125 P .* (dP - D)
//   dv   = P^T dO,   dk = dS^T q,   dq = dS k
//
// Two kinds of work item share one persistent kernel (item parity selects the kind, so both fill the machine together):
//   KV item (image, head, 128-key tile):   rows = keys.   S^T = k q_b^T, dP^T = v dO_b^T per 64-query block b (two
//            M128 x N64 x K64 products, both operands K-major from shared memory), P^T and dS^T re-split into bf16 planes
//            and written back over S^T / dP^T in TENSOR memory, then dv += P^T dO_b and dk += dS^T q_b (A from tensor
//            memory, B = the same q_b / dO_b tiles read as MN-major operands).  L and D are per COLUMN here: the item's
//            576 values of each sit in shared memory.
//   Q item  (image, head, 128-query tile): rows = queries.  S = q k_b^T, dP = dO v_b^T per 64-key block, dq += dS k_b.
//            L and D are per ROW (two registers).
// Every element of dqkv [n,576,576] is written exactly once (q columns by the Q items, k and v columns by the KV items):
// deterministic, no atomics, no zero-fill.  Operands are split-bf16 planes (x = x0 + x1, products a0 b0 + a0 b1 + a1 b0
// into one fp32 accumulator, correction terms first) = fp32 class, like the forward.
//
// CTA = 384 threads, two CTAs per SM (256 tensor-memory columns, 101 KB of shared memory each):
//   warp 0      TMA producer: the item's two row tiles (once per item), then the two 64-row blocks of every step
//   warp 1      MMA issuer + tensor-memory owner
//   warps 4-11  two threads per row (32 columns each): tcgen05.ld of S and dP, exponentials, plane split, tcgen05.st;
//               at the end of an item the epilogue (accumulators -> dqkv)
// One block is in flight per CTA (S -> P/dS -> products is a dependent chain and the second CTA of the SM fills the gaps);
// the products of block b release the block buffers to the TMA warp, so S of block b+1 can never overwrite P of block b
// while it is being read.
#include "tc_common.cuh"

namespace {

constexpr int NTOK = RP_NTOK, HD = RP_HDIM, EMB = RP_EMBED, HEADS = RP_HEADS;
constexpr int P = 2;                                   // split-bf16 planes (the training path is fp32 class only)
constexpr int BM = 128, BN = 64, NBLK = NTOK / BN, TILES = (NTOK + BM - 1) / BM;   // 9 blocks, 5 row tiles
constexpr int TILE_BYTES = BM * 128, BLK_BYTES = BN * 128;                         // one plane of a [rows x 64] bf16 tile
constexpr int OFF_TA = 0, OFF_TB = OFF_TA + P * TILE_BYTES, OFF_BA = OFF_TB + P * TILE_BYTES, OFF_BB = OFF_BA + P * BLK_BYTES;
constexpr int OFF_L = OFF_BB + P * BLK_BYTES, OFF_D = OFF_L + NTOK * 4, OFF_BAR = OFF_D + NTOK * 4;
constexpr int SMEM = OFF_BAR + 128 + 1024 /*align slack*/;
constexpr int CTRL_WARPS = 4, COMPUTE_WARPS = 8, THREADS = 32 * (CTRL_WARPS + COMPUTE_WARPS);
constexpr int S_COL = 0, DP_COL = 64, ACC1_COL = 128, ACC2_COL = 192, TMEM_COLS = 256;
constexpr int PLANE_COLS = BN / 2;                     // one bf16 plane of a 64-wide A operand = 32 columns
static_assert(NTOK % BN == 0 && OFF_BAR % 8 == 0 && 2 * SMEM <= 227 * 1024, "layout");

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// v[0..31] (fp32 bit patterns) -> P bf16 planes of this thread's 32 columns = 16 packed columns per plane at t_dst
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

__global__ void __launch_bounds__(THREADS, 2)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV128, const __grid_constant__ CUtensorMap tmQKV64,
                        const __grid_constant__ CUtensorMap tmDO128, const __grid_constant__ CUtensorMap tmDO64,
                        const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dqkv,
                        int n_img, float scale_log2, float scale) {
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
    float* sL = reinterpret_cast<float*>(smem + OFF_L);
    float* sD = reinterpret_cast<float*>(smem + OFF_D);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = 2 * n_img * HEADS * TILES;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmQKV128);
        tc::prefetch_tmap(&tmQKV64);
        tc::prefetch_tmap(&tmDO128);
        tc::prefetch_tmap(&tmDO64);
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

    // item -> (kind, image, head, row tile); kind 1 = KV item
    auto decode = [&](int item, int& kv, int& img, int& h, int& tile) {
        kv = item & 1;
        const int r = item >> 1;
        tile = r % TILES;
        h = (r / TILES) % HEADS;
        img = r / (TILES * HEADS);
    };

    if (warp < CTRL_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
        if (warp == 0) {
            // ------------------------------------------------------------------------ TMA producer (convergent warp)
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
                int kv, img, h, tile;
                decode(item, kv, img, h, tile);
                // KV item: tiles k, v; blocks q, dO.   Q item: tiles q, dO; blocks k, v.
                const CUtensorMap* mTB = kv ? &tmQKV128 : &tmDO128;
                const CUtensorMap* mBB = kv ? &tmDO64 : &tmQKV64;
                const int cTA = kv ? EMB + h * HD : h * HD;
                const int cTB = kv ? 2 * EMB + h * HD : h * HD;
                const int cBA = kv ? h * HD : EMB + h * HD;
                const int cBB = kv ? h * HD : 2 * EMB + h * HD;
                tc::mbar_wait(tile_free, (it & 1) ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(tile_full, 2 * P * TILE_BYTES);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        tc::tma_load_4d(smem + OFF_TA + p * TILE_BYTES, &tmQKV128, tile_full, cTA, tile * BM, img, p);
                        tc::tma_load_4d(smem + OFF_TB + p * TILE_BYTES, mTB, tile_full, cTB, tile * BM, img, p);
                    }
                }
                __syncwarp();
                for (int b = 0; b < NBLK; ++b, ++g) {
                    tc::mbar_wait(blk_free, (g & 1) ^ 1);
                    if (tc::elect_one_sync()) {
                        tc::mbar_expect_tx(blk_full, 2 * P * BLK_BYTES);
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            tc::tma_load_4d(smem + OFF_BA + p * BLK_BYTES, &tmQKV64, blk_full, cBA, b * BN, img, p);
                            tc::tma_load_4d(smem + OFF_BB + p * BLK_BYTES, mBB, blk_full, cBB, b * BN, img, p);
                        }
                    }
                    __syncwarp();
                }
            }
        } else if (warp == 1) {
            // ------------------------------------------------------------------------ MMA issuer (convergent warp)
            constexpr uint32_t idesc_s = tc::make_idesc_bf16(BM, BN);                       // A, B K-major
            constexpr uint32_t idesc_p = tc::make_idesc_bf16(BM, HD) | tc::IDESC_B_MN;      // A tensor memory, B MN-major
            const uint64_t ta0 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_TA));
            const uint64_t ta1 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_TA + TILE_BYTES));
            const uint64_t tb0 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_TB));
            const uint64_t tb1 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_TB + TILE_BYTES));
            const uint64_t ba0 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_BA));
            const uint64_t ba1 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_BA + BLK_BYTES));
            const uint64_t bb0 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_BB));
            const uint64_t bb1 = tc::make_kmajor_sw128_desc(tc::smem_u32(smem + OFF_BB + BLK_BYTES));
            // the same block tiles as MN-major operands ([row = contraction index][64 columns])
            const uint64_t ma0 = tc::make_mnmajor_sw128_desc(tc::smem_u32(smem + OFF_BA), 0);
            const uint64_t ma1 = tc::make_mnmajor_sw128_desc(tc::smem_u32(smem + OFF_BA + BLK_BYTES), 0);
            const uint64_t mb0 = tc::make_mnmajor_sw128_desc(tc::smem_u32(smem + OFF_BB), 0);
            const uint64_t mb1 = tc::make_mnmajor_sw128_desc(tc::smem_u32(smem + OFF_BB + BLK_BYTES), 0);
            const uint32_t dS = tmem_base + S_COL, dDP = tmem_base + DP_COL, dA1 = tmem_base + ACC1_COL, dA2 = tmem_base + ACC2_COL;
            uint32_t g = 0;
            int it = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
                const int kv = item & 1;
                tc::mbar_wait(tile_full, it & 1);
                for (int b = 0; b < NBLK; ++b, ++g) {
                    tc::mbar_wait(blk_full, g & 1);
                    tc::tcgen05_fence_after();
                    if (tc::elect_one_sync()) {
                        // S (or S^T) = tile A . block A^T,  dP (or dP^T) = tile B . block B^T:  K = 64 = 4 steps of 16
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
                            tc::umma_bf16(dDP, tb1 + 2 * k, bb0 + 2 * k, idesc_s, accum);
                            tc::umma_bf16(dDP, tb0 + 2 * k, bb1 + 2 * k, idesc_s, 1u);
                            accum = 1u;
                        }
#pragma unroll
                        for (int k = 0; k < HD / 16; ++k) tc::umma_bf16(dDP, tb0 + 2 * k, bb0 + 2 * k, idesc_s, 1u);
                        tc::umma_commit(s_full);
                    }
                    __syncwarp();
                    tc::mbar_wait(p_ready, g & 1);
                    if (b == 0) tc::mbar_wait(acc_free, (it & 1) ^ 1);      // the previous item's epilogue has read the accumulators
                    tc::tcgen05_fence_after();
                    if (tc::elect_one_sync()) {
                        const uint32_t first = b == 0 ? 0u : 1u;
                        // acc1 += dS . block A (dq / dk) ... for the KV item acc2 takes it and acc1 += P^T . block B (dv)
                        const uint32_t d_ds = kv ? dA2 : dA1;
                        uint32_t accum = first;
#pragma unroll
                        for (int kk = 0; kk < BN / 16; ++kk) {
                            const uint32_t b_off = (kk * 16 * 128) >> 4;
                            tc::umma_bf16_ts(d_ds, dDP + PLANE_COLS + 8 * kk, ma0 + b_off, idesc_p, accum);
                            tc::umma_bf16_ts(d_ds, dDP + 8 * kk, ma1 + b_off, idesc_p, 1u);
                            accum = 1u;
                        }
#pragma unroll
                        for (int kk = 0; kk < BN / 16; ++kk)
                            tc::umma_bf16_ts(d_ds, dDP + 8 * kk, ma0 + ((kk * 16 * 128) >> 4), idesc_p, 1u);
                        if (kv) {
                            accum = first;
#pragma unroll
                            for (int kk = 0; kk < BN / 16; ++kk) {
                                const uint32_t b_off = (kk * 16 * 128) >> 4;
                                tc::umma_bf16_ts(dA1, dS + PLANE_COLS + 8 * kk, mb0 + b_off, idesc_p, accum);
                                tc::umma_bf16_ts(dA1, dS + 8 * kk, mb1 + b_off, idesc_p, 1u);
                                accum = 1u;
                            }
#pragma unroll
                            for (int kk = 0; kk < BN / 16; ++kk)
                                tc::umma_bf16_ts(dA1, dS + 8 * kk, mb0 + ((kk * 16 * 128) >> 4), idesc_p, 1u);
                        }
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
        const int quarter = warp & 3;                      // tensor-memory lane quarter this warp may touch
        const int hsel = (warp - CTRL_WARPS) >> 2;         // which 32 of the 64 columns
        const int r = quarter * 32 + lane;                 // row inside the tile
        const int ctid = threadIdx.x - 32 * CTRL_WARPS;    // 0..255
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int bar_id = 1 + quarter;
        uint32_t g = 0;
        int it = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++it) {
            int kv, img, h, tile;
            decode(item, kv, img, h, tile);
            const int row = tile * BM + r;
            const size_t vec = ((size_t)img * HEADS + h) * NTOK;
            float Lr = 0.f, Dr = 0.f;
            if (kv) {
                asm volatile("bar.sync 5, 256;" ::: "memory");          // every warp is done with the previous item's vectors
                for (int i = ctid; i < NTOK; i += 32 * COMPUTE_WARPS) {
                    sL[i] = lse[vec + i];
                    sD[i] = delta[vec + i];
                }
                asm volatile("bar.sync 5, 256;" ::: "memory");
            } else if (row < NTOK) {
                Lr = lse[vec + row];
                Dr = delta[vec + row];
            }
            for (int b = 0; b < NBLK; ++b, ++g) {
                tc::mbar_wait(s_full, g & 1);
                tc::tcgen05_fence_after();
                uint32_t s[32], dp[32];
                tc::tmem_ld_32x32b_x32(t_lane + S_COL + hsel * 32, s);
                tc::tmem_ld_32x32b_x32(t_lane + DP_COL + hsel * 32, dp);
                tc::tmem_ld_wait();
                // both threads of a row have their S / dP columns in registers before either overwrites them with planes
                asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
                if (kv) {
                    const float* Lc = sL + b * BN + hsel * 32;
                    const float* Dc = sD + b * BN + hsel * 32;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float p = tc::fast_exp2(fmaf(__uint_as_float(s[i]), scale_log2, -Lc[i]));
                        s[i] = __float_as_uint(p);
                        dp[i] = __float_as_uint(scale * p * (__uint_as_float(dp[i]) - Dc[i]));
                    }
                    store_planes(t_lane + S_COL + hsel * (PLANE_COLS / 2), s);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float p = tc::fast_exp2(fmaf(__uint_as_float(s[i]), scale_log2, -Lr));
                        dp[i] = __float_as_uint(scale * p * (__uint_as_float(dp[i]) - Dr));
                    }
                }
                store_planes(t_lane + DP_COL + hsel * (PLANE_COLS / 2), dp);
                tc::tmem_st_wait();
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(p_ready);
            }
            // ------------------------------------------------------------------------ epilogue: accumulators -> dqkv
            tc::mbar_wait(acc_full, it & 1);
            tc::tcgen05_fence_after();
            float* dst = dqkv + ((size_t)img * NTOK + (row < NTOK ? row : 0)) * (3 * EMB) + h * HD + hsel * 32;
            {
                uint32_t a[32];
                tc::tmem_ld_32x32b_x32(t_lane + ACC1_COL + hsel * 32, a);
                tc::tmem_ld_wait();
                if (row < NTOK) {
                    float* d1 = dst + (kv ? 2 * EMB : 0);               // dv | dq
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(d1 + i) = make_float4(__uint_as_float(a[i]), __uint_as_float(a[i + 1]),
                                                                          __uint_as_float(a[i + 2]), __uint_as_float(a[i + 3]));
                }
            }
            if (kv) {
                uint32_t a[32];
                tc::tmem_ld_32x32b_x32(t_lane + ACC2_COL + hsel * 32, a);
                tc::tmem_ld_wait();
                if (row < NTOK) {
                    float* d2 = dst + EMB;                               // dk
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(d2 + i) = make_float4(__uint_as_float(a[i]), __uint_as_float(a[i + 1]),
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

// D[img][h][tok] = <dO, O> over the head's 64 columns, and the bf16 planes of dO.  One warp per token row.
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const float* __restrict__ dO, const float* __restrict__ O, __nv_bfloat16* __restrict__ dO_planes,
                     float* __restrict__ delta, long long rows) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const long long img = row / NTOK;
    const int tok = (int)(row % NTOK);
    const size_t plane = (size_t)rows * EMB;
#pragma unroll
    for (int h = 0; h < HEADS; ++h) {
        const size_t idx = (size_t)row * EMB + h * HD + 2 * lane;
        const float2 a = *reinterpret_cast<const float2*>(dO + idx);
        const float2 o = *reinterpret_cast<const float2*>(O + idx);
        const float d = rp::warp_sum(fmaf(a.x, o.x, a.y * o.y));
        if (lane == 0) delta[((size_t)img * HEADS + h) * NTOK + tok] = d;
        const uint32_t w0 = pack_bf16x2(a.x, a.y);
        const uint32_t w1 = pack_bf16x2(a.x - __uint_as_float(w0 << 16), a.y - __uint_as_float(w0 & 0xffff0000u));
        *reinterpret_cast<uint32_t*>(dO_planes + idx) = w0;
        *reinterpret_cast<uint32_t*>(dO_planes + plane + idx) = w1;
    }
}

// bf16 [P][n_img][576][ld] planes, box = [1][1][box_rows][64 columns], 128-byte swizzle, OOB rows -> 0
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
        rp::set_error("attention backward tensor map failed (CUresult %d) ld=%d n_img=%d box_rows=%d", (int)r, ld, n_img, box_rows);
        return RP_EINVAL;
    }
    return RP_OK;
}

}  // namespace

extern "C" int rp_attention_bwd_prep(const float* d_out, const float* out, void* d_out_planes, float* delta, int n_img,
                                     int device, void* stream) {
    RP_REQUIRE(d_out && out && d_out_planes && delta && n_img > 0, RP_EINVAL, "rp_attention_bwd_prep: bad argument");
    RP_REQUIRE(rp::aligned16(d_out) && rp::aligned16(out) && rp::aligned16(d_out_planes), RP_EALIGN,
               "rp_attention_bwd_prep: 16-byte alignment");
    RP_GUARD(device);
    const long long rows = (long long)n_img * NTOK;
    attn_bwd_prep_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        d_out, out, static_cast<__nv_bfloat16*>(d_out_planes), delta, rows);
    return rp::finish_launch("rp_attention_bwd_prep");
}

extern "C" int rp_attention_bwd_tc(const void* qkv_planes, const void* d_out_planes, const float* lse, const float* delta,
                                   float* d_qkv, int n_img, int device, void* stream) {
    RP_REQUIRE(qkv_planes && d_out_planes && lse && delta && d_qkv && n_img > 0, RP_EINVAL, "rp_attention_bwd_tc: bad argument");
    RP_REQUIRE(rp::aligned16(qkv_planes) && rp::aligned16(d_out_planes) && rp::aligned16(d_qkv), RP_EALIGN,
               "rp_attention_bwd_tc: 16-byte alignment");
    RP_GUARD(device);
    CUtensorMap tmQKV128, tmQKV64, tmDO128, tmDO64;
    int rc = make_tok_tmap(&tmQKV128, qkv_planes, 3 * EMB, n_img, BM);
    if (rc) return rc;
    rc = make_tok_tmap(&tmQKV64, qkv_planes, 3 * EMB, n_img, BN);
    if (rc) return rc;
    rc = make_tok_tmap(&tmDO128, d_out_planes, EMB, n_img, BM);
    if (rc) return rc;
    rc = make_tok_tmap(&tmDO64, d_out_planes, EMB, n_img, BN);
    if (rc) return rc;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) {
            rp::set_error("rp_attention_bwd_tc: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    const int nitems = 2 * n_img * HEADS * TILES;
    const int slots = 2 * rp::num_sms(device);
    const int grid = nitems < slots ? nitems : slots;
    const float scale = 0.125f;                                   // head_dim^-0.5
    rp::launch(attention_bwd_tc_kernel, dim3(grid), dim3(THREADS), (size_t)SMEM, (cudaStream_t)stream, tmQKV128, tmQKV64,
               tmDO128, tmDO64, lse, delta, d_qkv, n_img, scale * 1.4426950408889634f, scale);
    return rp::finish_launch("rp_attention_bwd_tc");
}
