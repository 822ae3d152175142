// 3x3 / stride 1 / pad 1 convolution with 64 input channels on tcgen05 -- the ResNet layer1 convolutions
// (src/model.py:127-131: four 64 -> 64 convolutions on 56 x 56 maps), "halo" variant of the implicit GEMM in
// gemm_tc.cu.
//
// gemm_tc.cu fetches, for every filter tap, its own [R rows x Wo pixels x 64 channels] activation window: nine
// overlapping TMA boxes per 112-pixel tile = 258 KB of L2 -> SM traffic per tile (bf16x3) next to 147 KB of filter
// taps, and ncu shows these layers pinned at the L2 -> SM bandwidth of the SM (39 of ~42 B/clk).  Here ONE box per
// tile brings the halo [(R+2) rows x (Wo+2) pixels x 64 channels] (59 KB, zero padding = TMA out-of-bounds fill),
// and every tap is the SAME shared-memory buffer read through a descriptor whose start address is shifted by
// (ky (Wo+2) + kx) pixels = rows of 128 bytes.  That works because the 128-byte swizzle is a function of the absolute
// shared-memory address: a K-major SWIZZLE_128B operand may start at any 128-byte-aligned row of a swizzled buffer
// with the descriptor's base-offset field left at 0 (tools/probes/umma_shift_probe.cu, checked on B200 for shifts
// 0..116 rows).  The M = 128 accumulator rows then run over R x (Wo+2) pixel SLOTS: the two slots per row that
// belong to the halo columns produce junk rows that the epilogue skips (each output row depends only on its own
// operand row, so junk never mixes into valid rows).
//
// Same operand planes (P = 1 bf16, P = 2 bf16x3), same epilogue contract as rp_conv2d_tc (folded BatchNorm
// scale/shift, residual before the activation, ReLU, float32 and/or bf16-plane outputs).
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, CIN = 64, TAPS = 9;
constexpr int EPI_WARPS = 16;
constexpr int NTHREADS = 32 * (2 + EPI_WARPS);
constexpr int TMEM_COLS = 512, ACC_STRIDE = 256;
constexpr int STG_LD = 16;
constexpr int STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;

struct HaloGeom {
    int n_img, H, W;       // stride 1, pad 1: Ho = H, Wo = W
    int R;                 // output rows per tile: R * (W + 2) <= 128
    int tiles_per_img;
    int pitch;             // W + 2 pixel slots per tile row
    int halo_bytes;        // one plane of one halo buffer, rounded up to 1 KiB
    int O;
};

struct HaloEpi {
    const float* scale;
    const float* shift;
    const float* res_pre;        // [M, O] or null
    float* out_f32;              // [M, O] or null
    __nv_bfloat16* out_planes;   // [p_out][M][O] or null
    int p_out;
    int act;
    long long M;                 // n_img * H * W
};

template <int P, int BN>
struct HCfg {
    static constexpr int B_TILE = BN * 64 * 2;            // one filter tap, one plane
    static constexpr int STAGE_BYTES = P * B_TILE;
    // filter-tap ring depth.  A tap (16 KiB in bf16x3) is consumed in ~580 cycles of tensor work but takes ~2900 cycles
    // to arrive (148 SMs fetch the same L2 lines): with 3 stages in flight the kernel ran at one tap per ~980 cycles
    // whatever the issuer or the activation traffic did.  4 stages is what fits next to two halo buffers (bf16x3).
    static constexpr int NSTAGE = (P == 1) ? 8 : 4;
};

template <int P, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_halo_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, HaloEpi ep, HaloGeom g) {
    using C = HCfg<P, BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a SHARED pointer (LDS / STS)
    // layout: halo[2][P][halo_bytes] | filter ring [NSTAGE][P][B_TILE] | staging | barriers.  Operand reads that run
    // past a halo buffer (junk accumulator rows) land in the next halo buffer / the ring: valid shared memory.
    const int halo_buf = P * g.halo_bytes;
    uint8_t* ring = smem + 2 * halo_buf;
    float* staging = reinterpret_cast<float*>(ring + C::NSTAGE * C::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(staging) + STG_BYTES);
    uint64_t* full = bars;                          // [NSTAGE]
    uint64_t* empty = bars + C::NSTAGE;             // [NSTAGE]
    uint64_t* hfull = bars + 2 * C::NSTAGE;         // [2]
    uint64_t* hempty = hfull + 2;                   // [2]
    uint64_t* tfull = hfull + 4;                    // [2]
    uint64_t* tempty = hfull + 6;                   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hfull + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = g.n_img * g.tiles_per_img;
    const uint32_t halo_tx = (uint32_t)(P * (g.R + 2) * g.pitch * 128);

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmB);
        for (int i = 0; i < C::NSTAGE; ++i) {
            tc::mbar_init(&full[i], 1);
            tc::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&hfull[i], 1);
            tc::mbar_init(&hempty[i], 1);
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

    auto halo = [&](int hb, int p) { return smem + hb * halo_buf + p * g.halo_bytes; };
    auto b_tile = [&](int stage, int p) { return ring + stage * C::STAGE_BYTES + p * C::B_TILE; };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (convergent warp)
        int stage = 0, phase = 0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int img = tile / g.tiles_per_img, oy0 = (tile % g.tiles_per_img) * g.R;
            const uint32_t hb = it & 1;
            tc::mbar_wait(&hempty[hb], ((it >> 1) & 1) ^ 1);
            if (tc::elect_one_sync()) {
                tc::mbar_expect_tx(&hfull[hb], halo_tx);
#pragma unroll
                for (int p = 0; p < P; ++p) tc::tma_load_5d(halo(hb, p), &tmA, &hfull[hb], 0, -1, oy0 - 1, img, p);
            }
            __syncwarp();
            for (int tap = 0; tap < TAPS; ++tap) {
                tc::mbar_wait(&empty[stage], phase ^ 1);
                if (tc::elect_one_sync()) {
                    tc::mbar_expect_tx(&full[stage], (uint32_t)C::STAGE_BYTES);
#pragma unroll
                    for (int p = 0; p < P; ++p) tc::tma_load_3d(b_tile(stage, p), &tmB, &full[stage], tap * CIN, 0, p);
                }
                __syncwarp();
                if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (convergent warp)
        // tap loop unrolled: halo shifts are constants, the filter descriptors are base + stage * stride
        constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
        int stage = 0;
        uint32_t phase = 0, it = 0;
        const uint64_t bd_base0 = tc::make_kmajor_sw128_desc(tc::smem_u32(b_tile(0, 0)));
        const uint64_t bd_base1 = tc::make_kmajor_sw128_desc(tc::smem_u32(b_tile(0, P - 1)));
        const uint32_t row_shift = (uint32_t)(g.pitch * 128) >> 4;     // one halo row, in descriptor units
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t hb = it & 1, acc = it & 1;
            tc::mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
            tc::mbar_wait(&hfull[hb], (it >> 1) & 1);
            tc::tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
            const uint64_t h0 = tc::make_kmajor_sw128_desc(tc::smem_u32(halo(hb, 0)));
            const uint64_t h1 = tc::make_kmajor_sw128_desc(tc::smem_u32(halo(hb, P - 1)));
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                tc::mbar_wait(&full[stage], phase);
                tc::tcgen05_fence_after();
                const int ky = tap / 3, kx = tap - 3 * ky;
                const uint32_t shift = ky * row_shift + (uint32_t)(kx * 128 >> 4);      // whole 128-byte rows
                const uint64_t a0 = h0 + shift, a1 = h1 + shift;
                const uint32_t soff = (uint32_t)stage * (uint32_t)(C::STAGE_BYTES >> 4);
                const uint64_t b0 = bd_base0 + soff, b1 = bd_base1 + soff;
                if (tc::elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t accum = (tap > 0 || k > 0) ? 1u : 0u;
                        if (P == 2) {            // smallest terms first (truncating fp32 accumulation)
                            tc::umma_bf16(d_tmem, a1 + 2 * k, b0 + 2 * k, idesc, accum);
                            tc::umma_bf16(d_tmem, a0 + 2 * k, b1 + 2 * k, idesc, 1u);
                            accum = 1u;
                        }
                        tc::umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, accum);
                    }
                    tc::umma_commit(&empty[stage]);
                    if (tap == TAPS - 1) {
                        tc::umma_commit(&tfull[acc]);
                        tc::umma_commit(&hempty[hb]);
                    }
                }
                __syncwarp();
                if (++stage == C::NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..17), as in gemm_tc.cu
        // except for the row map: accumulator row m = (tile row r, pixel slot x) with x < W valid.
        const int q = warp & 3, part = (warp - 2) >> 2;
        float* stg = staging + (warp - 2) * 32 * STG_LD;
        const int rr = lane >> 2, cq = lane & 3;
        constexpr int CH_PER_PART = (BN / 16) / 4;
        const int N = g.O;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1;
            const int img = tile / g.tiles_per_img, oy0 = (tile % g.tiles_per_img) * g.R;
            const int rows_valid = min(g.R, g.H - oy0);
            const long long row_base = ((long long)img * g.H + oy0) * g.W;
            // the four accumulator rows this lane handles on the coalesced side: slot -> output pixel
            long long grow[4];
            bool gok[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int m = q * 32 + t * 8 + rr;
                const int r = m / g.pitch, x = m - r * g.pitch;
                gok[t] = r < rows_valid && x < g.W;
                grow[t] = row_base + (long long)r * g.W + x;
            }
            const uint32_t t_row = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
            bool waited = false;
#pragma unroll 1
            for (int ci = 0; ci < CH_PER_PART; ++ci) {
                const int c0 = (part * CH_PER_PART + ci) * 16;
                const int col = c0 + cq * 4;
                // per-channel constants and the residual rows are requested BEFORE waiting for the accumulator: the
                // HBM round trip of the residual (~2 us) then hides behind this tile's MMAs
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ep.scale) sc = __ldg(reinterpret_cast<const float4*>(ep.scale + col));
                if (ep.shift) sh = __ldg(reinterpret_cast<const float4*>(ep.shift + col));
                float4 rpre[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    rpre[t] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (gok[t] && ep.res_pre) rpre[t] = __ldg(reinterpret_cast<const float4*>(ep.res_pre + grow[t] * N + col));
                }
                if (!waited) {
                    tc::mbar_wait(&tfull[acc], (it >> 1) & 1);
                    tc::tcgen05_fence_after();
                    waited = true;
                }
                uint32_t r[16];
                tc::tmem_ld_32x32b_x16(t_row + c0, r);
                tc::tmem_ld_wait();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<uint4*>(stg + lane * STG_LD + ((j ^ ((lane >> 1) & 3)) << 2)) =
                        make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (!gok[t]) continue;
                    const int rl = t * 8 + rr;
                    const size_t o = (size_t)grow[t] * N + col;
                    const float4 a = *reinterpret_cast<const float4*>(stg + rl * STG_LD + ((cq ^ ((rl >> 1) & 3)) << 2));
                    float v0 = fmaf(a.x, sc.x, sh.x) + rpre[t].x, v1 = fmaf(a.y, sc.y, sh.y) + rpre[t].y;
                    float v2 = fmaf(a.z, sc.z, sh.z) + rpre[t].z, v3 = fmaf(a.w, sc.w, sh.w) + rpre[t].w;
                    if (ep.act == RP_ACT_RELU) {
                        v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f);
                    }
                    if (ep.out_f32) *reinterpret_cast<float4*>(ep.out_f32 + o) = make_float4(v0, v1, v2, v3);
                    if (ep.out_planes) {
                        for (int p = 0; p < ep.p_out; ++p) {
                            __nv_bfloat162 h01 = __floats2bfloat162_rn(v0, v1), h23 = __floats2bfloat162_rn(v2, v3);
                            uint2 w;
                            w.x = *reinterpret_cast<uint32_t*>(&h01);
                            w.y = *reinterpret_cast<uint32_t*>(&h23);
                            *reinterpret_cast<uint2*>(ep.out_planes + (size_t)p * ep.M * N + o) = w;
                            v0 -= __uint_as_float(w.x << 16); v1 -= __uint_as_float(w.x & 0xffff0000u);
                            v2 -= __uint_as_float(w.y << 16); v3 -= __uint_as_float(w.y & 0xffff0000u);
                        }
                    }
                }
            }
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tempty[acc]);
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int P, int BN>
int launch_halo(const CUtensorMap& tmA, const CUtensorMap& tmB, const HaloEpi& ep, const HaloGeom& g, int device, cudaStream_t st) {
    using C = HCfg<P, BN>;
    const int smem = 2 * P * g.halo_bytes + C::NSTAGE * C::STAGE_BYTES + STG_BYTES + 256 + 1024;
    RP_REQUIRE(smem <= 227 * 1024, RP_EINVAL, "rp_conv3x3_halo_tc: %d bytes of shared memory needed", smem);
    static int attr_set[64] = {0};
    if (device >= 0 && device < 64 && attr_set[device] < smem) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_tc_kernel<P, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            rp::set_error("rp_conv3x3_halo_tc: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = smem;
    }
    const int ntiles = g.n_img * g.tiles_per_img;
    const int grid = ntiles < rp::num_sms(device) ? ntiles : rp::num_sms(device);
    rp::launch(conv3x3_halo_tc_kernel<P, BN>, dim3(grid), dim3(NTHREADS), (size_t)(smem), st, tmA, tmB, ep, g);
    return rp::finish_launch("rp_conv3x3_halo_tc");
}

}  // namespace

extern "C" int rp_conv3x3_halo_supported(int H, int W, int C, int O, int KH, int KW, int stride, int pad) {
    if (!(KH == 3 && KW == 3 && stride == 1 && pad == 1 && C == CIN && (O == 64 || O == 128) && W + 2 <= 128 && H > 0 && W >= 8))
        return 0;
    int R = 128 / (W + 2);
    if (R > H) R = H;
    // two double-plane halo buffers + filter ring + staging must fit the 227 KiB of shared memory
    const int halo_bytes = (((R + 2) * (W + 2) * 128) + 1023) & ~1023;
    return 4 * halo_bytes + 4 * 2 * O * 128 + STG_BYTES + 2048 <= 227 * 1024;
}

extern "C" int rp_conv3x3_halo_tc(const void* x_planes, const void* w_planes, const float* scale, const float* shift,
                                  const float* res_pre, float* out_f32, void* out_planes, int n_img, int H, int W, int C,
                                  int O, int P, int P_out, int act, int device, void* stream) {
    RP_REQUIRE(x_planes && w_planes && (out_f32 || out_planes), RP_EINVAL, "rp_conv3x3_halo_tc: null pointer");
    RP_REQUIRE(rp_conv3x3_halo_supported(H, W, C, O, 3, 3, 1, 1) && n_img > 0, RP_EINVAL,
               "rp_conv3x3_halo_tc: unsupported shape n=%d H=%d W=%d C=%d O=%d", n_img, H, W, C, O);
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_conv3x3_halo_tc: P must be 1 or 2");
    RP_REQUIRE(!out_planes || (P_out >= 1 && P_out <= 2), RP_EINVAL, "rp_conv3x3_halo_tc: bad P_out");
    RP_REQUIRE(act == RP_ACT_NONE || act == RP_ACT_RELU, RP_EINVAL, "rp_conv3x3_halo_tc: act must be none or relu");
    RP_REQUIRE(rp::aligned16(x_planes) && rp::aligned16(w_planes) && rp::aligned16(out_f32) && rp::aligned16(out_planes) &&
                   rp::aligned16(res_pre) && rp::aligned16(scale) && rp::aligned16(shift),
               RP_EALIGN, "rp_conv3x3_halo_tc: 16-byte alignment");
    HaloGeom g{};
    g.n_img = n_img; g.H = H; g.W = W; g.O = O;
    g.pitch = W + 2;
    g.R = 128 / g.pitch;
    if (g.R > H) g.R = H;
    g.tiles_per_img = (H + g.R - 1) / g.R;
    g.halo_bytes = (((g.R + 2) * g.pitch * 128) + 1023) & ~1023;
    RP_GUARD(device);
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    RP_REQUIRE(fn != nullptr, RP_EINVAL, "rp_conv3x3_halo_tc: cuTensorMapEncodeTiled entry point unavailable");
    CUtensorMap tmA, tmB;
    {
        // activations [plane][image][H][W][C]; box = the tile's halo: [1][1][R+2][W+2][64], out of bounds -> 0
        cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img, (cuuint64_t)P};
        cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                              (cuuint64_t)n_img * H * W * C * 2};
        cuuint32_t box[5] = {64, (cuuint32_t)g.pitch, (cuuint32_t)(g.R + 2), 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = fn(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x_planes), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RP_REQUIRE(r == CUDA_SUCCESS, RP_EINVAL, "rp_conv3x3_halo_tc: activation tensor map failed (CUresult %d)", (int)r);
    }
    int rc = tc::make_planes_tmap(&tmB, w_planes, P, O, TAPS * C, O);        // [P][O][9*64], box O rows x 64 K
    if (rc) return rc;
    HaloEpi ep{scale, shift, res_pre, out_f32, static_cast<__nv_bfloat16*>(out_planes), P_out, act, (long long)n_img * H * W};
    cudaStream_t st = (cudaStream_t)stream;
    if (P == 1) return O == 64 ? launch_halo<1, 64>(tmA, tmB, ep, g, device, st) : launch_halo<1, 128>(tmA, tmB, ep, g, device, st);
    return O == 64 ? launch_halo<2, 64>(tmA, tmB, ep, g, device, st) : launch_halo<2, 128>(tmA, tmB, ep, g, device, st);
}
