// ResNet stem in ONE launch (src/model.py:127-130: conv1 7x7/2 -> bn1 -> relu -> maxpool 3x3/2) on tcgen05 tensor cores.
//
// Before: rp_preprocess_stem_windows wrote 422 MB of space-to-depth windows for a 77 MB resized input, rp_conv2d_tc read
// them and wrote the 112 x 112 x 64 activation (411 MB float32), rp_maxpool3x3s2_planes read that again: ~2 GB of HBM
// traffic per 64 pairs for ~0.38 GB of algorithmic bytes.  Now:
//   * A1 writes the COMPACT space-to-depth image Zc[p][n][115][116][16] (108 MB): row yp = Y + 2, column xs = X + 2, 16
//     bf16 per 2x2 pixel block (12 used).  For a fixed tap row a, the operand row of output pixel (oy, ox) is the 64
//     consecutive bf16 starting at Zc[oy + a][ox]: rows of the A tile OVERLAP in memory.  The tensor map says exactly
//     that -- dimension 1 (ox) has a 32-byte stride under a 128-byte box row -- so one TMA box per (conv row, tap)
//     still lands a dense SWIZZLE_128B [112 x 64] K-major tile in shared memory; the window tensor is never written.
//     (rp_stem_pool_tc also accepts the old window layout: same kernel, ordinary strides.)
//   * the convolution's epilogue applies the folded BatchNorm + ReLU in the TMEM row layout (thread = output pixel),
//     keeps the VERTICAL 3-row maximum in registers across the consecutive conv rows a CTA walks, and once per pooled
//     row exchanges it through 28 KB of shared memory for the horizontal maximum: only the pooled 56 x 56 x 64 map
//     (float32 identity + bf16 planes) reaches HBM.
// Work unit = (image, band of 4 pooled rows) = conv rows 8b-1 .. 8b+7 (the first is recomputed: +12.5 % MMAs);
// persistent CTAs, 576 threads: warp 0 TMA producer (activation ring, weights once), warp 1 MMA issuer / TMEM owner,
// warps 2..17 epilogue (TMEM lane quarter x 16-channel quarter).  The 64 KB of filter planes stay resident in shared
// memory for the whole launch.  MMA order per conv row is the one rp_conv2d_tc uses (tap, k, [a1 b0, a0 b1, a0 b0]),
// so the result is bit-identical to the three-kernel sequence it replaces.
#include "tc_common.cuh"

namespace {

constexpr int HP = 115, WO = 112, HO = 112, PO = 56, CO = 64, TAPS = 4;
constexpr int WC = 116;                              // columns of the compact image (xs = 0..115; 115 is never read)
constexpr int A_ROWS_BYTES = WO * 128;               // 14 336: 112 rows of 128 bytes (14 swizzle atoms)
constexpr int W_TILE = CO * 128;                     // 8 KiB: [64 filters][64 K] of one tap
constexpr int EPI_WARPS = 16, NTHREADS = 32 * (2 + EPI_WARPS);
constexpr int STAGES = 4;
constexpr int BAND = 4;                              // pooled rows per work unit
constexpr int BANDS = PO / BAND;                     // 14
constexpr int VBUF_BYTES = WO * CO * 4;              // 28 672

template <int P>
struct SCfg {
    static constexpr int STAGE_BYTES = P * A_ROWS_BYTES;
    static constexpr int OFF_W = STAGES * STAGE_BYTES;                 // [TAPS][P][8 KiB], also the over-read pad of the last stage
    static constexpr int OFF_VBUF = OFF_W + TAPS * P * W_TILE;
    static constexpr int OFF_BAR = OFF_VBUF + VBUF_BYTES;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
};

template <int P>
__global__ void __launch_bounds__(NTHREADS, 1)
stem_pool_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out_f32,
                    __nv_bfloat16* __restrict__ out_planes, int p_out, int n_img) {
    using C = SCfg<P>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-aligned, still a SHARED pointer (LDS / STS)
    float* vbuf = reinterpret_cast<float*>(smem + C::OFF_VBUF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* full = bars;                      // [STAGES]
    uint64_t* empty = bars + STAGES;            // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;        // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;   // [2]
    uint64_t* wfull = bars + 2 * STAGES + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nunits = n_img * BANDS;

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmA);
        tc::prefetch_tmap(&tmW);
        for (int i = 0; i < STAGES; ++i) {
            tc::mbar_init(&full[i], 1);
            tc::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&tfull[i], 1);
            tc::mbar_init(&tempty[i], EPI_WARPS);
        }
        tc::mbar_init(wfull, 1);
        tc::fence_barrier_init();
    }
    rp::pdl_launch_dependents();                  // the next kernel may start its prologue (common.cuh)
    if (warp == 1) tc::tmem_alloc(tmem_slot, 128);
    rp::pdl_wait();                               // the previous kernel has completed: its outputs are visible
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto a_tile = [&](int stage, int p) { return smem + stage * C::STAGE_BYTES + p * A_ROWS_BYTES; };
    auto w_tile = [&](int tap, int p) { return smem + C::OFF_W + (tap * P + p) * W_TILE; };
    // unit -> (image, first conv row, number of conv rows)
    auto unit_rows = [&](int unit, int& img, int& band, int& y0, int& ny) {
        img = unit / BANDS;
        band = unit - img * BANDS;
        y0 = band == 0 ? 0 : 8 * band - 1;
        ny = band == 0 ? 8 : 9;
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (convergent warp)
        if (tc::elect_one_sync()) {
            tc::mbar_expect_tx(wfull, TAPS * P * W_TILE);
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap)
#pragma unroll
                for (int p = 0; p < P; ++p) tc::tma_load_3d(w_tile(tap, p), &tmW, wfull, tap * 64, 0, p);
        }
        __syncwarp();
        int stage = 0, phase = 0;
        for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
            int img, band, y0, ny;
            unit_rows(unit, img, band, y0, ny);
            for (int y = y0; y < y0 + ny; ++y) {
                for (int tap = 0; tap < TAPS; ++tap) {
                    tc::mbar_wait(&empty[stage], phase ^ 1);
                    if (tc::elect_one_sync()) {
                        tc::mbar_expect_tx(&full[stage], (uint32_t)C::STAGE_BYTES);
#pragma unroll
                        for (int p = 0; p < P; ++p) tc::tma_load_5d(a_tile(stage, p), &tmA, &full[stage], 0, 0, y + tap, img, p);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (convergent warp)
        // M = 128 instructions over a 112-row tile: rows 112..127 of the A operand are whatever follows in shared
        // memory (the next plane / stage / the filter planes: finite bf16); they only reach accumulator rows 112..127,
        // which no thread reads.
        constexpr uint32_t idesc = tc::make_idesc_bf16(128, CO);
        int stage = 0, phase = 0, acc = 0, acc_phase = 0;
        tc::mbar_wait(wfull, 0);
        for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
            int img, band, y0, ny;
            unit_rows(unit, img, band, y0, ny);
            for (int y = y0; y < y0 + ny; ++y) {
                tc::mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc::tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * CO;
                for (int tap = 0; tap < TAPS; ++tap) {
                    tc::mbar_wait(&full[stage], phase);
                    tc::tcgen05_fence_after();
                    const uint64_t a0 = tc::make_kmajor_sw128_desc(tc::smem_u32(a_tile(stage, 0)));
                    const uint64_t a1 = tc::make_kmajor_sw128_desc(tc::smem_u32(a_tile(stage, P - 1)));
                    const uint64_t b0 = tc::make_kmajor_sw128_desc(tc::smem_u32(w_tile(tap, 0)));
                    const uint64_t b1 = tc::make_kmajor_sw128_desc(tc::smem_u32(w_tile(tap, P - 1)));
                    if (tc::elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint32_t accum = (tap > 0 || k > 0) ? 1u : 0u;
                            if (P == 2) {
                                tc::umma_bf16(d_tmem, a1 + 2 * k, b0 + 2 * k, idesc, accum);
                                tc::umma_bf16(d_tmem, a0 + 2 * k, b1 + 2 * k, idesc, 1u);
                                accum = 1u;
                            }
                            tc::umma_bf16(d_tmem, a0 + 2 * k, b0 + 2 * k, idesc, accum);
                        }
                        tc::umma_commit(&empty[stage]);
                        if (tap + 1 == TAPS) tc::umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..17)
        const int q = warp & 3;                          // TMEM lane quarter
        const int part = (warp - 2) >> 2;                // 16-channel quarter
        const int x = q * 32 + lane;                     // output column of this thread's accumulator row
        const int et = threadIdx.x - 64;                 // 0..511 among the epilogue threads
        const int c0 = part * 16;
        float sc[16], sh[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            sc[i] = scale ? __ldg(scale + c0 + i) : 1.0f;
            sh[i] = shift ? __ldg(shift + c0 + i) : 0.0f;
        }
        int acc = 0, acc_phase = 0;
        const size_t plane = (size_t)n_img * PO * PO * CO;
        for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
            int img, band, y0, ny;
            unit_rows(unit, img, band, y0, ny);
            float vm[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) vm[i] = 0.0f;   // ReLU output is >= 0: 0 is neutral for the padded border
            for (int y = y0; y < y0 + ny; ++y) {
                tc::mbar_wait(&tfull[acc], acc_phase);
                tc::tcgen05_fence_after();
                uint32_t r[16];
                tc::tmem_ld_32x32b_x16(tmem_base + acc * CO + ((uint32_t)(q * 32) << 16) + c0, r);
                tc::tmem_ld_wait();
                tc::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&tempty[acc]);     // the accumulator is in registers
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(fmaf(__uint_as_float(r[i]), sc[i], sh[i]) + 0.0f, 0.0f);
                const bool odd = (y & 1) != 0;
                const bool first = (y == y0) && band > 0;          // row 2r-1 of the band's first pooled row
                if (first) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) vm[i] = v[i];
                    continue;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) vm[i] = fmaxf(vm[i], v[i]);
                if (!odd) continue;
                // conv row y = 2r + 1: pooled row r is complete in the vertical direction
                const int r_out = y >> 1;
                if (x < WO) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)        // 16-byte chunk (part*4 + j) of row x, XOR-swizzled by the row
                        *reinterpret_cast<float4*>(vbuf + x * CO + ((((part << 2) + j) ^ (x & 15)) << 2)) =
                            make_float4(vm[4 * j], vm[4 * j + 1], vm[4 * j + 2], vm[4 * j + 3]);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) vm[i] = v[i];         // row 2r+1 is also the top row of pooled row r+1
                asm volatile("bar.sync 1, 512;" ::: "memory");
                for (int idx = et; idx < PO * (CO / 4); idx += EPI_WARPS * 32) {
                    const int c4 = idx & 15, px = idx >> 4;
                    const int xa = 2 * px;
                    float4 m = *reinterpret_cast<const float4*>(vbuf + xa * CO + ((c4 ^ (xa & 15)) << 2));
                    {
                        const int xb = xa + 1;
                        const float4 t = *reinterpret_cast<const float4*>(vbuf + xb * CO + ((c4 ^ (xb & 15)) << 2));
                        m.x = fmaxf(m.x, t.x); m.y = fmaxf(m.y, t.y); m.z = fmaxf(m.z, t.z); m.w = fmaxf(m.w, t.w);
                    }
                    if (px > 0) {
                        const int xb = xa - 1;
                        const float4 t = *reinterpret_cast<const float4*>(vbuf + xb * CO + ((c4 ^ (xb & 15)) << 2));
                        m.x = fmaxf(m.x, t.x); m.y = fmaxf(m.y, t.y); m.z = fmaxf(m.z, t.z); m.w = fmaxf(m.w, t.w);
                    }
                    const size_t o = (((size_t)img * PO + r_out) * PO + px) * CO + c4 * 4;
                    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = m;
                    if (out_planes) {
                        float v0 = m.x, v1 = m.y, v2 = m.z, v3 = m.w;
                        for (int p = 0; p < p_out; ++p) {
                            __nv_bfloat162 h01 = __floats2bfloat162_rn(v0, v1), h23 = __floats2bfloat162_rn(v2, v3);
                            uint2 w;
                            w.x = *reinterpret_cast<uint32_t*>(&h01);
                            w.y = *reinterpret_cast<uint32_t*>(&h23);
                            *reinterpret_cast<uint2*>(out_planes + (size_t)p * plane + o) = w;
                            v0 -= __uint_as_float(w.x << 16); v1 -= __uint_as_float(w.x & 0xffff0000u);
                            v2 -= __uint_as_float(w.y << 16); v3 -= __uint_as_float(w.y & 0xffff0000u);
                        }
                    }
                }
                asm volatile("bar.sync 1, 512;" ::: "memory");     // vbuf may be rewritten
            }
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, 128);
    }
}

// ---- A1 into the compact space-to-depth layout ---------------------------------------------------------------------
// Zc[p][n][yp][xs][(dy*2+dx)*3 + c] = pixel (2(yp-2)+dy, 2(xs-2)+dx), channel c of the normalised 224 x 224 image
// (src/model.py:114-125: BGR -> RGB, /255, mean / std, legacy-nearest resize); 0 outside the image and in slots 12..15.
template <typename T>
__global__ void __launch_bounds__(256)
preprocess_stem_compact_kernel(const T* __restrict__ img, __nv_bfloat16* __restrict__ out, int n_img, int H, int W,
                               float scale_h, float scale_w, int P) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    const long long total = (long long)n_img * HP * WC;
    const long long plane = total * 16;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int xs = (int)(idx % WC);
        const int yp = (int)((idx / WC) % HP);
        const int n = (int)(idx / ((long long)WC * HP));
        const int Y = yp - 2, X = xs - 2;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.0f;
        if (Y >= 0 && Y < 112 && X >= 0 && X < 112) {
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int iy = min((int)floorf(__fmul_rn((float)(2 * Y + dy), scale_h)), H - 1);
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int ix = min((int)floorf(__fmul_rn((float)(2 * X + dx), scale_w)), W - 1);
                    const T* src = img + ((long long)n * 3) * H * W + (long long)iy * W + ix;
                    const float bb = (float)src[0], g = (float)src[(long long)H * W], r = (float)src[2ll * H * W];
                    float* q = v + (dy * 2 + dx) * 3;
                    q[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(r, 255.0f), 0.485f), 0.229f);
                    q[1] = __fdiv_rn(__fsub_rn(__fdiv_rn(g, 255.0f), 0.456f), 0.224f);
                    q[2] = __fdiv_rn(__fsub_rn(__fdiv_rn(bb, 255.0f), 0.406f), 0.225f);
                }
            }
        }
        for (int p = 0; p < P; ++p) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<uint32_t*>(&h);
                v[2 * i] -= __uint_as_float(w[i] << 16);
                v[2 * i + 1] -= __uint_as_float(w[i] & 0xffff0000u);
            }
            uint4* dst = reinterpret_cast<uint4*>(out + p * plane + idx * 16);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
}

// Activation map of the stem: [plane][image][115][112 windows][64].  compact: windows overlap (32-byte stride).
int make_stem_map(CUtensorMap* out, const void* base, int n_img, int P, bool compact, bool quiet) {
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    if (!fn) {
        if (!quiet) rp::set_error("cuTensorMapEncodeTiled entry point unavailable");
        return RP_EINVAL;
    }
    const cuuint64_t px = compact ? 32 : 128, row = compact ? (cuuint64_t)WC * 32 : (cuuint64_t)WO * 128;
    cuuint64_t gdim[5] = {64, (cuuint64_t)WO, (cuuint64_t)HP, (cuuint64_t)n_img, (cuuint64_t)P};
    cuuint64_t gstr[4] = {px, row, row * HP, row * HP * (cuuint64_t)n_img};
    cuuint32_t box[5] = {64, (cuuint32_t)WO, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (!quiet) rp::set_error("rp_stem_pool_tc: activation tensor map failed (CUresult %d, compact=%d)", (int)r, (int)compact);
        return RP_EINVAL;
    }
    return RP_OK;
}

template <int P>
int launch_stem_pool(const CUtensorMap& tmA, const CUtensorMap& tmW, const float* scale, const float* shift, float* out_f32,
                     void* out_planes, int p_out, int n_img, int device, cudaStream_t st) {
    using C = SCfg<P>;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(stem_pool_tc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            rp::set_error("rp_stem_pool_tc: cudaFuncSetAttribute(%d): %s", C::SMEM, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    const int nunits = n_img * BANDS;
    const int grid = nunits < rp::num_sms(device) ? nunits : rp::num_sms(device);
    rp::launch(stem_pool_tc_kernel<P>, dim3(grid), dim3(NTHREADS), (size_t)(C::SMEM), st, tmA, tmW, scale, shift, out_f32, static_cast<__nv_bfloat16*>(out_planes),
                                                           p_out, n_img);
    return rp::finish_launch("rp_stem_pool_tc");
}

template <typename T>
int stem_compact_launch(const T* images, void* planes, int n_img, int H, int W, int P, int device, void* stream) {
    RP_REQUIRE(images && planes && n_img > 0 && H > 0 && W > 0 && (P == 1 || P == 2), RP_EINVAL,
               "rp_preprocess_stem_compact: bad argument");
    RP_REQUIRE(rp::aligned16(planes), RP_EALIGN, "rp_preprocess_stem_compact: planes must be 16-byte aligned");
    RP_GUARD(device);
    const long long total = (long long)n_img * HP * WC;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)rp::num_sms(device) * 16;
    rp::launch(preprocess_stem_compact_kernel<T>, dim3((unsigned)(blocks < cap ? blocks : cap)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        images, static_cast<__nv_bfloat16*>(planes), n_img, H, W, (float)H / (float)224, (float)W / (float)224, P);
    return rp::finish_launch("rp_preprocess_stem_compact");
}

}  // namespace

extern "C" int rp_preprocess_stem_compact_f32(const float* images, void* planes, int n_img, int H, int W, int P, int device,
                                              void* stream) {
    return stem_compact_launch<float>(images, planes, n_img, H, W, P, device, stream);
}
extern "C" int rp_preprocess_stem_compact_u8(const uint8_t* images, void* planes, int n_img, int H, int W, int P, int device,
                                             void* stream) {
    return stem_compact_launch<uint8_t>(images, planes, n_img, H, W, P, device, stream);
}

// 1 when the driver accepts the overlapping-row tensor map of the compact layout (it encodes strides, it does not
// dereference the pointer), 0 otherwise: callers then keep the window layout.
extern "C" int rp_stem_compact_supported(int device) {
    (void)device;
    static int cached = -1;
    if (cached < 0) {
        CUtensorMap m;
        alignas(128) static char dummy[256];
        cached = make_stem_map(&m, dummy, 2, 1, true, true) == RP_OK ? 1 : 0;
    }
    return cached;
}

extern "C" int rp_stem_pool_tc(const void* z_planes, int compact, const void* w_planes, const float* scale, const float* shift,
                               float* out_f32, void* out_planes, int n_img, int P, int P_out, int device, void* stream) {
    RP_REQUIRE(z_planes && w_planes && (out_f32 || out_planes) && n_img > 0, RP_EINVAL, "rp_stem_pool_tc: bad argument");
    RP_REQUIRE(P == 1 || P == 2, RP_EINVAL, "rp_stem_pool_tc: P must be 1 (bf16) or 2 (bf16x3)");
    RP_REQUIRE(!out_planes || (P_out >= 1 && P_out <= 2), RP_EINVAL, "rp_stem_pool_tc: bad P_out");
    RP_REQUIRE(rp::aligned16(z_planes) && rp::aligned16(w_planes) && rp::aligned16(out_f32) && rp::aligned16(out_planes), RP_EALIGN,
               "rp_stem_pool_tc: 16-byte alignment");
    RP_GUARD(device);
    CUtensorMap tmA, tmW;
    int rc = make_stem_map(&tmA, z_planes, n_img, P, compact != 0, false);
    if (rc) return rc;
    rc = tc::make_planes_tmap(&tmW, w_planes, P, CO, TAPS * 64, CO);
    if (rc) return rc;
    if (P == 1) return launch_stem_pool<1>(tmA, tmW, scale, shift, out_f32, out_planes, P_out, n_img, device, (cudaStream_t)stream);
    return launch_stem_pool<2>(tmA, tmW, scale, shift, out_f32, out_planes, P_out, n_img, device, (cudaStream_t)stream);
}
