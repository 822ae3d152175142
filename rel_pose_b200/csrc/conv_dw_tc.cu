// Weight gradient of a convolution (stride 1 or 2) (A11: autograd of nn.Conv2d in torchvision BasicBlock / extractor.py:9-13)
// as an IMPLICIT GEMM on tcgen05 -- no im2col matrix, no transposed copies:
//
//   dW[o][kh][kw][c] = sum over (img, oy, ox) of  dy[img][oy][ox][o] * x[img][oy + kh - pad][ox + kw - pad][c]
//
// The contraction runs over PIXELS, and both operands are stored with the pixel index outermost (NHWC planes), i.e.
// both are "MN-major" UMMA operands exactly as TMA delivers them:
//   B = dy tile   [64 pixels (8 x 8 outputs)][O]        O / 64 spans of 64 channels, one TMA box each
//   A = x  tile   [64 pixels shifted by the tap][64 c]  M = 128 = TWO taps side by side (two 64-wide spans; the leading
//                                                       byte offset of the descriptor is the distance between them)
//   D[(tap, c)][o] (+)= A^T B    M = 128, N = O, K = 16 pixels per instruction, split-bf16 (a1 b0 + a0 b1 + a0 b0)
// HALO mode: ONE box per pixel tile brings the (8 + KH - 1) x (8 + KW - 1) halo of the 64 input channels; every tap
// is a shifted view of it (start address + ((kh) * halo_width + kw) rows of 128 bytes, stride byte offset = one halo
// row): the input tile is read once, not KH * KW times.  Stride-2 convolutions (and RELPOSE_DW_HALO=0, for A/B runs)
// load one 8 x 8 box per tap instead (element stride 2 in the tensor map).
//
// Work item (one CTA) = (64-channel chunk of C, group of G taps, slice of the pixel tiles): accumulators stay in
// tensor memory for the whole slice (ceil(G / 2) x O <= 512 columns), then go to a per-slice partial buffer; a
// fixed-order reduction over the slices writes dW as [O][KH][KW][C] (deterministic, no atomics).
// CTA = 256 threads: warp 0 TMA producer (ring of stages), warp 1 MMA issuer + tensor-memory owner, warps 4-7 epilogue.
#include "tc_common.cuh"

namespace {

constexpr int P = 2;                       // split-bf16 planes
constexpr int TP = 8;                      // pixel tile = TP x TP outputs = 64 = the K extent of one stage
constexpr int TILE_B = 64 * 128;           // bytes of one [64 pixels][64 channels] bf16 tile
constexpr int THREADS = 256;
constexpr int MAX_STAGES = 4;

struct DwGeom {
    int n_img, OH, OW, C, O, KH, KW, pad, stride_x, stride_y;
    int Osub, nog;                         // output channels per work item (<= 192) and number of such groups
    int tap_c;                             // 0: a tap shifts the pixel window (convolution); 64: a tap is the next 64-column block (nn.Linear)
    int HW, HH, halo_bytes;                // halo box (pixels) and its 1024-rounded size in shared memory
    int tiles_x, tiles_y, ntiles;
    int ncc, ntg, G, nps;                  // 64-channel chunks of C, tap groups, taps per group, pixel slices
    int noc;                               // O / 64
    int stages, stage_bytes, dy_bytes;
    int tmem_cols;
};

__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

template <bool HALO>
__global__ void __launch_bounds__(THREADS, 1)
conv_dw_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  float* __restrict__ partial, const DwGeom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + g.stages * g.stage_bytes);
    uint64_t* full = bars;                       // [stages]
    uint64_t* empty = bars + MAX_STAGES;         // [stages]
    uint64_t* acc_full = bars + 2 * MAX_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = g.KH * g.KW;
    // item -> (pixel slice, tap group, channel chunk)
    const int ps = blockIdx.x % g.nps;
    const int tg = (blockIdx.x / g.nps) % g.ntg;
    const int cc = (blockIdx.x / (g.nps * g.ntg)) % g.ncc;
    const int og = blockIdx.x / (g.nps * g.ntg * g.ncc);
    const int tap0 = tg * g.G;
    const int gcur = min(g.G, T - tap0);                     // taps of this group
    const int npairs = (gcur + 1) >> 1;
    const int t_begin = (int)((long long)ps * g.ntiles / g.nps), t_end = (int)((long long)(ps + 1) * g.ntiles / g.nps);

    if (threadIdx.x == 0) {
        tc::prefetch_tmap(&tmX);
        tc::prefetch_tmap(&tmDY);
        for (int i = 0; i < g.stages; ++i) {
            tc::mbar_init(&full[i], 1);
            tc::mbar_init(&empty[i], 1);
        }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    rp::pdl_launch_dependents();
    if (warp == 1) tc::tmem_alloc(tmem_slot, (uint32_t)g.tmem_cols);
    rp::pdl_wait();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------------------- TMA producer (convergent warp)
        int stage = 0, phase = 0;
        const uint32_t tx_bytes = (uint32_t)(P * g.noc * TILE_B + (HALO ? P * g.HH * g.HW * 128 : P * gcur * TILE_B));
        for (int tile = t_begin; tile < t_end; ++tile) {
            const int tx = tile % g.tiles_x, ty = (tile / g.tiles_x) % g.tiles_y, img = tile / (g.tiles_x * g.tiles_y);
            const int ox0 = tx * TP, oy0 = ty * TP;
            tc::mbar_wait(&empty[stage], phase ^ 1);
            if (tc::elect_one_sync()) {
                uint8_t* sb = smem + stage * g.stage_bytes;
                tc::mbar_expect_tx(&full[stage], tx_bytes);
                for (int p = 0; p < P; ++p)
                    for (int oc = 0; oc < g.noc; ++oc)
                        tc::tma_load_5d(sb + (p * g.noc + oc) * TILE_B, &tmDY, &full[stage], og * g.Osub + oc * 64, ox0, oy0, img, p);
                if (HALO) {
                    for (int p = 0; p < P; ++p)
                        tc::tma_load_5d(sb + g.dy_bytes + p * g.halo_bytes, &tmX, &full[stage], cc * 64, ox0 - g.pad, oy0 - g.pad, img, p);
                } else {
                    for (int t = 0; t < gcur; ++t) {
                        const int kh = (tap0 + t) / g.KW, kw = (tap0 + t) % g.KW;
                        for (int p = 0; p < P; ++p)
                            tc::tma_load_5d(sb + g.dy_bytes + (t * P + p) * TILE_B, &tmX, &full[stage], cc * 64 + kw * g.tap_c,
                                            ox0 * g.stride_x + (g.tap_c ? 0 : kw) - g.pad, oy0 * g.stride_y + kh - g.pad, img, p);
                    }
                }
            }
            __syncwarp();
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------------------- MMA issuer (convergent warp)
        const uint32_t idesc = tc::make_idesc_bf16(128, g.Osub) | tc::IDESC_A_MN | tc::IDESC_B_MN;
        int stage = 0, phase = 0;
        for (int tile = t_begin; tile < t_end; ++tile) {
            tc::mbar_wait(&full[stage], phase);
            tc::tcgen05_fence_after();
            if (tc::elect_one_sync()) {
                const uint32_t sb = tc::smem_u32(smem + stage * g.stage_bytes);
                const uint32_t first = tile == t_begin ? 0u : 1u;
                for (int q = 0; q < npairs; ++q) {
                    const int t0 = 2 * q, t1 = min(2 * q + 1, gcur - 1);
                    uint32_t a_off0, lbo, sbo;                          // bytes: first tap's view, span distance, 8-row group distance
                    if (HALO) {
                        const int kh0 = (tap0 + t0) / g.KW, kw0 = (tap0 + t0) % g.KW;
                        const int kh1 = (tap0 + t1) / g.KW, kw1 = (tap0 + t1) % g.KW;
                        a_off0 = (uint32_t)(kh0 * g.HW + kw0) * 128u;
                        lbo = (uint32_t)((kh1 - kh0) * g.HW + (kw1 - kw0)) * 128u;
                        sbo = (uint32_t)g.HW * 128u;
                    } else {
                        a_off0 = (uint32_t)(t0 * P) * TILE_B;
                        lbo = (uint32_t)((t1 - t0) * P) * TILE_B;
                        sbo = 1024u;
                    }
                    const uint32_t a_plane = HALO ? (uint32_t)g.halo_bytes : (uint32_t)TILE_B;
                    const uint32_t a_kstep = HALO ? 2u * sbo : 2048u;    // 16 pixels = two rows of the tile
                    const uint32_t a0 = sb + g.dy_bytes + a_off0, a1 = a0 + a_plane;
                    const uint32_t b0 = sb, b1 = sb + g.noc * TILE_B;
                    const uint32_t d = tmem_base + (uint32_t)(q * g.Osub);
                    uint32_t accum = first;
                    // correction terms of all four K steps first, main terms last (truncating fp32 accumulation)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        tc::umma_bf16_desc64(d, mn_desc(a1 + ks * a_kstep, lbo, sbo), mn_desc(b0 + ks * 2048, TILE_B, 1024), idesc, accum);
                        tc::umma_bf16_desc64(d, mn_desc(a0 + ks * a_kstep, lbo, sbo), mn_desc(b1 + ks * 2048, TILE_B, 1024), idesc, 1u);
                        accum = 1u;
                    }
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        tc::umma_bf16_desc64(d, mn_desc(a0 + ks * a_kstep, lbo, sbo), mn_desc(b0 + ks * 2048, TILE_B, 1024), idesc, 1u);
                }
                tc::umma_commit(&empty[stage]);
                if (tile + 1 == t_end) tc::umma_commit(acc_full);
            }
            __syncwarp();
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------------------- epilogue: accumulators -> partial[ps]
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                 // accumulator row = (span, channel)
        const int span = r >> 6, c = r & 63;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const size_t KC = (size_t)T * g.C;
        float* dst0 = partial + ((size_t)ps * g.O + (size_t)og * g.Osub) * KC + (size_t)cc * 64 + c;
        tc::mbar_wait(acc_full, 0);
        tc::tcgen05_fence_after();
        for (int q = 0; q < npairs; ++q) {
            const int t = 2 * q + span;
            const bool valid = t < gcur;                   // the odd group's last pair carries a duplicate second span
            float* dst = dst0 + (size_t)(tap0 + (valid ? t : 0)) * g.C;
            for (int o0 = 0; o0 < g.Osub; o0 += 32) {
                uint32_t v[32];
                tc::tmem_ld_32x32b_x32(t_lane + (uint32_t)(q * g.Osub + o0), v);
                tc::tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) dst[(size_t)(o0 + i) * KC] = __uint_as_float(v[i]);
                }
            }
        }
    }

    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tcgen05_fence_after();
        tc::tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
    }
}

// dw[i] = sum over the pixel slices of partial[s][i], fixed order
__global__ void __launch_bounds__(256)
conv_dw_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, long long n4, int nps) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n4) return;
    float4 acc = reinterpret_cast<const float4*>(partial)[i];
    for (int s = 1; s < nps; ++s) {
        const float4 v = reinterpret_cast<const float4*>(partial)[(long long)s * n4 + i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(dw)[i] = acc;
}

bool dw_halo_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RELPOSE_DW_HALO");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

int make_geom(DwGeom& g, int n_img, int H, int W, int C, int O, int KH, int KW, int pad, int stride_x, int stride_y, int device,
              bool halo, bool linear = false) {
    g.n_img = n_img; g.C = C; g.O = O; g.KH = KH; g.KW = KW; g.pad = pad; g.stride_x = stride_x; g.stride_y = stride_y;
    g.tap_c = linear ? 64 : 0;
    g.OH = (H + 2 * pad - KH) / stride_y + 1;
    g.OW = linear ? W : (W + 2 * pad - KW) / stride_x + 1;
    g.HW = TP + KW - 1; g.HH = TP + KH - 1;
    g.halo_bytes = (g.HH * g.HW * 128 + 1023) / 1024 * 1024;
    g.tiles_x = (g.OW + TP - 1) / TP; g.tiles_y = (g.OH + TP - 1) / TP;
    g.ntiles = n_img * g.tiles_x * g.tiles_y;
    g.Osub = O <= 192 ? O : (O % 192 == 0 ? 192 : 128);
    g.nog = O / g.Osub;
    g.ncc = C / 64; g.noc = g.Osub / 64;
    const int T = KH * KW;
    const int maxpairs = 512 / g.Osub;
    int G = T < 2 * maxpairs ? T : 2 * maxpairs;
    if (linear)                                                 // two stages in flight: the taps are not views of one halo tile
        while (G > 2 && 2 * (P * g.noc * TILE_B + P * G * TILE_B) > 200 * 1024) G -= 2;
    g.ntg = (T + G - 1) / G;
    g.G = (T + g.ntg - 1) / g.ntg;
    const int groups = g.ncc * g.ntg * g.nog;
    const int sms = rp::num_sms(device);
    int nps = (sms + groups / 2) / groups;
    if (nps > (g.ntiles + 3) / 4) nps = (g.ntiles + 3) / 4;      // at least ~4 pixel tiles per slice: the partial buffers are the traffic
    if (nps < 1) nps = 1;
    g.nps = nps;
    g.dy_bytes = P * g.noc * TILE_B;
    g.stage_bytes = g.dy_bytes + (halo ? P * g.halo_bytes : P * g.G * TILE_B);
    int stages = (200 * 1024) / g.stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    g.stages = stages;
    const int cols = ((g.G + 1) / 2) * g.Osub;
    g.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    return stages >= 1 ? RP_OK : RP_EINVAL;
}

// shared by the convolution and the linear entry points: tensor maps, launch, slice reduction
int launch_dw(const void* x_planes, const void* dy_planes, float* dw, const DwGeom& g, int H, int W, bool halo, void* workspace,
              size_t workspace_bytes, int device, cudaStream_t st, const char* what) {
    const int T = g.KH * g.KW;
    const size_t need = (size_t)g.nps * g.O * T * g.C * sizeof(float);
    RP_REQUIRE(workspace_bytes >= need, RP_EWORKSPACE, "%s: workspace too small (%zu < %zu)", what, workspace_bytes, need);
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    RP_REQUIRE(fn != nullptr, RP_EINVAL, "%s: cuTensorMapEncodeTiled entry point unavailable", what);
    CUtensorMap tmX, tmDY;
    {
        const cuuint64_t Cm = g.tap_c ? (cuuint64_t)64 * g.KW : (cuuint64_t)g.C;     // nn.Linear: a "pixel" is a whole row of X
        cuuint64_t gdim[5] = {Cm, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)g.n_img, (cuuint64_t)P};
        cuuint64_t gstr[4] = {Cm * 2, (cuuint64_t)W * Cm * 2, (cuuint64_t)H * W * Cm * 2, (cuuint64_t)g.n_img * H * W * Cm * 2};
        cuuint32_t box[5] = {64, (cuuint32_t)(halo ? g.HW : TP * g.stride_x), (cuuint32_t)(halo ? g.HH : TP * g.stride_y), 1, 1};
        cuuint32_t estr[5] = {1, (cuuint32_t)g.stride_x, (cuuint32_t)g.stride_y, 1, 1};
        CUresult r = fn(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x_planes), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RP_REQUIRE(r == CUDA_SUCCESS, RP_EINVAL, "%s: input tensor map failed (CUresult %d)", what, (int)r);
    }
    {
        cuuint64_t gdim[5] = {(cuuint64_t)g.O, (cuuint64_t)g.OW, (cuuint64_t)g.OH, (cuuint64_t)g.n_img, (cuuint64_t)P};
        cuuint64_t gstr[4] = {(cuuint64_t)g.O * 2, (cuuint64_t)g.OW * g.O * 2, (cuuint64_t)g.OH * g.OW * g.O * 2,
                              (cuuint64_t)g.n_img * g.OH * g.OW * g.O * 2};
        cuuint32_t box[5] = {64, TP, TP, 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = fn(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy_planes), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RP_REQUIRE(r == CUDA_SUCCESS, RP_EINVAL, "%s: gradient tensor map failed (CUresult %d)", what, (int)r);
    }
    const size_t smem = (size_t)g.stages * g.stage_bytes + 256 + 1024;
    auto kern = halo ? conv_dw_tc_kernel<true> : conv_dw_tc_kernel<false>;
    static bool attr_set[2][64] = {{false}};
    if (device >= 0 && device < 64 && !attr_set[halo][device]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) {
            rp::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[halo][device] = true;
    }
    const int grid = g.ncc * g.ntg * g.nog * g.nps;
    rp::launch(kern, dim3(grid), dim3(THREADS), smem, st, tmX, tmDY, static_cast<float*>(workspace), g);
    int rc = rp::finish_launch(what);
    if (rc) return rc;
    const long long n4 = (long long)g.O * T * g.C / 4;
    conv_dw_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(static_cast<const float*>(workspace), dw, n4, g.nps);
    return rp::finish_launch(what);
}

}  // namespace

extern "C" int rp_conv_dw_tc_supported(int C, int O, int KH, int KW, int stride) {
    return (stride == 1 || stride == 2) && C % 64 == 0 && C >= 64 && (O == 64 || O == 128 || O == 192) && KH >= 1 && KW >= 1 && KH <= 7 && KW <= 7;
}

extern "C" size_t rp_conv_dw_tc_workspace_bytes(int n_img, int H, int W, int C, int O, int KH, int KW, int pad, int stride,
                                                int device) {
    DwGeom g;
    if (!rp_conv_dw_tc_supported(C, O, KH, KW, stride) ||
        make_geom(g, n_img, H, W, C, O, KH, KW, pad, stride, stride, device, stride == 1 && dw_halo_enabled()))
        return 0;
    return (size_t)g.nps * O * KH * KW * C * sizeof(float);
}

extern "C" int rp_conv_dw_tc(const void* x_planes, const void* dy_planes, float* dw, int n_img, int H, int W, int C, int O,
                             int KH, int KW, int pad, int stride, void* workspace, size_t workspace_bytes, int device,
                             void* stream) {
    RP_REQUIRE(x_planes && dy_planes && dw && workspace && n_img > 0, RP_EINVAL, "rp_conv_dw_tc: bad argument");
    RP_REQUIRE(rp_conv_dw_tc_supported(C, O, KH, KW, stride), RP_EINVAL, "rp_conv_dw_tc: unsupported shape C=%d O=%d %dx%d/%d", C, O, KH, KW, stride);
    RP_REQUIRE(rp::aligned16(x_planes) && rp::aligned16(dy_planes) && rp::aligned16(dw) && rp::aligned16(workspace), RP_EALIGN,
               "rp_conv_dw_tc: 16-byte alignment");
    RP_GUARD(device);
    const bool halo = stride == 1 && dw_halo_enabled();      // a strided tap is not a shifted view of a dense halo
    DwGeom g;
    RP_REQUIRE(make_geom(g, n_img, H, W, C, O, KH, KW, pad, stride, stride, device, halo) == RP_OK && g.OH > 0 && g.OW > 0, RP_EINVAL,
               "rp_conv_dw_tc: geometry does not fit");
    return launch_dw(x_planes, dy_planes, dw, g, H, W, halo, workspace, workspace_bytes, device, (cudaStream_t)stream, "rp_conv_dw_tc");
}

// Weight gradient of nn.Linear, dW [N][K] = dY^T X with X [M][K], dY [M][N] (bf16 planes, row-major): the same kernel.  The
// rows are the "pixels" (an image of M / 8 rows of 8 pixels, a pixel = one row of X), the K / 64 column blocks of X are
// the "taps": tap t reads channels [64 t, 64 t + 64) of the same pixel window (tap_c = 64) -- so X and dY are read in
// place as MN-major operands and neither is transposed (the split-K path needed rp_transpose_split_planes_bf16 of both).
static bool linear_dw_geom(DwGeom& g, int M, int N, int K, int device) {
    if (M < 64 || M % 8 != 0 || K % 64 != 0 || K / 64 > 64) return false;
    if (!(N == 64 || N == 128 || N == 192 || (N > 192 && (N % 192 == 0 || N % 128 == 0)))) return false;
    return make_geom(g, 1, M / 8, 8, 64, N, 1, K / 64, 0, 1, 1, device, false, true) == RP_OK;
}
extern "C" int rp_linear_dw_tc_supported(int M, int N, int K) {
    DwGeom g;
    return linear_dw_geom(g, M, N, K, 0) ? 1 : 0;
}
extern "C" size_t rp_linear_dw_tc_workspace_bytes(int M, int N, int K, int device) {
    DwGeom g;
    if (!linear_dw_geom(g, M, N, K, device)) return 0;
    return (size_t)g.nps * N * K * sizeof(float);
}
extern "C" int rp_linear_dw_tc(const void* x_planes, const void* dy_planes, float* dw, int M, int N, int K, void* workspace,
                               size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(x_planes && dy_planes && dw && workspace, RP_EINVAL, "rp_linear_dw_tc: bad argument");
    RP_REQUIRE(rp::aligned16(x_planes) && rp::aligned16(dy_planes) && rp::aligned16(dw) && rp::aligned16(workspace), RP_EALIGN,
               "rp_linear_dw_tc: 16-byte alignment");
    RP_GUARD(device);
    DwGeom g;
    RP_REQUIRE(linear_dw_geom(g, M, N, K, device), RP_EINVAL, "rp_linear_dw_tc: unsupported shape M=%d N=%d K=%d", M, N, K);
    return launch_dw(x_planes, dy_planes, dw, g, M / 8, 8, false, workspace, workspace_bytes, device, (cudaStream_t)stream,
                     "rp_linear_dw_tc");
}
