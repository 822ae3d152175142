// Helpers of the CNN front end (A2/A3, src/model.py:127-134, extractor.py:51-65) around the
// implicit-GEMM convolution in gemm_simt.cu: NHWC preprocessing, max-pool, weight re-layout and
// eval-mode BatchNorm folding.  All HBM-bound, channel-innermost (coalesced, 16-byte vectors).
#include "common.cuh"

#include <cuda_bf16.h>

namespace {

// ---- stem on tensor cores -------------------------------------------------------------------------
// resnet.conv1 (7x7, stride 2, pad 3, 3 -> 64; src/model.py:127) is re-expressed as a 4x4 stride-1
// convolution over the 2x2 space-to-depth image Z[Y][X][(dy,dx,c)] (112 x 112 x 12, 12 padded to 16):
//   out[oy][ox] = sum_{a,b in 0..3} sum_q  W2[a][b][q] * Z[oy + a - 2][ox + b - 2][q],
//   W2[a][b][(dy,dx,c)] = W[c][2a+dy-1][2b+dx-1]   (zero where the 7x7 index falls outside 0..6).
// For a fixed tap row a, the four b taps of one output pixel are 64 consecutive bf16 of Z: the A1
// preprocessing therefore writes, for every padded row Yp = Y + 2 in 0..114 and output column ox, that
// 64-element window -- a [n][115][112][64] tensor that rp_conv2d_tc consumes as a KH=4, KW=1, C=64,
// stride-1, pad-0 convolution (every K step one rectangular TMA box; no gather inside the GEMM).
constexpr int STEM_HP = 115, STEM_W = 112, STEM_K = 64;

template <typename T>
__global__ void __launch_bounds__(256)
preprocess_stem_windows_kernel(const T* __restrict__ img, __nv_bfloat16* __restrict__ out, int n_img, int H, int W,
                               float scale_h, float scale_w, int P) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    const long long total = (long long)n_img * STEM_HP * STEM_W * 4;       // one thread = one s2d pixel of one window
    const long long plane = (long long)n_img * STEM_HP * STEM_W * STEM_K;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(idx & 3);
        const long long t = idx >> 2;
        const int ox = (int)(t % STEM_W);
        const int yp = (int)((t / STEM_W) % STEM_HP);
        const int n = (int)(t / ((long long)STEM_W * STEM_HP));
        const int Y = yp - 2, X = ox - 2 + b;
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

// conv1.weight [64][3][7][7] -> W2 [64][4 (a)][64 (b*16 + (dy*2+dx)*3 + c)] float32
__global__ void stem_weight_windows_kernel(const float* __restrict__ w, float* __restrict__ out, int O) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    const int total = O * 4 * 64;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int k = idx & 63, a = (idx >> 6) & 3, o = idx >> 8;
        const int b = k >> 4, q = k & 15;
        float val = 0.0f;
        if (q < 12) {
            const int c = q % 3, dx = (q / 3) & 1, dy = q / 6;
            const int ky = 2 * a + dy - 1, kx = 2 * b + dx - 1;
            if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) val = w[((o * 3 + c) * 7 + ky) * 7 + kx];
        }
        out[idx] = val;
    }
}

// A1 fused with the layout change: BGR NCHW image -> normalised RGB NHWC with C padded 3 -> 4.
template <typename T>
__global__ void __launch_bounds__(256) preprocess_nhwc4_kernel(const T* __restrict__ img, float4* __restrict__ out,
                                                                int n_img, int H, int W, float scale_h, float scale_w) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    const int OUT = 224;
    long long total = (long long)n_img * OUT * OUT;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(idx % OUT);
        int oy = (int)((idx / OUT) % OUT);
        int n = (int)(idx / (OUT * OUT));
        int ix = min((int)floorf(__fmul_rn((float)ox, scale_w)), W - 1);
        int iy = min((int)floorf(__fmul_rn((float)oy, scale_h)), H - 1);
        const T* src = img + ((long long)n * 3) * H * W + (long long)iy * W + ix;
        float b = (float)src[0], g = (float)src[(long long)H * W], r = (float)src[2ll * H * W];
        float4 v;
        v.x = __fdiv_rn(__fsub_rn(__fdiv_rn(r, 255.0f), 0.485f), 0.229f);
        v.y = __fdiv_rn(__fsub_rn(__fdiv_rn(g, 255.0f), 0.456f), 0.224f);
        v.z = __fdiv_rn(__fsub_rn(__fdiv_rn(b, 255.0f), 0.406f), 0.225f);
        v.w = 0.0f;
        out[idx] = v;
    }
}

// nn.MaxPool2d(3, stride 2, padding 1) on NHWC, one thread per 4 channels of an output pixel
__global__ void __launch_bounds__(256) maxpool3x3s2_nhwc_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                                 int n_img, int H, int W, int C4, int Ho, int Wo) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    long long total = (long long)n_img * Ho * Wo * C4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C4);
        long long p = idx / C4;
        int ox = (int)(p % Wo), oy = (int)((p / Wo) % Ho), n = (int)(p / ((long long)Wo * Ho));
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
        y[idx] = m;
    }
}

// [O][C][KH][KW] -> [O][KH][KW][Cp]  (Cp >= C, zero padded)
__global__ void permute_conv_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int O, int C, int KH,
                                           int KW, int Cp) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    long long total = (long long)O * KH * KW * Cp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % Cp);
        long long t = idx / Cp;
        int kx = (int)(t % KW), ky = (int)((t / KW) % KH), o = (int)(t / ((long long)KW * KH));
        out[idx] = c < C ? w[(((long long)o * C + c) * KH + ky) * KW + kx] : 0.0f;
    }
}

// BN(eval)(conv + bias) = conv*scale + shift;  scale = gamma/sqrt(var+eps), shift = (bias-mean)*scale + beta
__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               const float* __restrict__ conv_bias, float eps, float* __restrict__ scale,
                               float* __restrict__ shift, int C) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = gamma[c] / sqrtf(var[c] + eps);
    float b = conv_bias ? conv_bias[c] : 0.0f;
    scale[c] = s;
    shift[c] = (b - mean[c]) * s + beta[c];
}

int grid_for(long long total, int device) {
    long long b = (total + 255) / 256;
    long long cap = (long long)rp::num_sms(device) * 16;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" int rp_preprocess_nhwc4_f32(const float* images, float* out, int n_img, int H, int W, int device, void* stream) {
    RP_REQUIRE(images && out && n_img > 0 && H > 0 && W > 0, RP_EINVAL, "rp_preprocess_nhwc4: bad argument");
    RP_REQUIRE(rp::aligned16(out), RP_EALIGN, "rp_preprocess_nhwc4: out must be 16-byte aligned");
    RP_GUARD(device);
    long long total = (long long)n_img * 224 * 224;
    rp::launch(preprocess_nhwc4_kernel<float>, dim3(grid_for(total, device)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        images, reinterpret_cast<float4*>(out), n_img, H, W, (float)H / (float)224, (float)W / (float)224);
    return rp::finish_launch("rp_preprocess_nhwc4");
}

extern "C" int rp_preprocess_nhwc4_u8(const uint8_t* images, float* out, int n_img, int H, int W, int device, void* stream) {
    RP_REQUIRE(images && out && n_img > 0 && H > 0 && W > 0, RP_EINVAL, "rp_preprocess_nhwc4: bad argument");
    RP_REQUIRE(rp::aligned16(out), RP_EALIGN, "rp_preprocess_nhwc4: out must be 16-byte aligned");
    RP_GUARD(device);
    long long total = (long long)n_img * 224 * 224;
    rp::launch(preprocess_nhwc4_kernel<uint8_t>, dim3(grid_for(total, device)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        images, reinterpret_cast<float4*>(out), n_img, H, W, (float)H / (float)224, (float)W / (float)224);
    return rp::finish_launch("rp_preprocess_nhwc4");
}

template <typename T>
static int stem_windows_launch(const T* images, void* planes, int n_img, int H, int W, int P, int device, void* stream) {
    RP_REQUIRE(images && planes && n_img > 0 && H > 0 && W > 0 && (P == 1 || P == 2), RP_EINVAL,
               "rp_preprocess_stem_windows: bad argument");
    RP_REQUIRE(rp::aligned16(planes), RP_EALIGN, "rp_preprocess_stem_windows: planes must be 16-byte aligned");
    RP_GUARD(device);
    long long total = (long long)n_img * STEM_HP * STEM_W * 4;
    rp::launch(preprocess_stem_windows_kernel<T>, dim3(grid_for(total, device)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        images, static_cast<__nv_bfloat16*>(planes), n_img, H, W, (float)H / (float)224, (float)W / (float)224, P);
    return rp::finish_launch("rp_preprocess_stem_windows");
}

extern "C" int rp_preprocess_stem_windows_f32(const float* images, void* planes, int n_img, int H, int W, int P,
                                              int device, void* stream) {
    return stem_windows_launch<float>(images, planes, n_img, H, W, P, device, stream);
}
extern "C" int rp_preprocess_stem_windows_u8(const uint8_t* images, void* planes, int n_img, int H, int W, int P,
                                             int device, void* stream) {
    return stem_windows_launch<uint8_t>(images, planes, n_img, H, W, P, device, stream);
}

extern "C" int rp_stem_weight_windows_f32(const float* w, float* out, int O, int device, void* stream) {
    RP_REQUIRE(w && out && O > 0, RP_EINVAL, "rp_stem_weight_windows: bad argument");
    RP_GUARD(device);
    rp::launch(stem_weight_windows_kernel, dim3((O * 256 + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, w, out, O);
    return rp::finish_launch("rp_stem_weight_windows");
}

extern "C" int rp_maxpool3x3s2_nhwc_f32(const float* x, float* y, int n_img, int H, int W, int C, int device, void* stream) {
    RP_REQUIRE(x && y && n_img > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0, RP_EINVAL, "rp_maxpool3x3s2: bad argument");
    RP_REQUIRE(rp::aligned16(x) && rp::aligned16(y), RP_EALIGN, "rp_maxpool3x3s2: 16-byte alignment");
    RP_GUARD(device);
    int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    long long total = (long long)n_img * Ho * Wo * (C / 4);
    rp::launch(maxpool3x3s2_nhwc_kernel, dim3(grid_for(total, device)), dim3(256), (size_t)(0), (cudaStream_t)stream, 
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n_img, H, W, C / 4, Ho, Wo);
    return rp::finish_launch("rp_maxpool3x3s2");
}

extern "C" int rp_permute_conv_weight_f32(const float* w, float* out, int O, int C, int KH, int KW, int Cp, int device,
                                          void* stream) {
    RP_REQUIRE(w && out && O > 0 && C > 0 && KH > 0 && KW > 0 && Cp >= C, RP_EINVAL, "rp_permute_conv_weight: bad argument");
    RP_GUARD(device);
    long long total = (long long)O * KH * KW * Cp;
    rp::launch(permute_conv_weight_kernel, dim3(grid_for(total, device)), dim3(256), (size_t)(0), (cudaStream_t)stream, w, out, O, C, KH, KW, Cp);
    return rp::finish_launch("rp_permute_conv_weight");
}

extern "C" int rp_bn_fold_f32(const float* gamma, const float* beta, const float* mean, const float* var,
                              const float* conv_bias, float eps, float* scale, float* shift, int C, int device,
                              void* stream) {
    RP_REQUIRE(gamma && beta && mean && var && scale && shift && C > 0, RP_EINVAL, "rp_bn_fold: bad argument");
    RP_GUARD(device);
    rp::launch(bn_fold_kernel, dim3((C + 127) / 128), dim3(128), (size_t)(0), (cudaStream_t)stream, gamma, beta, mean, var, conv_bias, eps, scale, shift, C);
    return rp::finish_launch("rp_bn_fold");
}
