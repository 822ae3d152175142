// Helpers of the CNN front end (A2/A3, src/model.py:127-134, extractor.py:51-65) around the
// implicit-GEMM convolution in gemm_simt.cu: NHWC preprocessing, max-pool, weight re-layout and
// eval-mode BatchNorm folding.  All HBM-bound, channel-innermost (coalesced, 16-byte vectors).
#include "common.cuh"

namespace {

// A1 fused with the layout change: BGR NCHW image -> normalised RGB NHWC with C padded 3 -> 4.
template <typename T>
__global__ void __launch_bounds__(256) preprocess_nhwc4_kernel(const T* __restrict__ img, float4* __restrict__ out,
                                                                int n_img, int H, int W, float scale_h, float scale_w) {
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
    preprocess_nhwc4_kernel<float><<<grid_for(total, device), 256, 0, (cudaStream_t)stream>>>(
        images, reinterpret_cast<float4*>(out), n_img, H, W, (float)H / (float)224, (float)W / (float)224);
    return rp::finish_launch("rp_preprocess_nhwc4");
}

extern "C" int rp_preprocess_nhwc4_u8(const uint8_t* images, float* out, int n_img, int H, int W, int device, void* stream) {
    RP_REQUIRE(images && out && n_img > 0 && H > 0 && W > 0, RP_EINVAL, "rp_preprocess_nhwc4: bad argument");
    RP_REQUIRE(rp::aligned16(out), RP_EALIGN, "rp_preprocess_nhwc4: out must be 16-byte aligned");
    RP_GUARD(device);
    long long total = (long long)n_img * 224 * 224;
    preprocess_nhwc4_kernel<uint8_t><<<grid_for(total, device), 256, 0, (cudaStream_t)stream>>>(
        images, reinterpret_cast<float4*>(out), n_img, H, W, (float)H / (float)224, (float)W / (float)224);
    return rp::finish_launch("rp_preprocess_nhwc4");
}

extern "C" int rp_maxpool3x3s2_nhwc_f32(const float* x, float* y, int n_img, int H, int W, int C, int device, void* stream) {
    RP_REQUIRE(x && y && n_img > 0 && H > 0 && W > 0 && C > 0 && (C % 4) == 0, RP_EINVAL, "rp_maxpool3x3s2: bad argument");
    RP_REQUIRE(rp::aligned16(x) && rp::aligned16(y), RP_EALIGN, "rp_maxpool3x3s2: 16-byte alignment");
    RP_GUARD(device);
    int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    long long total = (long long)n_img * Ho * Wo * (C / 4);
    maxpool3x3s2_nhwc_kernel<<<grid_for(total, device), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), n_img, H, W, C / 4, Ho, Wo);
    return rp::finish_launch("rp_maxpool3x3s2");
}

extern "C" int rp_permute_conv_weight_f32(const float* w, float* out, int O, int C, int KH, int KW, int Cp, int device,
                                          void* stream) {
    RP_REQUIRE(w && out && O > 0 && C > 0 && KH > 0 && KW > 0 && Cp >= C, RP_EINVAL, "rp_permute_conv_weight: bad argument");
    RP_GUARD(device);
    long long total = (long long)O * KH * KW * Cp;
    permute_conv_weight_kernel<<<grid_for(total, device), 256, 0, (cudaStream_t)stream>>>(w, out, O, C, KH, KW, Cp);
    return rp::finish_launch("rp_permute_conv_weight");
}

extern "C" int rp_bn_fold_f32(const float* gamma, const float* beta, const float* mean, const float* var,
                              const float* conv_bias, float eps, float* scale, float* shift, int C, int device,
                              void* stream) {
    RP_REQUIRE(gamma && beta && mean && var && scale && shift && C > 0, RP_EINVAL, "rp_bn_fold: bad argument");
    RP_GUARD(device);
    bn_fold_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, mean, var, conv_bias, eps, scale, shift, C);
    return rp::finish_launch("rp_bn_fold");
}
