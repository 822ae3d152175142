// FP32 (true fp32 operands, fp32 accumulate) NT GEMM engine, SIMT FFMA path, with two front ends:
//   * nn.Linear:    C[M,N] = act(A[M,K] W[N,K]^T + bias) + residual            (rp_linear_f32)
//   * nn.Conv2d as implicit GEMM over NHWC activations, BatchNorm(eval) folded into a per-channel
//     scale/shift epilogue, optional residual before / after the activation     (rp_conv2d_nhwc_f32)
// This is the accuracy-first engine behind the fp32 parity configuration (BASELINE.json config 2):
// TF32 / single-pass bf16 tensor-core operands cannot hold the 1e-4 parity bar,
// and library convolutions (cuDNN picks FFT / Winograd algorithms per shape) were measured to break
// it as well (3.6e-2 abs error on the CNN tokens at B=2), so the CNN front end runs here too.
// Reference call sites: vision_transformer.py:323,331 (qkv, proj), vit_layers/mlp.py:21-24 (fc1+GELU,
// fc2), src/model.py:91-98 (pose regressor), src/model.py:127-134 + extractor.py:51-65 (convolutions).
//
// Tiling: 128x96 CTA tile, 16-wide K slabs, 3-stage cp.async (LDGSTS) ring, 256 threads each owning an
// 8x6 register tile with rows/cols interleaved by 16 so that every LDS.128 is conflict-free
// (row stride 20 floats: 16 consecutive rows cover all 32 banks exactly twice = the 2-wavefront
// minimum for 256 B).  Small-M problems (the 26880->512 regressor is weight-bandwidth-bound) are
// split along K so that every SM streams a slice of W exactly once; partial tiles are reduced in a
// fixed order (deterministic) by a second kernel that also applies the epilogue.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 96, BK = 16, TM = 8, TN = 6, TXN = BN / TN, TYN = BM / TM;
constexpr int LDT = BK + 4;   // padded row stride (floats); keeps 16-byte alignment for cp.async
constexpr int STAGES = 3;
constexpr int THREADS = TXN * TYN;
static_assert(THREADS == 256, "thread grid");
constexpr int SMEM_BYTES = STAGES * (BM + BN) * LDT * (int)sizeof(float);
constexpr int A_CHUNKS = BM * (BK / 4) / THREADS;   // 16-byte chunks of the A tile per thread (2)

struct Epilogue {
    const float* scale;      // [N] or null (1)
    const float* shift;      // [N] or null (0)      (the nn.Linear bias / folded BN shift)
    const float* res_pre;    // [M,N] or null: added BEFORE the activation (ResNet identity)
    const float* res_post;   // added AFTER the activation; row index taken modulo res_post_rows
    int res_post_rows;       // 0: same rows as C
    int act;
};

struct ConvGeom {            // NHWC input, weights [O][KH][KW][C]
    int H, W, C, KH, KW, stride, pad, Ho, Wo;
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == RP_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    if (act == RP_ACT_RELU) return fmaxf(v, 0.0f);
    return v;
}

__device__ __forceinline__ float epilogue_apply(float v, const Epilogue& e, int row, int col, int N) {
    if (e.scale) v *= e.scale[col];
    if (e.shift) v += e.shift[col];
    if (e.res_pre) v += e.res_pre[(size_t)row * N + col];
    v = apply_act(v, e.act);
    if (e.res_post) {
        int r = e.res_post_rows > 0 ? row % e.res_post_rows : row;
        v += e.res_post[(size_t)r * N + col];
    }
    return v;
}

template <bool CONV>
__global__ void __launch_bounds__(THREADS, 2)
sgemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ W, Epilogue ep, float* C,
                float* __restrict__ partial, int M, int N, int K, int k_per_split, ConvGeom g) {
    extern __shared__ __align__(16) float smem[];
    float(*As)[BM][LDT] = reinterpret_cast<float(*)[BM][LDT]>(smem);
    float(*Bs)[BN][LDT] = reinterpret_cast<float(*)[BN][LDT]>(smem + STAGES * BM * LDT);

    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int ktiles = (kend - kbeg + BK - 1) / BK;

    // implicit-GEMM row decomposition: each thread always gathers the same A_CHUNKS rows of the tile
    int a_iy0[A_CHUNKS], a_ix0[A_CHUNKS];
    const float* a_img[A_CHUNKS];
    if (CONV) {
#pragma unroll
        for (int i = 0; i < A_CHUNKS; ++i) {
            int r = (tid + i * THREADS) / (BK / 4);
            int gr = min(m0 + r, M - 1);
            int ox = gr % g.Wo, oy = (gr / g.Wo) % g.Ho, n = gr / (g.Wo * g.Ho);
            a_iy0[i] = oy * g.stride - g.pad;
            a_ix0[i] = ox * g.stride - g.pad;
            a_img[i] = A + (size_t)n * g.H * g.W * g.C;
        }
    }

    auto load_tile = [&](int stage, int kt) {
        const int k0 = kbeg + kt * BK;
#pragma unroll
        for (int i = 0; i < A_CHUNKS; ++i) {
            int c = tid + i * THREADS;
            int r = c / (BK / 4), kc = (c % (BK / 4)) * 4;
            int gr = m0 + r, gk = k0 + kc;
            bool ok = (gr < M) && (gk < kend);
            const float* src;
            if (CONV) {
                int gkc = min(gk, K - 4);
                int tap = gkc / g.C, ch = gkc - tap * g.C;       // k = (ky*KW + kx)*C + ch, C % 4 == 0
                int ky = tap / g.KW, kx = tap - ky * g.KW;
                int iy = a_iy0[i] + ky, ix = a_ix0[i] + kx;
                bool in = (iy >= 0) && (iy < g.H) && (ix >= 0) && (ix < g.W);
                ok = ok && in;                                   // zero padding = zero fill
                iy = min(max(iy, 0), g.H - 1);
                ix = min(max(ix, 0), g.W - 1);
                src = a_img[i] + ((size_t)iy * g.W + ix) * g.C + ch;
            } else {
                src = A + (size_t)min(gr, M - 1) * K + min(gk, K - 4);
            }
            rp::cp_async16_zfill(&As[stage][r][kc], src, ok);
        }
#pragma unroll
        for (int c = tid; c < BN * (BK / 4); c += THREADS) {
            int r = c / (BK / 4), kc = (c % (BK / 4)) * 4;
            int gr = n0 + r, gk = k0 + kc;
            bool ok = (gr < N) && (gk < kend);
            const float* src = W + (size_t)min(gr, N - 1) * K + min(gk, K - 4);
            rp::cp_async16_zfill(&Bs[stage][r][kc], src, ok);
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ktiles) load_tile(s, s);
        rp::cp_async_commit();
    }

    for (int kt = 0; kt < ktiles; ++kt) {
        rp::cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < ktiles) load_tile(nk % STAGES, nk);
            rp::cp_async_commit();
        }
        const int st = kt % STAGES;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            float4 a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(&As[st][ty + i * TYN][kk]);
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const float4*>(&Bs[st][tx + j * TXN][kk]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    float s = acc[i][j];
                    s = fmaf(a[i].x, b[j].x, s);
                    s = fmaf(a[i].y, b[j].y, s);
                    s = fmaf(a[i].z, b[j].z, s);
                    s = fmaf(a[i].w, b[j].w, s);
                    acc[i][j] = s;
                }
        }
    }
    rp::cp_async_wait<0>();

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int row = m0 + ty + i * TYN;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int col = n0 + tx + j * TXN;
            if (col >= N) continue;
            size_t o = (size_t)row * N + col;
            if (partial) partial[(size_t)blockIdx.z * M * N + o] = acc[i][j];
            else C[o] = epilogue_apply(acc[i][j], ep, row, col, N);
        }
    }
}

__global__ void __launch_bounds__(256)
splitk_epilogue_kernel(const float* __restrict__ partial, Epilogue ep, float* C, int M, int N, int splits) {
    size_t total = (size_t)M * N;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int s = 0; s < splits; ++s) v += partial[(size_t)s * total + o];   // fixed order: deterministic
        C[o] = epilogue_apply(v, ep, (int)(o / N), (int)(o % N), N);
    }
}

// The split-K plan is a pure function of the shape (B200: 148 SMs) so results are reproducible
// bit-for-bit from run to run and the workspace query needs no device.
constexpr int PLAN_SMS = 148;

struct Plan {
    int tiles_m, tiles_n, splits, k_per_split;
};

Plan make_plan(int M, int N, int K, int sms) {
    Plan p;
    p.tiles_m = (M + BM - 1) / BM;
    p.tiles_n = (N + BN - 1) / BN;
    int tiles = p.tiles_m * p.tiles_n;
    int ktiles = (K + BK - 1) / BK;
    p.splits = 1;
    if (tiles < sms && ktiles >= 16) {
        int want = (2 * sms + tiles - 1) / tiles;
        int maxs = ktiles / 8;                     // at least 8 K-slabs (128 k) per split
        p.splits = max(1, min(want, maxs));
    }
    int kt_per = (ktiles + p.splits - 1) / p.splits;
    p.k_per_split = kt_per * BK;
    p.splits = (K + p.k_per_split - 1) / p.k_per_split;
    return p;
}

template <bool CONV>
int launch_gemm(const float* A, const float* W, const Epilogue& ep, float* C, int M, int N, int K, ConvGeom g,
                void* workspace, size_t workspace_bytes, int device, cudaStream_t st, const char* what) {
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(sgemm_nt_kernel<CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) {
            rp::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    Plan p = make_plan(M, N, K, PLAN_SMS);
    float* partial = nullptr;
    if (p.splits > 1) {
        size_t need = (size_t)p.splits * M * N * sizeof(float);
        RP_REQUIRE(workspace && workspace_bytes >= need, RP_EWORKSPACE, "%s: workspace %zu < %zu bytes (split-K %d)",
                   what, workspace_bytes, need, p.splits);
        partial = static_cast<float*>(workspace);
    }
    dim3 grid(p.tiles_n, p.tiles_m, p.splits);
    sgemm_nt_kernel<CONV><<<grid, THREADS, SMEM_BYTES, st>>>(A, W, ep, C, partial, M, N, K, p.k_per_split, g);
    int rc = rp::finish_launch(what);
    if (rc != RP_OK) return rc;
    if (p.splits > 1) {
        size_t total = (size_t)M * N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 2048) blocks = 2048;
        splitk_epilogue_kernel<<<blocks, 256, 0, st>>>(partial, ep, C, M, N, p.splits);
        rc = rp::finish_launch(what);
    }
    return rc;
}

}  // namespace

extern "C" size_t rp_linear_workspace_bytes(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    Plan p = make_plan(M, N, K, PLAN_SMS);
    return p.splits > 1 ? (size_t)p.splits * M * N * sizeof(float) : 0;
}

extern "C" int rp_linear_f32(const float* A, const float* W, const float* bias, const float* residual, float* C,
                             int M, int N, int K, int act, void* workspace, size_t workspace_bytes, int device,
                             void* stream) {
    RP_REQUIRE(A && W && C, RP_EINVAL, "rp_linear: null pointer");
    RP_REQUIRE(M > 0 && N > 0 && K >= 4 && (K % 4) == 0, RP_EINVAL, "rp_linear: bad shape M=%d N=%d K=%d (K%%4==0)", M, N, K);
    RP_REQUIRE(act >= RP_ACT_NONE && act <= RP_ACT_RELU, RP_EINVAL, "rp_linear: bad act %d", act);
    RP_REQUIRE(rp::aligned16(A) && rp::aligned16(W), RP_EALIGN, "rp_linear: A/W must be 16-byte aligned");
    RP_GUARD(device);
    Epilogue ep{nullptr, bias, nullptr, residual, 0, act};
    return launch_gemm<false>(A, W, ep, C, M, N, K, ConvGeom{}, workspace, workspace_bytes, device,
                              (cudaStream_t)stream, "rp_linear");
}

extern "C" size_t rp_conv2d_workspace_bytes(int n_img, int H, int W, int C, int O, int KH, int KW, int stride, int pad) {
    if (n_img <= 0 || H <= 0 || W <= 0 || C <= 0 || O <= 0 || KH <= 0 || KW <= 0 || stride <= 0) return 0;
    int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    return rp_linear_workspace_bytes(n_img * Ho * Wo, O, KH * KW * C);
}

extern "C" int rp_conv2d_nhwc_f32(const float* x, const float* w, const float* scale, const float* shift,
                                  const float* res_pre, const float* res_post, int res_post_rows, float* y,
                                  int n_img, int H, int W, int C, int O, int KH, int KW, int stride, int pad, int act,
                                  void* workspace, size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(x && w && y, RP_EINVAL, "rp_conv2d: null pointer");
    RP_REQUIRE(n_img > 0 && H > 0 && W > 0 && C >= 4 && (C % 4) == 0 && O > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0,
               RP_EINVAL, "rp_conv2d: bad shape n=%d H=%d W=%d C=%d (C%%4==0) O=%d k=%dx%d s=%d p=%d", n_img, H, W, C, O, KH, KW, stride, pad);
    RP_REQUIRE(act >= RP_ACT_NONE && act <= RP_ACT_RELU, RP_EINVAL, "rp_conv2d: bad act %d", act);
    RP_REQUIRE(rp::aligned16(x) && rp::aligned16(w), RP_EALIGN, "rp_conv2d: x/w must be 16-byte aligned");
    ConvGeom g;
    g.H = H; g.W = W; g.C = C; g.KH = KH; g.KW = KW; g.stride = stride; g.pad = pad;
    g.Ho = (H + 2 * pad - KH) / stride + 1;
    g.Wo = (W + 2 * pad - KW) / stride + 1;
    RP_REQUIRE(g.Ho > 0 && g.Wo > 0, RP_EINVAL, "rp_conv2d: empty output");
    long long M = (long long)n_img * g.Ho * g.Wo;
    RP_REQUIRE(M < (1ll << 31), RP_EINVAL, "rp_conv2d: too many output pixels");
    RP_GUARD(device);
    Epilogue ep{scale, shift, res_pre, res_post, res_post_rows, act};
    return launch_gemm<true>(x, w, ep, y, (int)M, O, KH * KW * C, g, workspace, workspace_bytes, device,
                             (cudaStream_t)stream, "rp_conv2d");
}
