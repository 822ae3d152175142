// Training path (SURVEY.md 8(a) A11 / BASELINE.json config 5): the building blocks of the backward pass and of the
// train-mode forward (batch-statistics BatchNorm, saved softmax probabilities).  Everything here is fp32 SIMT and
// deterministic (no atomics; reductions are two-stage with a fixed order):
//   rp_gemm_f32                 strided-batched C = alpha op(A) op(B) + beta C, row-major, any transposes
//                               (dX = dY W, dW = dY^T X, attention / EM products on materialised 576x576 matrices)
//   rp_im2col_nhwc_f32 / rp_col2im_nhwc_f32     convolution weight / data gradients as GEMMs
//   rp_bn_train_* / rp_bn_bwd_*                 nn.BatchNorm2d in training mode (torchvision resnet18, extractor.py:24-28)
//   rp_layernorm_train_fwd / rp_layernorm_bwd   nn.LayerNorm(eps=1e-6) (vision_transformer.py:396)
//   rp_softmax_{rows,cols}_{fwd,bwd}            softmax(S,-1) / softmax(S,-2) (vision_transformer.py:326,205-206)
//   rp_gelu_{fwd,bwd}, rp_relu_bwd, rp_mul, rp_axpby, rp_colsum, rp_maxpool3x3s2_bwd, rp_normalize_pose_bwd ...
// Reference semantics being differentiated: src/model.py:114-191, src/modules/vision_transformer.py:188-354,
// src/modules/extractor.py:51-65, torchvision BasicBlock; autograd does the rest in the reference.
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

// ============================================================================================ generic GEMM
constexpr int GT = 64, GK = 16;      // 64 x 64 tile, 16-deep slabs, 256 threads x (4 x 4)

struct GemmArgs {
    const float* A; const float* B; float* C;
    int M, N, K, lda, ldb, ldc;
    float alpha, beta;
    int batch_inner;
    long long sAo, sAi, sBo, sBi, sCo, sCi;
};

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs g) {
    __shared__ float As[GK][GT + 4];
    __shared__ float Bs[GK][GT + 4];
    const int bz = blockIdx.z, bo = bz / g.batch_inner, bi = bz % g.batch_inner;
    const float* A = g.A + bo * g.sAo + bi * g.sAi;
    const float* B = g.B + bo * g.sBo + bi * g.sBi;
    float* C = g.C + bo * g.sCo + bi * g.sCi;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < g.K; k0 += GK) {
        // A slab: op(A)[m0..+64][k0..+16]; threads run along the contiguous index of the stored matrix
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = tid + e * 256, m, k;
            if (TA) { m = idx & 63; k = idx >> 6; } else { k = idx & 15; m = idx >> 4; }
            const int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < g.M && gk < g.K) v = TA ? A[(long long)gk * g.lda + gm] : A[(long long)gm * g.lda + gk];
            As[k][m] = v;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = tid + e * 256, n, k;
            if (TB) { k = idx & 15; n = idx >> 4; } else { n = idx & 63; k = idx >> 6; }
            const int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < g.N && gk < g.K) v = TB ? B[(long long)gn * g.ldb + gk] : B[(long long)gk * g.ldb + gn];
            Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty + 16 * i;
        if (gm >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx + 16 * j;
            if (gn >= g.N) continue;
            float* c = C + (long long)gm * g.ldc + gn;
            float v = g.alpha * acc[i][j];
            if (g.beta != 0.f) v += g.beta * *c;
            *c = v;
        }
    }
}

int grid1d(long long total, int device, int per = 256) {
    long long b = (total + per - 1) / per;
    long long cap = (long long)rp::num_sms(device) * 32;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

// ============================================================================================ elementwise
__global__ void gelu_fwd_kernel(const float* __restrict__ z, float* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = z[i];
        y[i] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    }
}
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, float* __restrict__ dz, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = z[i];
        const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
        const float pdf = 0.3989422804014327f * expf(-0.5f * v * v);
        dz[i] = dy[i] * (cdf + v * pdf);
    }
}
__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}
__global__ void mul_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        c[i] = a[i] * b[i];
}
__global__ void axpby_kernel(float alpha, const float* __restrict__ x, float beta, const float* __restrict__ y,
                             float* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = alpha * x[i] + (y ? beta * y[i] : 0.f);
}
// out[r, c] = a[r, c] + b[r % period, c]   (pos_embed broadcast add; period = rows of b)
__global__ void add_bcast_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                      long long rows, int cols, int period) {
    const long long n = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols;
        const int c = (int)(i - r * cols);
        out[i] = a[i] + b[(r % period) * cols + c];
    }
}

// column sums of A (optionally of A .* B) over rows: two deterministic stages.  partial [nblk][cols]
constexpr int CS_ROWS = 64;
__global__ void colsum_partial_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ partial,
                                      long long rows, int cols, int period) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const long long r0 = (long long)blockIdx.y * CS_ROWS, r1 = min(rows, r0 + CS_ROWS);
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) {
        float v = A[r * cols + c];
        if (B) v *= B[r * cols + c];
        s += v;
    }
    (void)period;
    partial[(long long)blockIdx.y * cols + c] = s;
}
// cols % 4 == 0: 256 threads = CW float4 columns x 256 / CW row lanes, lanes combined through shared memory in a fixed order
// (the scalar kernel above walked 256 rows per thread with one load in flight: 17 us per bias gradient)
__global__ void __launch_bounds__(256) colsum_partial_v4_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                                float* __restrict__ partial, long long rows, int cols, int CW) {
    __shared__ float4 sm[256];
    const int c4n = cols >> 2, nl = 256 / CW;
    const int cl = threadIdx.x % CW, rl = threadIdx.x / CW, col4 = blockIdx.x * CW + cl;
    const long long r0 = (long long)blockIdx.y * CS_ROWS, r1 = min(rows, r0 + CS_ROWS);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < nl && col4 < c4n) {
        for (long long r = r0 + rl; r < r1; r += nl) {
            float4 v = reinterpret_cast<const float4*>(A)[r * c4n + col4];
            if (B) {
                const float4 b = reinterpret_cast<const float4*>(B)[r * c4n + col4];
                v.x *= b.x; v.y *= b.y; v.z *= b.z; v.w *= b.w;
            }
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    sm[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < CW && col4 < c4n) {
        float4 t = sm[threadIdx.x];
        for (int l = 1; l < nl; ++l) {
            const float4 v = sm[l * CW + threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        reinterpret_cast<float4*>(partial + (long long)blockIdx.y * cols)[col4] = t;
    }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int nblk, int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += partial[(long long)b * cols + c];
    out[c] = s;
}
// out[p, c] = sum_k a[k * period + p, c]   (gradient of a row-broadcast add: pos_embed)
__global__ void sum_over_period_kernel(const float* __restrict__ a, float* __restrict__ out, int reps, int period, int cols) {
    const long long n = (long long)period * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < reps; ++k) s += a[(long long)k * n + i];
        out[i] = s;
    }
}

// ============================================================================================ LayerNorm
// one warp per row; cols <= 256
__global__ void __launch_bounds__(256) ln_train_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                           const float* __restrict__ b, float* __restrict__ y,
                                                           float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                           int rows, int cols, float eps) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * cols;
    float v[8], s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        v[i] = c < cols ? xr[c] : 0.f;
        s += v[i];
    }
    const float mean = rp::warp_sum(s) / (float)cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        const float d = c < cols ? v[i] - mean : 0.f;
        q += d * d;
    }
    const float rstd = 1.0f / sqrtf(rp::warp_sum(q) / (float)cols + eps);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        if (c < cols) y[(size_t)row * cols + c] = (v[i] - mean) * rstd * g[c] + b[c];
    }
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}
// dx per row; per-block partial dgamma / dbeta (block = LN_BWD_ROWS rows, four per warp) -> partial [nblk][2][cols]
constexpr int LN_BWD_ROWS = 32;
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ g, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, float* __restrict__ dx,
                                                     float* __restrict__ partial, int rows, int cols) {
    __shared__ float sg[8][256], sb[8][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float ga[8], ba[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { ga[i] = 0.f; ba[i] = 0.f; }
    for (int j = 0; j < LN_BWD_ROWS / 8; ++j) {
        const int row = blockIdx.x * LN_BWD_ROWS + j * 8 + w;
        if (row >= rows) break;
        float dyv[8], xh[8];
        const float m = mean[row], rs = rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            dyv[i] = 0.f; xh[i] = 0.f;
            if (c < cols) {
                dyv[i] = dy[(size_t)row * cols + c];
                xh[i] = (x[(size_t)row * cols + c] - m) * rs;
                const float gd = dyv[i] * g[c];
                s1 += gd;
                s2 += gd * xh[i];
            }
        }
        s1 = rp::warp_sum(s1) / (float)cols;
        s2 = rp::warp_sum(s2) / (float)cols;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
            if (c < cols) dx[(size_t)row * cols + c] = rs * (dyv[i] * g[c] - s1 - xh[i] * s2);
            ga[i] = fmaf(dyv[i], xh[i], ga[i]);
            ba[i] += dyv[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        if (c < 256) { sg[w][c] = ga[i]; sb[w][c] = ba[i]; }
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c < cols) {
        float a = 0.f, b2 = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) { a += sg[r][c]; b2 += sb[r][c]; }
        partial[((size_t)blockIdx.x * 2 + 0) * cols + c] = a;
        partial[((size_t)blockIdx.x * 2 + 1) * cols + c] = b2;
    }
}
// 32 columns per block, 8 threads per column each summing every 8th partial (fixed order), then a fixed-order
// shared-memory reduction: deterministic, and 8x the loads in flight of the one-thread-per-column version (which took
// 88 us per LayerNorm at 6912 rows: 26 of them were 1.1 ms of the training step).
__global__ void __launch_bounds__(256) ln_bwd_final_kernel(const float* __restrict__ partial, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int nblk, int cols) {
    __shared__ float sa[8][33], sb[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a = 0.f, b = 0.f;
    if (c < cols) {
        for (int k = ty; k < nblk; k += 8) {
            a += partial[((size_t)k * 2 + 0) * cols + c];
            b += partial[((size_t)k * 2 + 1) * cols + c];
        }
    }
    sa[ty][tx] = a; sb[ty][tx] = b;
    __syncthreads();
    if (ty == 0 && c < cols) {
        float ra = 0.f, rb = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) { ra += sa[r][tx]; rb += sb[r][tx]; }
        dgamma[c] = ra;
        dbeta[c] = rb;
    }
}

// ============================================================================================ softmax
// rows: one warp per row of length n (n <= 1024): p = softmax(scale * s)
__global__ void __launch_bounds__(256) softmax_rows_fwd_kernel(const float* __restrict__ s, float* __restrict__ p,
                                                               long long rows, int n, float scale) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* sr = s + row * n;
    float mx = -INFINITY;
    for (int c = lane; c < n; c += 32) mx = fmaxf(mx, sr[c] * scale);
    mx = rp::warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < n; c += 32) sum += expf(sr[c] * scale - mx);
    sum = rp::warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int c = lane; c < n; c += 32) p[row * n + c] = expf(sr[c] * scale - mx) * inv;
}
// ds = scale * p .* (dp - sum(dp .* p))     (accumulate: ds += ...)
__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ p,
                                                               float* __restrict__ ds, long long rows, int n, float scale,
                                                               int accumulate) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float dot = 0.f;
    for (int c = lane; c < n; c += 32) dot += dp[row * n + c] * p[row * n + c];
    dot = rp::warp_sum(dot);
    for (int c = lane; c < n; c += 32) {
        const float v = scale * p[row * n + c] * (dp[row * n + c] - dot);
        ds[row * n + c] = accumulate ? ds[row * n + c] + v : v;
    }
}
// cols: matrices [mats][n][n]; softmax over the ROW index (dim -2).  Block = 32 columns x 8 row phases.
__global__ void __launch_bounds__(256) softmax_cols_fwd_kernel(const float* __restrict__ s, float* __restrict__ p, int n,
                                                               float scale) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    const float* sm = s + (size_t)blockIdx.y * n * n;
    float* pm = p + (size_t)blockIdx.y * n * n;
    float mx = -INFINITY;
    if (col < n) for (int r = ry; r < n; r += 8) mx = fmaxf(mx, sm[(size_t)r * n + col] * scale);
    red[ry][cx] = mx;
    __syncthreads();
    mx = red[0][cx];
#pragma unroll
    for (int k = 1; k < 8; ++k) mx = fmaxf(mx, red[k][cx]);
    __syncthreads();
    float sum = 0.f;
    if (col < n) for (int r = ry; r < n; r += 8) sum += expf(sm[(size_t)r * n + col] * scale - mx);
    red[ry][cx] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += red[k][cx];
    const float inv = 1.0f / sum;
    if (col < n) for (int r = ry; r < n; r += 8) pm[(size_t)r * n + col] = expf(sm[(size_t)r * n + col] * scale - mx) * inv;
}
__global__ void __launch_bounds__(256) softmax_cols_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ p,
                                                               float* __restrict__ ds, int n, float scale, int accumulate) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cx;
    const size_t base = (size_t)blockIdx.y * n * n;
    float dot = 0.f;
    if (col < n) for (int r = ry; r < n; r += 8) dot += dp[base + (size_t)r * n + col] * p[base + (size_t)r * n + col];
    red[ry][cx] = dot;
    __syncthreads();
    dot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) dot += red[k][cx];
    if (col < n)
        for (int r = ry; r < n; r += 8) {
            const size_t i = base + (size_t)r * n + col;
            const float v = scale * p[i] * (dp[i] - dot);
            ds[i] = accumulate ? ds[i] + v : v;
        }
}

// ============================================================================================ BatchNorm (train)
// x [M][C] (NHWC flattened), C % 4 == 0.  Stage 1: per-block partial sums over BN_ROWS rows (256 threads = C/4 float4
// columns x 256/(C/4) row lanes, four loads in flight per thread, lanes combined through shared memory in a fixed
// order); stage 2: the blocks' partials in double (four lanes per channel), mean / biased variance / running statistics.
// Round 1's version walked 512 rows per 64-thread block with one scalar load in flight: 76 us for the 38 MB stem tensor.
constexpr int BN_ROWS = 128;
constexpr int BN_THREADS = 256;
constexpr int BN_FINAL_LANES = 16;                                   // threads per channel in the second stage (block = 64 channels x 16)

// one block's partial sums of (a, b) per channel: a = f0(row), b = f1(row), written to partial[blk][2][C]
template <typename F>
__device__ __forceinline__ void bn_block_sums(float* __restrict__ partial, long long M, int C, F&& load) {
    __shared__ float4 sm[2][BN_THREADS];
    const int c4n = C >> 2;
    const int nl = BN_THREADS / c4n;                                  // row lanes (C = 64: 16, 128: 8, 192: 5)
    const int col4 = threadIdx.x % c4n, rl = threadIdx.x / c4n;
    const long long r0 = (long long)blockIdx.x * BN_ROWS, r1 = min(M, r0 + BN_ROWS);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    if (rl < nl) {
        long long r = r0 + rl;
        for (; r + 3 * nl < r1; r += 4 * nl) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) load((r + u * nl) * c4n + col4, col4, a[u], b[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s.x += a[u].x; s.y += a[u].y; s.z += a[u].z; s.w += a[u].w;
                q.x += b[u].x; q.y += b[u].y; q.z += b[u].z; q.w += b[u].w;
            }
        }
        for (; r < r1; r += nl) {
            float4 a, b;
            load(r * c4n + col4, col4, a, b);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            q.x += b.x; q.y += b.y; q.z += b.z; q.w += b.w;
        }
    }
    sm[0][threadIdx.x] = s;
    sm[1][threadIdx.x] = q;
    __syncthreads();
    if (threadIdx.x < 2 * c4n) {                                     // threads [0,c4n): sums, [c4n,2 c4n): second sums
        const int which = threadIdx.x / c4n, c = threadIdx.x % c4n;
        float4 t = sm[which][c];
        for (int l = 1; l < nl; ++l) {
            const float4 v = sm[which][l * c4n + c];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * 2 + which) * C)[c] = t;
    }
}

__global__ void __launch_bounds__(BN_THREADS) bn_stats_partial_kernel(const float* __restrict__ x, float* __restrict__ partial,
                                                                       long long M, int C) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    bn_block_sums(partial, M, C, [&](long long i, int, float4& a, float4& b) {
        a = x4[i];
        b = make_float4(a.x * a.x, a.y * a.y, a.z * a.z, a.w * a.w);
    });
}
// partial [nblk][2][C] -> sums in double: thread = (channel, one of BN_FINAL_LANES lanes over the blocks), fixed combination order
__device__ __forceinline__ void bn_final_sums(const float* __restrict__ partial, int nblk, int C, double& s, double& q, int& c,
                                              bool& writer) {
    __shared__ double sm[2][BN_FINAL_LANES][64];
    const int cl = threadIdx.x & 63, l = threadIdx.x >> 6;
    c = blockIdx.x * 64 + cl;
    double ss = 0.0, qq = 0.0;
    if (c < C)
        for (int b = l; b < nblk; b += BN_FINAL_LANES) {
            ss += (double)partial[((long long)b * 2 + 0) * C + c];
            qq += (double)partial[((long long)b * 2 + 1) * C + c];
        }
    sm[0][l][cl] = ss;
    sm[1][l][cl] = qq;
    __syncthreads();
    writer = l == 0 && c < C;
    s = 0.0; q = 0.0;
    if (writer)
        for (int k = 0; k < BN_FINAL_LANES; ++k) { s += sm[0][k][cl]; q += sm[1][k][cl]; }
}
// mean / biased var in double (the one-pass sum of squares is safe there); also the running-stat update
__global__ void __launch_bounds__(64 * BN_FINAL_LANES) bn_stats_final_kernel(const float* __restrict__ partial, int nblk, long long M, int C,
                                                              float* __restrict__ mean, float* __restrict__ var,
                                                              float* __restrict__ running_mean, float* __restrict__ running_var,
                                                              float momentum) {
    double s, q;
    int c;
    bool writer;
    bn_final_sums(partial, nblk, C, s, q, c, writer);
    if (!writer) return;
    const double m = s / (double)M;
    double v = q / (double)M - m * m;
    if (v < 0.0) v = 0.0;
    mean[c] = (float)m;
    var[c] = (float)v;
    if (running_mean) {
        const double unbiased = M > 1 ? v * (double)M / (double)(M - 1) : v;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}
// y = act(gamma * (x - mean) * rstd + beta + residual), float4 per thread
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                                        const float* __restrict__ var, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ residual,
                                                        float* __restrict__ y, long long M, int C, float eps, int relu) {
    const long long n4 = M * C / 4;
    const int c4n = C >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        const float4 xv = reinterpret_cast<const float4*>(x)[i];
        const float4 mv = *reinterpret_cast<const float4*>(mean + c), vv = *reinterpret_cast<const float4*>(var + c);
        const float4 gv = *reinterpret_cast<const float4*>(gamma + c), bv = *reinterpret_cast<const float4*>(beta + c);
        float4 o;
        o.x = (xv.x - mv.x) * (1.0f / sqrtf(vv.x + eps)) * gv.x + bv.x;
        o.y = (xv.y - mv.y) * (1.0f / sqrtf(vv.y + eps)) * gv.y + bv.y;
        o.z = (xv.z - mv.z) * (1.0f / sqrtf(vv.z + eps)) * gv.z + bv.z;
        o.w = (xv.w - mv.w) * (1.0f / sqrtf(vv.w + eps)) * gv.w + bv.w;
        if (residual) {
            const float4 r = reinterpret_cast<const float4*>(residual)[i];
            o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        reinterpret_cast<float4*>(y)[i] = o;
    }
}
// partial [nblk][2][C]: sum(dz), sum(dz * xhat) with dz = dy .* (y > 0) when the block ended in a ReLU (y != NULL)
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_partial_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                                     const float* __restrict__ x, const float* __restrict__ mean,
                                                                     const float* __restrict__ var, float eps,
                                                                     float* __restrict__ partial, long long M, int C) {
    const float4* dy4 = reinterpret_cast<const float4*>(dy);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int col4 = threadIdx.x % (C >> 2);
    const float4 m = *reinterpret_cast<const float4*>(mean + 4 * col4), v = *reinterpret_cast<const float4*>(var + 4 * col4);
    const float4 rs = make_float4(1.0f / sqrtf(v.x + eps), 1.0f / sqrtf(v.y + eps), 1.0f / sqrtf(v.z + eps), 1.0f / sqrtf(v.w + eps));
    bn_block_sums(partial, M, C, [&](long long i, int, float4& a, float4& b) {
        a = dy4[i];
        if (y4) {
            const float4 yv = y4[i];
            a.x = yv.x > 0.f ? a.x : 0.f; a.y = yv.y > 0.f ? a.y : 0.f; a.z = yv.z > 0.f ? a.z : 0.f; a.w = yv.w > 0.f ? a.w : 0.f;
        }
        const float4 xv = x4[i];
        b = make_float4(a.x * (xv.x - m.x) * rs.x, a.y * (xv.y - m.y) * rs.y, a.z * (xv.z - m.z) * rs.z, a.w * (xv.w - m.w) * rs.w);
    });
}
__global__ void __launch_bounds__(64 * BN_FINAL_LANES) bn_bwd_final_kernel(const float* __restrict__ partial, int nblk, int C,
                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
    double s, q;
    int c;
    bool writer;
    bn_final_sums(partial, nblk, C, s, q, c, writer);
    if (!writer) return;
    dbeta[c] = (float)s;
    dgamma[c] = (float)q;
}
// dx = gamma * rstd * (dz - dbeta / M - xhat * dgamma / M); dz (the gradient of the residual branch) is written out on request
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                            const float* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ var, const float* __restrict__ gamma,
                                                            const float* __restrict__ dgamma, const float* __restrict__ dbeta,
                                                            float eps, float* __restrict__ dx, float* __restrict__ dz_out,
                                                            long long M, int C) {
    const long long n4 = M * C / 4;
    const int c4n = C >> 2;
    const float invM = 1.0f / (float)M;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        float4 d = reinterpret_cast<const float4*>(dy)[i];
        if (y) {
            const float4 yv = reinterpret_cast<const float4*>(y)[i];
            d.x = yv.x > 0.f ? d.x : 0.f; d.y = yv.y > 0.f ? d.y : 0.f; d.z = yv.z > 0.f ? d.z : 0.f; d.w = yv.w > 0.f ? d.w : 0.f;
        }
        if (dz_out) reinterpret_cast<float4*>(dz_out)[i] = d;
        const float4 xv = reinterpret_cast<const float4*>(x)[i];
        const float d4[4] = {d.x, d.y, d.z, d.w}, x4[4] = {xv.x, xv.y, xv.z, xv.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float rs = 1.0f / sqrtf(var[c + k] + eps);
            const float xh = (x4[k] - mean[c + k]) * rs;
            o[k] = gamma[c + k] * rs * (d4[k] - dbeta[c + k] * invM - xh * dgamma[c + k] * invM);
        }
        reinterpret_cast<float4*>(dx)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ============================================================================================ conv helpers (NHWC)
// cols [n*Ho*Wo][KH*KW*C], column = (ky*KW + kx)*C + c  (the layout of rp_permute_conv_weight_f32)
__global__ void im2col_kernel(const float* __restrict__ x, float* __restrict__ cols, int n, int H, int W, int C, int KH, int KW,
                              int stride, int pad, int Ho, int Wo) {
    const long long K = (long long)KH * KW * C, total = (long long)n * Ho * Wo * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / K;
        const int k = (int)(i - row * K);
        const int c = k % C, t = k / C, kx = t % KW, ky = t / KW;
        const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), im = (int)(row / ((long long)Wo * Ho));
        const int iy = oy * stride + ky - pad, ix = ox * stride + kx - pad;
        cols[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[(((long long)im * H + iy) * W + ix) * C + c] : 0.f;
    }
}
// The same gather written straight into the operand layout of the weight-gradient GEMM: bf16 planes of cols^T,
// planes[p][k][m] (m = output pixel, two per thread), for the convolutions rp_conv_dw_tc does not take (the 7x7 / 2 stem on
// 4 input channels): no float32 cols matrix, no separate transposing split pass.
__global__ void __launch_bounds__(256) im2col_t_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ planes, int P,
                                                               int n, int H, int W, int C, int KH, int KW, int stride, int pad,
                                                               int Ho, int Wo) {
    const long long M = (long long)n * Ho * Wo, m0 = 2 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (m0 >= M) return;
    const int k = blockIdx.y;
    const int c = k % C, t = k / C, kx = t % KW, ky = t / KW;
    float v[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const long long row = m0 + j;
        const int ox = (int)(row % Wo), oy = (int)((row / Wo) % Ho), im = (int)(row / ((long long)Wo * Ho));
        const int iy = oy * stride + ky - pad, ix = ox * stride + kx - pad;
        v[j] = (row < M && iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[(((long long)im * H + iy) * W + ix) * C + c] : 0.f;
    }
    const long long K = (long long)KH * KW * C;
    for (int p = 0; p < P; ++p) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(planes + ((long long)p * K + k) * M + m0) = h;
        v[0] -= __bfloat162float(h.x);
        v[1] -= __bfloat162float(h.y);
    }
}
// dx[n][iy][ix][c] = sum over (ky,kx) with a valid output pixel of dcols[(oy,ox)][(ky,kx,c)]   (gather: no atomics)
__global__ void col2im_kernel(const float* __restrict__ dcols, float* __restrict__ dx, int n, int H, int W, int C, int KH, int KW,
                              int stride, int pad, int Ho, int Wo) {
    const long long K = (long long)KH * KW * C, total = (long long)n * H * W * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const int ix = (int)(pix % W), iy = (int)((pix / W) % H), im = (int)(pix / ((long long)W * H));
        float s = 0.f;
        for (int ky = 0; ky < KH; ++ky) {
            const int ty = iy + pad - ky;
            if (ty < 0 || ty % stride) continue;
            const int oy = ty / stride;
            if (oy >= Ho) continue;
            for (int kx = 0; kx < KW; ++kx) {
                const int tx = ix + pad - kx;
                if (tx < 0 || tx % stride) continue;
                const int ox = tx / stride;
                if (ox >= Wo) continue;
                s += dcols[(((long long)im * Ho + oy) * Wo + ox) * K + (long long)(ky * KW + kx) * C + c];
            }
        }
        dx[i] = s;
    }
}
// nn.MaxPool2d(3,2,1) backward, NHWC.  PyTorch routes the gradient of a window to its FIRST maximum in (ky,kx)
// scan order; an input pixel gathers from the <= 4 windows that contain it.
// Two passes (the one-pass version re-scanned up to four 3x3 windows per input element: 36 loads each, 277 us for the stem):
//  1. arg[window][c] = position 0..8 of the FIRST maximum of the window (PyTorch's tie-break), 4 channels per thread
//  2. dx[pixel][c] = sum of dy over the <= 4 windows that contain the pixel and whose arg points at it (fixed order)
__global__ void __launch_bounds__(256) maxpool_arg_kernel(const float* __restrict__ x, uchar4* __restrict__ arg, int n, int H, int W,
                                                          int C4, int Ho, int Wo) {
    const long long total = (long long)n * Ho * Wo * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const long long win = i / C4;
        const int ox = (int)(win % Wo), oy = (int)((win / Wo) % Ho), im = (int)(win / ((long long)Wo * Ho));
        float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        uchar4 a = make_uchar4(255, 255, 255, 255);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = 2 * oy - 1 + ky;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = 2 * ox - 1 + kx;
                if (xx < 0 || xx >= W) continue;
                const float4 v = reinterpret_cast<const float4*>(x)[(((long long)im * H + yy) * W + xx) * C4 + c4];
                const unsigned char k = (unsigned char)(ky * 3 + kx);
                if (v.x > mx.x) { mx.x = v.x; a.x = k; }
                if (v.y > mx.y) { mx.y = v.y; a.y = k; }
                if (v.z > mx.z) { mx.z = v.z; a.z = k; }
                if (v.w > mx.w) { mx.w = v.w; a.w = k; }
            }
        }
        arg[i] = a;
    }
}
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ dy, const uchar4* __restrict__ arg,
                                                          float* __restrict__ dx, int n, int H, int W, int C4, int Ho, int Wo) {
    const long long total = (long long)n * H * W * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        const long long pix = i / C4;
        const int ix = (int)(pix % W), iy = (int)((pix / W) % H), im = (int)(pix / ((long long)W * H));
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int oy = iy / 2; oy <= (iy + 1) / 2; ++oy) {                        // windows with 2 oy - 1 <= iy <= 2 oy + 1
            if (oy >= Ho) continue;
            const int ky = iy - (2 * oy - 1);
            for (int ox = ix / 2; ox <= (ix + 1) / 2; ++ox) {
                if (ox >= Wo) continue;
                const unsigned char k = (unsigned char)(ky * 3 + (ix - (2 * ox - 1)));
                const long long w = (((long long)im * Ho + oy) * Wo + ox) * C4 + c4;
                const uchar4 a = arg[w];
                if (a.x == k || a.y == k || a.z == k || a.w == k) {
                    const float4 d = reinterpret_cast<const float4*>(dy)[w];
                    if (a.x == k) s.x += d.x;
                    if (a.y == k) s.y += d.y;
                    if (a.z == k) s.z += d.z;
                    if (a.w == k) s.w += d.w;
                }
            }
        }
        reinterpret_cast<float4*>(dx)[i] = s;
    }
}

// ============================================================================================ misc
// A10 backward: out[:,1,:3] = raw[:,1,:3]; out[:,1,3:] = q / max(|q|, 0.01); pose 0 comes from Gs (no gradient to raw)
__global__ void normalize_pose_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ raw, float* __restrict__ draw, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* r = raw + (long long)b * 14 + 7;
    const float* g = dout + (long long)b * 14 + 7;
    float* d = draw + (long long)b * 14;
#pragma unroll
    for (int i = 0; i < 7; ++i) d[i] = 0.f;
    d[7] = g[0]; d[8] = g[1]; d[9] = g[2];
    const float q[4] = {r[3], r[4], r[5], r[6]}, gq[4] = {g[3], g[4], g[5], g[6]};
    const float nn = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (nn > 0.01f) {
        const float dot = (gq[0] * q[0] + gq[1] * q[1] + gq[2] * q[2] + gq[3] * q[3]) / (nn * nn);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[10 + i] = (gq[i] - q[i] * dot) / nn;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) d[10 + i] = gq[i] * 100.0f;       // clamp active: q / 0.01
    }
}
// V' = [v | pos]: qkv [n_img][576][576] (v = columns 384 + h*64 ..), pos [B][576][6] -> vp [n_img][3][576][70]
__global__ void concat_vpos_kernel(const float* __restrict__ qkv, const float* __restrict__ pos, float* __restrict__ vp, int n_img) {
    const long long total = (long long)n_img * 3 * 576 * 70;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % 70);
        const int tok = (int)((i / 70) % 576), h = (int)((i / (70 * 576)) % 3), im = (int)(i / (70LL * 576 * 3));
        vp[i] = c < 64 ? qkv[((long long)im * 576 + tok) * 576 + 384 + h * 64 + c]
                       : pos[((long long)(im >> 1) * 576 + tok) * 6 + (c - 64)];
    }
}
// gradient of the v part of V' back into dqkv's v columns (the pos columns receive no gradient)
__global__ void scatter_dv_kernel(const float* __restrict__ dvp, float* __restrict__ dqkv, int n_img) {
    const long long total = (long long)n_img * 3 * 576 * 64;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % 64);
        const int tok = (int)((i / 64) % 576), h = (int)((i / (64 * 576)) % 3), im = (int)(i / (64LL * 576 * 3));
        dqkv[((long long)im * 576 + tok) * 576 + 384 + h * 64 + c] = dvp[(((long long)im * 3 + h) * 576 + tok) * 70 + c];
    }
}

}  // namespace

#define RP_LAUNCH1D(kernel, total, ...)                                                         \
    do {                                                                                        \
        kernel<<<grid1d((total), device), 256, 0, (cudaStream_t)stream>>>(__VA_ARGS__);          \
    } while (0)

extern "C" int rp_gemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                           int ldb, float beta, float* C, int ldc, int batch_outer, int batch_inner, int64_t sAo, int64_t sAi,
                           int64_t sBo, int64_t sBi, int64_t sCo, int64_t sCi, int device, void* stream) {
    RP_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch_outer > 0 && batch_inner > 0, RP_EINVAL, "rp_gemm: bad argument");
    RP_REQUIRE((long long)batch_outer * batch_inner <= 65535, RP_EINVAL, "rp_gemm: too many batches");
    RP_GUARD(device);
    GemmArgs g{A, B, C, M, N, K, lda, ldb, ldc, alpha, beta, batch_inner, sAo, sAi, sBo, sBi, sCo, sCi};
    dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT, batch_outer * batch_inner);
    cudaStream_t st = (cudaStream_t)stream;
    if (transA && transB) gemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
    else if (transA) gemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
    else if (transB) gemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
    else gemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
    return rp::finish_launch("rp_gemm");
}

extern "C" int rp_gelu_fwd_f32(const float* z, float* y, int64_t n, int device, void* stream) {
    RP_REQUIRE(z && y && n > 0, RP_EINVAL, "rp_gelu_fwd: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(gelu_fwd_kernel, n, z, y, n);
    return rp::finish_launch("rp_gelu_fwd");
}
extern "C" int rp_gelu_bwd_f32(const float* dy, const float* z, float* dz, int64_t n, int device, void* stream) {
    RP_REQUIRE(dy && z && dz && n > 0, RP_EINVAL, "rp_gelu_bwd: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(gelu_bwd_kernel, n, dy, z, dz, n);
    return rp::finish_launch("rp_gelu_bwd");
}
extern "C" int rp_relu_bwd_f32(const float* dy, const float* y, float* dx, int64_t n, int device, void* stream) {
    RP_REQUIRE(dy && y && dx && n > 0, RP_EINVAL, "rp_relu_bwd: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(relu_bwd_kernel, n, dy, y, dx, n);
    return rp::finish_launch("rp_relu_bwd");
}
extern "C" int rp_mul_f32(const float* a, const float* b, float* c, int64_t n, int device, void* stream) {
    RP_REQUIRE(a && b && c && n > 0, RP_EINVAL, "rp_mul: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(mul_kernel, n, a, b, c, n);
    return rp::finish_launch("rp_mul");
}
extern "C" int rp_axpby_f32(float alpha, const float* x, float beta, const float* y, float* out, int64_t n, int device, void* stream) {
    RP_REQUIRE(x && out && n > 0, RP_EINVAL, "rp_axpby: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(axpby_kernel, n, alpha, x, beta, y, out, n);
    return rp::finish_launch("rp_axpby");
}
extern "C" int rp_add_bcast_rows_f32(const float* a, const float* b, float* out, int64_t rows, int cols, int period, int device,
                                     void* stream) {
    RP_REQUIRE(a && b && out && rows > 0 && cols > 0 && period > 0, RP_EINVAL, "rp_add_bcast_rows: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(add_bcast_rows_kernel, rows * cols, a, b, out, rows, cols, period);
    return rp::finish_launch("rp_add_bcast_rows");
}
extern "C" int rp_sum_over_period_f32(const float* a, float* out, int reps, int period, int cols, int device, void* stream) {
    RP_REQUIRE(a && out && reps > 0 && period > 0 && cols > 0, RP_EINVAL, "rp_sum_over_period: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(sum_over_period_kernel, (long long)period * cols, a, out, reps, period, cols);
    return rp::finish_launch("rp_sum_over_period");
}
extern "C" size_t rp_colsum_workspace_bytes(int64_t rows, int cols) {
    if (rows <= 0 || cols <= 0) return 0;
    return (size_t)((rows + CS_ROWS - 1) / CS_ROWS) * cols * sizeof(float);
}
extern "C" int rp_colsum_f32(const float* A, const float* B, float* out, int64_t rows, int cols, void* workspace,
                             size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(A && out && rows > 0 && cols > 0, RP_EINVAL, "rp_colsum: bad argument");
    RP_REQUIRE(workspace && workspace_bytes >= rp_colsum_workspace_bytes(rows, cols), RP_EWORKSPACE, "rp_colsum: workspace too small");
    RP_GUARD(device);
    const int nblk = (int)((rows + CS_ROWS - 1) / CS_ROWS);
    if (cols % 4 == 0 && rp::aligned16(A) && rp::aligned16(B) && rp::aligned16(workspace)) {
        const int c4n = cols / 4, CW = c4n < 64 ? c4n : 64;
        colsum_partial_v4_kernel<<<dim3((c4n + CW - 1) / CW, nblk), 256, 0, (cudaStream_t)stream>>>(A, B, static_cast<float*>(workspace),
                                                                                                  rows, cols, CW);
    } else {
        dim3 grid((cols + 127) / 128, nblk);
        colsum_partial_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(A, B, static_cast<float*>(workspace), rows, cols, 0);
    }
    colsum_final_kernel<<<(cols + 127) / 128, 128, 0, (cudaStream_t)stream>>>(static_cast<float*>(workspace), out, nblk, cols);
    return rp::finish_launch("rp_colsum");
}

extern "C" int rp_layernorm_train_fwd_f32(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                                          int rows, int cols, float eps, int device, void* stream) {
    RP_REQUIRE(x && gamma && beta && y && mean && rstd && rows > 0 && cols > 0 && cols <= 256, RP_EINVAL, "rp_layernorm_train_fwd: bad argument");
    RP_GUARD(device);
    ln_train_fwd_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean, rstd, rows, cols, eps);
    return rp::finish_launch("rp_layernorm_train_fwd");
}
extern "C" size_t rp_layernorm_bwd_workspace_bytes(int rows, int cols) {
    if (rows <= 0 || cols <= 0) return 0;
    return (size_t)((rows + 7) / 8) * 2 * cols * sizeof(float);
}
extern "C" int rp_layernorm_bwd_f32(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                                    float* dx, float* dgamma, float* dbeta, int rows, int cols, void* workspace,
                                    size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && rows > 0 && cols > 0 && cols <= 256, RP_EINVAL,
               "rp_layernorm_bwd: bad argument");
    RP_REQUIRE(workspace && workspace_bytes >= rp_layernorm_bwd_workspace_bytes(rows, cols), RP_EWORKSPACE, "rp_layernorm_bwd: workspace too small");
    RP_GUARD(device);
    const int nblk = (rows + LN_BWD_ROWS - 1) / LN_BWD_ROWS;
    ln_bwd_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(dy, x, gamma, mean, rstd, dx, static_cast<float*>(workspace), rows, cols);
    ln_bwd_final_kernel<<<(cols + 31) / 32, 256, 0, (cudaStream_t)stream>>>(static_cast<float*>(workspace), dgamma, dbeta, nblk, cols);
    return rp::finish_launch("rp_layernorm_bwd");
}

extern "C" int rp_softmax_rows_fwd_f32(const float* s, float* p, int64_t rows, int n, float scale, int device, void* stream) {
    RP_REQUIRE(s && p && rows > 0 && n > 0, RP_EINVAL, "rp_softmax_rows_fwd: bad argument");
    RP_GUARD(device);
    softmax_rows_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(s, p, rows, n, scale);
    return rp::finish_launch("rp_softmax_rows_fwd");
}
extern "C" int rp_softmax_rows_bwd_f32(const float* dp, const float* p, float* ds, int64_t rows, int n, float scale, int accumulate,
                                       int device, void* stream) {
    RP_REQUIRE(dp && p && ds && rows > 0 && n > 0, RP_EINVAL, "rp_softmax_rows_bwd: bad argument");
    RP_GUARD(device);
    softmax_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(dp, p, ds, rows, n, scale, accumulate);
    return rp::finish_launch("rp_softmax_rows_bwd");
}
extern "C" int rp_softmax_cols_fwd_f32(const float* s, float* p, int mats, int n, float scale, int device, void* stream) {
    RP_REQUIRE(s && p && mats > 0 && mats <= 65535 && n > 0, RP_EINVAL, "rp_softmax_cols_fwd: bad argument");
    RP_GUARD(device);
    softmax_cols_fwd_kernel<<<dim3((n + 31) / 32, mats), 256, 0, (cudaStream_t)stream>>>(s, p, n, scale);
    return rp::finish_launch("rp_softmax_cols_fwd");
}
extern "C" int rp_softmax_cols_bwd_f32(const float* dp, const float* p, float* ds, int mats, int n, float scale, int accumulate,
                                       int device, void* stream) {
    RP_REQUIRE(dp && p && ds && mats > 0 && mats <= 65535 && n > 0, RP_EINVAL, "rp_softmax_cols_bwd: bad argument");
    RP_GUARD(device);
    softmax_cols_bwd_kernel<<<dim3((n + 31) / 32, mats), 256, 0, (cudaStream_t)stream>>>(dp, p, ds, n, scale, accumulate);
    return rp::finish_launch("rp_softmax_cols_bwd");
}

extern "C" size_t rp_bn_workspace_bytes(int64_t M, int C) {
    if (M <= 0 || C <= 0) return 0;
    return (size_t)((M + BN_ROWS - 1) / BN_ROWS) * 2 * C * sizeof(float);
}
extern "C" int rp_bn_train_stats_f32(const float* x, float* mean, float* var, float* running_mean, float* running_var,
                                     float momentum, int64_t M, int C, void* workspace, size_t workspace_bytes, int device,
                                     void* stream) {
    RP_REQUIRE(x && mean && var && M > 0 && C > 0, RP_EINVAL, "rp_bn_train_stats: bad argument");
    RP_REQUIRE(workspace && workspace_bytes >= rp_bn_workspace_bytes(M, C), RP_EWORKSPACE, "rp_bn_train_stats: workspace too small");
    RP_GUARD(device);
    RP_REQUIRE(C % 4 == 0 && C <= 2 * BN_THREADS && rp::aligned16(x), RP_EINVAL, "rp_bn_train_stats: C must be a multiple of 4 (<= 512)");
    const int nblk = (int)((M + BN_ROWS - 1) / BN_ROWS);
    bn_stats_partial_kernel<<<nblk, BN_THREADS, 0, (cudaStream_t)stream>>>(x, static_cast<float*>(workspace), M, C);
    bn_stats_final_kernel<<<(C + 63) / 64, 64 * BN_FINAL_LANES, 0, (cudaStream_t)stream>>>(static_cast<float*>(workspace), nblk, M, C, mean, var,
                                                                         running_mean, running_var, momentum);
    return rp::finish_launch("rp_bn_train_stats");
}
extern "C" int rp_bn_apply_f32(const float* x, const float* mean, const float* var, const float* gamma, const float* beta,
                               const float* residual, float* y, int64_t M, int C, float eps, int relu, int device, void* stream) {
    RP_REQUIRE(x && mean && var && gamma && beta && y && M > 0 && C > 0, RP_EINVAL, "rp_bn_apply: bad argument");
    RP_GUARD(device);
    RP_REQUIRE(C % 4 == 0 && rp::aligned16(x) && rp::aligned16(y) && rp::aligned16(residual), RP_EINVAL, "rp_bn_apply: C % 4, 16-byte alignment");
    RP_LAUNCH1D(bn_apply_kernel, M * C / 4, x, mean, var, gamma, beta, residual, y, M, C, eps, relu);
    return rp::finish_launch("rp_bn_apply");
}
extern "C" int rp_bn_bwd_f32(const float* dy, const float* y_relu, const float* x, const float* mean, const float* var,
                             const float* gamma, float eps, float* dx, float* dz_out, float* dgamma, float* dbeta, int64_t M, int C,
                             void* workspace, size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(dy && x && mean && var && gamma && dx && dgamma && dbeta && M > 0 && C > 0, RP_EINVAL, "rp_bn_bwd: bad argument");
    RP_REQUIRE(workspace && workspace_bytes >= rp_bn_workspace_bytes(M, C), RP_EWORKSPACE, "rp_bn_bwd: workspace too small");
    RP_GUARD(device);
    RP_REQUIRE(C % 4 == 0 && C <= 2 * BN_THREADS && rp::aligned16(dy) && rp::aligned16(x) && rp::aligned16(dx) && rp::aligned16(y_relu) &&
                   rp::aligned16(dz_out), RP_EINVAL, "rp_bn_bwd: C must be a multiple of 4 (<= 512), 16-byte alignment");
    const int nblk = (int)((M + BN_ROWS - 1) / BN_ROWS);
    bn_bwd_partial_kernel<<<nblk, BN_THREADS, 0, (cudaStream_t)stream>>>(dy, y_relu, x, mean, var, eps, static_cast<float*>(workspace), M, C);
    bn_bwd_final_kernel<<<(C + 63) / 64, 64 * BN_FINAL_LANES, 0, (cudaStream_t)stream>>>(static_cast<float*>(workspace), nblk, C, dgamma, dbeta);
    RP_LAUNCH1D(bn_bwd_apply_kernel, M * C / 4, dy, y_relu, x, mean, var, gamma, dgamma, dbeta, eps, dx, dz_out, M, C);
    return rp::finish_launch("rp_bn_bwd");
}

extern "C" int rp_im2col_nhwc_f32(const float* x, float* cols, int n, int H, int W, int C, int KH, int KW, int stride, int pad,
                                  int device, void* stream) {
    RP_REQUIRE(x && cols && n > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0, RP_EINVAL, "rp_im2col: bad argument");
    RP_GUARD(device);
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    RP_LAUNCH1D(im2col_kernel, (long long)n * Ho * Wo * KH * KW * C, x, cols, n, H, W, C, KH, KW, stride, pad, Ho, Wo);
    return rp::finish_launch("rp_im2col");
}
extern "C" int rp_im2col_t_planes_bf16(const float* x, void* planes, int P, int n, int H, int W, int C, int KH, int KW, int stride,
                                       int pad, int device, void* stream) {
    RP_REQUIRE(x && planes && n > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0 && P >= 1 && P <= 2,
               RP_EINVAL, "rp_im2col_t_planes: bad argument");
    RP_GUARD(device);
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    const long long M = (long long)n * Ho * Wo;
    RP_REQUIRE(M % 2 == 0 && KH * KW * C <= 65535, RP_EINVAL, "rp_im2col_t_planes: even pixel count, K <= 65535");
    im2col_t_planes_kernel<<<dim3((unsigned)((M / 2 + 255) / 256), KH * KW * C), 256, 0, (cudaStream_t)stream>>>(
        x, static_cast<__nv_bfloat16*>(planes), P, n, H, W, C, KH, KW, stride, pad, Ho, Wo);
    return rp::finish_launch("rp_im2col_t_planes");
}
extern "C" int rp_col2im_nhwc_f32(const float* dcols, float* dx, int n, int H, int W, int C, int KH, int KW, int stride, int pad,
                                  int device, void* stream) {
    RP_REQUIRE(dcols && dx && n > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0, RP_EINVAL, "rp_col2im: bad argument");
    RP_GUARD(device);
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    RP_LAUNCH1D(col2im_kernel, (long long)n * H * W * C, dcols, dx, n, H, W, C, KH, KW, stride, pad, Ho, Wo);
    return rp::finish_launch("rp_col2im");
}
extern "C" size_t rp_maxpool3x3s2_bwd_workspace_bytes(int n, int H, int W, int C) {
    if (n <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
    return (size_t)n * ((H - 1) / 2 + 1) * ((W - 1) / 2 + 1) * C;               // one byte per (window, channel)
}
extern "C" int rp_maxpool3x3s2_bwd_f32(const float* dy, const float* x, float* dx, int n, int H, int W, int C, void* workspace,
                                       size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(dy && x && dx && workspace && n > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, RP_EINVAL, "rp_maxpool3x3s2_bwd: bad argument");
    RP_REQUIRE(workspace_bytes >= rp_maxpool3x3s2_bwd_workspace_bytes(n, H, W, C), RP_EWORKSPACE, "rp_maxpool3x3s2_bwd: workspace too small");
    RP_REQUIRE(rp::aligned16(dy) && rp::aligned16(x) && rp::aligned16(dx) && rp::aligned16(workspace), RP_EALIGN, "rp_maxpool3x3s2_bwd: 16-byte alignment");
    RP_GUARD(device);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    uchar4* arg = static_cast<uchar4*>(workspace);
    RP_LAUNCH1D(maxpool_arg_kernel, (long long)n * Ho * Wo * (C / 4), x, arg, n, H, W, C / 4, Ho, Wo);
    RP_LAUNCH1D(maxpool_bwd_kernel, (long long)n * H * W * (C / 4), dy, arg, dx, n, H, W, C / 4, Ho, Wo);
    return rp::finish_launch("rp_maxpool3x3s2_bwd");
}
extern "C" int rp_normalize_pose_bwd_f32(const float* dout, const float* raw, float* draw, int B, int device, void* stream) {
    RP_REQUIRE(dout && raw && draw && B > 0, RP_EINVAL, "rp_normalize_pose_bwd: bad argument");
    RP_GUARD(device);
    normalize_pose_bwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dout, raw, draw, B);
    return rp::finish_launch("rp_normalize_pose_bwd");
}
extern "C" int rp_concat_vpos_f32(const float* qkv, const float* pos, float* vp, int n_img, int device, void* stream) {
    RP_REQUIRE(qkv && pos && vp && n_img > 0 && (n_img % 2) == 0, RP_EINVAL, "rp_concat_vpos: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(concat_vpos_kernel, (long long)n_img * 3 * 576 * 70, qkv, pos, vp, n_img);
    return rp::finish_launch("rp_concat_vpos");
}
extern "C" int rp_scatter_dv_f32(const float* dvp, float* dqkv, int n_img, int device, void* stream) {
    RP_REQUIRE(dvp && dqkv && n_img > 0, RP_EINVAL, "rp_scatter_dv: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(scatter_dv_kernel, (long long)n_img * 3 * 576 * 64, dvp, dqkv, n_img);
    return rp::finish_launch("rp_scatter_dv");
}
