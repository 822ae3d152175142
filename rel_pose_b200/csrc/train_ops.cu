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
constexpr int CS_ROWS = 256;
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
// dx per row; per-block partial dgamma / dbeta (block = 8 rows) -> partial [nblk][2][cols]
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                     const float* __restrict__ g, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, float* __restrict__ dx,
                                                     float* __restrict__ partial, int rows, int cols) {
    __shared__ float sg[8][256], sb[8][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + w;
    float dyv[8], xh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dyv[i] = 0.f; xh[i] = 0.f; }
    if (row < rows) {
        const float m = mean[row], rs = rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = lane + 32 * i;
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
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        if (c < 256) { sg[w][c] = dyv[i] * xh[i]; sb[w][c] = dyv[i]; }
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
// x [M][C] (NHWC flattened).  Stage 1: per-block partial sum / sum of squares; stage 2: mean, biased var.
constexpr int BN_ROWS = 512;
__global__ void bn_stats_partial_kernel(const float* __restrict__ x, float* __restrict__ partial, long long M, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const long long r0 = (long long)blockIdx.y * BN_ROWS, r1 = min(M, r0 + BN_ROWS);
    float s = 0.f, q = 0.f;
    for (long long r = r0; r < r1; ++r) {
        const float v = x[r * C + c];
        s += v;
        q += v * v;
    }
    partial[((long long)blockIdx.y * 2 + 0) * C + c] = s;
    partial[((long long)blockIdx.y * 2 + 1) * C + c] = q;
}
// mean / biased var in double (the one-pass sum of squares is safe there); also the running-stat update
__global__ void bn_stats_final_kernel(const float* __restrict__ partial, int nblk, long long M, int C, float* __restrict__ mean,
                                      float* __restrict__ var, float* __restrict__ running_mean, float* __restrict__ running_var,
                                      float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int b = 0; b < nblk; ++b) {
        s += (double)partial[((long long)b * 2 + 0) * C + c];
        q += (double)partial[((long long)b * 2 + 1) * C + c];
    }
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
// y = act(gamma * (x - mean) * rstd + beta + residual)
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ var,
                                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ residual,
                                float* __restrict__ y, long long M, int C, float eps, int relu) {
    const long long n = M * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float v = (x[i] - mean[c]) * (1.0f / sqrtf(var[c] + eps)) * gamma[c] + beta[c];
        if (residual) v += residual[i];
        y[i] = relu ? fmaxf(v, 0.f) : v;
    }
}
// partial [nblk][2][C]: sum(dy), sum(dy * xhat)
__global__ void bn_bwd_partial_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                      const float* __restrict__ var, float eps, float* __restrict__ partial, long long M, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const long long r0 = (long long)blockIdx.y * BN_ROWS, r1 = min(M, r0 + BN_ROWS);
    const float m = mean[c], rs = 1.0f / sqrtf(var[c] + eps);
    float s = 0.f, q = 0.f;
    for (long long r = r0; r < r1; ++r) {
        const float d = dy[r * C + c];
        s += d;
        q += d * (x[r * C + c] - m) * rs;
    }
    partial[((long long)blockIdx.y * 2 + 0) * C + c] = s;
    partial[((long long)blockIdx.y * 2 + 1) * C + c] = q;
}
__global__ void bn_bwd_final_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0, q = 0.0;
    for (int b = 0; b < nblk; ++b) {
        s += (double)partial[((long long)b * 2 + 0) * C + c];
        q += (double)partial[((long long)b * 2 + 1) * C + c];
    }
    dbeta[c] = (float)s;
    dgamma[c] = (float)q;
}
// dx = gamma * rstd * (dy - dbeta / M - xhat * dgamma / M)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ dgamma,
                                    const float* __restrict__ dbeta, float eps, float* __restrict__ dx, long long M, int C) {
    const long long n = M * C;
    const float invM = 1.0f / (float)M;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float rs = 1.0f / sqrtf(var[c] + eps);
        const float xh = (x[i] - mean[c]) * rs;
        dx[i] = gamma[c] * rs * (dy[i] - dbeta[c] * invM - xh * dgamma[c] * invM);
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
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, int n, int H,
                                   int W, int C, int Ho, int Wo) {
    const long long total = (long long)n * H * W * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const int ix = (int)(pix % W), iy = (int)((pix / W) % H), im = (int)(pix / ((long long)W * H));
        const float xv = x[i];
        float s = 0.f;
        for (int oy = (iy + 1 - 2 + 1) / 2; oy <= (iy + 1) / 2; ++oy) {          // windows rows with 2*oy-1 <= iy <= 2*oy+1
            if (oy < 0 || oy >= Ho) continue;
            for (int ox = (ix + 1 - 2 + 1) / 2; ox <= (ix + 1) / 2; ++ox) {
                if (ox < 0 || ox >= Wo) continue;
                // is (iy,ix) the first maximum of window (oy,ox)?
                bool first = true;
                float mx = -INFINITY;
                int ay = -1, ax = -1;
                for (int ky = 0; ky < 3; ++ky) {
                    const int yy = 2 * oy - 1 + ky;
                    if (yy < 0 || yy >= H) continue;
                    for (int kx = 0; kx < 3; ++kx) {
                        const int xx = 2 * ox - 1 + kx;
                        if (xx < 0 || xx >= W) continue;
                        const float v = x[(((long long)im * H + yy) * W + xx) * C + c];
                        if (v > mx) { mx = v; ay = yy; ax = xx; }
                    }
                }
                (void)first;
                if (ay == iy && ax == ix && xv == mx) s += dy[(((long long)im * Ho + oy) * Wo + ox) * C + c];
            }
        }
        dx[i] = s;
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
    dim3 grid((cols + 127) / 128, nblk);
    colsum_partial_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(A, B, static_cast<float*>(workspace), rows, cols, 0);
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
    const int nblk = (rows + 7) / 8;
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
    const int nblk = (int)((M + BN_ROWS - 1) / BN_ROWS);
    bn_stats_partial_kernel<<<dim3((C + 63) / 64, nblk), 64, 0, (cudaStream_t)stream>>>(x, static_cast<float*>(workspace), M, C);
    bn_stats_final_kernel<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>(static_cast<float*>(workspace), nblk, M, C, mean, var,
                                                                        running_mean, running_var, momentum);
    return rp::finish_launch("rp_bn_train_stats");
}
extern "C" int rp_bn_apply_f32(const float* x, const float* mean, const float* var, const float* gamma, const float* beta,
                               const float* residual, float* y, int64_t M, int C, float eps, int relu, int device, void* stream) {
    RP_REQUIRE(x && mean && var && gamma && beta && y && M > 0 && C > 0, RP_EINVAL, "rp_bn_apply: bad argument");
    RP_GUARD(device);
    RP_LAUNCH1D(bn_apply_kernel, M * C, x, mean, var, gamma, beta, residual, y, M, C, eps, relu);
    return rp::finish_launch("rp_bn_apply");
}
extern "C" int rp_bn_bwd_f32(const float* dy, const float* x, const float* mean, const float* var, const float* gamma, float eps,
                             float* dx, float* dgamma, float* dbeta, int64_t M, int C, void* workspace, size_t workspace_bytes,
                             int device, void* stream) {
    RP_REQUIRE(dy && x && mean && var && gamma && dx && dgamma && dbeta && M > 0 && C > 0, RP_EINVAL, "rp_bn_bwd: bad argument");
    RP_REQUIRE(workspace && workspace_bytes >= rp_bn_workspace_bytes(M, C), RP_EWORKSPACE, "rp_bn_bwd: workspace too small");
    RP_GUARD(device);
    const int nblk = (int)((M + BN_ROWS - 1) / BN_ROWS);
    bn_bwd_partial_kernel<<<dim3((C + 63) / 64, nblk), 64, 0, (cudaStream_t)stream>>>(dy, x, mean, var, eps, static_cast<float*>(workspace), M, C);
    bn_bwd_final_kernel<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>(static_cast<float*>(workspace), nblk, C, dgamma, dbeta);
    RP_LAUNCH1D(bn_bwd_apply_kernel, M * C, dy, x, mean, var, gamma, dgamma, dbeta, eps, dx, M, C);
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
extern "C" int rp_col2im_nhwc_f32(const float* dcols, float* dx, int n, int H, int W, int C, int KH, int KW, int stride, int pad,
                                  int device, void* stream) {
    RP_REQUIRE(dcols && dx && n > 0 && H > 0 && W > 0 && C > 0 && KH > 0 && KW > 0 && stride > 0 && pad >= 0, RP_EINVAL, "rp_col2im: bad argument");
    RP_GUARD(device);
    const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
    RP_LAUNCH1D(col2im_kernel, (long long)n * H * W * C, dcols, dx, n, H, W, C, KH, KW, stride, pad, Ho, Wo);
    return rp::finish_launch("rp_col2im");
}
extern "C" int rp_maxpool3x3s2_bwd_f32(const float* dy, const float* x, float* dx, int n, int H, int W, int C, int device, void* stream) {
    RP_REQUIRE(dy && x && dx && n > 0 && H > 0 && W > 0 && C > 0, RP_EINVAL, "rp_maxpool3x3s2_bwd: bad argument");
    RP_GUARD(device);
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    RP_LAUNCH1D(maxpool_bwd_kernel, (long long)n * H * W * C, dy, x, dx, n, H, W, C, Ho, Wo);
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
