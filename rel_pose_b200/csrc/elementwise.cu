// Elementwise / gather / normalisation kernels of the rel_pose hot path (HBM-bound work).
// Reference rows: A1 (src/model.py:114-125), A4 (:136-141,172), LayerNorm (vision_transformer.py:396),
// A6 (vision_transformer.py:90-158), A10 (src/model.py:145-152).
#include "common.cuh"
#include <cooperative_groups.h>

#include <stdlib.h>
#include <string.h>

namespace rp {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("RELPOSE_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
}  // namespace rp

extern "C" const char* rp_last_error(void) { return rp::g_err; }
extern "C" int rp_version(void) { return 100; }
extern "C" int rp_device_arch(int device) {
    int major = 0, minor = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (e != cudaSuccess) {
        rp::set_error("rp_device_arch: %s", cudaGetErrorString(e));
        return -(int)e;
    }
    return major * 10 + minor;
}

// ------------------------------------------------------------------------------------------ input pipeline (8 f-2)
// Host -> device copy of only the image ROWS the nearest resize of A1 (model.py:125) will read: legacy `nearest`
// picks source row floor(d * H / out_rows) for output row d, i.e. b = out_rows / g rows out of every a = H / g
// (g = gcd), the same pattern in every period and every plane.  One pitched 3-D DMA per residue j < b moves
// row (k a + floor(j a / b)) of the host tensor to row (k b + j) of a compact [planes][out_rows][W] device tensor:
// 224 of 384 rows = 58 % of the PCIe bytes, nothing touched by the CPU.  The compact tensor is then an ordinary
// input of the preprocessing kernels (its row map is the identity, its column map unchanged) -> identical pixels.
extern "C" int rp_copy_rows_h2d(void* dst, const void* src_pinned, int64_t planes, int H, int64_t row_bytes,
                                int out_rows, int device, void* stream) {
    RP_REQUIRE(dst && src_pinned && planes > 0 && H > 0 && row_bytes > 0 && out_rows > 0 && out_rows <= H, RP_EINVAL,
               "rp_copy_rows_h2d: bad argument (out_rows must not exceed H)");
    int g = H, t = out_rows;
    while (t) { const int r = g % t; g = t; t = r; }
    const int a = H / g, b = out_rows / g;
    RP_REQUIRE(b <= 64, RP_EINVAL, "rp_copy_rows_h2d: %d -> %d rows needs %d strided copies (limit 64); copy the whole tensor",
               H, out_rows, b);
    // The row map is evaluated exactly as the preprocessing kernels (and ATen's legacy nearest) do, in float32:
    // floorf(d * (float)H / out_rows).  Rounding of the scale can move a row down by one from some period on
    // (480 -> 224: output row 119 reads source row 254, not 255), so every residue j is copied in runs of periods
    // with a constant offset -- one pitched 3-D DMA (row, period, plane) per run.
    const float scale = (float)H / (float)out_rows;
    auto src_row = [&](int d) {
        int sr = (int)floorf((float)d * scale);
        return sr > H - 1 ? H - 1 : sr;
    };
    RP_GUARD(device);
    int ncopies = 0;
    for (int j = 0; j < b; ++j) {
        int k = 0;
        while (k < g) {
            const int off = src_row(k * b + j) - k * a;
            // off = -1 happens when float rounding pushes a row into the previous period: (period k, offset -1) is
            // (period k - 1, offset a - 1)
            RP_REQUIRE(off >= -a && off < a && (off >= 0 || k > 0), RP_EINVAL, "rp_copy_rows_h2d: unexpected row map (%d -> %d rows)",
                       H, out_rows);
            const int ky = off < 0 ? k - 1 : k, ox = off < 0 ? off + a : off;
            int k1 = k + 1;
            while (k1 < g && src_row(k1 * b + j) - k1 * a == off) ++k1;
            cudaMemcpy3DParms p;
            memset(&p, 0, sizeof(p));
            p.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src_pinned), (size_t)a * row_bytes, (size_t)a * row_bytes, (size_t)g);
            p.srcPos = make_cudaPos((size_t)ox * row_bytes, (size_t)ky, 0);
            p.dstPtr = make_cudaPitchedPtr(dst, (size_t)b * row_bytes, (size_t)b * row_bytes, (size_t)g);
            p.dstPos = make_cudaPos((size_t)j * row_bytes, (size_t)k, 0);
            p.extent = make_cudaExtent((size_t)row_bytes, (size_t)(k1 - k), (size_t)planes);
            p.kind = cudaMemcpyHostToDevice;
            cudaError_t e = cudaMemcpy3DAsync(&p, (cudaStream_t)stream);
            if (e != cudaSuccess) {
                rp::set_error("rp_copy_rows_h2d: cudaMemcpy3DAsync: %s", cudaGetErrorString(e));
                return (int)e;
            }
            RP_REQUIRE(++ncopies <= 256, RP_EINVAL, "rp_copy_rows_h2d: row map needs too many copies; copy the whole tensor");
            k = k1;
        }
    }
    return RP_OK;
}

// ------------------------------------------------------------------------------------------ A1
// One thread per output pixel (all three channels): out[n,c,oy,ox] = ((img[n,2-c,iy,ix]/255)-mean[c])/std[c].
// Writes are fully coalesced; reads are a strided gather inside one input row (L1/L2 absorb it).
template <typename T>
__global__ void __launch_bounds__(256) preprocess_kernel(const T* __restrict__ img, float* __restrict__ out,
                                                          int n_img, int H, int W, float scale_h, float scale_w) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    const int OUT = 224;
    const float mean[3] = {0.485f, 0.456f, 0.406f};
    const float stdv[3] = {0.229f, 0.224f, 0.225f};
    long long total = (long long)n_img * OUT * OUT;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int ox = (int)(idx % OUT);
        int oy = (int)((idx / OUT) % OUT);
        int n = (int)(idx / (OUT * OUT));
        int ix = min((int)floorf(__fmul_rn((float)ox, scale_w)), W - 1);
        int iy = min((int)floorf(__fmul_rn((float)oy, scale_h)), H - 1);
        const T* src = img + ((long long)n * 3) * H * W + (long long)iy * W + ix;
        float* dst = out + ((long long)n * 3) * OUT * OUT + (long long)oy * OUT + ox;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = (float)src[(long long)(2 - c) * H * W];       // BGR -> RGB
            v = __fdiv_rn(v, 255.0f);
            v = __fsub_rn(v, mean[c]);
            v = __fdiv_rn(v, stdv[c]);
            dst[(long long)c * OUT * OUT] = v;
        }
    }
}

template <typename T>
static int preprocess_launch(const T* images, float* out, int n_img, int H, int W, int device, void* stream) {
    RP_REQUIRE(images && out, RP_EINVAL, "rp_preprocess: null pointer");
    RP_REQUIRE(n_img > 0 && H > 0 && W > 0, RP_EINVAL, "rp_preprocess: bad shape n_img=%d H=%d W=%d", n_img, H, W);
    RP_GUARD(device);
    long long total = (long long)n_img * 224 * 224;
    int blocks = (int)min((total + 255) / 256, (long long)rp::num_sms(device) * 16);
    float sh = (float)H / (float)224, sw = (float)W / (float)224;   // ATen: scale = (float)in / out
    rp::launch(preprocess_kernel<T>, dim3(blocks), dim3(256), (size_t)(0), (cudaStream_t)stream, images, out, n_img, H, W, sh, sw);
    return rp::finish_launch("rp_preprocess");
}

extern "C" int rp_preprocess_f32(const float* images, float* out, int n_img, int H, int W, int device, void* stream) {
    return preprocess_launch<float>(images, out, n_img, H, W, device, stream);
}
extern "C" int rp_preprocess_u8(const uint8_t* images, float* out, int n_img, int H, int W, int device, void* stream) {
    return preprocess_launch<uint8_t>(images, out, n_img, H, W, device, stream);
}

// ------------------------------------------------------------------- intrinsics (model.py:100-109)
__global__ void intrinsics_prepare_kernel(float* __restrict__ intr, float* __restrict__ kxy, int* __restrict__ flags,
                                          int B, float sx, float sy) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float k[2][4];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        float4 q = reinterpret_cast<float4*>(intr)[b * 2 + v];
        k[v][0] = __fmul_rn(sx, q.x);   // fx
        k[v][1] = __fmul_rn(sy, q.y);   // fy
        k[v][2] = __fmul_rn(sx, q.z);   // cx
        k[v][3] = __fmul_rn(sy, q.w);   // cy
        reinterpret_cast<float4*>(intr)[b * 2 + v] = make_float4(k[v][0], k[v][1], k[v][2], k[v][3]);
    }
    int f = 0;
    if (k[0][0] != k[1][0] || k[0][1] != k[1][1] || k[0][2] != k[1][2] || k[0][3] != k[1][3]) f |= 1;
    if (b == 0 && __fmul_rn(k[0][2], k[0][3]) == 0.0f) f |= 2;
    if (f) atomicOr(flags, f);
    // K = diag(fx/cx, fy/cy, 1) after the reference's pixel->[-1,1] normalisation; K^-1 diagonal:
    float fxn = __fmul_rn(__fdiv_rn(k[0][0], __fmul_rn(k[0][2], 2.0f)), 2.0f);
    float fyn = __fmul_rn(__fdiv_rn(k[0][1], __fmul_rn(k[0][3], 2.0f)), 2.0f);
    kxy[b * 2 + 0] = __fdiv_rn(1.0f, fxn);
    kxy[b * 2 + 1] = __fdiv_rn(1.0f, fyn);
}

extern "C" int rp_intrinsics_prepare_f32(float* intrinsics, float* kxy, int* flags, int B, int H, int W, int device,
                                         void* stream) {
    RP_REQUIRE(intrinsics && kxy && flags, RP_EINVAL, "rp_intrinsics_prepare: null pointer");
    RP_REQUIRE(B > 0 && H > 0 && W > 0, RP_EINVAL, "rp_intrinsics_prepare: bad shape");
    RP_REQUIRE(rp::aligned16(intrinsics), RP_EALIGN, "rp_intrinsics_prepare: intrinsics not 16-byte aligned");
    RP_GUARD(device);
    float sx = (float)(24.0 / (double)W), sy = (float)(24.0 / (double)H);
    rp::launch(intrinsics_prepare_kernel, dim3((B + 127) / 128), dim3(128), (size_t)(0), (cudaStream_t)stream, intrinsics, kxy, flags, B, sx, sy);
    return rp::finish_launch("rp_intrinsics_prepare");
}

// ------------------------------------------------------------------------------------------ A4
// [n,192,576] -> [n,576,192] (+pos_embed): 32x32 shared-memory tile transpose, both sides coalesced.
__global__ void __launch_bounds__(256) tokens_posembed_kernel(const float* __restrict__ fmap,
                                                              const float* __restrict__ pos, float* __restrict__ x) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    __shared__ float tile[32][33];
    const int C = RP_EMBED, N = RP_NTOK;
    int n = blockIdx.z;
    int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const float* src = fmap + (long long)n * C * N;
#pragma unroll
    for (int r = ty; r < 32; r += 8) tile[r][tx] = src[(long long)(c0 + r) * N + t0 + tx];
    __syncthreads();
    float* dst = x + (long long)n * N * C;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        int t = t0 + r, c = c0 + tx;
        dst[(long long)t * C + c] = tile[tx][r] + pos[t * C + c];
    }
}

extern "C" int rp_tokens_posembed_f32(const float* fmap, const float* pos_embed, float* x, int n_img, int device,
                                      void* stream) {
    RP_REQUIRE(fmap && pos_embed && x && n_img > 0, RP_EINVAL, "rp_tokens_posembed: bad argument");
    RP_GUARD(device);
    dim3 grid(RP_NTOK / 32, RP_EMBED / 32, n_img);
    rp::launch(tokens_posembed_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, fmap, pos_embed, x);
    return rp::finish_launch("rp_tokens_posembed");
}

// ----------------------------------------------------------------------------------- LayerNorm
// One warp per row; the row lives in registers (cols <= 32*PER); two-pass mean / variance.
template <int PER>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                        const float* __restrict__ b, float* __restrict__ y, int rows,
                                                        int cols, float eps) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float* xr = x + (long long)warp * cols;
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        int c = lane + 32 * i;
        v[i] = (c < cols) ? xr[c] : 0.f;
        s += v[i];
    }
    float mean = rp::warp_sum(s) / (float)cols;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        int c = lane + 32 * i;
        float d = (c < cols) ? v[i] - mean : 0.f;
        q += d * d;
    }
    float rstd = 1.0f / sqrtf(rp::warp_sum(q) / (float)cols + eps);   // IEEE sqrt + div (no fast-math)
    float* yr = y + (long long)warp * cols;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        int c = lane + 32 * i;
        if (c < cols) yr[c] = (v[i] - mean) * rstd * g[c] + b[c];
    }
}

extern "C" int rp_layernorm_f32(const float* x, const float* gamma, const float* beta, float* y, int rows, int cols,
                                float eps, int device, void* stream) {
    RP_REQUIRE(x && gamma && beta && y, RP_EINVAL, "rp_layernorm: null pointer");
    RP_REQUIRE(rows > 0 && cols > 0 && cols <= 1024, RP_EINVAL, "rp_layernorm: bad shape rows=%d cols=%d", rows, cols);
    RP_GUARD(device);
    int blocks = (rows + 7) / 8;
    cudaStream_t st = (cudaStream_t)stream;
    if (cols <= 256)
        rp::launch(layernorm_kernel<8>, dim3(blocks), dim3(256), (size_t)(0), st, x, gamma, beta, y, rows, cols, eps);
    else
        rp::launch(layernorm_kernel<32>, dim3(blocks), dim3(256), (size_t)(0), st, x, gamma, beta, y, rows, cols, eps);
    return rp::finish_launch("rp_layernorm");
}

// ------------------------------------------------------------------------------------------ A6
struct Lin24 {
    float v[RP_GRID];
};

__global__ void posenc_kernel(const float* __restrict__ kxy, Lin24 lin, float* __restrict__ pos, int B, int l1) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * RP_NTOK) return;
    int b = idx / RP_NTOK, i = idx % RP_NTOK;
    float p3 = lin.v[i % RP_GRID];   // "y" runs with the FAST token index (transposed grid)
    float p4 = lin.v[i / RP_GRID];
    if (kxy) {
        p4 = __fmul_rn(kxy[b * 2 + 0], p4);
        p3 = __fmul_rn(kxy[b * 2 + 1], p3);
    }
    float* o = pos + (long long)idx * RP_NPOS;
    // --l1_pos_encoding (get_l1_positional_encodings, vision_transformer.py:36-87): the quadratic channels stay 1
    o[0] = l1 ? 1.0f : __fmul_rn(p3, p3);
    o[1] = l1 ? 1.0f : __fmul_rn(p4, p4);
    o[2] = l1 ? 1.0f : __fmul_rn(p3, p4);
    o[3] = p3;
    o[4] = p4;
    o[5] = 1.0f;
}

extern "C" int rp_posenc_f32(const float* kxy, const float* host_lin24, float* pos, int B, int device, void* stream) {
    return rp_posenc_ex_f32(kxy, host_lin24, pos, B, 0, device, stream);
}

extern "C" int rp_posenc_ex_f32(const float* kxy, const float* host_lin24, float* pos, int B, int l1, int device,
                                void* stream) {
    RP_REQUIRE(host_lin24 && pos && B > 0, RP_EINVAL, "rp_posenc: bad argument");
    RP_GUARD(device);
    Lin24 lin;
    memcpy(lin.v, host_lin24, sizeof(lin.v));
    int total = B * RP_NTOK;
    rp::launch(posenc_kernel, dim3((total + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, kxy, lin, pos, B, l1);
    return rp::finish_launch("rp_posenc");
}

// ----------------------------------------------------------------------------------------- A9 tail
// pose_regressor[2:5] (src/model.py:93-97): out = W2 relu(W1 h + b1) + b2 for h [B,512] -> [B,14], one launch.
// Two 64-row GEMMs of 17 MFLOP took 27 us each on the tiled SIMT engine (a 128 x 96 tile per 64 x 14 output, plus
// its split-K reduction launch); here a CTA takes RT_ROWS rows, thread j owns hidden unit j and walks the TRANSPOSED
// layer-1 weight (coalesced across the CTA), the 14 outputs are warp reductions.
constexpr int RT_ROWS = 1, RT_H = 512, RT_OUT = 14;     // one row per CTA: 64 CTAs at 64 pairs; the weight (1 MB) stays in L2

__global__ void __launch_bounds__(RT_H) regressor_tail_kernel(const float* __restrict__ h, const float* __restrict__ W1T,
                                                              const float* __restrict__ b1, const float* __restrict__ W2,
                                                              const float* __restrict__ b2, float* __restrict__ out, int B) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    __shared__ __align__(16) float hs[RT_ROWS][RT_H];
    __shared__ __align__(16) float h2[RT_ROWS][RT_H];
    const int j = threadIdx.x, row0 = blockIdx.x * RT_ROWS;
#pragma unroll
    for (int r = 0; r < RT_ROWS; ++r) hs[r][j] = row0 + r < B ? h[(size_t)(row0 + r) * RT_H + j] : 0.f;
    __syncthreads();
    float acc[RT_ROWS];
#pragma unroll
    for (int r = 0; r < RT_ROWS; ++r) acc[r] = 0.f;
    // 32 independent weight loads in flight per thread: the loop is a chain of L2 round trips otherwise
#pragma unroll 1
    for (int k = 0; k < RT_H; k += 32) {
        float w[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) w[u] = __ldg(W1T + (size_t)(k + u) * RT_H + j);
#pragma unroll
        for (int r = 0; r < RT_ROWS; ++r) {
#pragma unroll
            for (int u = 0; u < 32; u += 4) {
                const float4 x = *reinterpret_cast<const float4*>(&hs[r][k + u]);
                acc[r] = fmaf(x.x, w[u], acc[r]); acc[r] = fmaf(x.y, w[u + 1], acc[r]);
                acc[r] = fmaf(x.z, w[u + 2], acc[r]); acc[r] = fmaf(x.w, w[u + 3], acc[r]);
            }
        }
    }
    const float bj = b1[j];
#pragma unroll
    for (int r = 0; r < RT_ROWS; ++r) h2[r][j] = fmaxf(acc[r] + bj, 0.f);
    __syncthreads();
    const int warp = j >> 5, lane = j & 31;
    for (int item = warp; item < RT_ROWS * RT_OUT; item += RT_H / 32) {
        const int r = item / RT_OUT, n = item - r * RT_OUT;
        float s = 0.f;
        for (int k = lane; k < RT_H; k += 32) s = fmaf(h2[r][k], __ldg(W2 + (size_t)n * RT_H + k), s);
        s = rp::warp_sum(s);
        if (lane == 0 && row0 + r < B) out[(size_t)(row0 + r) * RT_OUT + n] = s + b2[n];
    }
}

// Second version (default; RELPOSE_REGRESSOR_TAIL_V1=1 selects the kernel above for A/B runs).  The first version gave every
// row its own CTA, so every CTA pulled the whole 1 MB layer-1 weight through its SM's L2 port (~26 k cycles = 13 us, the
// kernel's 32 us together with the serial layer-2 dots).  Here a thread-block CLUSTER of 8 CTAs takes 8 rows: CTA `rank` owns
// hidden units [64 rank, 64 rank + 64) for all 8 rows (128 KB of weights + 16 KB of h per CTA), four k-slices per CTA summed in
// a fixed order, and hands row r of its [8 x 64] ReLU block to CTA r through distributed shared memory; after one cluster
// barrier CTA r holds all 512 hidden values of row r and forms its 14 outputs (one warp per output, loads batched).
constexpr int RT2_ROWS = 8, RT2_J = RT_H / RT2_ROWS /* 64 hidden units per CTA */, RT2_KS = 4, RT2_THREADS = RT2_J * RT2_KS;
constexpr int RT2_SMEM = (RT_H * RT2_J + RT_H * RT2_ROWS + RT2_KS * RT2_ROWS * RT2_J + RT_H) * 4;   // Ws, hs, part, h2row
static_assert(RT2_ROWS == 8 && RT2_THREADS == 256 && RT_H % RT2_KS == 0 && (RT_H * RT2_J / 4) % RT2_THREADS == 0, "regressor_tail2 tiling");

__global__ void __cluster_dims__(RT2_ROWS, 1, 1) __launch_bounds__(RT2_THREADS)
regressor_tail2_kernel(const float* __restrict__ h, const float* __restrict__ W1T, const float* __restrict__ b1,
                       const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ out, int B) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) float rt_smem[];
    float (*Ws)[RT2_J] = reinterpret_cast<float (*)[RT2_J]>(rt_smem);                       // W1T[k][64 rank + jj], 128 KB
    float (*hs)[RT2_ROWS] = reinterpret_cast<float (*)[RT2_ROWS]>(rt_smem + RT_H * RT2_J);  // h of the 8 rows, k-major
    float (*part)[RT2_ROWS][RT2_J] = reinterpret_cast<float (*)[RT2_ROWS][RT2_J]>(rt_smem + RT_H * RT2_J + RT_H * RT2_ROWS);
    float* h2row = rt_smem + RT_H * RT2_J + RT_H * RT2_ROWS + RT2_KS * RT2_ROWS * RT2_J;   // row `rank` of relu(W1 h + b1)
    const int tid = threadIdx.x;
    const int rank = (int)cluster.block_rank();                      // = blockIdx.x % 8
    const int row0 = (blockIdx.x / RT2_ROWS) * RT2_ROWS;
    rp::pdl_launch_dependents();
    cluster.barrier_arrive();                                        // phase 1: "this CTA is running" (awaited before the first remote write)
    rp::pdl_wait();
    // the CTA's slice of the layer-1 weight: 8192 16-byte asynchronous copies, all in flight together under the loads of h
#pragma unroll
    for (int i = 0; i < RT_H * RT2_J / 4 / RT2_THREADS; ++i) {
        const int c = tid + i * RT2_THREADS, k = c / (RT2_J / 4), q = c % (RT2_J / 4);
        rp::cp_async16(&Ws[k][4 * q], W1T + (size_t)k * RT_H + rank * RT2_J + 4 * q);
    }
    rp::cp_async_commit();
    {
        float v[RT2_ROWS * RT_H / RT2_THREADS];                      // 16 loads in flight
#pragma unroll
        for (int i = 0; i < RT2_ROWS * RT_H / RT2_THREADS; ++i) {
            const int e = tid + i * RT2_THREADS, r = e / RT_H;
            v[i] = (row0 + r < B) ? h[(size_t)(row0 + r) * RT_H + (e % RT_H)] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < RT2_ROWS * RT_H / RT2_THREADS; ++i) {
            const int e = tid + i * RT2_THREADS;
            hs[e % RT_H][e / RT_H] = v[i];
        }
    }
    rp::cp_async_wait<0>();
    __syncthreads();
    const int jj = tid % RT2_J, ks = tid / RT2_J;
    constexpr int KSL = RT_H / RT2_KS;                               // 128 k per slice
    float acc[RT2_ROWS];
#pragma unroll
    for (int r = 0; r < RT2_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 8
    for (int k = ks * KSL; k < ks * KSL + KSL; ++k) {
        const float w = Ws[k][jj];
        const float4 x0 = *reinterpret_cast<const float4*>(&hs[k][0]);
        const float4 x1 = *reinterpret_cast<const float4*>(&hs[k][4]);
        acc[0] = fmaf(x0.x, w, acc[0]); acc[1] = fmaf(x0.y, w, acc[1]);
        acc[2] = fmaf(x0.z, w, acc[2]); acc[3] = fmaf(x0.w, w, acc[3]);
        acc[4] = fmaf(x1.x, w, acc[4]); acc[5] = fmaf(x1.y, w, acc[5]);
        acc[6] = fmaf(x1.z, w, acc[6]); acc[7] = fmaf(x1.w, w, acc[7]);
    }
#pragma unroll
    for (int r = 0; r < RT2_ROWS; ++r) part[ks][r][jj] = acc[r];
    __syncthreads();
    // [8 rows x 64 units] of this CTA: slices summed in order, bias, ReLU; row r goes to CTA r of the cluster.  Distributed shared
    // memory may only be written once the owning CTA has started: every CTA arrived at phase 1 as its first action.
    cluster.barrier_wait();
    for (int e = tid; e < RT2_ROWS * RT2_J; e += RT2_THREADS) {
        const int r = e / RT2_J, c = e % RT2_J;
        float v = part[0][r][c];
#pragma unroll
        for (int s2 = 1; s2 < RT2_KS; ++s2) v += part[s2][r][c];
        v = fmaxf(v + b1[rank * RT2_J + c], 0.f);
        float* remote = cluster.map_shared_rank(h2row, r);
        remote[rank * RT2_J + c] = v;
    }
    cluster.sync();                                                  // release / acquire: every CTA's h2row is complete and visible
    const int row = row0 + rank;
    if (row < B) {
        const int warp = tid >> 5, lane = tid & 31;
        for (int n = warp; n < RT_OUT; n += RT2_THREADS / 32) {
            float wv[RT_H / 32];
#pragma unroll
            for (int i = 0; i < RT_H / 32; ++i) wv[i] = __ldg(W2 + (size_t)n * RT_H + lane + 32 * i);
            float s2 = 0.f;
#pragma unroll
            for (int i = 0; i < RT_H / 32; ++i) s2 = fmaf(h2row[lane + 32 * i], wv[i], s2);
            s2 = rp::warp_sum(s2);
            if (lane == 0) out[(size_t)row * RT_OUT + n] = s2 + b2[n];
        }
    }
}

extern "C" int rp_regressor_tail_f32(const float* h, const float* W1T, const float* b1, const float* W2, const float* b2,
                                     float* out, int B, int hidden, int n_out, int device, void* stream) {
    RP_REQUIRE(h && W1T && b1 && W2 && b2 && out && B > 0, RP_EINVAL, "rp_regressor_tail: bad argument");
    RP_REQUIRE(hidden == RT_H && n_out == RT_OUT, RP_EINVAL, "rp_regressor_tail: built for 512 hidden units and 14 outputs (got %d, %d)",
               hidden, n_out);
    RP_GUARD(device);
    const char* env = getenv("RELPOSE_REGRESSOR_TAIL_V1");       // read per call: the GPU test toggles it in-process
    if ((env && env[0] == '1') || !rp::aligned16(W1T))            // v2 stages the weight with 16-byte asynchronous copies
        rp::launch(regressor_tail_kernel, dim3((B + RT_ROWS - 1) / RT_ROWS), dim3(RT_H), (size_t)(0), (cudaStream_t)stream, h, W1T, b1, W2, b2, out, B);
    else {
        static bool attr_set[64] = {false};
        if (device >= 0 && device < 64 && !attr_set[device]) {
            cudaError_t e = cudaFuncSetAttribute(regressor_tail2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RT2_SMEM);
            if (e != cudaSuccess) {
                rp::set_error("rp_regressor_tail: cudaFuncSetAttribute(%d): %s", RT2_SMEM, cudaGetErrorString(e));
                return (int)e;
            }
            attr_set[device] = true;
        }
        rp::launch(regressor_tail2_kernel, dim3(RT2_ROWS * ((B + RT2_ROWS - 1) / RT2_ROWS)), dim3(RT2_THREADS), (size_t)(RT2_SMEM), (cudaStream_t)stream, h, W1T,
                   b1, W2, b2, out, B);
    }
    return rp::finish_launch("rp_regressor_tail");
}

// ----------------------------------------------------------------------------------------- A10
__global__ void normalize_pose_kernel(const float* __restrict__ raw, const float* __restrict__ Gs,
                                      float* __restrict__ out, int B) {
    rp::pdl_launch_dependents();
    rp::pdl_wait();
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* g = Gs + (long long)b * 14;
    const float* r = raw + (long long)b * 14 + 7;
    float* o = out + (long long)b * 14;
#pragma unroll
    for (int i = 0; i < 7; ++i) o[i] = g[i];
    float qx = r[3], qy = r[4], qz = r[5], qw = r[6];
    float n = sqrtf(qx * qx + qy * qy + qz * qz + qw * qw);
    float d = fmaxf(n, 0.01f);
    o[7] = r[0];
    o[8] = r[1];
    o[9] = r[2];
    o[10] = __fdiv_rn(qx, d);
    o[11] = __fdiv_rn(qy, d);
    o[12] = __fdiv_rn(qz, d);
    o[13] = __fdiv_rn(qw, d);
}

extern "C" int rp_normalize_pose_f32(const float* raw, const float* Gs, float* out, int B, int device, void* stream) {
    RP_REQUIRE(raw && Gs && out && B > 0, RP_EINVAL, "rp_normalize_pose: bad argument");
    RP_GUARD(device);
    rp::launch(normalize_pose_kernel, dim3((B + 127) / 128), dim3(128), (size_t)(0), (cudaStream_t)stream, raw, Gs, out, B);
    return rp::finish_launch("rp_normalize_pose");
}
