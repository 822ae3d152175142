// FP32 SIMT attention kernels: fused self-attention (A5) and the Essential Matrix Module core (A7).
// Nothing of size 576x576 is ever written to HBM (the reference materialises [2B,3,576,576] five times
// for self-attention and eight times in the EM module: vision_transformer.py:325-329,198-223).
//
// Common structure: a CTA of 128 threads owns a 64-row tile; it walks the 9 column tiles of 64,
// computing S = Q K^T (K = head_dim = 64) from shared memory with a 4x8 register tile per thread.
//   self-attention : online softmax (running max / sum), O += P V, one pass.
//   EM module      : dual softmax needs row AND column statistics of the same S, so
//                      pass 1 (em_stats_kernel)   row log-sum-exp of S and of S^T  (S^T = K Q^T),
//                      pass 2 (em_accum_kernel)   A = exp(2S - lse_r[i] - lse_c[j]);  T = A [v|pos];
//                                                 F_part = [v|pos]_i^T T   (70x70 per row tile),
//                      pass 3 (em_reduce_kernel)  fixed-order sum of the 9 row-tile partials.
// Thread map: tid -> tx = tid&7 (columns), ty = tid>>3 (rows); a warp holds 4 consecutive ty, so all
// eight threads that share a row are lanes of one warp (row reductions are shuffles, P needs only
// __syncwarp between its producer and consumer).
#include "common.cuh"

namespace {

constexpr int NTOK = RP_NTOK, HD = RP_HDIM, LDQKV = 3 * RP_EMBED;   // 576
constexpr int TILE = 64, NTILES = NTOK / TILE;                        // 9
constexpr int LQ = 68;   // row stride of Q/K tiles: 4 adjacent rows -> 4 disjoint 16 B bank groups
constexpr int LP = 72;   // row stride of P / [v|pos] tiles (70 padded to 72)
constexpr int EMW = RP_EMW;
constexpr int THREADS = 128;

// [64 x 64] fp32 tile, global row stride 576 -> shared row stride LD, via 16-byte cp.async
template <int LD>
__device__ __forceinline__ void load_tile64(float* dst, const float* src, int tid) {
#pragma unroll
    for (int c = tid; c < TILE * 16; c += THREADS) {
        int r = c >> 4, k = (c & 15) * 4;
        rp::cp_async16(dst + r * LD + k, src + (size_t)r * LDQKV + k);
    }
}

// s[ii][jj] = sum_d Q[ty+16ii][d] * K[tx+8jj][d]
__device__ __forceinline__ void qk_tile(const float* __restrict__ Qs, const float* __restrict__ Ks, int ty, int tx,
                                        float (&s)[4][8]) {
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) s[ii][jj] = 0.f;
#pragma unroll 4
    for (int kk = 0; kk < HD; kk += 4) {
        float4 a[4], b[8];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) a[ii] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * ii) * LQ + kk);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) b[jj] = *reinterpret_cast<const float4*>(Ks + (tx + 8 * jj) * LQ + kk);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                float v = s[ii][jj];
                v = fmaf(a[ii].x, b[jj].x, v);
                v = fmaf(a[ii].y, b[jj].y, v);
                v = fmaf(a[ii].z, b[jj].z, v);
                v = fmaf(a[ii].w, b[jj].w, v);
                s[ii][jj] = v;
            }
    }
}

__device__ __forceinline__ float row_max8(float v) {   // across the 8 tx lanes of a row
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
    return v;
}
__device__ __forceinline__ float row_sum8(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

// ============================================================================ self-attention (A5)
constexpr int SA_SMEM = (TILE * LQ * 2 + TILE * HD + TILE * LP) * (int)sizeof(float);

__global__ void __launch_bounds__(THREADS, 3)
self_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int kv_xor) {
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;
    float* Ks = Qs + TILE * LQ;
    float* Vs = Ks + TILE * LQ;      // [64][64]
    float* Ps = Vs + TILE * HD;      // [64][LP]
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int qt = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
    // kv_xor = 1: keys/values of the OTHER view of the pair (plain cross attention of the --noess ablation)
    const float* base = qkv + (size_t)(n ^ kv_xor) * NTOK * LDQKV;

    load_tile64<LQ>(Qs, qkv + ((size_t)n * NTOK + qt * TILE) * LDQKV + h * HD, tid);
    rp::cp_async_commit();

    float m[4], l[4], o[4][8];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        m[ii] = -INFINITY;
        l[ii] = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) o[ii][c] = 0.f;
    }

    for (int kt = 0; kt < NTILES; ++kt) {
        __syncthreads();   // everyone is done with the previous K/V/P tiles
        load_tile64<LQ>(Ks, base + (size_t)(kt * TILE) * LDQKV + RP_EMBED + h * HD, tid);
        load_tile64<HD>(Vs, base + (size_t)(kt * TILE) * LDQKV + 2 * RP_EMBED + h * HD, tid);
        rp::cp_async_commit();
        rp::cp_async_wait<0>();
        __syncthreads();

        float s[4][8];
        qk_tile(Qs, Ks, ty, tx, s);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                s[ii][jj] *= 0.125f;                       // head_dim^-0.5, after the matmul (:325)
                mx = fmaxf(mx, s[ii][jj]);
            }
            mx = row_max8(mx);
            float mn = fmaxf(m[ii], mx);
            float corr = expf(m[ii] - mn);                 // exp(-inf) = 0 on the first tile
            float rs = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                float p = expf(s[ii][jj] - mn);
                rs += p;
                Ps[(ty + 16 * ii) * LP + tx + 8 * jj] = p;
            }
            l[ii] = l[ii] * corr + row_sum8(rs);
            m[ii] = mn;
#pragma unroll
            for (int c = 0; c < 8; ++c) o[ii][c] *= corr;
        }
        __syncwarp();
        // O[i][d] += sum_j P[i][j] V[j][d];   thread owns d in {4tx..4tx+3} U {32+4tx..32+4tx+3}
#pragma unroll 2
        for (int j = 0; j < TILE; j += 4) {
            float4 a[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) a[ii] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * ii) * LP + j);
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {
                float4 b0 = *reinterpret_cast<const float4*>(Vs + (j + jq) * HD + 4 * tx);
                float4 b1 = *reinterpret_cast<const float4*>(Vs + (j + jq) * HD + 32 + 4 * tx);
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    float p = jq == 0 ? a[ii].x : jq == 1 ? a[ii].y : jq == 2 ? a[ii].z : a[ii].w;
                    o[ii][0] = fmaf(p, b0.x, o[ii][0]);
                    o[ii][1] = fmaf(p, b0.y, o[ii][1]);
                    o[ii][2] = fmaf(p, b0.z, o[ii][2]);
                    o[ii][3] = fmaf(p, b0.w, o[ii][3]);
                    o[ii][4] = fmaf(p, b1.x, o[ii][4]);
                    o[ii][5] = fmaf(p, b1.y, o[ii][5]);
                    o[ii][6] = fmaf(p, b1.z, o[ii][6]);
                    o[ii][7] = fmaf(p, b1.w, o[ii][7]);
                }
            }
        }
    }
    // out[n, row, h*64 + d]  ((attn @ v).transpose(1,2).reshape(B,N,C), :329)
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        float inv = 1.0f / l[ii];
        float* dst = out + ((size_t)n * NTOK + qt * TILE + ty + 16 * ii) * RP_EMBED + h * HD;
        *reinterpret_cast<float4*>(dst + 4 * tx) = make_float4(o[ii][0] * inv, o[ii][1] * inv, o[ii][2] * inv, o[ii][3] * inv);
        *reinterpret_cast<float4*>(dst + 32 + 4 * tx) = make_float4(o[ii][4] * inv, o[ii][5] * inv, o[ii][6] * inv, o[ii][7] * inv);
    }
}

// ============================================================================ EM module (A7)
// dir 0: S1 = q2 k1^T, V = [v1|pos]  (queries from view 2b+1, keys/values from view 2b)
// dir 1: S2 = q1 k2^T, V = [v2|pos]  (queries from view 2b,   keys/values from view 2b+1)
// lse layout: [B][2 dir][2 which (0 = rows of S, 1 = columns of S)][3 heads][576]
constexpr int ST_SMEM = (TILE * LQ * 2) * (int)sizeof(float);

__global__ void __launch_bounds__(THREADS, 4)
em_stats_kernel(const float* __restrict__ qkv, float* __restrict__ lse) {
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;
    float* Ks = Qs + TILE * LQ;
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int rt = blockIdx.x, h = blockIdx.y;
    const int which = blockIdx.z & 1, dir = (blockIdx.z >> 1) & 1, b = blockIdx.z >> 2;
    const int q_img = 2 * b + (1 - dir), kv_img = 2 * b + dir;
    // which == 0: rows = queries (q of q_img), cols = keys (k of kv_img);  which == 1: swapped
    const float* rsrc = qkv + (size_t)(which == 0 ? q_img : kv_img) * NTOK * LDQKV + (which == 0 ? 0 : RP_EMBED) + h * HD;
    const float* csrc = qkv + (size_t)(which == 0 ? kv_img : q_img) * NTOK * LDQKV + (which == 0 ? RP_EMBED : 0) + h * HD;

    load_tile64<LQ>(Qs, rsrc + (size_t)(rt * TILE) * LDQKV, tid);
    rp::cp_async_commit();
    float m[4], l[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        m[ii] = -INFINITY;
        l[ii] = 0.f;
    }
    for (int ct = 0; ct < NTILES; ++ct) {
        __syncthreads();
        load_tile64<LQ>(Ks, csrc + (size_t)(ct * TILE) * LDQKV, tid);
        rp::cp_async_commit();
        rp::cp_async_wait<0>();
        __syncthreads();
        float s[4][8];
        qk_tile(Qs, Ks, ty, tx, s);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            float mx = -INFINITY;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                s[ii][jj] *= 0.125f;
                mx = fmaxf(mx, s[ii][jj]);
            }
            mx = row_max8(mx);
            float mn = fmaxf(m[ii], mx);
            float rs = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) rs += expf(s[ii][jj] - mn);
            l[ii] = l[ii] * expf(m[ii] - mn) + row_sum8(rs);
            m[ii] = mn;
        }
    }
    if (tx == 0) {
        float* dst = lse + ((((size_t)b * 2 + dir) * 2 + which) * RP_HEADS + h) * NTOK + rt * TILE;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) dst[ty + 16 * ii] = m[ii] + logf(l[ii]);
    }
}

// [64 x 72] tile of [v | pos | 0 0]
__device__ __forceinline__ void load_vpos_tile(float* dst, const float* vsrc, const float* pos /* [576][6] or null */,
                                               int row0, int tid) {
#pragma unroll
    for (int c = tid; c < TILE * 16; c += THREADS) {
        int r = c >> 4, k = (c & 15) * 4;
        rp::cp_async16(dst + r * LP + k, vsrc + (size_t)(row0 + r) * LDQKV + k);
    }
    for (int c = tid; c < TILE * 8; c += THREADS) {
        int r = c >> 3, k = c & 7;
        float v = 0.f;
        if (pos != nullptr && k < RP_NPOS) v = pos[(size_t)(row0 + r) * RP_NPOS + k];
        dst[r * LP + HD + k] = v;
    }
}

constexpr int AC_SMEM = (TILE * LQ * 2 + TILE * LP * 3 + TILE + NTOK) * (int)sizeof(float);

__global__ void __launch_bounds__(THREADS, 2)
em_accum_kernel(const float* __restrict__ qkv, const float* __restrict__ pos, const float* __restrict__ lse,
                float* __restrict__ part, int width, int flags) {
    extern __shared__ __align__(16) float sm[];
    float* Qs = sm;                    // [64][LQ]   q rows of this row tile
    float* Ks = Qs + TILE * LQ;        // [64][LQ]
    float* Vs = Ks + TILE * LQ;        // [64][LP]   [v|pos] of the column tile
    float* Ps = Vs + TILE * LP;        // [64][LP]   A tile, later T
    float* Vi = Ps + TILE * LP;        // [64][LP]   [v|pos] of the ROW tile (left factor of V^T A V)
    float* lr = Vi + TILE * LP;        // [64]
    float* lc = lr + TILE;             // [576]
    const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
    const int it = blockIdx.x, h = blockIdx.y;
    const int dir = blockIdx.z & 1, b = blockIdx.z >> 1;
    const int q_img = 2 * b + (1 - dir), kv_img = 2 * b + dir;
    const float* qsrc = qkv + (size_t)q_img * NTOK * LDQKV + h * HD;
    const float* ksrc = qkv + (size_t)kv_img * NTOK * LDQKV + RP_EMBED + h * HD;
    const float* vsrc = qkv + (size_t)kv_img * NTOK * LDQKV + 2 * RP_EMBED + h * HD;
    // left factor of V^T A V: the key/value image's own [v|pos] (vision_transformer.py:222-223) or, with
    // --cross_features, the query image's (:219-220)
    const float* vleft = (flags & RP_EM_CROSS_FEATURES) ? qkv + (size_t)q_img * NTOK * LDQKV + 2 * RP_EMBED + h * HD : vsrc;
    const bool single = (flags & RP_EM_SINGLE_SOFTMAX) != 0;       // --use_single_softmax (:201-203)
    const float* posb = pos ? pos + (size_t)b * NTOK * RP_NPOS : nullptr;
    const float* lse_r = lse + ((((size_t)b * 2 + dir) * 2 + 0) * RP_HEADS + h) * NTOK;
    const float* lse_c = lse + ((((size_t)b * 2 + dir) * 2 + 1) * RP_HEADS + h) * NTOK;

    load_tile64<LQ>(Qs, qsrc + (size_t)(it * TILE) * LDQKV, tid);
    load_vpos_tile(Vi, vleft, posb, it * TILE, tid);
    rp::cp_async_commit();
    for (int c = tid; c < TILE; c += THREADS) lr[c] = lse_r[it * TILE + c];
    for (int c = tid; c < NTOK; c += THREADS) lc[c] = lse_c[c];

    float t[4][9];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii)
#pragma unroll
        for (int c = 0; c < 9; ++c) t[ii][c] = 0.f;

    for (int jt = 0; jt < NTILES; ++jt) {
        __syncthreads();
        load_tile64<LQ>(Ks, ksrc + (size_t)(jt * TILE) * LDQKV, tid);
        load_vpos_tile(Vs, vsrc, posb, jt * TILE, tid);
        rp::cp_async_commit();
        rp::cp_async_wait<0>();
        __syncthreads();

        float s[4][8];
        qk_tile(Qs, Ks, ty, tx, s);
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            float r = lr[ty + 16 * ii];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                float sv = s[ii][jj] * 0.125f;
                // softmax(S,-1)*softmax(S,-2) = exp(S-lse_r) * exp(S-lse_c)      (:205-206)
                Ps[(ty + 16 * ii) * LP + tx + 8 * jj] =
                    single ? expf(sv - r) : expf((sv - r) + (sv - lc[jt * TILE + tx + 8 * jj]));
            }
        }
        __syncwarp();
        // T[i][c] += sum_j A[i][j] V[j][c];  thread owns c in {4tx..+3} U {32+4tx..+3} U {64+tx}
#pragma unroll 2
        for (int j = 0; j < TILE; j += 4) {
            float4 a[4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) a[ii] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * ii) * LP + j);
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {
                const float* vr = Vs + (j + jq) * LP;
                float4 b0 = *reinterpret_cast<const float4*>(vr + 4 * tx);
                float4 b1 = *reinterpret_cast<const float4*>(vr + 32 + 4 * tx);
                float b2 = vr[64 + tx];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    float p = jq == 0 ? a[ii].x : jq == 1 ? a[ii].y : jq == 2 ? a[ii].z : a[ii].w;
                    t[ii][0] = fmaf(p, b0.x, t[ii][0]);
                    t[ii][1] = fmaf(p, b0.y, t[ii][1]);
                    t[ii][2] = fmaf(p, b0.z, t[ii][2]);
                    t[ii][3] = fmaf(p, b0.w, t[ii][3]);
                    t[ii][4] = fmaf(p, b1.x, t[ii][4]);
                    t[ii][5] = fmaf(p, b1.y, t[ii][5]);
                    t[ii][6] = fmaf(p, b1.z, t[ii][6]);
                    t[ii][7] = fmaf(p, b1.w, t[ii][7]);
                    t[ii][8] = fmaf(p, b2, t[ii][8]);
                }
            }
        }
    }
    __syncthreads();   // all warps finished reading Ps as A
    float* Ts = Ps;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        float* tr = Ts + (ty + 16 * ii) * LP;
        *reinterpret_cast<float4*>(tr + 4 * tx) = make_float4(t[ii][0], t[ii][1], t[ii][2], t[ii][3]);
        *reinterpret_cast<float4*>(tr + 32 + 4 * tx) = make_float4(t[ii][4], t[ii][5], t[ii][6], t[ii][7]);
        tr[64 + tx] = t[ii][8];
    }
    __syncthreads();
    // F_part[a][c] = sum_{i in row tile} Vi[i][a] * T[i][c],  a,c < width.  14x7 threads, 5x10 outputs each.
    if (tid < 98) {
        const int ga = tid / 7, gc = tid % 7;
        float f[5][10];
#pragma unroll
        for (int u = 0; u < 5; ++u)
#pragma unroll
            for (int v = 0; v < 10; ++v) f[u][v] = 0.f;
        for (int i = 0; i < TILE; ++i) {
            float av[5], cv[10];
#pragma unroll
            for (int u = 0; u < 5; ++u) av[u] = Vi[i * LP + 5 * ga + u];
#pragma unroll
            for (int v = 0; v < 10; v += 2) {
                float2 q = *reinterpret_cast<const float2*>(Ts + i * LP + 10 * gc + v);
                cv[v] = q.x;
                cv[v + 1] = q.y;
            }
#pragma unroll
            for (int u = 0; u < 5; ++u)
#pragma unroll
                for (int v = 0; v < 10; ++v) f[u][v] = fmaf(av[u], cv[v], f[u][v]);
        }
        float* dst = part + ((((size_t)b * 2 + dir) * RP_HEADS + h) * NTILES + it) * (size_t)(width * width);
#pragma unroll
        for (int u = 0; u < 5; ++u)
#pragma unroll
            for (int v = 0; v < 10; ++v) {
                int a = 5 * ga + u, c = 10 * gc + v;
                if (a < width && c < width) dst[a * width + c] = f[u][v];
            }
    }
}

__global__ void __launch_bounds__(256) em_reduce_kernel(const float* __restrict__ part, float* __restrict__ bil,
                                                        size_t n_mats, int ww) {
    size_t total = n_mats * (size_t)ww;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        size_t mat = o / ww;
        int e = (int)(o % ww);
        const float* p = part + mat * NTILES * (size_t)ww + e;
        float v = 0.f;
#pragma unroll
        for (int t = 0; t < NTILES; ++t) v += p[(size_t)t * ww];
        bil[o] = v;
    }
}

// ============================================================================ proj_fundamental (A7 tail)
// out[2b + (1-dir)][c][o] = bias[o] + sum_{h,a} bil[b,dir,h,a,c] * W[o][h*70+a]
// 35 of the 70 columns c per CTA (the first version took 10: 896 CTAs each re-staging the whole 161 KB weight through
// shared memory = 144 MB of L2 traffic for a 0.36 GFLOP product, 72 us); K chunks of 14 keep the padded weight rows at
// an odd stride (15 floats, conflict free).  Same fmaf order per output as before: bit-identical results.
constexpr int PJ_CC = 35, PJ_KC = 14, PJ_K = RP_HEADS * EMW;   // 210

__global__ void __launch_bounds__(RP_EMBED)
em_project_kernel(const float* __restrict__ bil, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ out) {
    __shared__ __align__(16) float Zs[PJ_CC][PJ_K + 2];   // Z[c][h*70+a]; row pitch 212 floats keeps float2 reads aligned
    __shared__ float Ws[RP_EMBED][PJ_KC + 1];
    const int o = threadIdx.x;
    const int c0 = blockIdx.x * PJ_CC;
    const int dir = blockIdx.y & 1, b = blockIdx.y >> 1;
    const float* F = bil + ((size_t)b * 2 + dir) * RP_HEADS * EMW * EMW;
    for (int e = o; e < PJ_CC * PJ_K; e += RP_EMBED) {
        int k = e / PJ_CC, cc = e % PJ_CC;     // cc fastest: contiguous c in global
        Zs[cc][k] = F[(size_t)k * EMW + c0 + cc];   // F[h][a][c] with k = h*70+a
    }
    float acc[PJ_CC];
#pragma unroll
    for (int cc = 0; cc < PJ_CC; ++cc) acc[cc] = 0.f;
    for (int k0 = 0; k0 < PJ_K; k0 += PJ_KC) {
        __syncthreads();
        for (int e = o; e < RP_EMBED * PJ_KC; e += RP_EMBED) {
            int r = e / PJ_KC, kk = e % PJ_KC;
            Ws[r][kk] = W[(size_t)r * PJ_K + k0 + kk];
        }
        __syncthreads();
        // the loop is bound by shared-memory instructions (one broadcast read per FMA): read Z two k at a time
        static_assert(PJ_KC % 2 == 0, "float2 reads of Z");
#pragma unroll
        for (int kk = 0; kk < PJ_KC; kk += 2) {
            const float w0 = Ws[o][kk], w1 = Ws[o][kk + 1];
#pragma unroll
            for (int cc = 0; cc < PJ_CC; ++cc) {
                const float2 z = *reinterpret_cast<const float2*>(&Zs[cc][k0 + kk]);
                acc[cc] = fmaf(z.x, w0, acc[cc]);       // same order over k as before: bit-identical
                acc[cc] = fmaf(z.y, w1, acc[cc]);
            }
        }
    }
    float bo = bias[o];
    float* dst = out + ((size_t)(2 * b + (1 - dir)) * EMW + c0) * RP_EMBED + o;
#pragma unroll
    for (int cc = 0; cc < PJ_CC; ++cc) dst[(size_t)cc * RP_EMBED] = acc[cc] + bo;
}

// Second version (default; RELPOSE_EM_PROJECT_V1=1 selects the kernel above for A/B runs): ONE CTA per (pair, direction)
// matrix, 384 threads = four column groups (18, 18, 18, 16 of the 70 columns c) x 96 threads, TWO outputs (o, o + 96)
// per thread, Z read four k at a time.  The first version was bound by shared-memory instructions (one broadcast LDS.64
// per two FMAs = 25 % of the FMA rate at best, 60 us for 0.72 GFLOP at 64 pairs) and staged the weights twice per
// matrix; here every LDS.128 of Z feeds eight FMAs (18 + 8 shared-memory instructions per 144 FMAs: FMA bound), the
// twelve warps sit three per scheduler and the weights are staged once per matrix.  Each output accumulates over k in
// ascending order from zero exactly as before: results are bit-identical to the first version.
constexpr int PJ2_G = 4, PJ2_CC = 18, PJ2_OT = RP_EMBED / 2;             // column groups, columns per group, threads per group
constexpr int PJ2_THREADS = PJ2_G * PJ2_OT;                              // 384
constexpr int PJ2_KC = 28, PJ2_ZP = PJ_K + 2;                            // K chunk of the staged weights; pitch of a Z row (212)
constexpr int PJ2_SMEM = (PJ2_G * PJ2_CC * PJ2_ZP + RP_EMBED * (PJ2_KC + 1)) * 4;
static_assert(PJ2_G * PJ2_CC >= EMW && PJ2_ZP % 4 == 0 && PJ2_KC % 4 == 0 && (PJ_K % PJ2_KC) % 2 == 0, "em_project2 tiling");

constexpr int PJ2_WREG = RP_EMBED * PJ2_KC / PJ2_THREADS;                // staged weights per thread and chunk (14)
constexpr int PJ2_ZB = 13;                                               // loads of Z in flight per thread and round
static_assert(RP_EMBED * PJ2_KC % PJ2_THREADS == 0, "em_project2 staging");

// chunk [k0, k0 + KC) of W [192][210] -> registers: all loads of a thread are in flight together
template <int KC>
__device__ __forceinline__ void pj2_load_w(const float* __restrict__ W, int k0, int tid, float (&wn)[PJ2_WREG]) {
#pragma unroll
    for (int i = 0; i < PJ2_WREG; ++i) {
        const int e = tid + i * PJ2_THREADS;
        wn[i] = (e < RP_EMBED * KC) ? W[(size_t)(e / KC) * PJ_K + k0 + e % KC] : 0.f;
    }
}
template <int KC>
__device__ __forceinline__ void pj2_store_w(float (*Ws)[PJ2_KC + 1], int tid, const float (&wn)[PJ2_WREG]) {
#pragma unroll
    for (int i = 0; i < PJ2_WREG; ++i) {
        const int e = tid + i * PJ2_THREADS;
        if (e < RP_EMBED * KC) Ws[e / KC][e % KC] = wn[i];
    }
}

__global__ void __launch_bounds__(PJ2_THREADS)
em_project2_kernel(const float* __restrict__ bil, const float* __restrict__ W, const float* __restrict__ bias,
                   float* __restrict__ out) {
    extern __shared__ __align__(16) float pj_smem[];
    float (*Zs)[PJ2_ZP] = reinterpret_cast<float (*)[PJ2_ZP]>(pj_smem);                       // Z[c][h*70+a], rows >= 70 zero
    float (*Ws)[PJ2_KC + 1] = reinterpret_cast<float (*)[PJ2_KC + 1]>(pj_smem + PJ2_G * PJ2_CC * PJ2_ZP);
    const int tid = threadIdx.x;
    const int g = tid / PJ2_OT, o = tid % PJ2_OT;          // warp-uniform column group; outputs o and o + 96
    const int dir = blockIdx.x & 1, b = blockIdx.x >> 1;
    const float* F = bil + ((size_t)b * 2 + dir) * RP_HEADS * EMW * EMW;
    // One CTA per SM and twelve warps: global-memory latency is not hidden by occupancy, so every staging step issues all
    // of a thread's loads before it touches any of them (the first version of this kernel staged element by element and
    // spent 80 % of its time on the long scoreboard), and the next chunk of W is fetched while the current one is used.
    float wn[PJ2_WREG];
    pj2_load_w<PJ2_KC>(W, 0, tid, wn);
#pragma unroll 1
    for (int e0 = 0; e0 < PJ_K * EMW; e0 += PJ2_ZB * PJ2_THREADS) {      // three rounds of 13 loads in flight per thread
        float zr[PJ2_ZB];
#pragma unroll
        for (int i = 0; i < PJ2_ZB; ++i) {
            const int e = e0 + tid + i * PJ2_THREADS;       // F[h][a][c] flat = k * 70 + c with k = h*70+a: coalesced
            zr[i] = (e < PJ_K * EMW) ? F[e] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < PJ2_ZB; ++i) {
            const int e = e0 + tid + i * PJ2_THREADS;
            if (e < PJ_K * EMW) Zs[e % EMW][e / EMW] = zr[i];
        }
    }
    for (int e = tid; e < (PJ2_G * PJ2_CC - EMW) * PJ2_ZP; e += PJ2_THREADS) Zs[EMW + e / PJ2_ZP][e % PJ2_ZP] = 0.f;
    float acc0[PJ2_CC], acc1[PJ2_CC];
#pragma unroll
    for (int cc = 0; cc < PJ2_CC; ++cc) acc0[cc] = acc1[cc] = 0.f;
    const float (*Zg)[PJ2_ZP] = Zs + g * PJ2_CC;
    constexpr int NFULL = PJ_K / PJ2_KC, KTAIL = PJ_K - NFULL * PJ2_KC;      // 7 chunks of 28, then 14
    static_assert(KTAIL > 0 && KTAIL % 2 == 0, "tail chunk");
    for (int ch = 0; ch <= NFULL; ++ch) {
        const int k0 = ch * PJ2_KC;
        const int kc = ch < NFULL ? PJ2_KC : KTAIL;
        __syncthreads();                                    // the previous chunk of Ws has been consumed (first pass: nothing)
        if (ch < NFULL) pj2_store_w<PJ2_KC>(Ws, tid, wn); else pj2_store_w<KTAIL>(Ws, tid, wn);
        __syncthreads();                                    // also orders the fill of Zs before its first use
        if (ch + 1 < NFULL) pj2_load_w<PJ2_KC>(W, k0 + PJ2_KC, tid, wn);
        else if (ch + 1 == NFULL) pj2_load_w<KTAIL>(W, k0 + PJ2_KC, tid, wn);
        int kk = 0;
        for (; kk + 4 <= kc; kk += 4) {
            const float a0 = Ws[o][kk], a1 = Ws[o][kk + 1], a2 = Ws[o][kk + 2], a3 = Ws[o][kk + 3];
            const float b0 = Ws[o + PJ2_OT][kk], b1 = Ws[o + PJ2_OT][kk + 1], b2 = Ws[o + PJ2_OT][kk + 2], b3 = Ws[o + PJ2_OT][kk + 3];
#pragma unroll
            for (int cc = 0; cc < PJ2_CC; ++cc) {
                const float4 z = *reinterpret_cast<const float4*>(&Zg[cc][k0 + kk]);
                acc0[cc] = fmaf(z.x, a0, acc0[cc]); acc1[cc] = fmaf(z.x, b0, acc1[cc]);
                acc0[cc] = fmaf(z.y, a1, acc0[cc]); acc1[cc] = fmaf(z.y, b1, acc1[cc]);
                acc0[cc] = fmaf(z.z, a2, acc0[cc]); acc1[cc] = fmaf(z.z, b2, acc1[cc]);
                acc0[cc] = fmaf(z.w, a3, acc0[cc]); acc1[cc] = fmaf(z.w, b3, acc1[cc]);
            }
        }
        for (; kk + 2 <= kc; kk += 2) {                     // the last chunk (14) ends with one pair
            const float a0 = Ws[o][kk], a1 = Ws[o][kk + 1];
            const float b0 = Ws[o + PJ2_OT][kk], b1 = Ws[o + PJ2_OT][kk + 1];
#pragma unroll
            for (int cc = 0; cc < PJ2_CC; ++cc) {
                const float2 z = *reinterpret_cast<const float2*>(&Zg[cc][k0 + kk]);
                acc0[cc] = fmaf(z.x, a0, acc0[cc]); acc1[cc] = fmaf(z.x, b0, acc1[cc]);
                acc0[cc] = fmaf(z.y, a1, acc0[cc]); acc1[cc] = fmaf(z.y, b1, acc1[cc]);
            }
        }
    }
    const float bo0 = bias[o], bo1 = bias[o + PJ2_OT];
    const int c0 = g * PJ2_CC;
    float* dst = out + ((size_t)(2 * b + (1 - dir)) * EMW + c0) * RP_EMBED + o;
#pragma unroll
    for (int cc = 0; cc < PJ2_CC; ++cc) {
        if (c0 + cc < EMW) {
            dst[(size_t)cc * RP_EMBED] = acc0[cc] + bo0;
            dst[(size_t)cc * RP_EMBED + PJ2_OT] = acc1[cc] + bo1;
        }
    }
}

int set_smem(const void* fn, int bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        rp::set_error("%s: cudaFuncSetAttribute(%d): %s", what, bytes, cudaGetErrorString(e));
        return (int)e;
    }
    return RP_OK;
}

}  // namespace

extern "C" int rp_self_attention_f32(const float* qkv, float* out, int n_img, int device, void* stream) {
    RP_REQUIRE(qkv && out && n_img > 0, RP_EINVAL, "rp_self_attention: bad argument");
    RP_REQUIRE(rp::aligned16(qkv) && rp::aligned16(out), RP_EALIGN, "rp_self_attention: 16-byte alignment");
    RP_GUARD(device);
    int rc = set_smem((const void*)self_attention_kernel, SA_SMEM, "rp_self_attention");
    if (rc) return rc;
    dim3 grid(NTILES, RP_HEADS, n_img);
    self_attention_kernel<<<grid, THREADS, SA_SMEM, (cudaStream_t)stream>>>(qkv, out, 0);
    return rp::finish_launch("rp_self_attention");
}

extern "C" int rp_cross_attention_f32(const float* qkv, float* out, int n_img, int device, void* stream) {
    RP_REQUIRE(qkv && out && n_img > 0 && n_img % 2 == 0, RP_EINVAL, "rp_cross_attention: n_img must be a positive even number");
    RP_REQUIRE(rp::aligned16(qkv) && rp::aligned16(out), RP_EALIGN, "rp_cross_attention: 16-byte alignment");
    RP_GUARD(device);
    int rc = set_smem((const void*)self_attention_kernel, SA_SMEM, "rp_cross_attention");
    if (rc) return rc;
    dim3 grid(NTILES, RP_HEADS, n_img);
    self_attention_kernel<<<grid, THREADS, SA_SMEM, (cudaStream_t)stream>>>(qkv, out, 1);
    return rp::finish_launch("rp_cross_attention");
}

extern "C" size_t rp_essential_workspace_bytes(int B) {
    if (B <= 0) return 0;
    size_t lse = (size_t)B * 2 * 2 * RP_HEADS * NTOK;
    size_t part = (size_t)B * 2 * RP_HEADS * NTILES * EMW * EMW;
    return (lse + part) * sizeof(float);
}

extern "C" int rp_essential_f32(const float* qkv, const float* pos, float* bil, int B, void* workspace,
                                size_t workspace_bytes, int device, void* stream) {
    return rp_essential_ex_f32(qkv, pos, bil, B, 0, workspace, workspace_bytes, device, stream);
}

extern "C" int rp_essential_ex_f32(const float* qkv, const float* pos, float* bil, int B, int flags, void* workspace,
                                   size_t workspace_bytes, int device, void* stream) {
    RP_REQUIRE(qkv && bil && B > 0 && (flags & ~(RP_EM_SINGLE_SOFTMAX | RP_EM_CROSS_FEATURES)) == 0, RP_EINVAL,
               "rp_essential: bad argument");
    RP_REQUIRE(rp::aligned16(qkv), RP_EALIGN, "rp_essential: qkv must be 16-byte aligned");
    RP_REQUIRE(workspace && workspace_bytes >= rp_essential_workspace_bytes(B), RP_EWORKSPACE,
               "rp_essential: workspace %zu < %zu bytes", workspace_bytes, rp_essential_workspace_bytes(B));
    RP_REQUIRE(rp::aligned16(workspace), RP_EALIGN, "rp_essential: workspace must be 16-byte aligned");
    RP_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    const int width = pos ? EMW : HD;
    float* lse = static_cast<float*>(workspace);
    float* part = lse + (size_t)B * 2 * 2 * RP_HEADS * NTOK;
    int rc = set_smem((const void*)em_stats_kernel, ST_SMEM, "rp_essential(stats)");
    if (rc) return rc;
    rc = set_smem((const void*)em_accum_kernel, AC_SMEM, "rp_essential(accum)");
    if (rc) return rc;
    em_stats_kernel<<<dim3(NTILES, RP_HEADS, B * 4), THREADS, ST_SMEM, st>>>(qkv, lse);
    rc = rp::finish_launch("rp_essential(stats)");
    if (rc) return rc;
    em_accum_kernel<<<dim3(NTILES, RP_HEADS, B * 2), THREADS, AC_SMEM, st>>>(qkv, pos, lse, part, width, flags);
    rc = rp::finish_launch("rp_essential(accum)");
    if (rc) return rc;
    size_t n_mats = (size_t)B * 2 * RP_HEADS;
    size_t total = n_mats * width * width;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4096) blocks = 4096;
    em_reduce_kernel<<<blocks, 256, 0, st>>>(part, bil, n_mats, width * width);
    return rp::finish_launch("rp_essential(reduce)");
}

extern "C" int rp_em_project_f32(const float* bil, const float* W, const float* bias, float* out, int B, int device,
                                 void* stream) {
    RP_REQUIRE(bil && W && bias && out && B > 0, RP_EINVAL, "rp_em_project: bad argument");
    RP_GUARD(device);
    const char* env = std::getenv("RELPOSE_EM_PROJECT_V1");     // read per call: the GPU test toggles it in-process
    if (env && env[0] == '1') {
        em_project_kernel<<<dim3(EMW / PJ_CC, B * 2), RP_EMBED, 0, (cudaStream_t)stream>>>(bil, W, bias, out);
    } else {
        int rc = set_smem((const void*)em_project2_kernel, PJ2_SMEM, "rp_em_project");
        if (rc) return rc;
        em_project2_kernel<<<B * 2, PJ2_THREADS, PJ2_SMEM, (cudaStream_t)stream>>>(bil, W, bias, out);
    }
    return rp::finish_launch("rp_em_project");
}
