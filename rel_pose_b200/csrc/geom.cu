// SE3 group kernels (lietorch replacement, A12) and batched 3x3 SVD / essential -> (R,t) (config 3).
// HBM-bound: one thread per element, register resident; the AoS records (7 / 6 / 9 floats) are moved
// through shared memory so that every global access of a warp is a run of consecutive 4-byte words
// (odd record widths make the strided shared-memory reads conflict-free).
//
// Group formulas: lietorch so3.h / se3.h (pinned lietorch==0.2, /root/reference/environment.yml:20);
// call sites src/geom/losses.py:8-10.  Backward: gradient w.r.t. a LEFT perturbation exp(d)X, written
// to the first 6 of 7 slots, slot 7 = 0 (lietorch group_ops convention).  Parity unpinned by the
// reference (lietorch absent) -- checked against oracle/geom_oracle.py.
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int TPB = 128;
constexpr float EPSL = 1e-6f;   // lietorch EPS

struct V3 {
    float x, y, z;
};
struct Q4 {
    float x, y, z, w;
};
struct M3 {
    float m[9];
};

__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {   // lietorch SO3(const Scalar*) normalises on load
    float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    float r = 1.0f / n;
    return Q4{q.x * r, q.y * r, q.z * r, q.w * r};
}
__device__ __forceinline__ Q4 qconj(Q4 q) { return Q4{-q.x, -q.y, -q.z, q.w}; }
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    return Q4{a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
              a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
__device__ __forceinline__ V3 qrot(Q4 q, V3 p) {   // p + w*uv + qv x uv, uv = 2 qv x p
    V3 qv = v3(q.x, q.y, q.z);
    V3 uv = cross(qv, p);
    uv = uv + uv;
    return p + q.w * uv + cross(qv, uv);
}
__device__ __forceinline__ M3 hat(V3 v) {
    M3 r;
    r.m[0] = 0.f;  r.m[1] = -v.z; r.m[2] = v.y;
    r.m[3] = v.z;  r.m[4] = 0.f;  r.m[5] = -v.x;
    r.m[6] = -v.y; r.m[7] = v.x;  r.m[8] = 0.f;
    return r;
}
__device__ __forceinline__ M3 mmul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) r.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
    return r;
}
__device__ __forceinline__ M3 madd(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i];
    return r;
}
__device__ __forceinline__ M3 mscale(float s, const M3& a) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.m[i] = s * a.m[i];
    return r;
}
__device__ __forceinline__ V3 mvec(const M3& a, V3 v) {    // a v
    return v3(a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
              a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z);
}
__device__ __forceinline__ V3 mtvec(const M3& a, V3 v) {   // a^T v  (= row vector v times a)
    return v3(a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
              a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z);
}

__device__ __forceinline__ V3 so3_log(Q4 q) {
    float n2 = q.x * q.x + q.y * q.y + q.z * q.z;
    float w = q.w, f;
    if (n2 < EPSL * EPSL) {
        f = 2.0f / w - (2.0f / 3.0f) * n2 / (w * w * w);
    } else {
        float n = sqrtf(n2);
        if (fabsf(w) < EPSL) f = (w > 0.f ? 3.14159265358979323846f : -3.14159265358979323846f) / n;
        else f = 2.0f * atanf(n / w) / n;
    }
    return v3(f * q.x, f * q.y, f * q.z);
}
__device__ __forceinline__ Q4 so3_exp(V3 phi) {
    float th2 = dot(phi, phi), th = sqrtf(th2), im, re;
    if (th < EPSL) {
        float th4 = th2 * th2;
        im = 0.5f - th2 / 48.0f + th4 / 3840.0f;
        re = 1.0f - th2 / 8.0f + th4 / 384.0f;
    } else {
        float s, c;
        sincosf(0.5f * th, &s, &c);
        im = s / th;
        re = c;
    }
    return Q4{im * phi.x, im * phi.y, im * phi.z, re};
}
// I + c1*Phi + c2*Phi^2
__device__ __forceinline__ M3 poly_phi(V3 phi, float c1, float c2) {
    M3 P = hat(phi), P2 = mmul(P, P), r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.m[i] = c1 * P.m[i] + c2 * P2.m[i];
    r.m[0] += 1.f; r.m[4] += 1.f; r.m[8] += 1.f;
    return r;
}
__device__ __forceinline__ M3 so3_left_jacobian(V3 phi) {
    float th2 = dot(phi, phi), th = sqrtf(th2), c1, c2;
    if (th < EPSL) {
        c1 = 0.5f - th2 / 24.0f;
        c2 = 1.0f / 6.0f - th2 / 120.0f;
    } else {
        float s, c;
        sincosf(th, &s, &c);
        // 1-cos(th) = 2 sin^2(th/2): avoids cancellation in float32 for small angles
        float sh = sinf(0.5f * th);
        c1 = 2.0f * sh * sh / th2;
        c2 = (th - s) / (th2 * th);
        if (th < 0.05f) c2 = 1.0f / 6.0f - th2 / 120.0f + th2 * th2 / 5040.0f;   // series: (th - sin th) cancels
        (void)c;
    }
    return poly_phi(phi, c1, c2);
}
__device__ __forceinline__ M3 so3_left_jacobian_inverse(V3 phi) {
    float th2 = dot(phi, phi), th = sqrtf(th2), c;
    if (th < 0.02f) {   // lietorch switches at 1e-6; the closed form cancels badly in float32 below ~1e-2
        c = 1.0f / 12.0f + th2 / 720.0f;
    } else {
        float s, co;
        sincosf(0.5f * th, &s, &co);
        c = (1.0f - th * co / (2.0f * s)) / th2;
    }
    return poly_phi(phi, -0.5f, c);
}
// Barfoot's Q(tau, phi): upper-right block of the SE3 left jacobian
__device__ M3 calcQ(V3 tau, V3 phi) {
    M3 T = hat(tau), P = hat(phi);
    float th2 = dot(phi, phi), th = sqrtf(th2), th4 = th2 * th2, c1, c2, c3;
    if (th < 0.05f) {   // series (float32-safe; lietorch uses them below 1e-6 only)
        c1 = 1.0f / 6.0f - th2 / 120.0f + th4 / 5040.0f;
        c2 = 1.0f / 24.0f - th2 / 720.0f + th4 / 40320.0f;
        c3 = 1.0f / 120.0f - th2 / 2520.0f + th4 / 120960.0f;
    } else {
        float s, c;
        sincosf(th, &s, &c);
        c1 = (th - s) / (th2 * th);
        c2 = (th2 + 2.0f * c - 2.0f) / (2.0f * th4);
        c3 = (2.0f * th - 3.0f * s + th * c) / (2.0f * th4 * th);
        if (th < 0.6f) {   // the closed forms lose digits to cancellation in float32: extend the series
            float th6 = th4 * th2;
            c2 = 1.0f / 24.0f - th2 / 720.0f + th4 / 40320.0f - th6 / 3628800.0f;
            c3 = 1.0f / 120.0f - th2 / 2520.0f + th4 / 120960.0f - th6 / 9979200.0f;
            c1 = 1.0f / 6.0f - th2 / 120.0f + th4 / 5040.0f - th6 / 362880.0f;
        }
    }
    M3 PT = mmul(P, T), TP = mmul(T, P), PTP = mmul(PT, P);
    M3 PP = mmul(P, P);
    M3 PPT = mmul(PP, T), TPP = mmul(TP, P);
    M3 PTPP = mmul(PTP, P), PPTP = mmul(PP, TP);
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; ++i)
        r.m[i] = 0.5f * T.m[i] + c1 * (PT.m[i] + TP.m[i] + PTP.m[i]) + c2 * (PPT.m[i] + TPP.m[i] - 3.0f * PTP.m[i]) +
                 c3 * (PTPP.m[i] + PPTP.m[i]);
    return r;
}

struct SE3e {
    V3 t;
    Q4 q;
};

// ------------------------------------------------------------------ coalesced AoS <-> registers
// VEC: a full block moves its TPB*W contiguous floats as 128-bit words with streaming hints (one address computation
// per thread, no per-element bounds test: ~6 instead of ~45 instructions for W = 9); the last partial block and
// unaligned bases take the scalar path.  Measured A/B on one B200 (N = 2^20, profiles/r01_geom_staging_ab.md): the
// instruction-heavy kernels gain (svd3 39.7 -> 34.5 us, essential_to_rt 43.4 -> 37.8 us), the short bandwidth-bound SE3
// kernels LOSE (se3_mul 16.6 -> 20.0 us with streaming hints, 17.2 us without; se3_inv 11.6 -> 13.1 us): with only
// 2-4 wide requests per thread fewer bytes are in flight than with 7-14 scalar ones.  So VEC is per kernel.
template <int W, bool VEC>
__device__ __forceinline__ bool stage_fast(const float* g, int64_t base_elem, int64_t n) {
    return VEC && base_elem + TPB <= n && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
}
template <int W, bool VEC = false>
__device__ __forceinline__ void stage_in(const float* __restrict__ g, float* s, int64_t base_elem, int64_t n) {
    constexpr int NV = TPB * W / 4;
    static_assert((TPB * W) % 4 == 0, "a block's slab must be a whole number of float4");
    if (stage_fast<W, VEC>(g, base_elem, n)) {
        const float4* g4 = reinterpret_cast<const float4*>(g + base_elem * W);
        float4* s4 = reinterpret_cast<float4*>(s);
#pragma unroll
        for (int k = 0; k < (NV + TPB - 1) / TPB; ++k) {
            const int i = k * TPB + threadIdx.x;
            if ((k + 1) * TPB <= NV || i < NV) s4[i] = __ldcs(g4 + i);
        }
        return;
    }
    int64_t first = base_elem * W;
    int64_t total = n * W;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        int64_t i = first + (int64_t)k * TPB + threadIdx.x;
        s[k * TPB + threadIdx.x] = (i < total) ? g[i] : 0.f;
    }
}
template <int W, bool VEC = false>
__device__ __forceinline__ void stage_out(float* __restrict__ g, const float* s, int64_t base_elem, int64_t n) {
    constexpr int NV = TPB * W / 4;
    if (stage_fast<W, VEC>(g, base_elem, n)) {
        float4* g4 = reinterpret_cast<float4*>(g + base_elem * W);
        const float4* s4 = reinterpret_cast<const float4*>(s);
#pragma unroll
        for (int k = 0; k < (NV + TPB - 1) / TPB; ++k) {
            const int i = k * TPB + threadIdx.x;
            if ((k + 1) * TPB <= NV || i < NV) __stcs(g4 + i, s4[i]);
        }
        return;
    }
    int64_t first = base_elem * W;
    int64_t total = n * W;
#pragma unroll
    for (int k = 0; k < W; ++k) {
        int64_t i = first + (int64_t)k * TPB + threadIdx.x;
        if (i < total) g[i] = s[k * TPB + threadIdx.x];
    }
}
__device__ __forceinline__ SE3e read_se3(const float* s) {
    const float* p = s + threadIdx.x * 7;
    SE3e e;
    e.t = v3(p[0], p[1], p[2]);
    e.q = qnormalize(Q4{p[3], p[4], p[5], p[6]});
    return e;
}
__device__ __forceinline__ void write7(float* s, V3 a, float b0, float b1, float b2, float b3) {
    float* p = s + threadIdx.x * 7;
    p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = b0; p[4] = b1; p[5] = b2; p[6] = b3;
}
__device__ __forceinline__ void write6(float* s, V3 a, V3 b) {
    float* p = s + threadIdx.x * 6;
    p[0] = a.x; p[1] = a.y; p[2] = a.z; p[3] = b.x; p[4] = b.y; p[5] = b.z;
}

// ------------------------------------------------------------------ SE3 kernels
__global__ void __launch_bounds__(TPB) se3_mul_fwd_kernel(const float* X, const float* Y, float* Z, int64_t n) {
    __shared__ __align__(16) float sx[TPB * 7], sy[TPB * 7];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<7>(X, sx, base, n);
    stage_in<7>(Y, sy, base, n);
    __syncthreads();
    SE3e a = read_se3(sx), b = read_se3(sy);
    __syncthreads();
    V3 t = a.t + qrot(a.q, b.t);
    Q4 q = qmul(a.q, b.q);
    write7(sx, t, q.x, q.y, q.z, q.w);
    __syncthreads();
    stage_out<7>(Z, sx, base, n);
}

__global__ void __launch_bounds__(TPB)
se3_mul_bwd_kernel(const float* dZ, const float* X, const float* Y, float* dX, float* dY, int64_t n) {
    __shared__ __align__(16) float sg[TPB * 7], sx[TPB * 7];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<7>(dZ, sg, base, n);
    stage_in<7>(X, sx, base, n);
    __syncthreads();
    const float* g = sg + threadIdx.x * 7;
    V3 gt = v3(g[0], g[1], g[2]), gp = v3(g[3], g[4], g[5]);
    SE3e a = read_se3(sx);
    __syncthreads();
    // dX = dZ ;  dY = dZ * Ad(X):  tau' = R^T g_tau, phi' = R^T (g_tau x t + g_phi)
    Q4 qc = qconj(a.q);
    V3 yt = qrot(qc, gt);
    V3 yp = qrot(qc, cross(gt, a.t) + gp);
    write7(sg, gt, gp.x, gp.y, gp.z, 0.f);
    write7(sx, yt, yp.x, yp.y, yp.z, 0.f);
    __syncthreads();
    stage_out<7>(dX, sg, base, n);
    stage_out<7>(dY, sx, base, n);
}

__global__ void __launch_bounds__(TPB) se3_inv_fwd_kernel(const float* X, float* Y, int64_t n) {
    __shared__ __align__(16) float sx[TPB * 7];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<7>(X, sx, base, n);
    __syncthreads();
    SE3e a = read_se3(sx);
    __syncthreads();
    Q4 qi = qconj(a.q);
    V3 t = neg(qrot(qi, a.t));
    write7(sx, t, qi.x, qi.y, qi.z, qi.w);
    __syncthreads();
    stage_out<7>(Y, sx, base, n);
}

__global__ void __launch_bounds__(TPB) se3_inv_bwd_kernel(const float* dY, const float* X, float* dX, int64_t n) {
    __shared__ __align__(16) float sg[TPB * 7], sx[TPB * 7];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<7>(dY, sg, base, n);
    stage_in<7>(X, sx, base, n);
    __syncthreads();
    const float* g = sg + threadIdx.x * 7;
    V3 gt = v3(g[0], g[1], g[2]), gp = v3(g[3], g[4], g[5]);
    SE3e a = read_se3(sx);
    __syncthreads();
    // dX = -dY * Ad(X^-1):  tau' = -R g_tau ;  phi' = (R g_tau) x t - R g_phi
    V3 rgt = qrot(a.q, gt);
    V3 xt = neg(rgt);
    V3 xp = cross(rgt, a.t) - qrot(a.q, gp);
    write7(sg, xt, xp.x, xp.y, xp.z, 0.f);
    __syncthreads();
    stage_out<7>(dX, sg, base, n);
}

__global__ void __launch_bounds__(TPB) se3_log_fwd_kernel(const float* X, float* A, int64_t n) {
    __shared__ __align__(16) float sx[TPB * 7];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<7>(X, sx, base, n);
    __syncthreads();
    SE3e a = read_se3(sx);
    __syncthreads();
    V3 phi = so3_log(a.q);
    V3 tau = mvec(so3_left_jacobian_inverse(phi), a.t);
    write6(sx, tau, phi);
    __syncthreads();
    stage_out<6>(A, sx, base, n);
}

__global__ void __launch_bounds__(TPB) se3_log_bwd_kernel(const float* dA, const float* X, float* dX, int64_t n) {
    __shared__ __align__(16) float sg[TPB * 7], sx[TPB * 7];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<6>(dA, sg, base, n);
    stage_in<7>(X, sx, base, n);
    __syncthreads();
    const float* g = sg + threadIdx.x * 6;
    V3 gt = v3(g[0], g[1], g[2]), gp = v3(g[3], g[4], g[5]);
    SE3e a = read_se3(sx);
    __syncthreads();
    V3 phi = so3_log(a.q);
    M3 Ji = so3_left_jacobian_inverse(phi);
    V3 tau = mvec(Ji, a.t);
    M3 Q = calcQ(tau, phi);
    // dX = da * [[Ji, -Ji Q Ji],[0, Ji]]
    V3 u = mtvec(Ji, gt);                          // Ji^T g_tau
    V3 xp = mtvec(Ji, gp - mtvec(Q, u));           // Ji^T (g_phi - Q^T Ji^T g_tau)
    write7(sx, u, xp.x, xp.y, xp.z, 0.f);
    __syncthreads();
    stage_out<7>(dX, sx, base, n);
}

__global__ void __launch_bounds__(TPB) se3_exp_fwd_kernel(const float* A, float* X, int64_t n) {
    __shared__ __align__(16) float sx[TPB * 7];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<6>(A, sx, base, n);
    __syncthreads();
    const float* p = sx + threadIdx.x * 6;
    V3 tau = v3(p[0], p[1], p[2]), phi = v3(p[3], p[4], p[5]);
    __syncthreads();
    Q4 q = so3_exp(phi);
    V3 t = mvec(so3_left_jacobian(phi), tau);
    write7(sx, t, q.x, q.y, q.z, q.w);
    __syncthreads();
    stage_out<7>(X, sx, base, n);
}

__global__ void __launch_bounds__(TPB) se3_exp_bwd_kernel(const float* dX, const float* A, float* dA, int64_t n) {
    __shared__ __align__(16) float sg[TPB * 7], sa[TPB * 6];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<7>(dX, sg, base, n);
    stage_in<6>(A, sa, base, n);
    __syncthreads();
    const float* g = sg + threadIdx.x * 7;
    const float* p = sa + threadIdx.x * 6;
    V3 gt = v3(g[0], g[1], g[2]), gp = v3(g[3], g[4], g[5]);
    V3 tau = v3(p[0], p[1], p[2]), phi = v3(p[3], p[4], p[5]);
    __syncthreads();
    M3 J = so3_left_jacobian(phi);
    M3 Q = calcQ(tau, phi);
    // da = dX[:6] * [[J, Q],[0, J]]
    V3 at = mtvec(J, gt);
    V3 ap = mtvec(Q, gt) + mtvec(J, gp);
    write6(sa, at, ap);
    __syncthreads();
    stage_out<6>(dA, sa, base, n);
}

// ------------------------------------------------------------------ 3x3 SVD (one-sided Jacobi)
// Columns of A are rotated in pairs until mutually orthogonal: A V = U diag(S).  Register resident,
// fixed sweep count with a per-rotation skip, so warps do not diverge in control flow.
struct Svd3 {
    V3 u[3];
    float s[3];
    V3 v[3];   // columns
    float det_u, det_v;   // +-1, known from the construction (no determinant is evaluated)
};

// Bare MUFU.RCP / MUFU.RSQ.  `__fdividef` and `rsqrtf` wrap each MUFU in a denormal range check (FSETP + two FMUL:
// 12 of the 72 instructions of one rotation); the rotation only feeds them r >= 1, 1 + t^2 >= 1 and quantities whose
// flush-to-zero limit is a harmless identity rotation (2*gamma -> 0 gives t = 0, c = 1, s = 0).
__device__ __forceinline__ float mufu_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float mufu_rsq(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float mufu_sqrt(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One Jacobi rotation of columns p, q of A.  MUFU-only arithmetic (rcp / rsqrt / sqrt approximations): a slightly
// inexact angle costs nothing -- the next rotation removes what is left -- while c, s are normalised consistently so
// the accumulated V stays orthonormal to ~1e-7.  Returns false if the pair is already orthogonal (c = 1, s = 0).
// `flags`: bit 0 = another sweep is needed because of this pair, bit 1 = the carried norms are inconsistent.
//
// The squared column norms alpha, beta are CARRIED, not recomputed: a rotation by the root t of t^2 + 2 zeta t - 1 = 0
// changes them by -+ t gamma (Golub & Van Loan 8.4), one FFMA + clamp each instead of two dot products.  They only
// steer the angle and the convergence test (their drift, ~2e-7 of the larger norm per rotation, is far below both);
// the singular values come from norms recomputed after the last sweep.  A carried norm that has drifted BELOW what
// Cauchy-Schwarz allows (gamma^2 > alpha beta: the clamped norm of a (near-)null column, i.e. every essential matrix)
// must not reach the tests: the product is raised to gamma^2 here (one FMNMX, the rotation then proceeds and asks for
// another sweep) and the sweep loop recomputes the three norms before that sweep.  |zeta| may overflow to inf for
// columns 19 orders of magnitude apart; sqrt.approx keeps that an identity rotation (t = 1 / inf = 0) where
// r * rsqrt(r) would be inf * 0.  The kernels are instruction-issue bound (~0.9 IPC per scheduler): every
// instruction saved per rotation is time.
__device__ __forceinline__ bool jacobi_rot(V3& ap, V3& aq, float& alpha, float& beta, float& c, float& sn, unsigned& flags) {
    const float gamma = dot(ap, aq);
    const float g2 = gamma * gamma, ab0 = alpha * beta;
    const float ab = fmaxf(ab0, g2);
    if (g2 > ab0) flags |= 2u;
    // already orthogonal to fp32 precision: gamma^2 <= (1e-7)^2 alpha beta  (no sqrt)
    if (g2 <= 1e-14f * ab) return false;
    // Jacobi converges quadratically: a pair whose cosine is below 3e-4 BEFORE its rotation is orthogonal to ~1e-7
    // after it, so such a rotation does not ask for another sweep
    if (g2 > 1e-7f * ab) flags |= 1u;
    const float zeta = (beta - alpha) * mufu_rcp(2.0f * gamma);
    const float t = copysignf(mufu_rcp(fabsf(zeta) + mufu_sqrt(fmaf(zeta, zeta, 1.0f))), zeta);
    c = mufu_rsq(fmaf(t, t, 1.0f));
    sn = c * t;
    const V3 np_ = c * ap - sn * aq, nq = sn * ap + c * aq;
    ap = np_; aq = nq;
    alpha = fmaxf(fmaf(-t, gamma, alpha), 0.0f);
    beta = fmaxf(fmaf(t, gamma, beta), 0.0f);
    return true;
}
// A V = U Sigma: the sweeps orthogonalise the COLUMNS of A only (6 of the 12 rotated vectors' multiply-adds per
// rotation; V is not accumulated).  Afterwards U comes from the normalised columns and V from V = A^T U Sigma^-1:
// v0 = normalise(A^T u0), v1 = normalise(A^T u1 - (.,v0) v0), v2 = +-(v0 x v1) with the sign that makes sigma_2 =
// u2 . A v2 non-negative.  V is orthonormal by construction; the direction of v1 carries an error ~eps sigma_0 / sigma_1
// (instead of Jacobi's high relative accuracy), which reaches U Sigma V^T only multiplied by sigma_1, i.e. at eps
// sigma_0 like every other rounding error of the factorisation -- and essential matrices have sigma_0 = sigma_1.
// Saves ~90 of ~920 instructions per matrix in kernels that are instruction-issue bound.
__device__ __forceinline__ Svd3 svd3(const float* e /* row-major 3x3 */) {
    V3 a0 = v3(e[0], e[3], e[6]), a1 = v3(e[1], e[4], e[7]), a2 = v3(e[2], e[5], e[8]);   // columns
    float n0 = dot(a0, a0), n1 = dot(a1, a1), n2 = dot(a2, a2);
    // Cyclic one-sided Jacobi converges quadratically: 3-4 sweeps reach fp32 precision for almost every matrix.
    // The loop ends as soon as no lane of the warp saw a large rotation during a sweep (warp-uniform exit, no
    // divergence); 8 sweeps is a safety bound.
    unsigned flags = 1u;
#pragma unroll 1
    for (int sweep = 0; sweep < 8; ++sweep) {
        if (!__any_sync(0xffffffffu, flags != 0)) break;
        if (flags & 2u) { n0 = dot(a0, a0); n1 = dot(a1, a1); n2 = dot(a2, a2); }   // a carried norm drifted: refresh
        flags = 0;
        float c, sn;
        jacobi_rot(a0, a1, n0, n1, c, sn, flags);
        jacobi_rot(a0, a2, n0, n2, c, sn, flags);
        jacobi_rot(a1, a2, n1, n2, c, sn, flags);
    }
    n0 = dot(a0, a0); n1 = dot(a1, a1); n2 = dot(a2, a2);
    // the two dominant columns, in order (the third is only needed through sigma_2 below)
    if (n0 < n1) { V3 t = a0; a0 = a1; a1 = t; float f = n0; n0 = n1; n1 = f; }
    if (n0 < n2) { V3 t = a0; a0 = a2; a2 = t; float f = n0; n0 = n2; n2 = f; }
    if (n1 < n2) { a1 = a2; n1 = n2; }
    Svd3 r;
    // bare MUFU.RSQ (the range-checked rsqrtf costs ~8 instructions a call): zero norms are excluded explicitly, and a
    // flushed denormal norm is a zero column to fp32 anyway
    const float i0 = n0 > 0.f ? mufu_rsq(n0) : 0.f, i1 = n1 > 0.f ? mufu_rsq(n1) : 0.f;
    r.s[0] = n0 * i0; r.s[1] = n1 * i1;                  // |a| = n / sqrt(n)
    const bool rank2 = r.s[1] > 1e-12f * r.s[0] && r.s[1] > 0.f;
    // U: normalise the two dominant columns, complete by a cross product (rank-deficient safe)
    const V3 u0 = r.s[0] > 0.f ? i0 * a0 : v3(1, 0, 0);
    V3 u1;
    if (rank2) {
        u1 = i1 * a1;
        u1 = u1 - dot(u1, u0) * u0;                      // one Gram-Schmidt polish
        u1 = mufu_rsq(dot(u1, u1)) * u1;                 // |u1| ~ 1 here
    } else {                                             // rank <= 1: any unit vector orthogonal to u0
        const V3 ax = fabsf(u0.x) < 0.6f ? v3(1, 0, 0) : v3(0, 1, 0);
        u1 = cross(u0, ax);
        u1 = mufu_rsq(dot(u1, u1)) * u1;                 // |u0 x ax|^2 >= 0.64
    }
    const V3 u2 = cross(u0, u1);                         // det [u0 u1 u2] = +1
    // V = A^T U Sigma^-1, orthonormalised.  The original columns are read again here (shared memory) instead of being
    // kept in nine registers through the sweeps: 32 registers per thread = 16 resident blocks per SM instead of 12.
    const V3 A0 = v3(e[0], e[3], e[6]), A1 = v3(e[1], e[4], e[7]), A2 = v3(e[2], e[5], e[8]);
    V3 v0 = v3(dot(A0, u0), dot(A1, u0), dot(A2, u0));   // = sigma_0 v0
    const float q0 = dot(v0, v0);
    v0 = q0 > 0.f ? mufu_rsq(q0) * v0 : v3(1, 0, 0);
    V3 v1 = v3(dot(A0, u1), dot(A1, u1), dot(A2, u1));   // = sigma_1 v1 (+ rounding noise ~eps sigma_0)
    v1 = v1 - dot(v1, v0) * v0;
    const float q1 = dot(v1, v1);
    if (rank2 && q1 > 0.f) {
        v1 = mufu_rsq(q1) * v1;
    } else {                                             // sigma_1 = 0 to fp32: any unit vector orthogonal to v0
        const V3 ax = fabsf(v0.x) < 0.6f ? v3(1, 0, 0) : v3(0, 1, 0);
        v1 = cross(v0, ax);
        v1 = mufu_rsq(dot(v1, v1)) * v1;
    }
    V3 v2 = cross(v0, v1);                               // det [v0 v1 v2] = +1 ...
    const V3 av2 = v2.x * A0 + v2.y * A1 + v2.z * A2;    // A v2 = sigma_2 u2
    float sg = dot(u2, av2);
    r.det_u = 1.0f; r.det_v = 1.0f;
    if (sg < 0.f) { v2 = neg(v2); sg = -sg; r.det_v = -1.0f; }   // ... unless the third column is flipped
    r.s[2] = sg;
    r.u[0] = u0; r.u[1] = u1; r.u[2] = u2;
    r.v[0] = v0; r.v[1] = v1; r.v[2] = v2;
    return r;
}

__global__ void __launch_bounds__(TPB, 16) svd3_kernel(const float* E, float* U, float* S, float* V, int64_t n) {
    __shared__ __align__(16) float se[TPB * 9], sv[TPB * 9], ss[TPB * 3];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<9, true>(E, se, base, n);
    __syncthreads();
    Svd3 r = svd3(se + threadIdx.x * 9);
    __syncthreads();
    float* pu = se + threadIdx.x * 9;
    float* pv = sv + threadIdx.x * 9;
#pragma unroll
    for (int c = 0; c < 3; ++c) {   // row-major 3x3 with the vectors as COLUMNS
        pu[0 + c] = r.u[c].x; pu[3 + c] = r.u[c].y; pu[6 + c] = r.u[c].z;
        pv[0 + c] = r.v[c].x; pv[3 + c] = r.v[c].y; pv[6 + c] = r.v[c].z;
        ss[threadIdx.x * 3 + c] = r.s[c];
    }
    __syncthreads();
    stage_out<9, true>(U, se, base, n);
    stage_out<9, true>(V, sv, base, n);
    stage_out<3, true>(S, ss, base, n);
}

__device__ __forceinline__ float det3(V3 a, V3 b, V3 c) { return dot(a, cross(b, c)); }

__global__ void __launch_bounds__(TPB, 16) essential_to_rt_kernel(const float* E, float* R1, float* R2, float* T, int64_t n) {
    __shared__ __align__(16) float se[TPB * 9], s2[TPB * 9], st[TPB * 3];
    int64_t base = (int64_t)blockIdx.x * TPB;
    stage_in<9, true>(E, se, base, n);
    __syncthreads();
    Svd3 r = svd3(se + threadIdx.x * 9);
    __syncthreads();
    // The decomposition needs PROPER rotations U, V with E = U diag(s, s, ~0) V^T.  svd3 builds u2 = u0 x u1, so U is
    // proper; taking v2 = v0 x v1 makes V proper as well -- the sign svd3 may have put on its v2 belongs to sigma_2,
    // which the decomposition discards.  (Reading only v0, v1 here lets the compiler drop svd3's sigma_2 / sign code.)
    const V3 u0 = r.u[0], u1 = r.u[1], u2 = r.u[2];
    const V3 v0 = r.v[0], v1 = r.v[1], v2 = cross(r.v[0], r.v[1]);
    // U W = [u1, -u0, u2] ; U W^T = [-u1, u0, u2] ;  R = (U W) V^T = sum_k (UW)_k v_k^T
    V3 w0 = u1, w1 = neg(u0);
    float* p1 = se + threadIdx.x * 9;
    float* p2 = s2 + threadIdx.x * 9;
    const float uw[3][3] = {{w0.x, w1.x, u2.x}, {w0.y, w1.y, u2.y}, {w0.z, w1.z, u2.z}};
    const float vv[3][3] = {{v0.x, v1.x, v2.x}, {v0.y, v1.y, v2.y}, {v0.z, v1.z, v2.z}};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float a = fmaf(uw[i][0], vv[j][0], uw[i][1] * vv[j][1]);
            p1[i * 3 + j] = fmaf(uw[i][2], vv[j][2], a);       // U W V^T
            p2[i * 3 + j] = fmaf(uw[i][2], vv[j][2], -a);      // U W^T V^T
        }
    st[threadIdx.x * 3 + 0] = u2.x; st[threadIdx.x * 3 + 1] = u2.y; st[threadIdx.x * 3 + 2] = u2.z;
    __syncthreads();
    stage_out<9, true>(R1, se, base, n);
    stage_out<9, true>(R2, s2, base, n);
    stage_out<3, true>(T, st, base, n);
}

inline int nblocks(int64_t n) { return (int)((n + TPB - 1) / TPB); }

}  // namespace

#define RP_GEOM_PROLOGUE(name, cond)                                           \
    RP_REQUIRE((cond) && n > 0, RP_EINVAL, name ": bad argument");             \
    RP_REQUIRE(n <= (int64_t)2147483647 * TPB, RP_EINVAL, name ": n too large"); \
    RP_GUARD(device);                                                          \
    cudaStream_t st = (cudaStream_t)stream;

extern "C" int rp_se3_mul_fwd_f32(const float* X, const float* Y, float* Z, int64_t n, int device, void* stream) {
    RP_GEOM_PROLOGUE("rp_se3_mul_fwd", X && Y && Z)
    se3_mul_fwd_kernel<<<nblocks(n), TPB, 0, st>>>(X, Y, Z, n);
    return rp::finish_launch("rp_se3_mul_fwd");
}
extern "C" int rp_se3_mul_bwd_f32(const float* dZ, const float* X, const float* Y, float* dX, float* dY, int64_t n,
                                  int device, void* stream) {
    RP_GEOM_PROLOGUE("rp_se3_mul_bwd", dZ && X && Y && dX && dY)
    se3_mul_bwd_kernel<<<nblocks(n), TPB, 0, st>>>(dZ, X, Y, dX, dY, n);
    return rp::finish_launch("rp_se3_mul_bwd");
}
extern "C" int rp_se3_inv_fwd_f32(const float* X, float* Y, int64_t n, int device, void* stream) {
    RP_GEOM_PROLOGUE("rp_se3_inv_fwd", X && Y)
    se3_inv_fwd_kernel<<<nblocks(n), TPB, 0, st>>>(X, Y, n);
    return rp::finish_launch("rp_se3_inv_fwd");
}
extern "C" int rp_se3_inv_bwd_f32(const float* dY, const float* X, float* dX, int64_t n, int device, void* stream) {
    RP_GEOM_PROLOGUE("rp_se3_inv_bwd", dY && X && dX)
    se3_inv_bwd_kernel<<<nblocks(n), TPB, 0, st>>>(dY, X, dX, n);
    return rp::finish_launch("rp_se3_inv_bwd");
}
extern "C" int rp_se3_log_fwd_f32(const float* X, float* a, int64_t n, int device, void* stream) {
    RP_GEOM_PROLOGUE("rp_se3_log_fwd", X && a)
    se3_log_fwd_kernel<<<nblocks(n), TPB, 0, st>>>(X, a, n);
    return rp::finish_launch("rp_se3_log_fwd");
}
extern "C" int rp_se3_log_bwd_f32(const float* da, const float* X, float* dX, int64_t n, int device, void* stream) {
    RP_GEOM_PROLOGUE("rp_se3_log_bwd", da && X && dX)
    se3_log_bwd_kernel<<<nblocks(n), TPB, 0, st>>>(da, X, dX, n);
    return rp::finish_launch("rp_se3_log_bwd");
}
extern "C" int rp_se3_exp_fwd_f32(const float* a, float* X, int64_t n, int device, void* stream) {
    RP_GEOM_PROLOGUE("rp_se3_exp_fwd", a && X)
    se3_exp_fwd_kernel<<<nblocks(n), TPB, 0, st>>>(a, X, n);
    return rp::finish_launch("rp_se3_exp_fwd");
}
extern "C" int rp_se3_exp_bwd_f32(const float* dX, const float* a, float* da, int64_t n, int device, void* stream) {
    RP_GEOM_PROLOGUE("rp_se3_exp_bwd", dX && a && da)
    se3_exp_bwd_kernel<<<nblocks(n), TPB, 0, st>>>(dX, a, da, n);
    return rp::finish_launch("rp_se3_exp_bwd");
}
extern "C" int rp_svd3_f32(const float* E, float* U, float* S, float* V, int64_t n, int device, void* stream) {
    RP_GEOM_PROLOGUE("rp_svd3", E && U && S && V)
    svd3_kernel<<<nblocks(n), TPB, 0, st>>>(E, U, S, V, n);
    return rp::finish_launch("rp_svd3");
}
extern "C" int rp_essential_to_rt_f32(const float* E, float* R1, float* R2, float* t, int64_t n, int device,
                                      void* stream) {
    RP_GEOM_PROLOGUE("rp_essential_to_rt", E && R1 && R2 && t)
    essential_to_rt_kernel<<<nblocks(n), TPB, 0, st>>>(E, R1, R2, t, n);
    return rp::finish_launch("rp_essential_to_rt");
}
