// FP32 (true fp32 operands, fp32 accumulate) NT GEMM with fused nn.Linear epilogue, SIMT FFMA path.
// This is the accuracy-first engine behind the fp32 parity configuration (BASELINE.json config 2):
// TF32 / bf16 tensor-core operands cannot hold the 1e-4 parity bar (SURVEY.md section 7 "hard parts").
//   C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) + residual[M,N]
// Reference call sites: vision_transformer.py:323,331 (qkv, proj), vit_layers/mlp.py:21-24 (fc1+GELU,
// fc2), src/model.py:91-98 (pose regressor, ReLU).
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

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == RP_ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    if (act == RP_ACT_RELU) return fmaxf(v, 0.0f);
    return v;
}

__global__ void __launch_bounds__(THREADS, 2)
sgemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                const float* residual, float* C, float* __restrict__ partial, int M, int N, int K, int act,
                int k_per_split) {
    extern __shared__ __align__(16) float smem[];
    float(*As)[BM][LDT] = reinterpret_cast<float(*)[BM][LDT]>(smem);
    float(*Bs)[BN][LDT] = reinterpret_cast<float(*)[BN][LDT]>(smem + STAGES * BM * LDT);

    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    const int ktiles = (kend - kbeg + BK - 1) / BK;

    auto load_tile = [&](int stage, int kt) {
        const int k0 = kbeg + kt * BK;
#pragma unroll
        for (int c = tid; c < BM * (BK / 4); c += THREADS) {
            int r = c / (BK / 4), kc = (c % (BK / 4)) * 4;
            int gr = m0 + r, gk = k0 + kc;
            bool ok = (gr < M) && (gk < kend);
            const float* src = A + (size_t)min(gr, M - 1) * K + min(gk, K - 4);
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
            float v = acc[i][j];
            size_t o = (size_t)row * N + col;
            if (partial) {
                partial[(size_t)blockIdx.z * M * N + o] = v;
            } else {
                if (bias) v += bias[col];
                v = apply_act(v, act);
                if (residual) v += residual[o];
                C[o] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
splitk_epilogue_kernel(const float* __restrict__ partial, const float* __restrict__ bias, const float* residual,
                       float* C, int M, int N, int splits, int act) {
    size_t total = (size_t)M * N;
    for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int s = 0; s < splits; ++s) v += partial[(size_t)s * total + o];   // fixed order: deterministic
        int col = (int)(o % N);
        if (bias) v += bias[col];
        v = apply_act(v, act);
        if (residual) v += residual[o];
        C[o] = v;
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
    cudaStream_t st = (cudaStream_t)stream;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        cudaError_t e = cudaFuncSetAttribute(sgemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) {
            rp::set_error("rp_linear: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        attr_set[device] = true;
    }
    Plan p = make_plan(M, N, K, PLAN_SMS);
    float* partial = nullptr;
    if (p.splits > 1) {
        size_t need = (size_t)p.splits * M * N * sizeof(float);
        RP_REQUIRE(workspace && workspace_bytes >= need, RP_EWORKSPACE,
                   "rp_linear: workspace %zu < %zu bytes (split-K %d)", workspace_bytes, need, p.splits);
        partial = static_cast<float*>(workspace);
    }
    dim3 grid(p.tiles_n, p.tiles_m, p.splits);
    sgemm_nt_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(A, W, bias, residual, C, partial, M, N, K, act, p.k_per_split);
    int rc = rp::finish_launch("rp_linear(sgemm)");
    if (rc != RP_OK) return rc;
    if (p.splits > 1) {
        size_t total = (size_t)M * N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 2048) blocks = 2048;
        splitk_epilogue_kernel<<<blocks, 256, 0, st>>>(partial, bias, residual, C, M, N, p.splits, act);
        rc = rp::finish_launch("rp_linear(splitk epilogue)");
    }
    return rc;
}
