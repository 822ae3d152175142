// LayerNorm of 8 rows per warp, written straight into the K-major SWIZZLE_128B A-operand planes of a [128 x 192] tile
// (three [128 x 64] K blocks per plane) -- the LayerNorm stage of the fused LN+QKV and LN+MLP kernels
// (nn.LayerNorm(192, eps=1e-6), vision_transformer.py:396; two passes: mean, then centred squares).
//
// Two rows per warp instruction: lanes 0-15 own row 2k, lanes 16-31 row 2k+1 of the warp's k-th row pair; lane h (= lane
// & 15) owns the four columns 64 i + 4 h .. 4 h + 3 of each K block i.  Against the first version (one row per warp,
// two columns per lane and K block) that is 12 LDG.128 instead of 24 LDG.64 per warp, 4-step reductions that serve two
// rows at once (32 instead of 80 shuffles per 8 rows) and 24 STS.64 instead of 48 STS.32 -- the stage is latency bound
// (all sixteen warps run it in lockstep while the tensor pipe waits for the A operand), so fewer dependent steps is time.
#pragma once
#include "tc_common.cuh"

namespace lnrows {

constexpr int D = 192, KB = 3, TILE16K = 128 * 64 * 2;

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// rows row0 + 8 ew .. + 7 of x [M,192]; v[pair][12]: K block i -> v[pair][4 i .. 4 i + 3]
__device__ __forceinline__ void load8(const float* __restrict__ x, int M, int tile_row0, int ew, int lane, float (&v)[4][12]) {
    const int half = lane >> 4, h = lane & 15;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int grow = tile_row0 + ew * 8 + 2 * k + half;
#pragma unroll
        for (int i = 0; i < KB; ++i) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow < M) a = __ldg(reinterpret_cast<const float4*>(x + (size_t)grow * D + 64 * i + 4 * h));
            v[k][4 * i] = a.x; v[k][4 * i + 1] = a.y; v[k][4 * i + 2] = a.z; v[k][4 * i + 3] = a.w;
        }
    }
}

__device__ __forceinline__ float half_warp_sum(float s) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// xn: base of the A-operand planes, tile (p, i) at xn + (p * 3 + i) * 16 KiB
template <int P>
__device__ __forceinline__ void finish8(float (&v)[4][12], const float* __restrict__ gamma, const float* __restrict__ beta,
                                        float eps, uint8_t* xn, int ew, int lane) {
    const int half = lane >> 4, h = lane & 15;
    float4 g[KB], bt[KB];
#pragma unroll
    for (int i = 0; i < KB; ++i) {
        g[i] = __ldg(reinterpret_cast<const float4*>(gamma + 64 * i + 4 * h));
        bt[i] = __ldg(reinterpret_cast<const float4*>(beta + 64 * i + 4 * h));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int rl = ew * 8 + 2 * k + half;             // row inside the 128-row tile
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) s += v[k][i];
        const float mean = half_warp_sum(s) * (1.0f / D);
        float qv = 0.f;
#pragma unroll
        for (int i = 0; i < 12; ++i) { const float dlt = v[k][i] - mean; qv = fmaf(dlt, dlt, qv); }
        const float rstd = 1.0f / sqrtf(half_warp_sum(qv) * (1.0f / D) + eps);
        // 16-byte chunk h >> 1 of the 128-byte row, XOR-swizzled by the row; 8 bytes (4 bf16) per lane
        const uint32_t off = (uint32_t)(rl >> 3) * 1024 + (uint32_t)(rl & 7) * 128 +
                             ((((uint32_t)h >> 1) ^ (uint32_t)(rl & 7)) << 4) + (uint32_t)(h & 1) * 8;
#pragma unroll
        for (int i = 0; i < KB; ++i) {
            float y0 = (v[k][4 * i] - mean) * rstd * g[i].x + bt[i].x, y1 = (v[k][4 * i + 1] - mean) * rstd * g[i].y + bt[i].y;
            float y2 = (v[k][4 * i + 2] - mean) * rstd * g[i].z + bt[i].z, y3 = (v[k][4 * i + 3] - mean) * rstd * g[i].w + bt[i].w;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                uint2 w;
                w.x = pack2(y0, y1);
                w.y = pack2(y2, y3);
                *reinterpret_cast<uint2*>(xn + (p * KB + i) * TILE16K + off) = w;
                y0 -= __uint_as_float(w.x << 16); y1 -= __uint_as_float(w.x & 0xffff0000u);
                y2 -= __uint_as_float(w.y << 16); y3 -= __uint_as_float(w.y & 0xffff0000u);
            }
        }
    }
}

}  // namespace lnrows
