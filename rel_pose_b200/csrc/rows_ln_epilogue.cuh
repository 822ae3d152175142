// Epilogue shared by the two kernels that finish a residual-stream update for COMPLETE 192-wide rows of a 128-row tile
//     x' = acc + bias + x                 (attention projection: vision_transformer.py:331,351; fused MLP: :352-353)
// and, optionally, hand the NEXT LayerNorm's output to its consumer as bf16 planes
//     xn = LayerNorm(x') * gamma + beta   (norm2 of the same Block for the projection, norm1 of the next Block for the MLP)
// so that the consuming GEMM kernel receives its A operand by TMA instead of computing the LayerNorm itself.
//
// Why: ncu (profiles/r02_ncu_summary_call6.md) shows the in-kernel LayerNorm of the fused MLP and of the LayerNorm+QKV
// kernel holding 30 % / 36 % of their sixteen epilogue warps' time for ~5 % of their instructions: all warps enter it in
// lockstep (global loads, two dependent warp reductions, shared-memory writes) with nothing else to issue, and the
// tensor pipe drains behind them because the single A-operand buffer is being rewritten.  In a producer's epilogue the
// row is already in registers, the statistics are two 4-way exchanges, and the consumer's tensor pipe no longer waits
// for SIMT work at row-tile boundaries.
//
// Thread layout (16 epilogue warps): warp (q, part) owns TMEM lanes 32q..32q+31 and columns 48 part .. 48 part + 47.
// tcgen05.ld hands a thread one ROW; the 16-column groups are transposed through the warp's private staging patch so
// that residual loads / output stores run on coalesced 128-byte row segments: lane (rr = lane >> 2, cq = lane & 3) then
// holds, for the four rows rt = 32q + 8t + rr, the twelve columns 48 part + 16 gi + 4 cq + {0..3}.
#pragma once
#include "tc_common.cuh"

namespace rowsln {

constexpr int D = 192, STG_LD = 16, PATCH = 32 * STG_LD;   // floats per warp-private staging patch

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// t_acc: tensor-memory address of the accumulator's column 0 in this warp's lane quarter.
// staging: base of the 16 patches (patch index = 4 part + q).  wait_acc(): blocks until the accumulator is complete.
// release(): called once the accumulator is in registers.
// P_ln = 0: no LayerNorm planes.
template <int P_LN, typename WaitFn, typename ReleaseFn>
__device__ __forceinline__ void epilogue(uint32_t t_acc, float* staging, int q, int part, int lane, const float* __restrict__ bias,
                                         const float* __restrict__ res, float* __restrict__ out,
                                         __nv_bfloat16* __restrict__ ln_planes, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float eps, int row_base, int valid_rows,
                                         size_t plane_stride, WaitFn wait_acc, ReleaseFn release) {
    const int rr = lane >> 2, cq = lane & 3;
    float* stg = staging + (part * 4 + q) * PATCH;
    // Phase 1: accumulator -> registers (transposed), then release it at once.  The residual rows are requested only
    // afterwards: their L2 / HBM latency used to sit between the wait and the release, i.e. on the tensor pipe's
    // critical path (the next tile's products wait for this accumulator), now it overlaps with those products.
    float4 o[3][4];
    wait_acc();
#pragma unroll
    for (int gi = 0; gi < 3; ++gi) {
        const int c0 = part * 48 + gi * 16;
        uint32_t a[16];
        tc::tmem_ld_32x32b_x16(t_acc + c0, a);
        tc::tmem_ld_wait();
        if (gi == 2) release();
        __syncwarp();                                    // the previous group's readers are done with the patch
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
            *reinterpret_cast<uint4*>(stg + lane * STG_LD + ((jj ^ ((lane >> 1) & 3)) << 2)) =
                make_uint4(a[4 * jj], a[4 * jj + 1], a[4 * jj + 2], a[4 * jj + 3]);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int rl = t * 8 + rr;
            o[gi][t] = *reinterpret_cast<const float4*>(stg + rl * STG_LD + ((cq ^ ((rl >> 1) & 3)) << 2));
        }
    }
    // Phase 2: + bias + residual, float32 output (coalesced 128-byte row segments)
#pragma unroll
    for (int gi = 0; gi < 3; ++gi) {
        const int col = part * 48 + gi * 16 + cq * 4;
        const float4 sh = __ldg(reinterpret_cast<const float4*>(bias + col));
        float4 r4[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int rt = q * 32 + t * 8 + rr;
            r4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rt < valid_rows) r4[t] = __ldg(reinterpret_cast<const float4*>(res + (size_t)(row_base + rt) * D + col));
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int rt = q * 32 + t * 8 + rr;
            o[gi][t].x = (o[gi][t].x + sh.x) + r4[t].x; o[gi][t].y = (o[gi][t].y + sh.y) + r4[t].y;
            o[gi][t].z = (o[gi][t].z + sh.z) + r4[t].z; o[gi][t].w = (o[gi][t].w + sh.w) + r4[t].w;
            if (rt < valid_rows) *reinterpret_cast<float4*>(out + (size_t)(row_base + rt) * D + col) = o[gi][t];
        }
    }
    if constexpr (P_LN > 0) {
        // LayerNorm statistics of the finished rows: in-lane sum of 12 columns, two shuffles over the four lanes of a row
        // segment, then a 4-way exchange between the part-warps of this lane quarter through the (now idle) patches.
        // Two passes (mean, then centred squares) as in nn.LayerNorm; the four partials are added in the same order by
        // every warp, so all threads of a row agree bit for bit.
        const int bar_id = 2 + q;
        float s[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float v = 0.f;
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) v += (o[gi][t].x + o[gi][t].y) + (o[gi][t].z + o[gi][t].w);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            s[t] = v;
        }
        __syncwarp();                                    // last transposed reads of the patch are complete
        if (cq == 0) {
#pragma unroll
            for (int t = 0; t < 4; ++t) stg[t * 8 + rr] = s[t];
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        float mean[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float v = 0.f;
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) v += staging[(pp * 4 + q) * PATCH + t * 8 + rr];
            mean[t] = v * (1.0f / D);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float v = 0.f;
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) {
                o[gi][t].x -= mean[t]; o[gi][t].y -= mean[t]; o[gi][t].z -= mean[t]; o[gi][t].w -= mean[t];
                v += (o[gi][t].x * o[gi][t].x + o[gi][t].y * o[gi][t].y) + (o[gi][t].z * o[gi][t].z + o[gi][t].w * o[gi][t].w);
            }
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            s[t] = v;
        }
        if (cq == 0) {
#pragma unroll
            for (int t = 0; t < 4; ++t) stg[32 + t * 8 + rr] = s[t];
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        float rstd[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float v = 0.f;
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) v += staging[(pp * 4 + q) * PATCH + 32 + t * 8 + rr];
            rstd[t] = 1.0f / sqrtf(v * (1.0f / D) + eps);
        }
        // a sibling warp may run ahead into its next tile and transpose through its patch: everyone has read first
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
        for (int gi = 0; gi < 3; ++gi) {
            const int col = part * 48 + gi * 16 + cq * 4;
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + col));
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + col));
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int rt = q * 32 + t * 8 + rr;
                float y0 = o[gi][t].x * rstd[t] * g4.x + b4.x, y1 = o[gi][t].y * rstd[t] * g4.y + b4.y;
                float y2 = o[gi][t].z * rstd[t] * g4.z + b4.z, y3 = o[gi][t].w * rstd[t] * g4.w + b4.w;
                __nv_bfloat16* dst = ln_planes + (size_t)(row_base + rt) * D + col;
#pragma unroll
                for (int p = 0; p < P_LN; ++p) {
                    uint2 w;
                    w.x = pack2(y0, y1);
                    w.y = pack2(y2, y3);
                    if (rt < valid_rows) *reinterpret_cast<uint2*>(dst + p * plane_stride) = w;
                    y0 -= __uint_as_float(w.x << 16); y1 -= __uint_as_float(w.x & 0xffff0000u);
                    y2 -= __uint_as_float(w.y << 16); y3 -= __uint_as_float(w.y & 0xffff0000u);
                }
            }
        }
    }
}

}  // namespace rowsln
