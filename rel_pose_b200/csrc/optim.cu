// Training step after the path: gradient clipping by global norm, Adam with L2 weight
// decay and the learning rate of the step in TWO multi-tensor launches (+ one tiny reduction), no host sync:
//   train.py:161  torch.nn.utils.clip_grad_norm_(model.parameters(), 2.5)
//   train.py:162  optimizer.step()              torch.optim.Adam(lr, weight_decay)  (train.py:69)
//   train.py:165  scheduler.step()              OneCycleLR: the host passes this step's lr (rel_pose_b200/optim.py)
// The reference's eager sequence is ~6 elementwise passes over 19.3 M parameters in ~1200 small launches (123 tensors)
// plus a device->host sync for the norm; here every parameter, gradient and moment is read once and written once.
// Arithmetic follows torch/optim/adam.py::_single_tensor_adam operation by operation (fp32).
#include "common.cuh"

namespace {

struct TensorDesc {
    float* p;          // parameter
    const float* g;    // gradient
    float* m;          // exp_avg
    float* v;          // exp_avg_sq
    long long n;
};

constexpr int OPT_TPB = 256;

// sum of squares of one chunk of one gradient tensor -> partial[block] (fixed summation order: reproducible)
__global__ void __launch_bounds__(OPT_TPB) grad_sqnorm_multi_kernel(const TensorDesc* __restrict__ td,
                                                                   const int* __restrict__ blk_tensor,
                                                                   const long long* __restrict__ blk_off, int chunk,
                                                                   float* __restrict__ partial) {
    const TensorDesc t = td[blk_tensor[blockIdx.x]];
    const long long off = blk_off[blockIdx.x];
    const long long end = (off + chunk < t.n) ? off + chunk : t.n;
    const float* g = t.g;
    float s = 0.f;
    const bool vec = ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) && ((off & 3) == 0);
    long long i = off + (long long)threadIdx.x * 4;
    if (vec) {
        for (; i + 4 <= end; i += (long long)OPT_TPB * 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(g + i));
            s += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        }
        for (long long j = ((end - off) & ~3ll) + off + threadIdx.x; j < end; j += OPT_TPB) s += g[j] * g[j];
    } else {
        for (long long j = off + threadIdx.x; j < end; j += OPT_TPB) s += g[j] * g[j];
    }
    __shared__ float red[OPT_TPB / 32];
    s = rp::warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = threadIdx.x < OPT_TPB / 32 ? red[threadIdx.x] : 0.f;
        r = rp::warp_sum(r);
        if (threadIdx.x == 0) partial[blockIdx.x] = r;
    }
}

// total norm = sqrt(sum of the partials), one block, fixed order
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += (double)partial[i];
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double r = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (threadIdx.x == 0) out[0] = (float)sqrt(r);
    }
}

struct AdamArgs {
    float max_norm;        // <= 0: no clipping
    float lr_over_bc1;     // lr / (1 - beta1^t)
    float bc2_sqrt;        // sqrt(1 - beta2^t)
    float beta1, beta2, eps, weight_decay;
    float omb1, omb2;      // 1 - beta1, 1 - beta2 evaluated in double like torch's Python scalars
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float clip, const AdamArgs& a) {
    g *= clip;                                         // clip_grad_norm_: grad.mul_(clip_coef_clamped)
    g = fmaf(a.weight_decay, p, g);                    // grad.add(param, alpha=weight_decay)
    m = fmaf(g - m, a.omb1, m);                     // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(a.omb2, g * g, v * a.beta2);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) / a.bc2_sqrt + a.eps; // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = fmaf(-a.lr_over_bc1, m / denom, p);            // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(OPT_TPB) adam_clip_multi_kernel(const TensorDesc* __restrict__ td,
                                                                 const int* __restrict__ blk_tensor,
                                                                 const long long* __restrict__ blk_off, int chunk,
                                                                 const float* __restrict__ total_norm, AdamArgs a,
                                                                 const float* __restrict__ step_hyper) {
    // CUDA-graph replays: the two per-step scalars come from device memory (written by a copy that precedes the replay)
    if (step_hyper) { a.lr_over_bc1 = __ldg(step_hyper); a.bc2_sqrt = __ldg(step_hyper + 1); }
    const TensorDesc t = td[blk_tensor[blockIdx.x]];
    const long long off = blk_off[blockIdx.x];
    const long long end = (off + chunk < t.n) ? off + chunk : t.n;
    float clip = 1.0f;
    if (a.max_norm > 0.f) {
        clip = a.max_norm / (__ldg(total_norm) + 1e-6f);   // clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
        clip = fminf(clip, 1.0f);
    }
    const bool vec = (((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                        reinterpret_cast<uintptr_t>(t.v)) & 15u) == 0) && ((off & 3) == 0);
    if (vec) {
        const long long vend = off + ((end - off) & ~3ll);
        for (long long i = off + (long long)threadIdx.x * 4; i < vend; i += (long long)OPT_TPB * 4) {
            float4 p = *reinterpret_cast<float4*>(t.p + i);
            const float4 g = __ldg(reinterpret_cast<const float4*>(t.g + i));
            float4 m = *reinterpret_cast<float4*>(t.m + i);
            float4 v = *reinterpret_cast<float4*>(t.v + i);
            adam_one(p.x, g.x, m.x, v.x, clip, a);
            adam_one(p.y, g.y, m.y, v.y, clip, a);
            adam_one(p.z, g.z, m.z, v.z, clip, a);
            adam_one(p.w, g.w, m.w, v.w, clip, a);
            *reinterpret_cast<float4*>(t.p + i) = p;
            *reinterpret_cast<float4*>(t.m + i) = m;
            *reinterpret_cast<float4*>(t.v + i) = v;
        }
        for (long long j = vend + threadIdx.x; j < end; j += OPT_TPB) adam_one(t.p[j], t.g[j], t.m[j], t.v[j], clip, a);
    } else {
        for (long long j = off + threadIdx.x; j < end; j += OPT_TPB) adam_one(t.p[j], t.g[j], t.m[j], t.v[j], clip, a);
    }
}

}  // namespace

extern "C" int rp_grad_norm_multi(const void* descs, const int* blk_tensor, const int64_t* blk_off, int nblk, int chunk,
                                  float* partial, float* norm_out, int device, void* stream) {
    RP_REQUIRE(descs && blk_tensor && blk_off && partial && norm_out && nblk > 0 && chunk > 0 && (chunk % 4) == 0, RP_EINVAL,
               "rp_grad_norm_multi: bad argument");
    RP_GUARD(device);
    cudaStream_t st = (cudaStream_t)stream;
    grad_sqnorm_multi_kernel<<<nblk, OPT_TPB, 0, st>>>(static_cast<const TensorDesc*>(descs), blk_tensor,
                                                      reinterpret_cast<const long long*>(blk_off), chunk, partial);
    reduce_partials_kernel<<<1, 1024, 0, st>>>(partial, nblk, norm_out);
    return rp::finish_launch("rp_grad_norm_multi");
}

extern "C" int rp_adam_clip_step_multi(const void* descs, const int* blk_tensor, const int64_t* blk_off, int nblk, int chunk,
                                       const float* total_norm, double max_norm, double lr, double beta1, double beta2, double eps,
                                       double weight_decay, int step, int device, void* stream) {
    RP_REQUIRE(descs && blk_tensor && blk_off && nblk > 0 && chunk > 0 && (chunk % 4) == 0 && step >= 1, RP_EINVAL,
               "rp_adam_clip_step_multi: bad argument");
    RP_REQUIRE(max_norm <= 0.f || total_norm, RP_EINVAL, "rp_adam_clip_step_multi: clipping needs the total norm");
    RP_GUARD(device);
    AdamArgs a;
    // hyper-parameters arrive as doubles (Python floats) and are rounded to fp32 once, exactly where torch does it
    a.max_norm = (float)max_norm;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    a.lr_over_bc1 = (float)(lr / bc1);
    a.bc2_sqrt = (float)sqrt(bc2);
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.weight_decay = (float)weight_decay;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    adam_clip_multi_kernel<<<nblk, OPT_TPB, 0, (cudaStream_t)stream>>>(static_cast<const TensorDesc*>(descs), blk_tensor,
                                                                      reinterpret_cast<const long long*>(blk_off), chunk,
                                                                      total_norm, a, nullptr);
    return rp::finish_launch("rp_adam_clip_step_multi");
}

// Same update with the two per-step scalars {lr / (1 - beta1^t), sqrt(1 - beta2^t)} read from DEVICE memory
// (step_hyper[2], float32), so that the launch can be captured in a CUDA graph once and replayed every step: the
// host writes the pair for step t into pinned memory and enqueues a 8-byte copy in front of the replay.
extern "C" int rp_adam_clip_step_multi_dev(const void* descs, const int* blk_tensor, const int64_t* blk_off, int nblk, int chunk,
                                           const float* total_norm, double max_norm, const float* step_hyper, double beta1,
                                           double beta2, double eps, double weight_decay, int device, void* stream) {
    RP_REQUIRE(descs && blk_tensor && blk_off && step_hyper && nblk > 0 && chunk > 0 && (chunk % 4) == 0, RP_EINVAL,
               "rp_adam_clip_step_multi_dev: bad argument");
    RP_REQUIRE(max_norm <= 0.f || total_norm, RP_EINVAL, "rp_adam_clip_step_multi_dev: clipping needs the total norm");
    RP_GUARD(device);
    AdamArgs a;
    a.max_norm = (float)max_norm;
    a.lr_over_bc1 = 0.f; a.bc2_sqrt = 1.f;
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.weight_decay = (float)weight_decay;
    a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    adam_clip_multi_kernel<<<nblk, OPT_TPB, 0, (cudaStream_t)stream>>>(static_cast<const TensorDesc*>(descs), blk_tensor,
                                                                      reinterpret_cast<const long long*>(blk_off), chunk,
                                                                      total_norm, a, step_hyper);
    return rp::finish_launch("rp_adam_clip_step_multi_dev");
}
