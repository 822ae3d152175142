// Probe (not part of the library): does tcgen05.mma accept a K-major SWIZZLE_128B A operand whose start address is
// 128-byte aligned but NOT 1024-byte aligned (a window of rows shifted by k rows inside a larger swizzled buffer)?
// Variants: descriptor base_offset field (bits 49-51) = 0 or = (addr >> 7) & 7.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I rel_pose_b200/csrc -o tools/probes/umma_shift_probe tools/probes/umma_shift_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "tc_common.cuh"
namespace rp { void set_error(const char*, ...) {} }

constexpr int ROWS = 256, N = 64, K = 64;

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* X, const __nv_bfloat16* W, float* out, int shift, int use_base_offset) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* xa = smem;                    // 256 rows x 128 B, swizzled by absolute row
    uint8_t* wb = smem + ROWS * 128;       // 64 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < ROWS * 8; i += 128) {
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(xa + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(X + r * K + c * 8);
    }
    for (int i = threadIdx.x; i < N * 8; i += 128) {
        int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(wb + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(W + r * K + c * 8);
    }
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc(&slot, 64);
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint32_t a_addr = tc::smem_u32(xa) + shift * 128;
        uint64_t da = tc::make_kmajor_sw128_desc(a_addr);
        if (use_base_offset) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
        const uint64_t db = tc::make_kmajor_sw128_desc(tc::smem_u32(wb));
        if (tc::elect_one_sync()) {
            for (int k = 0; k < K / 16; ++k) tc::umma_bf16(tm, da + 2 * k, db + 2 * k, tc::make_idesc_bf16(128, N), k > 0);
            tc::umma_commit(&bar);
        }
        __syncwarp();
    }
    tc::mbar_wait(&bar, 0);
    tc::tcgen05_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
        tc::tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(r[i]);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tm, 64);
}

int main() {
    std::vector<__nv_bfloat16> X(ROWS * K), W(N * K);
    std::vector<float> Xf(ROWS * K), Wf(N * K);
    srand(1);
    for (int i = 0; i < ROWS * K; ++i) { float v = (rand() % 255 - 127) / 64.0f; X[i] = __float2bfloat16(v); Xf[i] = __bfloat162float(X[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 255 - 127) / 64.0f; W[i] = __float2bfloat16(v); Wf[i] = __bfloat162float(W[i]); }
    __nv_bfloat16 *dX, *dW; float* dO;
    cudaMalloc(&dX, X.size() * 2); cudaMalloc(&dW, W.size() * 2); cudaMalloc(&dO, 128 * N * 4);
    cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice);
    const int smem = ROWS * 128 + N * 128 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> out(128 * N);
    for (int ubo = 0; ubo < 2; ++ubo)
        for (int shift : {0, 1, 2, 3, 5, 7, 8, 9, 58, 59, 116}) {
            cudaMemset(dO, 0, 128 * N * 4);
            probe<<<1, 128, smem>>>(dX, dW, dO, shift, ubo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("base_offset=%d shift=%d: CUDA error %s\n", ubo, shift, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0; int bad = 0;
            for (int i = 0; i < 128; ++i)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)Xf[(shift + i) * K + k] * Wf[n * K + k];
                    double err = fabs(ref - out[i * N + n]);
                    if (err > maxerr) maxerr = err;
                    if (err > 1e-3) ++bad;
                }
            printf("base_offset_field=%d shift=%3d rows: max_err=%.3e bad=%d/%d %s\n", ubo, shift, maxerr, bad, 128 * N, bad ? "MISMATCH" : "ok");
        }
    return 0;
}
