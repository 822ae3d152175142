// Probe (not part of the library): tcgen05.mma issue / execution rate on B200 for M=128, K=16, bf16, operands in
// shared memory (SWIZZLE_128B K-major), as a function of N and of whether consecutive MMAs accumulate into the SAME
// tensor-memory accumulator (a dependent chain, what a K loop is) or alternate between NACC accumulators.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I rel_pose_b200/csrc -o tools/probes/umma_rate_probe tools/probes/umma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
namespace rp { void set_error(const char*, ...) {} }

template <int N, int NACC, bool TS>
__global__ void __launch_bounds__(128) rate(long long* out, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a = smem;                 // [128][64] bf16
    uint8_t* b = smem + 16384;         // [256][64] bf16
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        const uint64_t da = tc::make_kmajor_sw128_desc(tc::smem_u32(a));
        const uint64_t db = tc::make_kmajor_sw128_desc(tc::smem_u32(b));
        constexpr uint32_t idesc = tc::make_idesc_bf16(128, N);
        long long t0 = 0, t1 = 0;
        if (tc::elect_one_sync()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int c = 0; c < NACC; ++c) {
                        if (TS) tc::umma_bf16_ts(tm + c * (N > 128 ? 256 : 128), tm + 448 + 8 * k, db + 2 * k, idesc, 1u);
                        else tc::umma_bf16(tm + c * (N > 128 ? 256 : 128), da + 2 * k, db + 2 * k, idesc, 1u);
                    }
                }
            }
            tc::umma_commit(&bar);
        }
        __syncwarp();
        tc::mbar_wait(&bar, 0);
        t1 = clock64();
        if (threadIdx.x == 0 || t0 != 0) { if (t0 != 0) out[0] = t1 - t0; }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tm, 512);
}

template <int N, int NACC, bool TS>
void run(const char* name) {
    long long* d; cudaMalloc(&d, 8);
    const int smem = 16384 + 32768 + 1024, iters = 2000;
    cudaFuncSetAttribute(rate<N, NACC, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate<N, NACC, TS><<<1, 128, smem>>>(d, 10);
    cudaDeviceSynchronize();
    rate<N, NACC, TS><<<1, 128, smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 4.0 * NACC);
    printf("%-34s N=%3d accumulators=%d : %7.1f cycles per MMA (nominal %3d)  %s\n", name, N, NACC, per, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<64, 1, false>("A,B smem, one accumulator");
    run<64, 2, false>("A,B smem, alternating accumulators");
    run<64, 1, true>("A tmem, one accumulator");
    run<64, 2, true>("A tmem, alternating accumulators");
    run<96, 1, false>("A,B smem, one accumulator");
    run<96, 2, false>("A,B smem, alternating accumulators");
    run<128, 1, false>("A,B smem, one accumulator");
    run<128, 2, false>("A,B smem, alternating accumulators");
    run<128, 1, true>("A tmem, one accumulator");
    run<192, 1, false>("A,B smem, one accumulator");
    run<192, 2, false>("A,B smem, alternating accumulators");
    run<192, 1, true>("A tmem, one accumulator");
    run<256, 1, false>("A,B smem, one accumulator");
    return 0;
}
