// Probe (not part of the library): does tcgen05.ld (tensor memory -> registers) share bandwidth with tcgen05.mma?
// One CTA: warp 0 issues a chain of M=128 MMAs (N and operand source as template parameters), warps 1..NLD read
// 32 lanes x 32 columns of OTHER tensor-memory columns in a loop.  Reported: cycles per MMA alone / with the loaders
// running, bytes per cycle of the loaders alone / with the MMAs running.
// build: make -C tools/probes      run on a B200: tools/probes/tmem_ld_probe
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
namespace rp { void set_error(const char*, ...) {} }

template <int N, bool TS, int NLD>
__global__ void __launch_bounds__(32 * (1 + NLD)) probe(long long* out, int mma_iters, int ld_iters) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a = smem;                 // [128][64] bf16
    uint8_t* b = smem + 16384;         // [256][64] bf16
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
    tc::fence_proxy_async_smem();
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        const uint64_t da = tc::make_kmajor_sw128_desc(tc::smem_u32(a));
        const uint64_t db = tc::make_kmajor_sw128_desc(tc::smem_u32(b));
        constexpr uint32_t idesc = tc::make_idesc_bf16(128, N);
        long long t0 = 0, t1 = 0;
        if (mma_iters > 0) {
            if (tc::elect_one_sync()) {
                t0 = clock64();
                for (int i = 0; i < mma_iters; ++i) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (TS) tc::umma_bf16_ts(tm, tm + 448 + 8 * k, db + 2 * k, idesc, 1u);
                        else tc::umma_bf16(tm, da + 2 * k, db + 2 * k, idesc, 1u);
                    }
                }
                tc::umma_commit(&bar);
            }
            __syncwarp();
            tc::mbar_wait(&bar, 0);
            t1 = clock64();
            if (t0 != 0) out[0] = t1 - t0;
        }
    } else if (ld_iters > 0) {
        const uint32_t t_lane = tm + ((uint32_t)(((warp - 1) & 3) * 32) << 16) + 256 + ((warp - 1) >> 2) * 32;
        uint32_t acc = 0;
        const long long t0 = clock64();
        for (int i = 0; i < ld_iters; ++i) {
            uint32_t r[32];
            tc::tmem_ld_32x32b_x32(t_lane, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= r[j];
        }
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
        if (acc == 0x12345678u) out[15] = acc;
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tm, 512);
}

template <int N, bool TS, int NLD>
void run(const char* name) {
    long long* d; cudaMalloc(&d, 16 * 8);
    const int smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(probe<N, TS, NLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[16];
    const int mi = 2000, li = 4000;
    auto go = [&](int m, int l) {
        cudaMemset(d, 0, 16 * 8);
        probe<N, TS, NLD><<<1, 32 * (1 + NLD), smem>>>(d, m, l);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
        cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
    };
    go(10, 10);
    go(mi, 0);
    const double mma_alone = (double)h[0] / (mi * 4.0);
    go(0, li);
    long long mx = 0; for (int w = 1; w <= NLD; ++w) mx = h[w] > mx ? h[w] : mx;
    const double ld_alone = (double)NLD * li * 4096.0 / (double)mx;
    // both: size the loops so that they run about equally long
    const int li2 = (int)(mma_alone * mi * 4.0 / ((double)mx / li));
    go(mi, li2 > 0 ? li2 : 1);
    mx = 0; for (int w = 1; w <= NLD; ++w) mx = h[w] > mx ? h[w] : mx;
    const double mma_both = (double)h[0] / (mi * 4.0), ld_both = (double)NLD * li2 * 4096.0 / (double)mx;
    printf("%-10s N=%3d loaders=%d : MMA %6.1f cyc alone, %6.1f with loads (nominal %3d) | tcgen05.ld %6.1f B/clk alone, %6.1f with MMAs\n",
           name, N, NLD, mma_alone, mma_both, 128 * N / 256, ld_alone, ld_both);
    cudaFree(d);
}

int main() {
    run<64, false, 4>("A,B smem");
    run<64, true, 4>("A tmem");
    run<96, false, 4>("A,B smem");
    run<96, false, 8>("A,B smem");
    run<192, false, 4>("A,B smem");
    run<192, true, 4>("A tmem");
    run<192, true, 8>("A tmem");
    run<256, false, 8>("A,B smem");
    return 0;
}
