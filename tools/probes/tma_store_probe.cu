// Probe (not part of the library): throughput of shared -> global bulk stores / bulk reduce-adds issued by ONE thread per CTA,
// as a function of the box shape, with every SM running the same loop (grid = #SMs) -- the epilogues of the fused MLP and
// LayerNorm+QKV kernels wait for these reads before they may reuse their staging patches.
// build: make -C tools/probes tma_store_probe       run: tools/probes/tma_store_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"
namespace rp { void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); } }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int ROWS_PER_CTA = 1024;     // each CTA owns 1024 rows x 192 float32 of the global matrix

__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(tc::smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_red_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(tc::smem_u32(ssrc)), "r"(bytes) : "memory");
}

// mode: 0 tensor store, 1 tensor reduce-add, 2 1-D store, 3 1-D reduce-add.  box_cols x box_rows float32 per op.
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, float* g, int mode, int box_cols, int box_rows,
                                            int iters, int wait_each, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    tc::fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int row0 = blockIdx.x * ROWS_PER_CTA;
        const int ncolblk = 192 / box_cols, nrowblk = ROWS_PER_CTA / box_rows;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const int cb = i % ncolblk, rb = (i / ncolblk) % nrowblk;
            const int c0 = cb * box_cols, r0 = row0 + rb * box_rows;
            if (mode == 0) tc::tma_store_2d(&tm, smem, c0, r0);
            else if (mode == 1) tc::tma_reduce_add_2d(&tm, smem, c0, r0);
            else if (mode == 2) bulk_store_1d(g + (size_t)r0 * 192, smem, (uint32_t)(box_cols * box_rows * 4));
            else bulk_red_1d(g + (size_t)r0 * 192, smem, (uint32_t)(box_cols * box_rows * 4));
            tc::tma_store_commit();
            if (wait_each) tc::tma_store_wait_read();
        }
        tc::tma_store_wait_read();
        const long long t1 = clock64();
        tc::tma_store_wait_all();
        const long long t2 = clock64();
        cycles[2 * blockIdx.x] = t1 - t0;
        cycles[2 * blockIdx.x + 1] = t2 - t0;
    }
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t rows = (size_t)sms * ROWS_PER_CTA;
    float* g; CK(cudaMalloc(&g, rows * 192 * 4)); CK(cudaMemset(g, 0, rows * 192 * 4));
    long long* cyc; CK(cudaMalloc(&cyc, sms * 16));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    tc::EncodeTiledFn fn = tc::get_encode_fn();
    struct Cfg { int cols, rows; CUtensorMapSwizzle sw; const char* name; };
    const Cfg cfgs[] = {{16, 32, CU_TENSOR_MAP_SWIZZLE_64B, "16x32 sw64 (2 KB)"},   {32, 32, CU_TENSOR_MAP_SWIZZLE_128B, "32x32 sw128 (4 KB)"},
                        {32, 128, CU_TENSOR_MAP_SWIZZLE_128B, "32x128 sw128 (16 KB)"}, {64, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "64x32 none (8 KB)"},
                        {192, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "192x32 none (24 KB)"}, {192, 128, CU_TENSOR_MAP_SWIZZLE_NONE, "192x128 none (96 KB)"}};
    for (int grid : {1, sms}) {
        for (const Cfg& c : cfgs) {
            CUtensorMap tm;
            cuuint64_t gdim[2] = {192, rows}; cuuint64_t gstr[1] = {768};
            cuuint32_t box[2] = {(cuuint32_t)c.cols, (cuuint32_t)c.rows}; cuuint32_t es[2] = {1, 1};
            CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("%s: tensor map failed %d\n", c.name, (int)r); continue; }
            for (int mode = 0; mode < 4; ++mode) {
                if (mode >= 2 && c.cols != 192) continue;             // 1-D copies: whole rows only (contiguous in global memory)
                for (int wait_each = 0; wait_each < 2; ++wait_each) {
                    const int bytes = c.cols * c.rows * 4;
                    const int iters = 4 * 1024 * 1024 / bytes;         // 4 MB per CTA
                    for (int rep = 0; rep < 2; ++rep) {
                        probe<<<grid, 128, 100 * 1024>>>(tm, g, mode, c.cols, c.rows, iters, wait_each, cyc);
                        CK(cudaDeviceSynchronize());
                    }
                    std::vector<long long> h(2 * grid);
                    CK(cudaMemcpy(h.data(), cyc, grid * 16, cudaMemcpyDeviceToHost));
                    double rd = 0, all = 0;
                    for (int i = 0; i < grid; ++i) { rd += h[2 * i]; all += h[2 * i + 1]; }
                    rd /= grid; all /= grid;
                    static const char* mn[] = {"tensor store", "tensor red.add", "1-D store", "1-D red.add"};
                    printf("grid %3d  %-22s %-14s %s: %7.0f cycles/op (smem read), %6.1f B/clk/SM read, %6.1f B/clk/SM complete, %5.1f cycles per box row\n",
                           grid, c.name, mn[mode], wait_each ? "wait each " : "back2back ", rd / iters, bytes * (double)iters / rd,
                           bytes * (double)iters / all, rd / iters / c.rows);
                }
            }
        }
    }
    return 0;
}
