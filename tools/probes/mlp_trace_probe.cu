// Probe (not part of the library): timeline of ONE CTA of the fused MLP kernel (mlp_tc.cu compiled with RP_MLP_TRACE).
// Lane 0 of every warp records clock64 at the barrier waits / issues of the chunk pipeline; the host prints, per hidden
// chunk of the steady-state tiles, when each stage happened (cycles relative to the tile's first event).
// build: make -C tools/probes mlp_trace_probe      run: tools/probes/mlp_trace_probe [pairs=64] [P=2] [events=200] [inplace=0]
#define RP_MLP_TRACE 1
#define RP_MLP_TRACE_CTA 5
#include "../../rel_pose_b200/csrc/mlp_tc.cu"
#include <cstdlib>
#include <vector>
#include <map>
namespace rp {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
bool pdl_enabled() { return false; }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv) {
    const int pairs = argc > 1 ? atoi(argv[1]) : 64;
    const int P = argc > 2 ? atoi(argv[2]) : 2;
    const bool inplace = argc > 4 && atoi(argv[4]) != 0;
    const int M = pairs * 2 * 576;
    float *x, *out, *g, *b, *b1, *b2; void *w1, *w2;
    CK(cudaMalloc(&x, (size_t)M * 192 * 4)); CK(cudaMalloc(&out, (size_t)M * 192 * 4));
    CK(cudaMalloc(&g, 768)); CK(cudaMalloc(&b, 768)); CK(cudaMalloc(&b1, 3072)); CK(cudaMalloc(&b2, 768));
    CK(cudaMalloc(&w1, (size_t)P * 768 * 192 * 2)); CK(cudaMalloc(&w2, (size_t)P * 768 * 192 * 2));
    std::vector<float> hx((size_t)M * 192);
    for (size_t i = 0; i < hx.size(); ++i) hx[i] = (float)((i * 2654435761u >> 8) & 0xffff) / 65536.f - 0.5f;
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> ones(768, 1.f);
    CK(cudaMemcpy(g, ones.data(), 768, cudaMemcpyHostToDevice));
    CK(cudaMemset(b, 0, 768)); CK(cudaMemset(b1, 0, 3072)); CK(cudaMemset(b2, 0, 768));
    std::vector<uint16_t> hw((size_t)P * 768 * 192);
    for (size_t i = 0; i < hw.size(); ++i) hw[i] = (uint16_t)(0x3c00 + (i * 40503u & 0xff) + ((i & 1) << 15));   // small bf16 values, both signs
    CK(cudaMemcpy(w1, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(w2, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 3; ++it) {
        cudaEventRecord(e0);
        int rc = rp_mlp_tc(x, g, b, 1e-6f, w1, b1, w2, b2, inplace ? x : out, M, 192, 768, P, 0, nullptr);
        cudaEventRecord(e1);
        if (rc) { printf("rp_mlp_tc rc=%d\n", rc); return 1; }
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("launch %d: %.1f us (traced build)\n", it, ms * 1000);
    }
    static long long tr[NTHREADS / 32][TRACE_EV];
    CK(cudaMemcpyFromSymbol(tr, g_mlp_trace, sizeof(tr)));
    // the last launch's events overwrite the earlier ones (same count), so tr holds launch 2
    long long t0 = tr[4][0] >> 8;
    auto dump = [&](int w, const char* name, int maxev) {
        printf("---- warp %d (%s)\n", w, name);
        long long prev = t0;
        for (int i = 0; i < maxev && tr[w][i]; ++i) {
            const long long t = tr[w][i] >> 8; const int tag = (int)(tr[w][i] & 0xff);
            printf("  ev %4d tag %2d  t=%8lld  (+%6lld)\n", i, tag, t - t0, t - prev);
            prev = t;
        }
    };
    const int nev = argc > 3 ? atoi(argv[3]) : 200;
    dump(1, "fc1 issuer even: 1 before acc1_empty wait, 2 after, 3 chunk issued", nev);
    dump(2, "fc1 issuer odd", nev);
    dump(3, "fc2 issuer: 4 before h_full wait, 5 after, 6 issued, 7 weights reloaded", nev);
    dump(4, "GELU warp 0: 10 before acc1_full wait, 11 after, 12 GELU done, 13 h_empty ok, 14 h_full arrived, 16 LN done, 17/18 output epilogue", nev * 2);
    dump(19, "GELU warp 15", nev * 2);
    return 0;
}
