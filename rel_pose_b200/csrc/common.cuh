// Shared helpers for the relpose_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/relpose_b200.h"

namespace rp {

void set_error(const char* fmt, ...);

// RAII: make `device` current for the duration of an API call (autograd worker threads do not
// inherit the caller's current device).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int finish_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return RP_OK;
}

#define RP_REQUIRE(cond, code, ...)       \
    do {                                  \
        if (!(cond)) {                    \
            rp::set_error(__VA_ARGS__);   \
            return (code);                \
        }                                 \
    } while (0)

#define RP_GUARD(device)                                                          \
    rp::DeviceGuard _rp_guard(device);                                            \
    if (_rp_guard.err != cudaSuccess) {                                           \
        rp::set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(_rp_guard.err)); \
        return (int)_rp_guard.err;                                                \
    }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 16-byte cp.async (LDGSTS), global -> shared
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int bytes = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// A forward pass is ~50 dependent launches on one stream, most of them persistent one-CTA-per-SM kernels whose
// prologue (mbarrier init, tensor-memory allocation, tensor-map prefetch, first weight loads) and launch latency
// used to sit exposed between the last CTA of one kernel and the first useful instruction of the next.  Every
// kernel of the library is launched with programmaticStreamSerialization: it may become resident as soon as ALL
// CTAs of its predecessor have executed pdl_launch_dependents() (the first instruction of every kernel) and an SM
// has room, runs its prologue, and blocks in pdl_wait() until the predecessor grid has completed and its memory is
// visible.  Rule: no global memory produced by an earlier kernel is read, and nothing an earlier kernel may still
// read or write is written, before pdl_wait(); all threads execute it.  RELPOSE_PDL=0 launches without the attribute
// (pdl_wait() then returns at once) for A/B runs.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int num_sms(int device) {
    static int cached[64] = {0};
    if (device < 0 || device >= 64) return 148;
    if (cached[device] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
        cached[device] = n;
    }
    return cached[device];
}

}  // namespace rp
