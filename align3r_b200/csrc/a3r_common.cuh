// a3r_common.cuh - error plumbing and small device helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

// thread-local message of the last failure (returned by a3r_last_error)
char* a3r_error_buffer();
int a3r_fail(int code, const char* fmt, ...);

#define A3R_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return a3r_fail(A3R_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#if defined(__CUDACC__)
// Programmatic dependent launch: a kernel launched with the attribute may start while its predecessor in the stream is
// still running; pdl_wait() returns once the predecessor has completed and its memory is visible (a no-op for a normal
// launch), pdl_trigger() lets the successor's CTAs be scheduled from here on.  Every kernel launched with the attribute
// calls pdl_wait() before it touches anything its predecessors wrote.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA; every thread gets the result; `red` is a shared array of >= 32 elements.
// Fixed reduction tree -> bitwise reproducible for a given blockDim.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    T r = (lane < nw) ? red[lane] : T(0);
    r = warp_sum(r);
    __syncthreads();
    return r;
}

// streaming (read-once) global loads: bypass L1 allocation so the small hot blocks stay cached
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ld_stream2(const float2* p) {
    float2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
#endif
