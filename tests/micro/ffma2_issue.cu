// Microbenchmark (diagnostic, not product): does packed FFMA2 relieve the issue slot on sm_100a?
// Per loop iteration and thread: NF fp32 FMAs (as FFMA or as FFMA2 pairs), NA ALU-pipe ops (FMNMX), ND FADDs.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, float b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(pk(b,b)), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 c){ u64 r; asm volatile("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(c)); return r; }
__device__ __forceinline__ float ffma1(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd1(float a, float c){ float r; asm volatile("add.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(c)); return r; }
__device__ __forceinline__ float fmin1(float a, float b){ float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
template <bool PACK, int NF, int NA, int ND>
__global__ void __launch_bounds__(512, 1) k(float* out, int iters, float s) {
    float a[16], m[16], d[16];
    u64 A[8], D[8];
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 1e-3f + i; m[i] = a[i]; d[i] = a[i]; }
    for (int i = 0; i < 8; ++i) { A[i] = pk(a[2*i], a[2*i+1]); D[i] = A[i]; }
    const u64 one2 = *reinterpret_cast<const u64*>(out + 2 * (threadIdx.x & 1));
    const float onef = out[threadIdx.x & 1];
    s = out[2 + (threadIdx.x & 1)];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (PACK) { if (i < NF / 2) A[i % 8] = ffma2(A[i % 8], s, one2); }
            else if (i < NF) a[i] = ffma1(a[i], s, onef);
            if (i < NA) m[i] = fmin1(m[i], s);
            if (PACK) { if (i < ND / 2) D[i % 8] = fadd2(D[i % 8], one2); }
            else if (i < ND) d[i] = fadd1(d[i], s);
        }
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) r += a[i] + m[i] + d[i];
    for (int i = 0; i < 8; ++i) { float x, y; upk(A[i], x, y); r += x + y; upk(D[i], x, y); r += x + y; }
    if (r == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <bool PACK, int NF, int NA, int ND> void run(float* d) {
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<PACK, NF, NA, ND><<<148, 512>>>(d, 100, 0.999f);
    cudaEventRecord(e0);
    k<PACK, NF, NA, ND><<<148, 512>>>(d, iters, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const int slots = PACK ? NF / 2 + NA + ND / 2 : NF + NA + ND;
    printf("%s fma=%2d alu=%2d fadd=%2d | issue slots/iter/warp %2d | measured cycles/iter/warp = %.1f (4 warps per SMSP, 1.965 GHz assumed)\n",
           PACK ? "packed" : "scalar", NF, NA, ND, slots, ms * 1e-3 * 1.965e9 / iters / 4);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 512 * 4); { float h[4] = {1.f, 1.f, 0.999f, 0.999f}; cudaMemcpy(d, h, 16, cudaMemcpyHostToDevice); }
    run<false, 16, 0, 0>(d);  run<true, 16, 0, 0>(d);
    run<false, 16, 8, 0>(d);  run<true, 16, 8, 0>(d);
    run<false, 16, 16, 0>(d); run<true, 16, 16, 0>(d);
    run<false, 16, 4, 0>(d);  run<true, 16, 4, 0>(d);
    run<false, 16, 8, 8>(d);  run<true, 16, 8, 8>(d);
    run<false, 16, 8, 16>(d); run<true, 16, 8, 16>(d);
    run<false, 0, 0, 16>(d);  run<true, 0, 0, 16>(d);
    run<false, 0, 16, 0>(d);
    return 0;
}
