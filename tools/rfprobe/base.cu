// Base kernels for the register-file probe (tools/rfprobe): a loop of NI independent-looking FFMA2 / FFMA instructions whose
// REGISTER FIELDS, stall counts and reuse flags are rewritten afterwards in the cubin by patch.py -- fixed-latency pipes have
// no interlock, so a patched stream issues at whatever rate the register file allows, whatever its data dependences say.
// out[block] = clock64 ticks warp 0 spent in the loop.
#include <cuda_runtime.h>
typedef unsigned long long u64;
#ifndef NI
#define NI 96
#endif
extern "C" __global__ void __launch_bounds__(256, 1) k_ffma2(long long* out, const float* in, int iters) {
    u64 w[48], x[12], a[16];
#pragma unroll
    for (int i = 0; i < 48; ++i) asm volatile("ld.global.b64 %0, [%1];" : "=l"(w[i]) : "l"(in + 2 * (i + threadIdx.x)));
#pragma unroll
    for (int i = 0; i < 12; ++i) asm volatile("ld.global.b64 %0, [%1];" : "=l"(x[i]) : "l"(in + 800 + 2 * (i + threadIdx.x)));
#pragma unroll
    for (int i = 0; i < 16; ++i) asm volatile("ld.global.b64 %0, [%1];" : "=l"(a[i]) : "l"(in + 1600 + 2 * (i + threadIdx.x)));
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NI; ++j)
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a[j % 16]) : "l"(w[(j * 7) % 48]), "l"(x[(j * 5) % 12]));
    }
    const long long t1 = clock64();
    u64 s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= a[i];
    if (s == 0x1234567u) out[1000 + threadIdx.x] = (long long)s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
extern "C" __global__ void __launch_bounds__(256, 1) k_ffma(long long* out, const float* in, int iters) {
    float w[96], x[24], a[32];
#pragma unroll
    for (int i = 0; i < 96; ++i) asm volatile("ld.global.f32 %0, [%1];" : "=f"(w[i]) : "l"(in + i + threadIdx.x));
#pragma unroll
    for (int i = 0; i < 24; ++i) asm volatile("ld.global.f32 %0, [%1];" : "=f"(x[i]) : "l"(in + 200 + i + threadIdx.x));
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("ld.global.f32 %0, [%1];" : "=f"(a[i]) : "l"(in + 300 + i + threadIdx.x));
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NI; ++j)
            asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[j % 32]) : "f"(w[(j * 7) % 96]), "f"(x[(j * 5) % 24]));
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += a[i];
    if (s == 1.2345f) out[1000 + threadIdx.x] = (long long)s;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
