// FFMA / FFMA2 throughput vs number of independent dependent-chains per thread (ILP), at 1 and 2 warps per SMSP.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int ILP> __global__ void k1(float* out, int iters, float w0) {
    float a[ILP], w[8];
    for (int i = 0; i < 8; ++i) w[i] = w0 + i;
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = ffma1(w[k], a[(i + 1) % ILP == i ? i : i], a[i]);
    float s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k2(float* out, int iters, float w0) {
    u64 a[ILP], w[8];
    for (int i = 0; i < 8; ++i) { float x = w0 + i; asm("mov.b64 %0, {%1,%1};" : "=l"(w[i]) : "f"(x)); }
    for (int i = 0; i < ILP; ++i) { float x = threadIdx.x + i; asm("mov.b64 %0, {%1,%1};" : "=l"(a[i]) : "f"(x)); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = ffma2(w[k], a[i], a[i]);
    u64 s = 0; for (int i = 0; i < ILP; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
}
template <typename F> void run(const char* n, F f, int thr, int ilp, int lanes_per_instr, float* out) {
    int iters = 20000;
    f<<<148, thr>>>(out, 100, 0.5f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); f<<<148, thr>>>(out, iters, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr_per_warp = (double)iters * 8 * ilp;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%s ILP=%2d thr=%4d: %.1f cyc per dependent step (chain latency bound), %.3f warp-instr/clk/SMSP, %.1f TFMA/s\n", n, ilp, thr,
           cyc / (iters * 8.0), instr_per_warp * (thr / 32 / 4.0) / cyc, 148.0 * thr * instr_per_warp * lanes_per_instr / 32.0 / ms / 1e9 * 32 / 32);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    for (int thr : {128, 256}) {
        run("FFMA ", k1<1>, thr, 1, 1, out); run("FFMA ", k1<2>, thr, 2, 1, out); run("FFMA ", k1<4>, thr, 4, 1, out); run("FFMA ", k1<8>, thr, 8, 1, out); run("FFMA ", k1<16>, thr, 16, 1, out);
        run("FFMA2", k2<1>, thr, 1, 2, out); run("FFMA2", k2<2>, thr, 2, 2, out); run("FFMA2", k2<4>, thr, 4, 2, out); run("FFMA2", k2<8>, thr, 8, 2, out); run("FFMA2", k2<16>, thr, 16, 2, out);
    }
    return 0;
}
