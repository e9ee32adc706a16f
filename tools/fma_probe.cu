// FFMA vs FFMA2 (fma.rn.f32x2) issue/throughput probe at 2 and 4 warps per SMSP, register-heavy like the CSPN kernel.
#include <cuda_runtime.h>
#include <cstdio>

template <int NACC>
__global__ void __launch_bounds__(1024, 1) k_ffma(float* out, const float* in, int iters) {
    float w[NACC], a[NACC], e[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { w[i] = in[i]; a[i] = in[NACC + i]; e[i] = in[2 * NACC + i] + threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int i = 0; i < NACC; ++i) a[i] = fmaf(w[(i + k) % NACC], e[(i + 3 * k + 1) % NACC], a[i]);
#pragma unroll
        for (int i = 0; i < NACC; ++i) e[i] = a[i] * 0.5f;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>  // NACC pairs
__global__ void __launch_bounds__(1024, 1) k_ffma2(float* out, const float* in, int iters) {
    float2 w[NACC], a[NACC], e[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        w[i] = make_float2(in[i], in[i + 1]); a[i] = make_float2(in[NACC + i], in[NACC + i + 1]);
        e[i] = make_float2(in[2 * NACC + i] + threadIdx.x, in[2 * NACC + i + 1]);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int i = 0; i < NACC; ++i) a[i] = __ffma2_rn(w[(i + k) % NACC], e[(i + 3 * k + 1) % NACC], a[i]);
#pragma unroll
        for (int i = 0; i < NACC; ++i) e[i] = make_float2(a[i].x * 0.5f, a[i].y * 0.5f);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
void run(const char* name, F kern, int threads, double fma_per_thread_iter, float* out, float* in) {
    int iters = 4000;
    kern<<<148, threads>>>(out, in, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<148, threads>>>(out, in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = 148.0 * threads * fma_per_thread_iter * iters;
    printf("%-28s threads=%4d: %.3f ms  %.2f TFMA/s (%.1f%% of 148*128*1.965e9)  err=%s\n", name, threads, ms, fma / ms / 1e9,
           100.0 * fma / (ms * 1e-3) / (148.0 * 128 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *out, *in;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&in, 4096);
    cudaMemset(in, 0, 4096);
    for (int thr : {128, 256, 512, 1024}) {
        run("FFMA  16 acc", k_ffma<16>, thr, 16 * 8, out, in);
        run("FFMA2 8 pair (16 acc)", k_ffma2<8>, thr, 16 * 8, out, in);
        run("FFMA2 16 pair (32 acc)", k_ffma2<16>, thr, 32 * 8, out, in);
    }
    return 0;
}
