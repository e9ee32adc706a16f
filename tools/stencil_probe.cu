// Register-resident 3x3 stencil inner loop in isolation (no exchange, no shuffles): FFMA vs FFMA2, to find what the
// FMA pipe sustains with 80 distinct weight pairs / 160 weights per thread at 8 warps (256 threads, 1 CTA/SM).
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mk(float x, float y) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }

constexpr int PR = 5, PCH = 2;
__host__ __device__ constexpr int tap_of(int dy, int dx) { return dy == 1 ? 1 - dx : (dy == 0 ? (dx == 1 ? 3 : 4) : 6 - dx); }

template <bool SCATTER>
__global__ void __launch_bounds__(256, 1) k_pair(float* out, const float* in, int iters) {
    u64 w[PR][PCH][8], d[PR + 2][PCH + 2], n[PR][PCH];
    for (int r = 0; r < PR; ++r) for (int j = 0; j < PCH; ++j) for (int k = 0; k < 8; ++k) w[r][j][k] = mk(in[(r * 2 + j) * 8 + k], in[k + 1]);
    for (int r = 0; r < PR + 2; ++r) for (int j = 0; j < PCH + 2; ++j) d[r][j] = mk(in[r * 4 + j] + threadIdx.x, in[j]);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < PR; ++r)
#pragma unroll
            for (int j = 0; j < PCH; ++j) n[r][j] = d[r + 1][j + 1];
        if (SCATTER) {
#pragma unroll
            for (int rs = 0; rs < PR + 2; ++rs)
#pragma unroll
                for (int jx = 0; jx < PCH + 2; ++jx)
#pragma unroll
                    for (int dy = 1; dy >= -1; --dy)
#pragma unroll
                        for (int dx = 1; dx >= -1; --dx) {
                            const int r = rs - 1 - dy, j = jx - 1 - dx;
                            if (r < 0 || r >= PR || j < 0 || j >= PCH || (dy == 0 && dx == 0)) continue;
                            n[r][j] = ffma2(w[r][j][tap_of(dy, dx)], d[rs][jx], n[r][j]);
                        }
        } else {
#pragma unroll
            for (int r = 0; r < PR; ++r)
#pragma unroll
                for (int j = 0; j < PCH; ++j)
#pragma unroll
                    for (int dy = 1; dy >= -1; --dy)
#pragma unroll
                        for (int dx = 1; dx >= -1; --dx) {
                            if (dy == 0 && dx == 0) continue;
                            n[r][j] = ffma2(w[r][j][tap_of(dy, dx)], d[r + 1 + dy][j + 1 + dx], n[r][j]);
                        }
        }
#pragma unroll
        for (int r = 0; r < PR; ++r)
#pragma unroll
            for (int j = 0; j < PCH; ++j) d[r + 1][j + 1] = n[r][j];
    }
    u64 s = 0;
    for (int r = 0; r < PR; ++r) for (int j = 0; j < PCH; ++j) s ^= d[r + 1][j + 1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
}

constexpr int SPR = 5, SPC = 4;
template <bool SCATTER>
__global__ void __launch_bounds__(256, 1) k_scalar(float* out, const float* in, int iters) {
    float w[SPR][SPC][8], d[SPR + 2][SPC + 2], n[SPR][SPC];
    for (int r = 0; r < SPR; ++r) for (int j = 0; j < SPC; ++j) for (int k = 0; k < 8; ++k) w[r][j][k] = in[(r * 4 + j) * 8 + k];
    for (int r = 0; r < SPR + 2; ++r) for (int j = 0; j < SPC + 2; ++j) d[r][j] = in[r * 6 + j] + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < SPR; ++r)
#pragma unroll
            for (int j = 0; j < SPC; ++j) n[r][j] = d[r + 1][j + 1];
        if (SCATTER) {
#pragma unroll
            for (int rs = 0; rs < SPR + 2; ++rs)
#pragma unroll
                for (int jx = 0; jx < SPC + 2; ++jx)
#pragma unroll
                    for (int dy = 1; dy >= -1; --dy)
#pragma unroll
                        for (int dx = 1; dx >= -1; --dx) {
                            const int r = rs - 1 - dy, j = jx - 1 - dx;
                            if (r < 0 || r >= SPR || j < 0 || j >= SPC || (dy == 0 && dx == 0)) continue;
                            n[r][j] = fmaf(w[r][j][tap_of(dy, dx)], d[rs][jx], n[r][j]);
                        }
        } else {
#pragma unroll
            for (int r = 0; r < SPR; ++r)
#pragma unroll
                for (int j = 0; j < SPC; ++j)
#pragma unroll
                    for (int dy = 1; dy >= -1; --dy)
#pragma unroll
                        for (int dx = 1; dx >= -1; --dx) {
                            if (dy == 0 && dx == 0) continue;
                            n[r][j] = fmaf(w[r][j][tap_of(dy, dx)], d[r + 1 + dy][j + 1 + dx], n[r][j]);
                        }
        }
#pragma unroll
        for (int r = 0; r < SPR; ++r)
#pragma unroll
            for (int j = 0; j < SPC; ++j) d[r + 1][j + 1] = n[r][j];
    }
    float s = 0;
    for (int r = 0; r < SPR; ++r) for (int j = 0; j < SPC; ++j) s += d[r + 1][j + 1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> void run(const char* n, F f, float* out, float* in) {
    const int iters = 2000;
    f<<<148, 256>>>(out, in, 10);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); f<<<148, 256>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double cyc = ms * 1e-3 * 1.965e9 / iters;
    printf("%-22s %.0f cycles per stencil iteration (20 px/thread, 2 warps/SMSP) -> %.2f px-iter/clk/SM   %s\n", n, cyc, 256.0 * 20 / cyc,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *out, *in; cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    run("FFMA2 gather order", k_pair<false>, out, in);
    run("FFMA2 scatter order", k_pair<true>, out, in);
    run("FFMA  gather order", k_scalar<false>, out, in);
    run("FFMA  scatter order", k_scalar<true>, out, in);
    return 0;
}
