// Standalone TMA probe: which (tensor, box, coords) combinations does cp.async.bulk.tensor.3d accept on this box?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void k3d(const __grid_constant__ CUtensorMap tm, float* out, int bx, int by, int x, int y, int z) {
    extern __shared__ __align__(1024) unsigned char sm[];
    float* dst = reinterpret_cast<float*>(sm);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + (size_t)bx * by * 4);
    uint32_t b32 = (uint32_t)__cvta_generic_to_shared(bar), d32 = (uint32_t)__cvta_generic_to_shared(dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b32) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(bx * by * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d32),
                     "l"(&tm), "r"(x), "r"(y), "r"(z), "r"(b32) : "memory");
    }
    __syncthreads();
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b32) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < bx * by; i += blockDim.x) out[i] = dst[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    if (argc < 10) { printf("usage: W H P bx by x y z l2promo\n"); return 2; }
    int W = atoi(argv[1]), H = atoi(argv[2]), P = atoi(argv[3]), bx = atoi(argv[4]), by = atoi(argv[5]);
    int x = atoi(argv[6]), y = atoi(argv[7]), z = atoi(argv[8]), l2 = atoi(argv[9]);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    size_t n = (size_t)W * H * P;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 100003) + 1.f;
    float *g, *o; cudaMalloc(&g, n * 4); cudaMalloc(&o, (size_t)bx * by * 4);
    cudaMemcpy(g, h.data(), n * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P};
    cuuint64_t str[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode=%d ", (int)r);
    size_t smem = (size_t)bx * by * 4 + 64;
    cudaFuncSetAttribute(k3d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k3d<<<1, 128, smem>>>(tm, o, bx, by, x, y, z);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> ho((size_t)bx * by);
        cudaMemcpy(ho.data(), o, ho.size() * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int j = 0; j < by; ++j)
            for (int i = 0; i < bx; ++i) {
                int xx = x + i, yy = y + j;
                float exp = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? h[((size_t)z * H + yy) * W + xx] : 0.f;
                if (ho[(size_t)j * bx + i] != exp) ++bad;
            }
        printf("mismatches=%ld", bad);
    }
    printf("  (W=%d H=%d P=%d box=%dx%d at %d,%d,%d l2=%d)\n", W, H, P, bx, by, x, y, z, l2);
    return 0;
}
