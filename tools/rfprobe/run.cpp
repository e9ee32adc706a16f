// Loads patched cubins through the driver API and reports issue cycles per instruction and sub-partition.
//   run <kernel> <threads> <iters> <ni> file.cubin [file.cubin ...]
#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <vector>
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); fprintf(stderr, "%s -> %s\n", #x, s_); exit(1); } } while (0)
int main(int argc, char** argv) {
    if (argc < 6) return 1;
    const char* kname = argv[1];
    int threads = atoi(argv[2]), iters = atoi(argv[3]), ni = atoi(argv[4]);
    CK(cuInit(0));
    CUdevice dev; CK(cuDeviceGet(&dev, 0));
    CUcontext ctx; CK(cuCtxCreate(&ctx, 0, dev));
    CUdeviceptr out, in;
    CK(cuMemAlloc(&out, 1 << 20)); CK(cuMemAlloc(&in, 1 << 20));
    CK(cuMemsetD8(in, 0, 1 << 20));
    for (int f = 5; f < argc; ++f) {
        std::ifstream is(argv[f], std::ios::binary);
        std::vector<char> img((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
        CUmodule mod; CK(cuModuleLoadData(&mod, img.data()));
        CUfunction fn; CK(cuModuleGetFunction(&fn, mod, kname));
        void* args[3] = {&out, &in, &iters};
        for (int rep = 0; rep < 2; ++rep) CK(cuLaunchKernel(fn, 148, 1, 1, threads, 1, 1, 0, 0, args, nullptr));
        CK(cuCtxSynchronize());
        std::vector<long long> h(148);
        CK(cuMemcpyDtoH(h.data(), out, 148 * 8));
        std::sort(h.begin(), h.end());
        const double wps = threads / 128.0;
        printf("%-44s thr=%4d  %.3f cycles/instr/SMSP (median SM; min %.3f max %.3f)  %.2f cyc/instr/warp\n", argv[f], threads,
               h[74] / ((double)iters * ni * wps), h[0] / ((double)iters * ni * wps), h[147] / ((double)iters * ni * wps), h[74] / ((double)iters * ni));
        CK(cuModuleUnload(mod));
    }
    return 0;
}
