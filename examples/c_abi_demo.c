/* Minimal C consumer of the cspn_b200 C ABI (include/cspn_b200.h): what a non-Python front-end (a C++ inference
 * server, a Paddle custom op for cspn_paddle/demo.py's affinity_propagate) links against.
 *
 *   gcc -I include examples/c_abi_demo.c -o c_abi_demo -ldl && ./c_abi_demo cspn_b200/_build/libcspn_b200.so
 *
 * Without a GPU it exercises everything that needs none (version, planning, argument checking); with one it also
 * runs a small propagation through the host-buffer entry point. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cspn_b200.h"

#define LOAD(name)                                                      \
    *(void**)(&p_##name) = dlsym(lib, #name);                           \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "cspn_b200/_build/libcspn_b200.so";
    void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    int (*p_cspn_version)(void);
    const char* (*p_cspn_last_error)(void);
    size_t (*p_cspn2d_workspace_bytes)(int, int, int, int, int, int);
    int (*p_cspn2d_describe_plan)(int, int, int, int, int, int, char*, int);
    int (*p_cspn2d_fwd_f32)(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, void*,
                            size_t, cspn_stream_t);
    int (*p_cspn2d_fwd_f32_host)(const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int);
    LOAD(cspn_version) LOAD(cspn_last_error) LOAD(cspn2d_workspace_bytes) LOAD(cspn2d_describe_plan) LOAD(cspn2d_fwd_f32)
    LOAD(cspn2d_fwd_f32_host)

    printf("cspn_b200 version %d\n", p_cspn_version());
    char plan[1024];
    p_cspn2d_describe_plan(32, 1, 352, 1216, 24, CSPN_ALGO_AUTO, plan, (int)sizeof(plan));
    printf("plan for 32x1x352x1216, 24 steps: %s\n", plan);
    printf("workspace: %zu bytes (24 steps), %zu bytes (48 steps: multi-pass)\n",
           p_cspn2d_workspace_bytes(32, 1, 352, 1216, 24, CSPN_ALGO_AUTO), p_cspn2d_workspace_bytes(32, 1, 352, 1216, 48, CSPN_ALGO_AUTO));

    /* argument checking mirrors the reference's assertions (cspn.py:33,36,91-98) and needs no device */
    float dummy[4] = {0, 0, 0, 0};
    int rc = p_cspn2d_fwd_f32(dummy, dummy, NULL, dummy, 1, 1, 2, 2, /*guidance_channels=*/7, 3, CSPN_NORM_8SUM, CSPN_ALGO_AUTO, NULL, 0, NULL);
    printf("7 guidance channels -> status %d (%s)\n", rc, p_cspn_last_error());
    if (rc != CSPN_ERR_INVALID_ARGUMENT) return 1;
    rc = p_cspn2d_fwd_f32(dummy, dummy, NULL, dummy, 1, 1, 2, 2, 8, 3, /*norm_type=*/5, CSPN_ALGO_AUTO, NULL, 0, NULL);
    printf("unknown norm_type   -> status %d (%s)\n", rc, p_cspn_last_error());
    if (rc != CSPN_ERR_INVALID_ARGUMENT) return 1;

    if (argc > 2 && strcmp(argv[2], "--run") == 0) {   /* needs a GPU: constant depth under '8sum_abs' stays constant */
        enum { H = 16, W = 32, N = 6 };
        float* g = malloc(sizeof(float) * 8 * H * W);
        float* d = malloc(sizeof(float) * H * W);
        float* o = malloc(sizeof(float) * H * W);
        for (int i = 0; i < 8 * H * W; ++i) g[i] = (float)((i * 2654435761u) % 1000) / 1000.f + 0.1f;
        for (int i = 0; i < H * W; ++i) d[i] = 2.5f;
        rc = p_cspn2d_fwd_f32_host(g, d, NULL, o, 1, 1, H, W, 8, N, CSPN_NORM_8SUM_ABS, CSPN_ALGO_AUTO, 0);
        if (rc != CSPN_OK) { fprintf(stderr, "forward failed: %s\n", p_cspn_last_error()); return 1; }
        /* interior pixels: the weights sum to 1 and the centre term vanishes, so 2.5 propagates to 2.5 */
        printf("out[H/2][W/2] = %f (expected 2.5)\n", o[(H / 2) * W + W / 2]);
        free(g); free(d); free(o);
    }
    dlclose(lib);
    return 0;
}
