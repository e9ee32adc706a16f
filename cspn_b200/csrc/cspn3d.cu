// 3D CSPN (26 neighbours): prep (normalisation) + one 27-point stencil launch per iteration.
//
// Reference: call sites only -- /root/reference/cspn_paddle/demo.py:20-54; the Paddle op's
// source is not in the tree, so the arithmetic is the definition of SURVEY.md Appendix A.3
// (PARITY UNPINNED, see oracle/cspn_numpy.py).  Per-voxel state is 26 weights + kappa + value
// (112 B of input per voxel); with N = 12 and a 3D halo no on-chip temporal blocking fits an
// SM's 64K registers, so this path streams the normalised weights once per iteration from
// L2/HBM, one volume at a time (the 27*D*H*W workspace of one volume is reused for the batch).
#include "common.cuh"

namespace cspn {

namespace {

// mode 0 '26sum', 1 '26sum_abs': gathered + centre term.  mode 2 'paddle': own location, no centre.
__global__ void __launch_bounds__(256)
prep3d_kernel(const float* __restrict__ g, float* __restrict__ wk, int D, int H, int W, int mode) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W, V = (size_t)D * HW;
    const size_t p = (size_t)z * HW + (size_t)y * W + x;
    float a[26], S = 0.f;
#pragma unroll
    for (int k = 0; k < 26; ++k) {
        float v;
        if (mode == 2) {
            v = fabsf(__ldg(g + k * V + p));
        } else {
            const int zz = z + off3_dz(k), yy = y + off3_dy(k), xx = x + off3_dx(k);
            v = 0.f;
            if (zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W) {
                v = __ldg(g + k * V + (size_t)zz * HW + (size_t)yy * W + xx);
                if (mode == 1) v = fabsf(v);
            }
        }
        a[k] = v;
        S += fabsf(v);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 26; ++k) {
        const float w = __fdiv_rn(a[k], S);
        s += w;
        wk[k * V + p] = w;
    }
    wk[26 * V + p] = (mode == 2) ? 0.f : 1.f - s;
}

__global__ void __launch_bounds__(256)
step3d_kernel(const float* __restrict__ wk, const float* __restrict__ d0, const float* __restrict__ cur,
              float* __restrict__ dst, int D, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z % D, c = blockIdx.z / D;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W, V = (size_t)D * HW;
    const size_t p = (size_t)z * HW + (size_t)y * W + x;
    const float* cc = cur + (size_t)c * V;
    float acc = __ldg(wk + 26 * V + p) * __ldg(d0 + (size_t)c * V + p);
#pragma unroll
    for (int k = 0; k < 26; ++k) {
        const int zz = z + off3_dz(k), yy = y + off3_dy(k), xx = x + off3_dx(k);
        const float dv = (zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W)
                             ? __ldg(cc + (size_t)zz * HW + (size_t)yy * W + xx) : 0.f;
        acc = fmaf(__ldg(wk + k * V + p), dv, acc);
    }
    dst[(size_t)c * V + p] = acc;
}

// Vectorised step for W % 4 == 0: one thread produces 4 consecutive voxels of a row.  The 27 weight planes stream
// through as LDG.128 (they are the HBM traffic of this path: 108 B per voxel and step), the 9 neighbouring rows of the
// current volume are read as one aligned float4 plus two edge scalars each (L1/L2 hits).
__global__ void __launch_bounds__(128)
step3d_vec4_kernel(const float* __restrict__ wk, const float* __restrict__ d0, const float* __restrict__ cur,
                   float* __restrict__ dst, int D, int H, int W) {
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z % D, c = blockIdx.z / D;
    if (x >= W || y >= H) return;
    const size_t HW = (size_t)H * W, V = (size_t)D * HW;
    const size_t p = (size_t)z * HW + (size_t)y * W + x;
    const float* cc = cur + (size_t)c * V;
    // rows[dz+1][dy+1][0..5] = cur(z+dz, y+dy, x-1 .. x+4), zero outside the volume
    float rows[3][3][6];
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int zz = z + dz, yy = y + dy;
            float* r = rows[dz + 1][dy + 1];
            if (zz >= 0 && zz < D && yy >= 0 && yy < H) {
                const float* src = cc + (size_t)zz * HW + (size_t)yy * W + x;
                const float4 v = __ldg(reinterpret_cast<const float4*>(src));
                r[1] = v.x; r[2] = v.y; r[3] = v.z; r[4] = v.w;
                r[0] = x > 0 ? __ldg(src - 1) : 0.f;
                r[5] = x + 4 < W ? __ldg(src + 4) : 0.f;
            } else {
#pragma unroll
                for (int i = 0; i < 6; ++i) r[i] = 0.f;
            }
        }
    const float4 kap = __ldg(reinterpret_cast<const float4*>(wk + 26 * V + p));
    const float4 dz0 = __ldg(reinterpret_cast<const float4*>(d0 + (size_t)c * V + p));
    float acc[4] = {kap.x * dz0.x, kap.y * dz0.y, kap.z * dz0.z, kap.w * dz0.w};
#pragma unroll
    for (int k = 0; k < 26; ++k) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wk + k * V + p));
        const float* r = rows[off3_dz(k) + 1][off3_dy(k) + 1];
        const int o = 1 + off3_dx(k);
        acc[0] = fmaf(w.x, r[o], acc[0]);
        acc[1] = fmaf(w.y, r[o + 1], acc[1]);
        acc[2] = fmaf(w.z, r[o + 2], acc[2]);
        acc[3] = fmaf(w.w, r[o + 3], acc[3]);
    }
    *reinterpret_cast<float4*>(dst + (size_t)c * V + p) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

}  // namespace

size_t generic3d_workspace_bytes(int B, int C, int D, int H, int W, int iters) {
    if (iters <= 0) return 0;
    const size_t V = (size_t)D * H * W;
    return sizeof(float) * (27 * V + (iters > 1 ? (size_t)C * V : 0));
}

int generic3d_forward(const float* guidance, const float* feat, float* out, int B, int C, int D, int H, int W,
                      int iters, int mode, void* ws, size_t ws_bytes, cudaStream_t stream, int* launches) {
    const size_t V = (size_t)D * H * W;
    if (iters == 0) {
        CSPN_CUDA_TRY(cudaMemcpyAsync(out, feat, (size_t)B * C * V * sizeof(float), cudaMemcpyDeviceToDevice,
                                      stream));
        return CSPN_OK;
    }
    const size_t need = generic3d_workspace_bytes(B, C, D, H, W, iters);
    if (!ws || ws_bytes < need) {
        set_error("3D path needs %zu workspace bytes, got %zu", need, ws ? ws_bytes : (size_t)0);
        return CSPN_ERR_WORKSPACE;
    }
    if ((size_t)D * C > 65535 || D > 65535) {
        set_error("3D path: D*C=%zu exceeds gridDim.z", (size_t)D * C);
        return CSPN_ERR_UNSUPPORTED;
    }
    float* wk = static_cast<float*>(ws);
    float* tmp = wk + 27 * V;
    const dim3 block(32, 8);
    const dim3 g2((W + 31) / 32, (H + 7) / 8);
    for (int b = 0; b < B; ++b) {
        const float* d0 = feat + (size_t)b * C * V;
        float* o = out + (size_t)b * C * V;
        prep3d_kernel<<<dim3(g2.x, g2.y, D), block, 0, stream>>>(guidance + (size_t)b * 26 * V, wk, D, H, W, mode);
        ++*launches;
        const float* cur = d0;
        float* dst = (iters & 1) ? o : tmp;
        const bool vec4 = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) | reinterpret_cast<uintptr_t>(out) |
                                              reinterpret_cast<uintptr_t>(ws)) % 16 == 0);
        for (int it = 0; it < iters; ++it) {
            if (vec4)
                step3d_vec4_kernel<<<dim3((W / 4 + 31) / 32, (H + 3) / 4, D * C), dim3(32, 4), 0, stream>>>(wk, d0, cur, dst, D, H, W);
            else
                step3d_kernel<<<dim3(g2.x, g2.y, D * C), block, 0, stream>>>(wk, d0, cur, dst, D, H, W);
            ++*launches;
            cur = dst;
            dst = (dst == o) ? tmp : o;
        }
    }
    CSPN_CUDA_TRY(cudaGetLastError());
    return CSPN_OK;
}

}  // namespace cspn
