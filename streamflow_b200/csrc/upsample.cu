// convex 8x flow upsampling (SKFlow_MF8.upsample_flow, core/models/streamflow.py:82-93):
//   out[n, c, 8y+i, 8x+j] = sum_k softmax_k(mask[n, k*64 + i*8 + j, y, x]) * 8 * flow[n, c, y + k/3 - 1, x + k%3 - 1]
// (3x3 neighbourhood, zero padded, softmax over the 9 neighbours).  The reference materialises the softmax, the
// unfolded flow and their product (5 tensor passes); this is one pass: 576 mask values in, 128 flow values out per
// coarse pixel.  Thread = (coarse pixel, sub-row i): 72 independent coalesced mask loads (lanes run along x), fp32
// softmax, and 8 consecutive outputs per channel so a warp writes 1 KB contiguous.
#include <cuda_bf16.h>

#include "sf_internal.h"

namespace sf {

namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p) {
    return static_cast<float>(*p);
}
template <>
__device__ __forceinline__ float ldf<__half>(const __half* p) {
    return __half2float(*p);
}
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}

// block = (32 x-positions, 8 sub-rows); grid = (ceil(W/32), H, N)
template <typename T>
__global__ void __launch_bounds__(256) upsample_flow_kernel(const float* __restrict__ flow, const T* __restrict__ mask,
                                                            float* __restrict__ out, int H, int W) {
    pdl_launch();
    pdl_wait();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y, n = blockIdx.z, i = threadIdx.y;
    if (x >= W) return;
    const long long hw = static_cast<long long>(H) * W;
    const T* m = mask + static_cast<long long>(n) * 576 * hw + static_cast<long long>(y) * W + x;
    const float* f = flow + static_cast<long long>(n) * 2 * hw;
    float fx[9], fy[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
        fx[k] = ok ? 8.0f * __ldg(f + static_cast<long long>(yy) * W + xx) : 0.f;
        fy[k] = ok ? 8.0f * __ldg(f + hw + static_cast<long long>(yy) * W + xx) : 0.f;
    }
    float ox[8], oy[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v[9], mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            v[k] = ldf<T>(m + static_cast<long long>(k * 64 + i * 8 + j) * hw);
            mx = fmaxf(mx, v[k]);
        }
        float s = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float e = __expf(v[k] - mx);
            s += e;
            ax = fmaf(e, fx[k], ax);
            ay = fmaf(e, fy[k], ay);
        }
        const float r = 1.0f / s;
        ox[j] = ax * r;
        oy[j] = ay * r;
    }
    const long long HW8 = 64 * hw;                         // (8H) * (8W)
    float* o = out + static_cast<long long>(n) * 2 * HW8 + static_cast<long long>(8 * y + i) * (8 * W) + 8 * x;
    *reinterpret_cast<float4*>(o) = make_float4(ox[0], ox[1], ox[2], ox[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(ox[4], ox[5], ox[6], ox[7]);
    *reinterpret_cast<float4*>(o + HW8) = make_float4(oy[0], oy[1], oy[2], oy[3]);
    *reinterpret_cast<float4*>(o + HW8 + 4) = make_float4(oy[4], oy[5], oy[6], oy[7]);
}

}  // namespace

int launch_upsample_flow(const float* flow, const void* mask, int mask_dtype, float* out, int64_t N, int64_t H,
                         int64_t W, cudaStream_t s) {
    dim3 grid(static_cast<unsigned>((W + 31) / 32), static_cast<unsigned>(H), static_cast<unsigned>(N));
    dim3 block(32, 8);
    prof_before(SF_KERNEL_UPSAMPLE, s);
    switch (mask_dtype) {
        case SF_DT_F32:
            SF_CUDA_CHECK(launch_kernel(upsample_flow_kernel<float>, grid, block, 0, s, flow,
                                        static_cast<const float*>(mask), out, (int)H, (int)W));
            break;
        case SF_DT_F16:
            SF_CUDA_CHECK(launch_kernel(upsample_flow_kernel<__half>, grid, block, 0, s, flow,
                                        static_cast<const __half*>(mask), out, (int)H, (int)W));
            break;
        case SF_DT_BF16:
            SF_CUDA_CHECK(launch_kernel(upsample_flow_kernel<__nv_bfloat16>, grid, block, 0, s, flow,
                                        static_cast<const __nv_bfloat16*>(mask), out, (int)H, (int)W));
            break;
        default: set_error("upsample_flow: unsupported mask dtype %d", mask_dtype); return SF_ERR_INVALID;
    }
    prof_after(SF_KERNEL_UPSAMPLE, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
