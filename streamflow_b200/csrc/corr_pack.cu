// Operand preparation for the correlation GEMM (part of CorrBlock.__init__, core/corr.py:7-21).
//
//  absmax2_kernel   per-tensor |max| of fmap1 / fmap2 (device-side, no host sync) -> power-of-two scale.
//  corr_pack_kernel reads the fp32 feature maps through their ORIGINAL strides (the model hands over
//                   channels-last views, core/models/streamflow.py:107,110) and writes K-major fp16 operand
//                   matrices the TMA can tile:  A = fmap1 as [B, N, Kp];  B_l = avg-pooled fmap2 at level l
//                   as [B, th_l*tw_l*16, Kp] in 4x4-tiled cell order (pad cells are zero rows).  Pooling the operand instead of
//                   the volume uses linearity: avg_pool(f1^T f2) == f1^T avg_pool(f2) (core/corr.py:19-21).
//                   With split != 0 each value is stored as hi/lo fp16 parts concatenated along K so that one
//                   GEMM over Kp = 3*D accumulates  hi*hi + hi*lo + lo*hi  (fp32-faithful mode).
#include <algorithm>
#include <utility>

#include "sf_internal.h"

namespace sf {

namespace {

struct Strided4 {
    const float* p;
    long long sb, sk, sy, sx;
};

__global__ void absmax2_kernel(Strided4 t0, Strided4 t1, int D, int h, int w, long long per_batch,
                               long long total, unsigned* out_bits) {
    const Strided4 t = blockIdx.y == 0 ? t0 : t1;
    float m = 0.f;
    unsigned low = 0u;       // OR of the 13 mantissa bits an fp16 cannot hold: 0 <=> every value is fp16-representable
    const int hw = h * w;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / per_batch;
        const int rem = static_cast<int>(i - b * per_batch);
        // enumerate in an order that is contiguous for both NCHW (sx == 1) and channels-last (sk == 1)
        int k, y, x;
        if (t.sk == 1) {
            k = rem % D;
            const int n = rem / D;
            y = n / w;
            x = n - y * w;
        } else {
            k = rem / hw;
            const int n = rem - k * hw;
            y = n / w;
            x = n - y * w;
        }
        const float v = __ldg(t.p + b * t.sb + k * t.sk + y * t.sy + x * t.sx);
        m = fmaxf(m, fabsf(v));
        low |= __float_as_uint(v) & 0x1FFFu;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
        low |= __shfl_xor_sync(0xffffffffu, low, s);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out_bits + blockIdx.y, __float_as_uint(m));
        if (low) atomicOr(out_bits + 2, low);
    }
}

// dense per-batch blocks (NCHW-contiguous or channels-last): plain vectorised sweep of the storage.  grid = (blocks, 2 operands,
// B batches): every thread issues ALL its 16-byte loads before it reduces (the first version walked the batches one after the other
// with four loads in flight: 14 us for the 29 MB of a Sintel clip), one atomic pair per CTA.
__global__ void __launch_bounds__(256) absmax2_flat_kernel(const float* f0, const float* f1, long long sb0, long long sb1, int B,
                                                           long long per_batch, unsigned* out_bits) {
    __shared__ float s_m[8];
    __shared__ unsigned s_low[8];
    pdl_launch();
    pdl_wait();
    const float* f = blockIdx.y == 0 ? f0 : f1;
    const long long sb = blockIdx.y == 0 ? sb0 : sb1;
    const long long n4 = per_batch >> 2;
    const float4* v = reinterpret_cast<const float4*>(f + static_cast<long long>(blockIdx.z) * sb);
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    float m = 0.f;
    unsigned low = 0u;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += 8 * stride) {
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)         // eight independent 16-byte loads in flight per thread
            q[u] = (i + u * stride < n4) ? __ldg(v + i + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            m = fmaxf(fmaxf(m, fmaxf(fabsf(q[u].x), fabsf(q[u].y))), fmaxf(fabsf(q[u].z), fabsf(q[u].w)));
            low |= (__float_as_uint(q[u].x) | __float_as_uint(q[u].y) | __float_as_uint(q[u].z) |
                    __float_as_uint(q[u].w)) & 0x1FFFu;
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
        low |= __shfl_xor_sync(0xffffffffu, low, s);
    }
    if ((threadIdx.x & 31) == 0) {
        s_m[threadIdx.x >> 5] = m;
        s_low[threadIdx.x >> 5] = low;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 1; w < 8; ++w) {
            m = fmaxf(m, s_m[w]);
            low |= s_low[w];
        }
        if (m > 0.f) atomicMax(out_bits + blockIdx.y, __float_as_uint(m));
        if (low) atomicOr(out_bits + 2, low);
    }
}

// CTA = one 8x8 source-pixel block x 64 channels of fmap1 (operand A) or fmap2 (operands B_0..B_3).
// Every source value is loaded ONCE; the three pooled levels are built hierarchically in shared memory with
// the reference's own nesting (avg_pool2d applied l times, floor mode), so a block yields 64 / 16 / 4 / 1 operand
// rows of levels 0 / 1 / 2 / 3.  (The first version gathered 4^l source pixels per pooled element: the level-3
// rows lived in 12 CTAs with 512 serial loads per thread and set the kernel time at 37-46 us.)
__global__ void __launch_bounds__(256) corr_pack_kernel(const __grid_constant__ PackParams p) {
    __shared__ float t0[64][65];
    __shared__ float t1[16][65];
    __shared__ float t2[4][65];
    __shared__ float t3[1][65];
    __shared__ float s_scale;
    __shared__ int s_inexact;

    const int tid = threadIdx.x;
    const int which = blockIdx.z & 1;              // 0: A from fmap1, 1: B levels from fmap2
    const int b = blockIdx.z >> 1;
    const int byi = blockIdx.x / p.bx, bxi = blockIdx.x - byi * p.bx;
    const int y0 = byi * 8, x0 = bxi * 8, k0 = blockIdx.y * 64;
    const long long sk = p.sk[which], sy = p.sy[which], sx = p.sx[which];
    const float* src = p.src[which] + b * p.sb[which];
    pdl_launch();
    pdl_wait();
    if (tid == 0) {
        s_scale = exp2f(static_cast<float>(scale_exponent_from_bits(p.amax_bits[which])));
        s_inexact = (p.split == 2) ? (p.amax_bits[2] != 0u) : (p.split == 1);
    }
    __syncthreads();
    const float scale = s_scale;
    const bool inexact = s_inexact != 0;

    const bool vec = (sk == 1) && ((sx & 3) == 0) && ((sy & 3) == 0) && ((p.sb[which] & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.src[which]) & 15) == 0) && (k0 + 64 <= p.D);
    if (vec) {
        float4 q[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {           // 16 threads x float4 = the 64 channels of one pixel (256 B)
            const int px = (tid >> 4) + 16 * it;
            const int y = y0 + (px >> 3), x = x0 + (px & 7);
            q[it] = (y < p.h && x < p.w) ? __ldg(reinterpret_cast<const float4*>(src + y * sy + x * sx + k0) + (tid & 15))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int px = (tid >> 4) + 16 * it, c = 4 * (tid & 15);
            t0[px][c + 0] = q[it].x * scale;
            t0[px][c + 1] = q[it].y * scale;
            t0[px][c + 2] = q[it].z * scale;
            t0[px][c + 3] = q[it].w * scale;
        }
    } else {
        const bool xfast = (sx == 1);
#pragma unroll 4
        for (int i = tid; i < 64 * 64; i += 256) {
            const int px = xfast ? (i & 63) : (i >> 6);
            const int kk = xfast ? (i >> 6) : (i & 63);
            const int y = y0 + (px >> 3), x = x0 + (px & 7);
            float v = 0.f;
            if (y < p.h && x < p.w && k0 + kk < p.D) v = __ldg(src + y * sy + x * sx + (k0 + kk) * sk) * scale;
            t0[px][kk] = v;
        }
    }
    __syncthreads();

    if (which == 1) {   // pooled levels, nested exactly like the reference: level l+1 = avg_pool2d(level l, 2, 2)
        for (int i = tid; i < 16 * 64; i += 256) {
            const int c = i >> 6, kk = i & 63, cy = c >> 2, cx = c & 3;
            const int o = (2 * cy) * 8 + 2 * cx;
            t1[c][kk] = (((t0[o][kk] + t0[o + 1][kk]) + t0[o + 8][kk]) + t0[o + 9][kk]) * 0.25f;
        }
        __syncthreads();
        {
            const int c = tid >> 6, kk = tid & 63, cy = c >> 1, cx = c & 1;
            const int o = (2 * cy) * 4 + 2 * cx;
            t2[c][kk] = (((t1[o][kk] + t1[o + 1][kk]) + t1[o + 4][kk]) + t1[o + 5][kk]) * 0.25f;
        }
        __syncthreads();
        if (tid < 64) t3[0][tid] = (((t2[0][tid] + t2[1][tid]) + t2[2][tid]) + t2[3][tid]) * 0.25f;
        __syncthreads();
    }

    const int Kp = p.kp;
    const int nlev = which ? SF_NUM_LEVELS : 1;
    for (int l = 0; l < nlev; ++l) {
        // k-blocks to write behind [hi]: all three (three-product mode), or for fp16-exact inputs in auto mode only
        // what hi*hi + hi*lo of the pooled levels needs: A's hi*2^-11 block and the lo*2^11 block of B_1..B_3
        const bool second = inexact || (p.split == 2 && (which == 0 || l > 0));
        const bool third = inexact;
        const int side = 8 >> l;                                   // cells per block edge at this level
        const int hl = which ? p.hl[l] : p.h, wl = which ? p.wl[l] : p.w;
        // A rows are the dense query index y*w + x; B rows follow the 4x4-tiled image layout of the pyramid
        // (the GEMM's column index IS the offset inside the query's correlation image), pad cells = zero rows
        const int vmax = which ? p.th[l] * 4 : p.h, umax = which ? p.tw[l] * 4 : p.w;
        const long long rows = which ? p.rows[l] : static_cast<long long>(p.h) * p.w;
        __half* dst = (which ? p.dst_b[l] : p.dst_a) + static_cast<long long>(b) * rows * Kp;
        const float(*t)[65] = (l == 0) ? t0 : (l == 1) ? t1 : (l == 2) ? t2 : t3;
        const int v0 = y0 >> l, u0 = x0 >> l;
        for (int i = tid; i < side * side * 32; i += 256) {
            const int c = i >> 5, kk = 2 * (i & 31);
            const int v = v0 + c / side, u = u0 + c % side;
            if (v >= vmax || u >= umax || k0 + kk >= p.D) continue;
            const bool valid = (v < hl) && (u < wl);
            const float f0 = valid ? t[c][kk] : 0.f, f1 = valid ? t[c][kk + 1] : 0.f;
            const __half h0 = __float2half_rn(f0), h1 = __float2half_rn(f1);
            const long long m = which ? tiled_offset(v, u, p.tw[l]) : static_cast<long long>(v) * p.w + u;
            __half* drow = dst + m * Kp + k0 + kk;
            *reinterpret_cast<__half2*>(drow) = __halves2half2(h0, h1);
            if (second) {
                const float l0 = (f0 - __half2float(h0)) * 2048.f, l1 = (f1 - __half2float(h1)) * 2048.f;
                const __half2 lo = __halves2half2(__float2half_rn(l0), __float2half_rn(l1));
                const __half2 hs = __halves2half2(__float2half_rn(__half2float(h0) * (1.f / 2048.f)),
                                                  __float2half_rn(__half2float(h1) * (1.f / 2048.f)));
                *reinterpret_cast<__half2*>(drow + p.D) = which ? lo : hs;
                if (third) *reinterpret_cast<__half2*>(drow + 2 * p.D) = which ? hs : lo;
            }
        }
    }
}

}  // namespace

int launch_absmax2(const float* f1, const float* f2, int64_t B, int64_t D, int64_t h, int64_t w,
                   const int64_t s1[4], const int64_t s2[4], unsigned* amax_bits, cudaStream_t s) {
    SF_CUDA_CHECK(cudaMemsetAsync(amax_bits, 0, 3 * sizeof(unsigned), s));
    Strided4 t0{f1, s1[0], s1[1], s1[2], s1[3]}, t1{f2, s2[0], s2[1], s2[2], s2[3]};
    const long long per_batch = D * h * w, total = B * per_batch;
    auto dense = [&](const int64_t st[4]) {
        int64_t sz[3] = {D, h, w}, sd[3] = {st[1], st[2], st[3]};
        for (int i = 0; i < 3; ++i)
            for (int j = i + 1; j < 3; ++j)
                if (sd[j] < sd[i]) { std::swap(sd[i], sd[j]); std::swap(sz[i], sz[j]); }
        return sd[0] == 1 && sd[1] == sz[0] && sd[2] == sz[0] * sz[1];
    };
    if (dense(s1) && dense(s2) && per_batch % 4 == 0 && s1[0] % 4 == 0 && s2[0] % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(f1) & 15) == 0 && (reinterpret_cast<uintptr_t>(f2) & 15) == 0) {
        // one float4 x 8 per thread where the tensor is large enough; B <= 65535 batches ride in grid.z
        const int fb = static_cast<int>(std::max<long long>(1, std::min<long long>((per_batch / 4 + 2047) / 2048, 592)));
        SF_REQUIRE(B <= 65535, "corr_build: batch %lld too large", (long long)B);
        prof_before(SF_KERNEL_CORR_PACK, s);
        SF_CUDA_CHECK(launch_kernel(absmax2_flat_kernel, dim3(fb, 2, static_cast<unsigned>(B)), dim3(256), 0, s, f1, f2, (long long)s1[0],
                                    (long long)s2[0], static_cast<int>(B), per_batch, amax_bits));
        prof_after(SF_KERNEL_CORR_PACK, s);
        SF_CUDA_CHECK(cudaGetLastError());
        return SF_OK;
    }
    const int blocks = static_cast<int>(std::min<long long>((total + 1023) / 1024, 1184));
    prof_before(SF_KERNEL_CORR_PACK, s);
    absmax2_kernel<<<dim3(blocks, 2), 256, 0, s>>>(t0, t1, static_cast<int>(D), static_cast<int>(h),
                                                   static_cast<int>(w), per_batch, total, amax_bits);
    prof_after(SF_KERNEL_CORR_PACK, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

int launch_corr_pack(const PackParams& p, int64_t B, cudaStream_t s) {
    dim3 grid(p.bx * p.by, (p.D + 63) / 64, static_cast<unsigned>(2 * B));
    prof_before(SF_KERNEL_CORR_PACK, s);
    SF_CUDA_CHECK(launch_kernel(corr_pack_kernel, grid, dim3(256), 0, s, p));
    prof_after(SF_KERNEL_CORR_PACK, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
