// Operand preparation for the correlation GEMM (part of CorrBlock.__init__, core/corr.py:7-21).
//
//  absmax2_kernel   per-tensor |max| of fmap1 / fmap2 (device-side, no host sync) -> power-of-two scale.
//  corr_pack_kernel reads the fp32 feature maps through their ORIGINAL strides (the model hands over
//                   channels-last views, core/models/streamflow.py:107,110) and writes K-major fp16 operand
//                   matrices the TMA can tile:  A = fmap1 as [B, N, Kp];  B_l = avg-pooled fmap2 at level l
//                   as [B, h_l*pitch_l, Kp] (rows at pad columns are zero).  Pooling the operand instead of
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
        m = fmaxf(m, fabsf(__ldg(t.p + b * t.sb + k * t.sk + y * t.sy + x * t.sx)));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits + blockIdx.y, __float_as_uint(m));
}

// dense per-batch blocks (NCHW-contiguous or channels-last): plain vectorised sweep of the storage
__global__ void absmax2_flat_kernel(const float* f0, const float* f1, long long sb0, long long sb1, int B,
                                    long long per_batch, unsigned* out_bits) {
    const float* f = blockIdx.y == 0 ? f0 : f1;
    const long long sb = blockIdx.y == 0 ? sb0 : sb1;
    const long long n4 = per_batch >> 2;
    float m = 0.f;
    for (int b = 0; b < B; ++b) {
        const float4* v = reinterpret_cast<const float4*>(f + b * sb);
        const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
        for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += 4 * stride) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)     // four independent 16-byte loads in flight per thread
                q[u] = (i + u * stride < n4) ? __ldg(v + i + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                m = fmaxf(fmaxf(m, fmaxf(fabsf(q[u].x), fabsf(q[u].y))), fmaxf(fabsf(q[u].z), fabsf(q[u].w)));
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits + blockIdx.y, __float_as_uint(m));
}

// CTA = 32 rows x 64 channels of one segment / batch element.  blockDim = (32, 8).
// Index arithmetic is done once per row (pixel offset) in shared memory; the per-element work is one add, the
// loads of the 2^l x 2^l pooling block and one multiply (the first version spent ~110 instructions per element
// on 64-bit stride arithmetic and was issue-bound at 37 us).
__global__ void __launch_bounds__(256) corr_pack_kernel(const __grid_constant__ PackParams p) {
    __shared__ float tile[64][33];
    __shared__ long long s_pix[32];     // element offset of the top-left source pixel of the row's pooling block
    __shared__ float s_scale;

    int si = 0;
#pragma unroll
    for (int i = 1; i < 1 + SF_NUM_LEVELS; ++i)
        if (i < p.nseg && static_cast<int>(blockIdx.x) >= p.seg[i].tile0) si = i;
    const PackSeg& s = p.seg[si];
    const int row0 = (static_cast<int>(blockIdx.x) - s.tile0) * 32;
    const int k0 = blockIdx.y * 64;
    const int b = blockIdx.z;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const bool kfast = (s.sk == 1);
    const int level = s.level;
    if (ty == 0) {
        const int m = row0 + tx;
        const int v = m / s.pitch, u = m - v * s.pitch;
        const bool ok = (m < s.rows) && (u < s.wl);
        s_pix[tx] = ok ? (static_cast<long long>(v << level) * s.sy + static_cast<long long>(u << level) * s.sx) : -1;
        if (tx == 0) s_scale = exp2f(static_cast<float>(scale_exponent_from_bits(p.amax_bits[s.amax_slot])));
    }
    __syncthreads();
    const float* src = s.src + b * s.sb;
    const int side = 1 << level;
    const float scale = s_scale / static_cast<float>(side * side);

#pragma unroll
    for (int it = 0; it < 8; ++it) {
        int kk, rr;
        if (kfast) {
            kk = tx + 32 * (it & 1);
            rr = ty + 8 * (it >> 1);
        } else {
            rr = tx;
            kk = ty + 8 * it;
        }
        const long long pix = s_pix[rr];
        float val = 0.f;
        if (pix >= 0 && k0 + kk < p.D) {
            const float* q = src + pix + static_cast<long long>(k0 + kk) * s.sk;
            if (level == 0) {
                val = __ldg(q);
            } else {
                // mean over the 2^l x 2^l block == l nested 2x2 average pools (floor mode); exact up to fp32
                // summation order, far below the fp16 operand rounding that follows
                for (int dy = 0; dy < side; ++dy)
                    for (int dx = 0; dx < side; ++dx) val += __ldg(q + dy * s.sy + dx * s.sx);
            }
            val *= scale;
        }
        tile[kk][rr] = val;
    }
    __syncthreads();

    const int Kp = p.split ? 3 * p.D : p.D;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int rr = ty + 8 * it;
        const int m = row0 + rr;
        const int kk = 2 * tx;
        if (m >= s.rows || k0 + kk >= p.D) continue;
        const float v0 = tile[kk][rr], v1 = tile[kk + 1][rr];
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        __half* drow = s.dst + (static_cast<long long>(b) * s.rows + m) * Kp + k0 + kk;
        *reinterpret_cast<__half2*>(drow) = __halves2half2(h0, h1);
        if (p.split) {
            const float l0 = (v0 - __half2float(h0)) * 2048.f, l1 = (v1 - __half2float(h1)) * 2048.f;
            const __half2 lo = __halves2half2(__float2half_rn(l0), __float2half_rn(l1));
            const __half2 hs = __halves2half2(__float2half_rn(__half2float(h0) * (1.f / 2048.f)),
                                              __float2half_rn(__half2float(h1) * (1.f / 2048.f)));
            *reinterpret_cast<__half2*>(drow + p.D) = s.is_b ? lo : hs;
            *reinterpret_cast<__half2*>(drow + 2 * p.D) = s.is_b ? hs : lo;
        }
    }
}

}  // namespace

int launch_absmax2(const float* f1, const float* f2, int64_t B, int64_t D, int64_t h, int64_t w,
                   const int64_t s1[4], const int64_t s2[4], unsigned* amax_bits, cudaStream_t s) {
    SF_CUDA_CHECK(cudaMemsetAsync(amax_bits, 0, 2 * sizeof(unsigned), s));
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
        const int fb = static_cast<int>(std::min<long long>((per_batch / 4 + 255) / 256, 592));
        prof_before(SF_KERNEL_CORR_PACK, s);
        absmax2_flat_kernel<<<dim3(fb, 2), 256, 0, s>>>(f1, f2, s1[0], s2[0], static_cast<int>(B), per_batch,
                                                        amax_bits);
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

int launch_corr_pack(const PackParams& p, int total_tiles, int64_t B, cudaStream_t s) {
    dim3 grid(total_tiles, (p.D + 63) / 64, static_cast<unsigned>(B));
    prof_before(SF_KERNEL_CORR_PACK, s);
    corr_pack_kernel<<<grid, dim3(32, 8), 0, s>>>(p);
    prof_after(SF_KERNEL_CORR_PACK, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
