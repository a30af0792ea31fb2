// SF_PREC_FP32_SIMT: the reference's own arithmetic (fp32 FFMA dot products, then three successive
// 2x2 average-pool passes, core/corr.py:13-21,46-54) on CUDA cores.  It is the strict-fp32 mode and the
// in-library cross-check for the tensor-core path; it is NOT a fallback (selected only on request).
#include "sf_internal.h"

namespace sf {

namespace {

// [B, D, h, w] strided -> dense [B, D, N]
__global__ void gather_dense_kernel(const float* __restrict__ src, long long sb, long long sk, long long sy,
                                    long long sx, float* __restrict__ dst, int D, int h, int w, long long total) {
    const int hw = h * w;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / (static_cast<long long>(D) * hw);
        const int rem = static_cast<int>(i - b * D * hw);
        const int k = rem / hw, n = rem - k * hw;
        const int y = n / w, x = n - y * w;
        dst[i] = __ldg(src + b * sb + k * sk + y * sy + x * sx);
    }
}

constexpr int TM = 64, TN = 64, TK = 16;

// C[n, m] = sum_k A[k, n] * Bm[k, m] / alpha;  A, Bm dense [D, N];  C row n is a 4x4-tiled image (tw tiles per row).
__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                       float* __restrict__ C, int D, int N, int w, int tw,
                                                       long long img, float alpha) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int b = blockIdx.z;
    A += static_cast<long long>(b) * D * N;
    Bm += static_cast<long long>(b) * D * N;
    C += static_cast<long long>(b) * N * img;
    const int n0 = blockIdx.y * TM, m0 = blockIdx.x * TN;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 16 x 16 threads, 4 x 4 outputs each
    float acc[4][4] = {};
    for (int k0 = 0; k0 < D; k0 += TK) {
        for (int i = threadIdx.x; i < TK * TM; i += 256) {
            const int kk = i / TM, c = i - kk * TM;
            const int k = k0 + kk;
            As[kk][c] = (k < D && n0 + c < N) ? __ldg(A + static_cast<long long>(k) * N + n0 + c) : 0.f;
            Bs[kk][c] = (k < D && m0 + c < N) ? __ldg(Bm + static_cast<long long>(k) * N + m0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + tx * 4 + j;
            if (m >= N) continue;
            const int v = m / w, u = m - v * w;
            C[static_cast<long long>(n) * img + tiled_offset(v, u, tw)] = acc[i][j] / alpha;
        }
    }
}

// level l -> level l+1 (both 4x4-tiled), one thread per output cell of the tile grid; pad cells are written as 0
__global__ void pool2x2_kernel(const float* __restrict__ src, float* __restrict__ dst, int tw_s, long long img_s,
                               int ho, int wo, int tw_o, long long img_o, long long rows) {
    const long long total = rows * img_o;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / img_o;
        const int rem = static_cast<int>(i - r * img_o);
        const int tile = rem >> 4, in = rem & 15;
        const int v = (tile / tw_o) * 4 + (in >> 2), u = (tile % tw_o) * 4 + (in & 3);
        float out = 0.f;
        if (u < wo && v < ho) {
            const float* s = src + r * img_s;
            out = (((s[tiled_offset(2 * v, 2 * u, tw_s)] + s[tiled_offset(2 * v, 2 * u + 1, tw_s)]) +
                    s[tiled_offset(2 * v + 1, 2 * u, tw_s)]) + s[tiled_offset(2 * v + 1, 2 * u + 1, tw_s)]) * 0.25f;
        }
        dst[i] = out;
    }
}

}  // namespace

int launch_corr_simt(const float* f1, const float* f2, int64_t B, int64_t D, int64_t h, int64_t w,
                     const int64_t s1[4], const int64_t s2[4], float* const levels[SF_NUM_LEVELS], float* ws_a,
                     float* ws_b, cudaStream_t s) {
    const LevelGeom g = make_level_geom(h, w);
    const long long N = h * w, total = B * D * N;
    const int gb = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    prof_before(SF_KERNEL_CORR_SIMT, s);
    prof_before(0, s); prof_before(0, s); prof_before(0, s); prof_before(0, s); prof_before(0, s);   // 6 launches
    gather_dense_kernel<<<gb, 256, 0, s>>>(f1, s1[0], s1[1], s1[2], s1[3], ws_a, (int)D, (int)h, (int)w, total);
    gather_dense_kernel<<<gb, 256, 0, s>>>(f2, s2[0], s2[1], s2[2], s2[3], ws_b, (int)D, (int)h, (int)w, total);
    if ((g.h[0] & 3) || (g.w[0] & 3))
        SF_CUDA_CHECK(cudaMemsetAsync(levels[0], 0, sizeof(float) * B * N * g.img[0], s));
    dim3 grid((unsigned)((N + TN - 1) / TN), (unsigned)((N + TM - 1) / TM), (unsigned)B);
    sgemm_tn_kernel<<<grid, 256, 0, s>>>(ws_a, ws_b, levels[0], (int)D, (int)N, (int)w, g.tw[0], g.img[0],
                                         sqrtf(static_cast<float>(D)));
    for (int l = 0; l + 1 < SF_NUM_LEVELS; ++l) {
        const long long rows = B * N, tot = rows * g.img[l + 1];
        const int pb = static_cast<int>(std::min<long long>((tot + 255) / 256, 148 * 32));
        pool2x2_kernel<<<pb, 256, 0, s>>>(levels[l], levels[l + 1], g.tw[l], g.img[l], g.h[l + 1], g.w[l + 1],
                                          g.tw[l + 1], g.img[l + 1], rows);
    }
    prof_after(SF_KERNEL_CORR_SIMT, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
