// G2: radius-4 bilinear pyramid lookup (replaces CorrBlock.__call__, core/corr.py:23-44 and
// bilinear_sampler / grid_sample, core/utils/utils.py:65-79).
//
// HBM-bound gather.  Algorithmic bytes per query: 4 levels x 10x10 fp32 window (1600 B) + 8 B coords
// + 4 x 81 fp32 outputs (1296 B) = 2904 B.
//
// Each query's correlation image is stored as 4x4 tiles of 64 B (see sf_internal.h): a 10x10 window at an
// arbitrary offset touches on average 3.25 x 3.25 tiles = 676 B of 64-byte DRAM fetches, against 1024 B for a
// row-major image (10 rows x 1.6 blocks; measured 87.7 MB of DRAM reads for 33.8 MB of window bytes).
//
// Mapping: one CTA = 32 consecutive queries x ONE pyramid level (grid = tiles x 4 levels x groups).
//   phase 0  warp 0, lane = query: coords -> integer window origin (x0, y0) and the single fractional
//            pair (ax, ay) shared by all 81 taps of the level (window offsets are integers);
//   phase 1  all 128 threads, thread = (query, tile column): the up-to 4x4 tiles under the window are fetched as
//            16-byte cp.async.cg chunks (zero-filled outside the image), up to 13 loads in flight per thread;
//   phase 2  lane = query: conflict-free LDS.128 of its own rows, horizontal then vertical lerp in
//            registers, and one 128-byte coalesced store per output channel straight into the NCHW result
//            (channel = l*81 + i*9 + j, i moves x, j moves y).  Warps split the 9 y-offsets.
#include "sf_internal.h"

namespace sf {

namespace {

constexpr int kQ = 32;              // queries per CTA
constexpr int kRowFloats = 16;      // 4 tile columns of 4 floats per staged row
constexpr int kRows = 2 * SF_RADIUS + 2;                 // 10 window rows / columns
constexpr int kStageRows = kRows + 3;                    // window may start at row 0..3 of its first tile
constexpr int kWinStride = kStageRows * kRowFloats + 4;  // 212 floats: 8 consecutive queries -> distinct bank quads
constexpr int kSide = 2 * SF_RADIUS + 1;                 // 9

__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, bool pred) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst));
    const int bytes = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

template <bool kHalfOut>
__global__ void __launch_bounds__(128) corr_lookup_kernel(const __grid_constant__ LookupParams p) {
    __shared__ __align__(16) float win[kQ * kWinStride];
    __shared__ float s_ax[kQ], s_ay[kQ];
    __shared__ int s_x0[kQ], s_y0[kQ];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int lvl = blockIdx.y;
    const int grp = blockIdx.z;
    const long long q0 = static_cast<long long>(blockIdx.x) * kQ;

    const int hl = p.hl[lvl], wl = p.wl[lvl], th = p.th[lvl], tw = p.tw[lvl];

    if (warp == 0) {
        const long long qid = q0 + lane;
        float X0 = -16.f, Y0 = -16.f;
        if (qid < p.BN) {
            const long long b = qid / p.N;
            const long long n = qid - b * p.N;
            const float* c = p.coords[grp] + b * 2 * p.N + n;
            const float inv = 1.0f / static_cast<float>(1 << lvl);
            X0 = __ldg(c) * inv - static_cast<float>(SF_RADIUS);
            Y0 = __ldg(c + p.N) * inv - static_cast<float>(SF_RADIUS);
        }
        // Far outside the image every tap is zero; clamping keeps the int conversion defined
        // (NaN coordinates clamp to the lower bound and yield zeros).
        X0 = fminf(fmaxf(X0, -16.f), static_cast<float>(wl + 8));
        Y0 = fminf(fmaxf(Y0, -16.f), static_cast<float>(hl + 8));
        const float xf = floorf(X0), yf = floorf(Y0);
        s_ax[lane] = X0 - xf;
        s_ay[lane] = Y0 - yf;
        s_x0[lane] = static_cast<int>(xf);
        s_y0[lane] = static_cast<int>(yf);
    }
    __syncthreads();

    {   // phase 1: thread = (query, tile column)
        const int q = tid >> 2, j = tid & 3;
        const long long qid = q0 + q;
        const int x0 = s_x0[q], y0 = s_y0[q];
        const int ox = x0 & 3, oy = y0 & 3;                 // window origin inside its first tile
        const int txc = (x0 >> 2) + j, ty0 = y0 >> 2;
        if (qid < p.BN && 4 * j < ox + kRows) {             // tile column j overlaps window columns ox .. ox+9
            const bool colok = (txc >= 0) && (txc < tw);
            const float* base = p.lvl[grp][lvl] + qid * p.img[lvl];
            float* dst = win + q * kWinStride + j * 4;
#pragma unroll
            for (int R = 0; R < kStageRows; ++R) {          // staged row R = tile row R>>2, row R&3 inside the tile
                if (R < oy || R >= oy + kRows) continue;    // outside the window (still in the same 64 B block)
                const int ty = ty0 + (R >> 2);
                const bool ok = colok && (ty >= 0) && (ty < th);
                const float* src = ok ? base + ((static_cast<long long>(ty) * tw + txc) << 4) + ((R & 3) << 2) : base;
                cp_async16_zfill(dst + R * kRowFloats, src, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    // phase 2: lane = query; warp w owns y-offsets j in [jb, je]
    const long long qid = q0 + lane;
    if (qid >= p.BN) return;
    const int jb = (warp == 0) ? 0 : (2 * warp + 1);       // 0,3,5,7
    const int je = (warp == 0) ? 2 : (2 * warp + 2);       // 2,4,6,8
    const float ax = s_ax[lane], ay = s_ay[lane];
    const int o = s_x0[lane] & 3;
    const float4* wq = reinterpret_cast<const float4*>(win + lane * kWinStride) + (s_y0[lane] & 3) * 4;

    const long long b = qid / p.N;
    const long long n = qid - b * p.N;
    const long long chan0 = b * (SF_NUM_LEVELS * kSide * kSide) + lvl * (kSide * kSide);
    float hprev[kSide];

#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        if (r < jb || r > je + 1) continue;
        float t[16];
        {
            const float4 c0 = wq[r * 4 + 0], c1 = wq[r * 4 + 1], c2 = wq[r * 4 + 2], c3 = wq[r * 4 + 3];
            t[0] = c0.x; t[1] = c0.y; t[2] = c0.z; t[3] = c0.w;
            t[4] = c1.x; t[5] = c1.y; t[6] = c1.z; t[7] = c1.w;
            t[8] = c2.x; t[9] = c2.y; t[10] = c2.z; t[11] = c2.w;
            t[12] = c3.x; t[13] = c3.y; t[14] = c3.z; t[15] = c3.w;
        }
        float s1[12], u[10], hcur[kSide];
#pragma unroll
        for (int i = 0; i < 12; ++i) s1[i] = (o & 1) ? t[i + 1] : t[i];
#pragma unroll
        for (int i = 0; i < 10; ++i) u[i] = (o & 2) ? s1[i + 2] : s1[i];
#pragma unroll
        for (int i = 0; i < kSide; ++i) hcur[i] = fmaf(ax, u[i + 1] - u[i], u[i]);
        if (r > jb) {
            const int j = r - 1;
#pragma unroll
            for (int i = 0; i < kSide; ++i) {
                const float v = fmaf(ay, hcur[i] - hprev[i], hprev[i]);
                const long long idx = (chan0 + i * kSide + j) * p.N + n;
                if (kHalfOut) {
                    reinterpret_cast<__half*>(p.out[grp])[idx] = __float2half_rn(v);
                } else {
                    __stcs(reinterpret_cast<float*>(p.out[grp]) + idx, v);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kSide; ++i) hprev[i] = hcur[i];
    }
}

}  // namespace

int launch_corr_lookup(const LookupParams& p, int groups, cudaStream_t s) {
    const long long tiles = (p.BN + kQ - 1) / kQ;
    SF_REQUIRE(tiles > 0 && tiles < (1ll << 31), "corr_lookup: bad query count %lld", p.BN);
    dim3 grid(static_cast<unsigned>(tiles), SF_NUM_LEVELS, static_cast<unsigned>(groups));
    prof_before(SF_KERNEL_LOOKUP, s);
    if (p.out_f16) {
        corr_lookup_kernel<true><<<grid, 128, 0, s>>>(p);
    } else {
        corr_lookup_kernel<false><<<grid, 128, 0, s>>>(p);
    }
    prof_after(SF_KERNEL_LOOKUP, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
