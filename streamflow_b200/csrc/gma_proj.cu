// 1x1-convolution projections of GMA (to_qk, core/gma.py:47,56; to_v, core/gma.py:80,94) as a small
// tensor-core GEMM:  out[o, n] = scale * sum_c W[o, c] * X[c, n]   (X = NCHW feature map, n = y*w + x).
//
// 0.23 GFLOP per map: latency-, not throughput-critical, so this uses the warp-level mma.sync path
// (m16n8k16, fp16 in / fp32 accumulate) with ldmatrix.trans for the n-contiguous X operand.  Results are
// written in the layouts the tcgen05 kernels consume through TMA:
//   token-major   [P, N, ld]    (q, k; optionally [hi | lo] split along K for fp32-faithful logits)
//   channel-major [P, O, ld]    (v; ld = Npad, pad columns written as zero)
#include <cuda_bf16.h>

#include "sf_internal.h"

namespace sf {

namespace {

constexpr int kTok = 64;          // tokens per CTA
constexpr int kXPad = kTok + 8;   // Xs row pitch (halfs): 144 B rows keep ldmatrix conflict-free

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p) {
    return static_cast<float>(*p);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p) {
    return __half2float(*p);
}
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}

__device__ __forceinline__ void ldmatrix_x4_trans(unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3,
                                                  const void* smem_ptr) {
    const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(smem_ptr));
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(a));
}

__device__ __forceinline__ void mma_16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// CTA: 256 threads = 8 warps; tile = 128 outputs x 64 tokens; warp w owns outputs [16w, 16w+16).
// Requires O == 128 per launch (rows o0..o0+127 of W), C % 16 == 0, C <= 256.
template <typename T>
__global__ void __launch_bounds__(256, 2) gma_proj_kernel(const __grid_constant__ GmaProjParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int C = p.C;
    const bool second = blockIdx.z != 0;                             // q and k of one attention call share a launch
    const void* px = second ? p.x2 : p.x;
    const float* pw = second ? p.w2 : p.w;
    const float pscale = second ? p.scale2 : p.scale;
    __half* pout = second ? p.out2 : p.out;
    const int wpad = C + 8;
    __half* Xs = reinterpret_cast<__half*>(smem);                    // [C][kXPad]
    __half* Ws = Xs + C * kXPad;                                     // [128][C + 8]
    // fp32-faithful projection (q, k): x = x_hi + x_lo, w = w_hi + w_lo in fp16, three MMA products
    const bool faithful = p.split != 0;
    __half* Xl = Ws + 128 * wpad;                                    // [C][kXPad]      (faithful only)
    __half* Wl = Xl + C * kXPad;                                     // [128][C + 8]    (faithful only)
    // output staging (128 x 72 or 64 x 136 halfs): reuses the X tile once the MMAs are done (two CTAs per SM then fit
    // at C = 128); a narrower X tile is too small, so it gets its own space behind everything else
    __half* Ds = (C >= 128) ? Xs : (faithful ? Wl + 128 * wpad : Xl);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * kTok;
    const int pb = blockIdx.y;
    const T* X = reinterpret_cast<const T*>(px) + static_cast<long long>(pb) * C * p.N;

    pdl_launch();
    pdl_wait();
    if (!second && tid < kTok && n0 + tid < p.N) {
        const long long r = static_cast<long long>(pb) * p.N + n0 + tid;
        if (p.zero_u32 != nullptr) p.zero_u32[r] = 0u;
        if (p.zero_u64 != nullptr) p.zero_u64[r] = 0ull;
    }
    // X tile [C][64 tokens]: all global loads of a batch are issued before any conversion so they overlap
    const bool vec_ok = (sizeof(T) == 4) && ((p.N & 3) == 0) && (n0 + kTok <= p.N) &&
                        ((reinterpret_cast<uintptr_t>(px) & 15) == 0);
    for (int i0 = tid; i0 < C * (kTok / 4); i0 += 256 * 8) {
        float v[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 256;
            const int c = i / (kTok / 4), n4 = (i - c * (kTok / 4)) * 4;
            if (i < C * (kTok / 4)) {
                if (vec_ok) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(
                        reinterpret_cast<const float*>(X) + static_cast<long long>(c) * p.N + n0 + n4));
                    v[u][0] = q.x; v[u][1] = q.y; v[u][2] = q.z; v[u][3] = q.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int n = n0 + n4 + e;
                        v[u][e] = (n < p.N) ? load_as_float<T>(X + static_cast<long long>(c) * p.N + n) : 0.f;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 256;
            if (i >= C * (kTok / 4)) continue;
            const int c = i / (kTok / 4), n4 = (i - c * (kTok / 4)) * 4;
            __half2* d = reinterpret_cast<__half2*>(Xs + c * kXPad + n4);
            d[0] = __floats2half2_rn(v[u][0], v[u][1]);
            d[1] = __floats2half2_rn(v[u][2], v[u][3]);
            if (faithful) {
                const float2 h0 = __half22float2(d[0]), h1 = __half22float2(d[1]);
                __half2* dl = reinterpret_cast<__half2*>(Xl + c * kXPad + n4);
                dl[0] = __floats2half2_rn(v[u][0] - h0.x, v[u][1] - h0.y);
                dl[1] = __floats2half2_rn(v[u][2] - h1.x, v[u][3] - h1.y);
            }
        }
    }
    // W [128][C] fp32 -> fp16 (C % 16 == 0, rows 16-byte aligned)
    for (int i0 = tid; i0 < 128 * (C / 4); i0 += 256 * 8) {
        float4 wv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 256;
            if (i < 128 * (C / 4)) wv[u] = __ldg(reinterpret_cast<const float4*>(pw) + i);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 256;
            if (i >= 128 * (C / 4)) continue;
            const int o = i / (C / 4), c4 = (i - o * (C / 4)) * 4;
            const __half2 w0 = __floats2half2_rn(wv[u].x, wv[u].y), w1 = __floats2half2_rn(wv[u].z, wv[u].w);
            __half2* d = reinterpret_cast<__half2*>(Ws + o * wpad + c4);
            d[0] = w0;
            d[1] = w1;
            if (faithful) {
                const float2 f0 = __half22float2(w0), f1 = __half22float2(w1);
                __half2* dl = reinterpret_cast<__half2*>(Wl + o * wpad + c4);
                dl[0] = __floats2half2_rn(wv[u].x - f0.x, wv[u].y - f0.y);
                dl[1] = __floats2half2_rn(wv[u].z - f1.x, wv[u].w - f1.y);
            }
        }
    }
    __syncthreads();

    float acc[8][4] = {};
    const int g = lane >> 2, t = lane & 3;
    const __half* wrow = Ws + (warp * 16 + g) * wpad;
    for (int k0 = 0; k0 < C; k0 += 16) {
        unsigned a[4];
        a[0] = *reinterpret_cast<const unsigned*>(wrow + k0 + 2 * t);
        a[1] = *reinterpret_cast<const unsigned*>(wrow + 8 * wpad + k0 + 2 * t);
        a[2] = *reinterpret_cast<const unsigned*>(wrow + k0 + 8 + 2 * t);
        a[3] = *reinterpret_cast<const unsigned*>(wrow + 8 * wpad + k0 + 8 + 2 * t);
#pragma unroll
        for (int jt = 0; jt < 8; jt += 2) {
            // matrices: (k 0-7, tile jt) (k 8-15, tile jt) (k 0-7, tile jt+1) (k 8-15, tile jt+1)
            const int mat = lane >> 3, r = lane & 7;
            const __half* src = Xs + (k0 + (mat & 1) * 8 + r) * kXPad + (jt + (mat >> 1)) * 8;
            unsigned b0, b1, b2, b3;
            ldmatrix_x4_trans(b0, b1, b2, b3, src);
            mma_16816(acc[jt], a, b0, b1);
            mma_16816(acc[jt + 1], a, b2, b3);
        }
        if (faithful) {
            const __half* wlrow = Wl + (warp * 16 + g) * wpad;
            unsigned al[4];
            al[0] = *reinterpret_cast<const unsigned*>(wlrow + k0 + 2 * t);
            al[1] = *reinterpret_cast<const unsigned*>(wlrow + 8 * wpad + k0 + 2 * t);
            al[2] = *reinterpret_cast<const unsigned*>(wlrow + k0 + 8 + 2 * t);
            al[3] = *reinterpret_cast<const unsigned*>(wlrow + 8 * wpad + k0 + 8 + 2 * t);
#pragma unroll
            for (int jt = 0; jt < 8; jt += 2) {
                const int mat = lane >> 3, r = lane & 7;
                const int off = (k0 + (mat & 1) * 8 + r) * kXPad + (jt + (mat >> 1)) * 8;
                unsigned b0, b1, b2, b3;
                ldmatrix_x4_trans(b0, b1, b2, b3, Xs + off);      // w_lo * x_hi
                mma_16816(acc[jt], al, b0, b1);
                mma_16816(acc[jt + 1], al, b2, b3);
                ldmatrix_x4_trans(b0, b1, b2, b3, Xl + off);      // w_hi * x_lo
                mma_16816(acc[jt], a, b0, b1);
                mma_16816(acc[jt + 1], a, b2, b3);
            }
        }
    }

    // D fragment: acc[jt][0..1] -> (o = 16w + g, n = 8jt + 2t, +1); acc[jt][2..3] -> o + 8
    // split output: [hi | lo] along K, lo = fp16(v - hi) unscaled (see gma_stats_kernel for the operand schedule)
    const int parts = (p.token_major && p.split) ? 2 : 1;
    for (int part = 0; part < parts; ++part) {
        __syncthreads();   // Ds reuse (and Xs/Ws reads finished on the first pass)
#pragma unroll
        for (int jt = 0; jt < 8; ++jt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int o = warp * 16 + g + (e >> 1) * 8;
                const int n = jt * 8 + 2 * t + (e & 1);
                const float v = acc[jt][e] * pscale;
                const __half hi = __float2half_rn(v);
                const __half val = (part == 0) ? hi : __float2half_rn(v - __half2float(hi));
                if (p.token_major)
                    Ds[n * 136 + o] = val;
                else
                    Ds[o * kXPad + n] = val;
            }
        __syncthreads();
        if (p.token_major) {      // rows = tokens, 128 outputs = 256 B per row
            __half* out = pout + static_cast<long long>(pb) * p.out_batch_stride + part * 128;
            for (int i = tid; i < kTok * 16; i += 256) {
                const int n = i >> 4, seg = i & 15;
                if (n0 + n < p.N)
                    *reinterpret_cast<int4*>(out + static_cast<long long>(n0 + n) * p.ld + seg * 8) =
                        *reinterpret_cast<const int4*>(Ds + n * 136 + seg * 8);
            }
        } else {                  // rows = outputs, 64 tokens = 128 B per row; pad columns get the zeros
            __half* out = pout + static_cast<long long>(pb) * p.out_batch_stride;
            for (int i = tid; i < 128 * 8; i += 256) {
                const int o = i >> 3, seg = i & 7;
                if (n0 + seg * 8 < p.ld)
                    *reinterpret_cast<int4*>(out + static_cast<long long>(o) * p.ld + n0 + seg * 8) =
                        *reinterpret_cast<const int4*>(Ds + o * kXPad + seg * 8);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Operand preparation of the aggregate, every refinement iteration: the motion features X [rows = P * C][N] (fp32 in
// the model, core/update.py:339) -> fp16 [rows][Npad] with zero pad columns -- the K-major-over-keys A operand that
// gma_aggregate_kernel streams by TMA -- and W_v -> fp16 (the reference's autocast casts both the same way,
// core/gma.py:94 under streamflow.py:135).  Pure streaming: 8 elements per thread, 16-byte stores.
template <typename T>
__global__ void __launch_bounds__(256) gma_cast_kernel(const T* __restrict__ x, __half* __restrict__ x16, int N, int Npad,
                                                       int blocks_per_row, long long x_blocks, const void* w,
                                                       int w_is_f32, __half* __restrict__ w16, int w_elems) {
    pdl_launch();
    pdl_wait();
    const long long b = blockIdx.x;
    if (b >= x_blocks) {                                             // trailing blocks: the weights
        const int i = (static_cast<int>(b - x_blocks) * 256 + threadIdx.x) * 8;
        if (i < w_elems) {
            alignas(16) __half h[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
                h[e] = w_is_f32 ? __float2half_rn(static_cast<const float*>(w)[i + e]) : static_cast<const __half*>(w)[i + e];
            *reinterpret_cast<uint4*>(w16 + i) = *reinterpret_cast<const uint4*>(h);
        }
        return;
    }
    // a block covers 4096 columns of one row: thread t takes the 8-element groups t, t + 256 (coalesced 16-byte stores)
    const long long row = b / blocks_per_row;
    const int c0 = static_cast<int>(b - row * blocks_per_row) * 4096 + threadIdx.x * 8;
    const T* src = x + row * N;
    float v[2][8];
    const bool vec = sizeof(T) == 4 && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int n0 = c0 + g * 2048;
        if (vec) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 a = (n0 < N) ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + n0) : z;
            const float4 c = (n0 + 4 < N) ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + n0 + 4) : z;
            v[g][0] = a.x; v[g][1] = a.y; v[g][2] = a.z; v[g][3] = a.w;
            v[g][4] = c.x; v[g][5] = c.y; v[g][6] = c.z; v[g][7] = c.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[g][e] = (n0 + e < N) ? load_as_float<T>(src + n0 + e) : 0.f;
        }
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int n0 = c0 + g * 2048;
        if (n0 >= Npad) continue;
        alignas(16) __half2 h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[g][2 * e], v[g][2 * e + 1]);
        *reinterpret_cast<uint4*>(x16 + row * Npad + n0) = *reinterpret_cast<const uint4*>(h);
    }
}

}  // namespace

int launch_gma_proj(const GmaProjParams& p, cudaStream_t s) {
    SF_REQUIRE(p.O == 128, "gma_proj: 128 output channels per launch (got %d)", p.O);
    SF_REQUIRE(p.C % 16 == 0 && p.C >= 16 && p.C <= 256, "gma_proj: C must be a multiple of 16 in [16, 256] (got %d)",
               p.C);
    SF_REQUIRE(p.ld % 8 == 0, "gma_proj: output pitch must be a multiple of 8");
    SF_REQUIRE((reinterpret_cast<uintptr_t>(p.w) & 15) == 0, "gma_proj: weight pointer must be 16-byte aligned");
    const int base = (p.C * kXPad + 128 * (p.C + 8)) * 2;
    const int smem = base + (p.split ? base : 0) + (p.C >= 128 ? 0 : 128 * kXPad * 2);
    const int cols = p.token_major ? p.N : p.ld;
    dim3 grid((cols + kTok - 1) / kTok, p.P, p.x2 != nullptr ? 2 : 1);
    auto launch = [&](auto kernel) -> int {
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem)) return rc;
        prof_before(SF_KERNEL_GMA_PROJ, s);
        SF_CUDA_CHECK(launch_kernel(kernel, grid, dim3(256), static_cast<size_t>(smem), s, p));
        prof_after(SF_KERNEL_GMA_PROJ, s);
        SF_CUDA_CHECK(cudaGetLastError());
        return SF_OK;
    };
    switch (p.x_dtype) {
        case SF_DT_F32: return launch(gma_proj_kernel<float>);
        case SF_DT_F16: return launch(gma_proj_kernel<__half>);
        case SF_DT_BF16: return launch(gma_proj_kernel<__nv_bfloat16>);
        default: set_error("gma_proj: unsupported dtype %d", p.x_dtype); return SF_ERR_INVALID;
    }
}

int launch_gma_cast(const void* x, int x_dtype, __half* x16, int64_t rows, int64_t N, int64_t Npad, const void* w,
                    int w_dtype, __half* w16, int64_t w_elems, cudaStream_t s) {
    SF_REQUIRE(Npad % 8 == 0 && w_elems % 8 == 0, "gma_cast: Npad and the weight count must be multiples of 8");
    SF_REQUIRE((reinterpret_cast<uintptr_t>(x16) & 15) == 0 && (reinterpret_cast<uintptr_t>(w16) & 15) == 0,
               "gma_cast: outputs must be 16-byte aligned");
    const int bpr = static_cast<int>((Npad + 4095) / 4096);
    const long long x_blocks = rows * bpr;
    const long long w_blocks = (w_elems + 2047) / 2048;
    SF_REQUIRE(x_blocks + w_blocks < (1ll << 31), "gma_cast: shape too large");
    const dim3 grid(static_cast<unsigned>(x_blocks + w_blocks));
    auto launch = [&](auto kernel, auto* xp) -> int {
        prof_before(SF_KERNEL_GMA_PROJ, s);
        SF_CUDA_CHECK(launch_kernel(kernel, grid, dim3(256), 0, s, xp, x16, static_cast<int>(N), static_cast<int>(Npad), bpr,
                                    x_blocks, w, static_cast<int>(w_dtype == SF_DT_F32), w16, static_cast<int>(w_elems)));
        prof_after(SF_KERNEL_GMA_PROJ, s);
        SF_CUDA_CHECK(cudaGetLastError());
        return SF_OK;
    };
    switch (x_dtype) {
        case SF_DT_F32: return launch(gma_cast_kernel<float>, static_cast<const float*>(x));
        case SF_DT_F16: return launch(gma_cast_kernel<__half>, static_cast<const __half*>(x));
        case SF_DT_BF16: return launch(gma_cast_kernel<__nv_bfloat16>, static_cast<const __nv_bfloat16*>(x));
        default: set_error("gma_cast: unsupported dtype %d", x_dtype); return SF_ERR_INVALID;
    }
}

namespace {
__global__ void gma_identity_kernel(float* dst, int d) {
    pdl_launch();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d * d) dst[i] = (i / d == i % d) ? 1.0f : 0.0f;
}
}  // namespace

int launch_gma_identity(float* dst, int d, cudaStream_t s) {
    SF_CUDA_CHECK(launch_kernel(gma_identity_kernel, dim3((d * d + 255) / 256), dim3(256), 0, s, dst, d));
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
