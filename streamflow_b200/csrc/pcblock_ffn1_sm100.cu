// Motion-encoder entry: the first stage of the reference's PCBlock4_Deep_nopool_res
// (core/update.py:18-22,31),
//     y = gelu(x + W2 . gelu(W1 . x + b1) + b2)          per pixel, W1: [1.5 C, C], W2: [C, 1.5 C]   (1x1 convolutions)
// as ONE tcgen05 kernel.  `convc1` applies it to the 324-channel correlation feature the lookup has just written
// (core/update.py:320,330): the reference runs four eager kernels there (conv, GELU, conv, add + GELU) plus the fp32 ->
// fp16 casts of autocast, i.e. the 324 / 486-channel intermediates cross HBM six times; here x is read once (+ once
// more from L2 for the fp32 residual) and y is written once.
//
// One CTA = one tile of 128 pixels of one map, 19 warps:
//   warp 0      TMA producer of W1: chunk h = [128 hidden x Kp] as 128x64 blocks of 16 KB (ring of 3)
//   warp 18     TMA producer of W2: [N2 x 32 hidden] quarters of a chunk (64-byte swizzle, 21 KB, ring of 3)
//   warp 1      MMA issuer (elected lane), owns the 512 TMEM columns
//   warps 2-17  512 workers = 4 per pixel (accumulator row; a warp reads the TMEM lanes of quad = warp & 3), which split the
//               channels of the staging pass and the columns of every accumulator four ways:
//                 stage x (fp32 / fp16 NCHW, coalesced over pixels) as the K-major fp16 A operand, 128B-swizzled;
//                 per hidden chunk h: D1 (TMEM) -> registers -> + b1 -> GELU -> fp16 -> G (the A operand of the second GEMM), one
//                 64-hidden half at a time;
//                 finally Y (TMEM) -> + b2 + x -> GELU -> y (NCHW, coalesced over pixels)
// GEMM1(h):   D1[128 x 128] = X[128 x Kp] . W1_h[128 x Kp]^T            (hidden chunk h of 128; Kp = C rounded up to 64)
// GEMM2(h,s): Y[128 x N2] += G_hs[128 x 64] . W2_hs[N2 x 64]^T           (s = 0, 1: the two 64-hidden halves of the chunk;
//                                                                         N2 = C rounded up to 16, as N = 256 + (N2 - 256))
// issued as G1(0) G1(1) G2(0,0) G2(0,1) G1(2) G2(1,0) ...: the workers copy all of D1(h) into registers first, so the tensor
// pipe computes chunk h+1 while they run GELU on chunk h.  A shared-memory-operand MMA of M = 128 costs ~120 cycles whatever
// its N (measured: 24 N = 64 MMAs took 1.5 us; the A tile read sets the floor), hence N = 128 for GEMM1 -- all the TMEM left
// next to Y allows -- and only two MMAs per k-step for GEMM2.
// TMEM: 128 columns of D1 + N2 <= 384 columns of Y.  The kernel is bound by the 128 x (Hp + N2) GELU evaluations per
// tile on the FP32 pipe (erf by Abramowitz-Stegun 7.1.28, |err| < 3e-7, one MUFU.RCP; evaluated two at a time with the
// packed FFMA2 / FMUL2 instructions of sm_100: 62.5 -> 52.1 us), not by the tensor pipe.
#include <algorithm>

#include "sf_internal.h"
#include "sm100_ptx.cuh"

namespace sf {

namespace {

constexpr int kTile = 128;                  // pixels per CTA
constexpr int kBlk = kTile * 128;           // one 64-k block of a 128-row K-major operand: 16 KB
constexpr int kW1Blk = 128 * 128;           // 128 hidden rows x 64 k: 16 KB
constexpr int kW1Stages = 3;
constexpr int kW2Bufs = 3;                  // W2 travels in 32-hidden quarters of a chunk (64-byte swizzle)
constexpr int kSubs = 4;                    // worker threads per pixel
constexpr int kWorkers = 128 * kSubs;       // 16 warps: 4 per TMEM lane quad
constexpr int kThreadsFfn = 64 + kWorkers + 32;     // + the W2 producer warp
constexpr int kTmemColsFfn = 512;
constexpr int kYCol = 128;                  // Y starts behind D1

struct FfnArgs {
    CUtensorMap tm_w1, tm_w2a, tm_w2b;
    const void* x;
    void* out;
    const float* b1;        // [Hp], zero past Hd
    const float* b2;        // [C]
    int x_f16, out_f16;
    int C, N, kb1, nh, n2, n2a, n2b, w2_bytes, tiles_per_map;
    unsigned long long* trace;   // debug: %globaltimer stamps of CTA 0's first worker thread (null in production)
};

__device__ __forceinline__ void stamp(const FfnArgs& a, int slot) {
    if (a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 64) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        a.trace[slot] = t;
    }
}

__device__ __forceinline__ float gelu_erf(float v) {
    // 0.5 v (1 + erf(v / sqrt 2)); erf(a) = 1 - (1 + a1 a + ... + a6 a^6)^-16 for a >= 0 (Abramowitz & Stegun 7.1.28)
    const float a = fabsf(v) * 0.70710678118654752f;
    float p = fmaf(a, 0.0000430638f, 0.0002765672f);
    p = fmaf(a, p, 0.0001520143f);
    p = fmaf(a, p, 0.0092705272f);
    p = fmaf(a, p, 0.0422820123f);
    p = fmaf(a, p, 0.0705230784f);
    p = fmaf(a, p, 1.0f);
    p *= p;
    p *= p;
    p *= p;
    p *= p;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
    const float e = copysignf(1.0f - r, v);
    const float hv = 0.5f * v;
    return fmaf(hv, e, hv);
}

// Two GELUs at once on the packed-fp32 pipe (FFMA2 / FMUL2, sm_100): the same arithmetic as gelu_erf, 20 instructions per pair
// instead of 32.
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ void gelu_erf2(float v0, float v1, float& g0, float& g1) {
    const unsigned long long a = mul2(pack2(fabsf(v0), fabsf(v1)), pack2(0.70710678118654752f, 0.70710678118654752f));
    unsigned long long p = fma2(a, pack2(0.0000430638f, 0.0000430638f), pack2(0.0002765672f, 0.0002765672f));
    p = fma2(a, p, pack2(0.0001520143f, 0.0001520143f));
    p = fma2(a, p, pack2(0.0092705272f, 0.0092705272f));
    p = fma2(a, p, pack2(0.0422820123f, 0.0422820123f));
    p = fma2(a, p, pack2(0.0705230784f, 0.0705230784f));
    p = fma2(a, p, pack2(1.0f, 1.0f));
    p = mul2(p, p);
    p = mul2(p, p);
    p = mul2(p, p);
    p = mul2(p, p);
    float p0, p1, r0, r1;
    unpack2(p, p0, p1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(p0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(p1));
    const unsigned long long e = fma2(pack2(r0, r1), pack2(-1.0f, -1.0f), pack2(1.0f, 1.0f));
    float e0, e1;
    unpack2(e, e0, e1);
    const unsigned long long hv = mul2(pack2(v0, v1), pack2(0.5f, 0.5f));
    const unsigned long long g = fma2(hv, pack2(copysignf(e0, v0), copysignf(e1, v1)), hv);
    unpack2(g, g0, g1);
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
__device__ __forceinline__ float ld_nc(const float* p) { return __ldg(p); }
__device__ __forceinline__ __half ld_nc(const __half* p) { return __ldg(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(__half* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
    tma_load_3d(m, bar, dst, c0, c1, 0);
}

template <typename XT, typename OT>
__global__ void __launch_bounds__(kThreadsFfn, 1) pcblock_ffn1_kernel(const __grid_constant__ FfnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* xs = smem;                                        // kb1 blocks of 16 KB
    uint8_t* gs = xs + a.kb1 * kBlk;                           // 16 KB
    uint8_t* w2s = gs + kBlk;                                  // kW2Bufs x w2_bytes (one 32-hidden half of a chunk each)
    uint8_t* w1s = w2s + kW2Bufs * a.w2_bytes;                 // ring of 8 KB blocks
    uint64_t* bars = reinterpret_cast<uint64_t*>(w1s + kW1Stages * kW1Blk);
    uint64_t* w1_full = bars;              // [3]
    uint64_t* w1_empty = bars + 3;         // [3]
    uint64_t* w2_full = bars + 6;          // [3]
    uint64_t* w2_empty = bars + 9;         // [3]
    uint64_t* d1_full = bars + 12;
    uint64_t* d1_empty = bars + 13;
    uint64_t* x_full = bars + 22;
    uint64_t* g_full = bars + 23;
    uint64_t* g_empty = bars + 24;
    uint64_t* y_full = bars + 25;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);
    float* b1s = reinterpret_cast<float*>(bars + 32);          // [nh * 64]
    float* b2s = b1s + a.nh * 128;                             // [n2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int map = blockIdx.x / a.tiles_per_map;
    const int n0 = (blockIdx.x - map * a.tiles_per_map) * kTile;

    if (tid == 0) {
        tma_prefetch_desc(&a.tm_w1);
        tma_prefetch_desc(&a.tm_w2a);
        tma_prefetch_desc(&a.tm_w2b);
        for (int i = 0; i < kW1Stages; ++i) {
            mbar_init(&w1_full[i], 1);
            mbar_init(&w1_empty[i], 1);
        }
        for (int i = 0; i < kW2Bufs; ++i) {
            mbar_init(&w2_full[i], 1);
            mbar_init(&w2_empty[i], 1);
        }
        mbar_init(d1_full, 1);
        mbar_init(d1_empty, kWorkers / 32);
        mbar_init(x_full, kWorkers);
        mbar_init(g_full, kWorkers);
        mbar_init(g_empty, 1);
        mbar_init(y_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kTmemColsFfn>(tmem_slot);
    pdl_launch();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        int it = 0;
        for (int h = 0; h < a.nh; ++h) {
            for (int kb = 0; kb < a.kb1; ++kb, ++it) {
                const int s = it % kW1Stages;
                mbar_wait(&w1_empty[s], ((it / kW1Stages) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&w1_full[s], kW1Blk);
                    tma_load_2d(&a.tm_w1, &w1_full[s], w1s + s * kW1Blk, kb * 64, h * 128);
                }
                __syncwarp();
            }
        }
    } else if (warp == kThreadsFfn / 32 - 1) {
        // ------------------------------------------------------------------ W2 producer (its own warp: a W2 buffer that is
        // not free yet must not hold back the W1 ring, which needs a full chunk of look-ahead to hide the L2 latency)
        for (int idx = 0; idx < 4 * a.nh; ++idx) {             // quarter idx = hidden units 32 idx .. 32 idx + 31
            const int b = idx % kW2Bufs;
            mbar_wait(&w2_empty[b], ((idx / kW2Bufs) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&w2_full[b], static_cast<uint32_t>(a.n2 * 64));
                tma_load_2d(&a.tm_w2a, &w2_full[b], w2s + b * a.w2_bytes, idx * 32, 0);
                if (a.n2b > 0) tma_load_2d(&a.tm_w2b, &w2_full[b], w2s + b * a.w2_bytes + a.n2a * 64, idx * 32, a.n2a);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc1 = make_idesc_f16_f32(kTile, 128);
        const uint32_t idesc2a = make_idesc_f16_f32(kTile, a.n2a);
        const uint32_t idesc2b = make_idesc_f16_f32(kTile, a.n2b > 0 ? a.n2b : 16);
        int it = 0;
        auto gemm1 = [&](int h) {
            mbar_wait(d1_empty, (h & 1) ^ 1);                  // the workers hold chunk h-1 in registers
            tc_fence_after();
            for (int kb = 0; kb < a.kb1; ++kb, ++it) {
                const int s = it % kW1Stages;
                mbar_wait(&w1_full[s], (it / kW1Stages) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = make_kmajor_sw128_desc(smem_u32(xs + kb * kBlk));
                    const uint64_t db = make_kmajor_sw128_desc(smem_u32(w1s + s * kW1Blk));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc1, (kb | k) != 0);
                    umma_commit(&w1_empty[s]);
                    if (kb == a.kb1 - 1) umma_commit(d1_full);
                }
                __syncwarp();
            }
        };
        mbar_wait(x_full, 0);                                  // the A operand is staged
        tc_fence_after();
        gemm1(0);
        for (int h = 0; h < a.nh; ++h) {
            if (h + 1 < a.nh) gemm1(h + 1);
            for (int half = 0; half < 2; ++half) {
                mbar_wait(g_full, half);                       // GELU(chunk h, half) is in shared memory
                for (int q = 0; q < 2; ++q) {
                    const int idx = 4 * h + 2 * half + q, b = idx % kW2Bufs;
                    mbar_wait(&w2_full[b], (idx / kW2Bufs) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t da = make_kmajor_sw128_desc(smem_u32(gs)) + 4 * q;      // k-steps 2q, 2q+1 of G
                        const uint64_t db0 = make_kmajor_sw64_desc(smem_u32(w2s + b * a.w2_bytes));
                        const uint64_t db1 = make_kmajor_sw64_desc(smem_u32(w2s + b * a.w2_bytes + a.n2a * 64));
                        const uint32_t first = static_cast<uint32_t>(h | half | q);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            umma_f16_ss(tmem_base + kYCol, da + 2 * k, db0 + 2 * k, idesc2a, (first | k) != 0);
                            if (a.n2b > 0)
                                umma_f16_ss(tmem_base + kYCol + a.n2a, da + 2 * k, db1 + 2 * k, idesc2b, (first | k) != 0);
                        }
                        umma_commit(&w2_empty[b]);
                        if (q == 1) {
                            umma_commit(g_empty);
                            if (h == a.nh - 1 && half == 1) umma_commit(y_full);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ workers: 4 threads per pixel
        const int quad = warp & 3;                             // TMEM lane group this warp may access
        const int sub = (warp - 2) >> 2;                       // 0..3: which share of the k-blocks / columns
        const int r = quad * 32 + lane;                        // accumulator row = pixel of the tile
        const int n = min(n0 + r, a.N - 1);                    // rows past the map repeat its last pixel and are never stored
        const bool live = n0 + r < a.N;
        const long long map_off = static_cast<long long>(map) * a.C * a.N;
        const XT* xg = reinterpret_cast<const XT*>(a.x) + map_off + n;
        const size_t cstride = static_cast<size_t>(a.N);
        stamp(a, 0);
        // biases -> shared memory (broadcast reads later)
        for (int i = tid - 64; i < a.nh * 128; i += kWorkers) b1s[i] = __ldg(a.b1 + i);
        for (int i = tid - 64; i < a.n2; i += kWorkers) b2s[i] = i < a.C ? __ldg(a.b2 + i) : 0.f;
        // stage x: row r of the 32-channel half blocks sub, sub + 4, ...; 16-byte chunk ch at ch ^ (r & 7)
        for (int hb = sub; hb < 2 * a.kb1; hb += kSubs) {
            const int c0 = hb * 32;
            float v[32];
            if (c0 + 32 <= a.C) {
                const XT* src = xg + static_cast<size_t>(c0) * cstride;
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = to_float(ld_nc(src + e * cstride));
            } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = (c0 + e < a.C) ? to_float(ld_nc(xg + static_cast<size_t>(c0 + e) * cstride)) : 0.f;
            }
            uint8_t* row = xs + (hb >> 1) * kBlk + r * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ch = (hb & 1) * 4 + q;
                uint4 pk;
                __half2 t;
                t = __floats2half2_rn(v[8 * q + 0], v[8 * q + 1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2half2_rn(v[8 * q + 2], v[8 * q + 3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2half2_rn(v[8 * q + 4], v[8 * q + 5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
                t = __floats2half2_rn(v[8 * q + 6], v[8 * q + 7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
                *reinterpret_cast<uint4*>(row + ((ch ^ (r & 7)) << 4)) = pk;
            }
        }
        fence_proxy_async_smem();
        mbar_arrive(x_full);
        stamp(a, 1);
        asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");       // the biases are in shared memory

        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const int hc0 = sub * 16;                               // this thread's 16 columns of every 64-hidden half
        for (int h = 0; h < a.nh; ++h) {
            mbar_wait(d1_full, h & 1);
            stamp(a, 2 + 3 * h);
            tc_fence_after();
            uint32_t acc[2][16];
            tmem_ld_32x16(t_lane + hc0, acc[0]);
            tmem_ld_32x16(t_lane + 64 + hc0, acc[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(d1_empty);              // D1 may be overwritten: GEMM1(h+1) runs under the GELUs below
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const float* bs = b1s + h * 128 + half * 64 + hc0;
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 bb = *reinterpret_cast<const float2*>(bs + 2 * j);
                    float g0, g1;
                    gelu_erf2(__uint_as_float(acc[half][2 * j]) + bb.x, __uint_as_float(acc[half][2 * j + 1]) + bb.y, g0, g1);
                    __half2 t = __floats2half2_rn(g0, g1);
                    pk[j] = *reinterpret_cast<uint32_t*>(&t);
                }
                if (half == 0) stamp(a, 3 + 3 * h);
                mbar_wait(g_empty, half ^ 1);                  // GEMM2 of the previous half has read G
                if (half == 0) stamp(a, 4 + 3 * h);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int ch = sub * 2 + q;
                    *reinterpret_cast<uint4*>(gs + r * 128 + ((ch ^ (r & 7)) << 4)) =
                        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
                fence_proxy_async_smem();
                mbar_arrive(g_full);
            }
        }

        // final pass: Y -> + b2 + x -> GELU -> y; 16-column blocks sub, sub + 4, ...
        stamp(a, 26);
        mbar_wait(y_full, 0);
        stamp(a, 27);
        tc_fence_after();
        OT* og = reinterpret_cast<OT*>(a.out) + map_off + n;
        const int nblk = a.n2 / 16;
        auto load_res = [&](int cb, float (&xv)[16]) {         // the fp32 residual x of one 16-channel block
            const int c0 = cb * 16;
            if (cb < nblk && c0 + 16 <= a.C) {
                const XT* src = xg + static_cast<size_t>(c0) * cstride;
#pragma unroll
                for (int j = 0; j < 16; ++j) xv[j] = to_float(ld_nc(src + j * cstride));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    xv[j] = (cb < nblk && c0 + j < a.C) ? to_float(ld_nc(xg + static_cast<size_t>(c0 + j) * cstride)) : 0.f;
            }
        };
        float xcur[16], xnext[16];
        load_res(sub, xcur);
        for (int cb = sub; cb < nblk; cb += kSubs) {           // software-pipelined: the next block's residual is in flight
            load_res(cb + kSubs, xnext);
            uint32_t acc[16];
            tmem_ld_32x16(t_lane + kYCol + cb * 16, acc);
            tmem_ld_wait();
            const int c0 = cb * 16;
            OT* dst = og + static_cast<size_t>(c0) * cstride;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float y0, y1;
                gelu_erf2(__uint_as_float(acc[j]) + b2s[c0 + j] + xcur[j], __uint_as_float(acc[j + 1]) + b2s[c0 + j + 1] + xcur[j + 1],
                          y0, y1);
                if (live && c0 + j < a.C) st_stream(dst + j * cstride, y0);
                if (live && c0 + j + 1 < a.C) st_stream(dst + (j + 1) * cstride, y1);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) xcur[j] = xnext[j];
        }
        stamp(a, 28);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemColsFfn>(tmem_base);
    }
}

unsigned long long* g_ffn1_trace = nullptr;

}  // namespace

void set_ffn1_trace(void* p) { g_ffn1_trace = static_cast<unsigned long long*>(p); }

int launch_pcblock_ffn1(const void* x, int x_dtype, const void* w1p, const float* b1p, const void* w2p, const float* b2,
                        void* out, int out_dtype, int64_t P, int64_t C, int64_t Hd, int64_t N, cudaStream_t s) {
    const int kb1 = static_cast<int>((C + 63) / 64), nh = static_cast<int>((Hd + 127) / 128);     // hidden chunks of 128
    const int n2 = static_cast<int>(align_up(C, 16));
    SF_REQUIRE(P >= 1 && N >= 1 && C >= 16 && Hd >= 16, "pcblock_ffn1: bad shape P=%lld C=%lld hidden=%lld N=%lld",
               (long long)P, (long long)C, (long long)Hd, (long long)N);
    SF_REQUIRE(n2 <= 384 && nh * 128 <= 512,
               "pcblock_ffn1: specialised for C <= 384 channels and <= 512 hidden units (TMEM: 128 + C columns); got C=%lld, "
               "hidden=%lld -- no generic fallback", (long long)C, (long long)Hd);
    SF_REQUIRE(P * C * N < (1ll << 40) && N < (1ll << 31), "pcblock_ffn1: shape too large");
    FfnArgs a{};
    a.x = x; a.out = out; a.b1 = b1p; a.b2 = b2;
    a.x_f16 = (x_dtype == SF_DT_F16); a.out_f16 = (out_dtype == SF_DT_F16);
    a.C = static_cast<int>(C); a.N = static_cast<int>(N);
    a.kb1 = kb1; a.nh = nh; a.n2 = n2;
    a.n2a = std::min(n2, 256);
    a.n2b = n2 - a.n2a;
    a.w2_bytes = static_cast<int>(align_up(static_cast<int64_t>(n2) * 64, 1024));     // one 32-hidden half, 64 B per row
    a.tiles_per_map = static_cast<int>((N + kTile - 1) / kTile);
    a.trace = g_ffn1_trace;
    const uint64_t Kp = static_cast<uint64_t>(kb1) * 64, Hp = static_cast<uint64_t>(nh) * 128;
    // W1p: [Hp, Kp] fp16 (rows = hidden units, K = input channels); W2p: [n2, Hp] fp16 (rows = output channels, K = hidden)
    if (int rc = make_tmap3(&a.tm_w1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, w1p, Kp, Hp, 1, Kp * 2, Hp * Kp * 2, 64, 128, "W1"))
        return rc;
    if (int rc = make_tmap3(&a.tm_w2a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, w2p, Hp, n2, 1, Hp * 2, n2 * Hp * 2, 32,
                            static_cast<uint32_t>(a.n2a), "W2a", 64))
        return rc;
    if (int rc = make_tmap3(&a.tm_w2b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, w2p, Hp, n2, 1, Hp * 2, n2 * Hp * 2, 32,
                            static_cast<uint32_t>(a.n2b > 0 ? a.n2b : 16), "W2b", 64))
        return rc;
    const int smem = kb1 * kBlk + kBlk + kW2Bufs * a.w2_bytes + kW1Stages * kW1Blk + 256 + (nh * 128 + n2) * 4;
    SF_REQUIRE(smem <= 232448, "pcblock_ffn1: %d bytes of shared memory needed", smem);
    const long long grid = P * a.tiles_per_map;
    SF_REQUIRE(grid < (1ll << 31), "pcblock_ffn1: too many tiles");
    auto launch = [&](auto kernel) -> int {
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), smem)) return rc;
        prof_before(SF_KERNEL_PCBLOCK_FFN1, s);
        SF_CUDA_CHECK(launch_kernel(kernel, dim3(static_cast<unsigned>(grid)), dim3(kThreadsFfn), static_cast<size_t>(smem), s, a));
        prof_after(SF_KERNEL_PCBLOCK_FFN1, s);
        SF_CUDA_CHECK(cudaGetLastError());
        return SF_OK;
    };
    if (a.x_f16) return a.out_f16 ? launch(pcblock_ffn1_kernel<__half, __half>) : launch(pcblock_ffn1_kernel<__half, float>);
    return a.out_f16 ? launch(pcblock_ffn1_kernel<float, __half>) : launch(pcblock_ffn1_kernel<float, float>);
}

}  // namespace sf
