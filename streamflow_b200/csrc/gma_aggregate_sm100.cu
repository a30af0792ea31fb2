// G3b: per-iteration GMA aggregation (core/gma.py:91-104) as a streaming tcgen05 GEMM
//
//     acc[p, c, n] += sum_j E[p, n, j] * V[p, c, j]          (gma_aggregate_kernel, this file)
//     out[p, c, n]  = fmap[p, c, n] + acc[p, c, n] * gamma / rowsum[p, n];  acc <- 0     (gma_finalize_kernel)
//
// HBM-bound on the fp16 softmax numerators E (297 MB per Sintel clip and iteration).  Design points:
//   * one CTA owns TWO 128-query tiles (M = 256) per key block, so each 16 KB V tile fetched from L2 feeds two MMAs:
//     SM ingest is 1.5 B per E byte instead of 2 -- with M = 128 the kernel sat on the L2->SM bandwidth cap
//     (measured: ~10.4 TB/s of L2 reads for 4.9 TB/s of HBM), not on HBM;
//   * stream-K: the linear (map, tile-pair, key-block) space is cut into gridDim equal contiguous ranges, so every SM
//     streams the same number of bytes (165 tiles over 148 SMs would otherwise quantise to 2 waves);
//   * E tiles are 16 KB contiguous blocks (tile-major layout written by gma_stats_kernel) on a deep mbarrier ring
//     (5 x 32 KB in flight per SM), V on a shallow one (3 x 16 KB);
//   * split tiles are reduced with red.global.add.f32 into a channel-major fp32 buffer: thread = query row, so for
//     each channel the 32 lanes of a warp hit one 128 B line.  (A fused "last arriver" fix-up epilogue was tried
//     and rejected: with in-order stream-K every CTA ends on a shared tile, so the 512 KB/tile fix-up traffic is
//     issued by 128 threads at the very end of the kernel, latency-bound and fully exposed -- 159 us vs 62 us.)
//
// warps: 0 = E producer (TMA), 1 = TMEM alloc + MMA issuer, 2 = V producer (TMA), 3-6 = epilogue.
#include <cuda_bf16.h>

#include "sf_internal.h"
#include "sm100_ptx.cuh"

namespace sf {

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kEStages = 5, kVStages = 3;
constexpr int kTileBytes = BM * BK * 2;                          // 16 KB
constexpr int kEStageBytes = 2 * kTileBytes;                     // two query tiles per stage
constexpr int kSmemBytes = kEStages * kEStageBytes + kVStages * kTileBytes + 1024 + 512;
constexpr int kTmemCols = 512;                                   // 2 buffers x (2 tiles x 128 columns)
constexpr int kThreads = 224;

struct GmaAggArgs {
    CUtensorMap tm_e, tm_v;
    GmaAggParams p;
};

__global__ void __launch_bounds__(kThreads, 1) gma_aggregate_kernel(const __grid_constant__ GmaAggArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* e_base = smem;
    uint8_t* v_base = smem + kEStages * kEStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(v_base + kVStages * kTileBytes);
    uint64_t* e_full = bars;
    uint64_t* e_empty = e_full + kEStages;
    uint64_t* v_full = e_empty + kEStages;
    uint64_t* v_empty = v_full + kVStages;
    uint64_t* tfull = v_empty + kVStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const GmaAggParams& p = args.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long KB = p.k_blocks;
    const long long G = gridDim.x;
    const long long work = static_cast<long long>(p.P) * p.pair_tiles * KB;
    const long long w_begin = work * blockIdx.x / G;
    const long long w_end = work * (blockIdx.x + 1) / G;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tm_e);
        tma_prefetch_desc(&args.tm_v);
        for (int i = 0; i < kEStages; ++i) {
            mbar_init(&e_full[i], 1);
            mbar_init(&e_empty[i], 1);
        }
        for (int i = 0; i < kVStages; ++i) {
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    pdl_launch();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {                                       // ---- E producer
            int stage = 0;
            uint32_t phase = 0;
            for (long long pos = w_begin; pos < w_end; ++pos) {
                const long long pt = pos / KB;
                const int kb = static_cast<int>(pos - pt * KB);
                const int pb = static_cast<int>(pt / p.pair_tiles);
                const int mp = static_cast<int>(pt - static_cast<long long>(pb) * p.pair_tiles);
                mbar_wait(&e_empty[stage], phase ^ 1);
                uint8_t* dst = e_base + stage * kEStageBytes;
                mbar_expect_tx(&e_full[stage], kEStageBytes);
                // tile-major E: (m-tile, key-block) -> 128 consecutive 128-byte rows; a tile index past the last
                // m-tile (odd tile count) is out of bounds for the tensor map and arrives as zeros
                const long long r0 = (static_cast<long long>(2 * mp) * KB + kb) * BM;
                const long long r1 = (static_cast<long long>(2 * mp + 1) * KB + kb) * BM;
                tma_load_3d_hint(&args.tm_e, &e_full[stage], dst, 0, static_cast<int>(r0), pb, kEvictFirst);
                tma_load_3d_hint(&args.tm_e, &e_full[stage], dst + kTileBytes, 0, static_cast<int>(r1), pb,
                                 kEvictFirst);
                if (++stage == kEStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {                                       // ---- V producer
            int stage = 0;
            uint32_t phase = 0;
            for (long long pos = w_begin; pos < w_end; ++pos) {
                const long long pt = pos / KB;
                const int kb = static_cast<int>(pos - pt * KB);
                const int pb = static_cast<int>(pt / p.pair_tiles);
                mbar_wait(&v_empty[stage], phase ^ 1);
                mbar_expect_tx(&v_full[stage], kTileBytes);
                tma_load_3d_hint(&args.tm_v, &v_full[stage], v_base + stage * kTileBytes, kb * BK, 0, pb, kEvictLast);
                if (++stage == kVStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                       // ---- MMA issuer
            constexpr uint32_t idesc = make_idesc_f16_f32(BM, BN);
            int es = 0, vs = 0, local = 0;
            uint32_t ephase = 0, vphase = 0;
            long long pos = w_begin;
            while (pos < w_end) {
                const long long pt = pos / KB;
                const long long seg_end = min(w_end, (pt + 1) * KB);
                const int acc = local & 1;
                mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (2 * BN);
                bool first = true;
                for (; pos < seg_end; ++pos) {
                    mbar_wait(&v_full[vs], vphase);
                    mbar_wait(&e_full[es], ephase);
                    tc_fence_after();
                    const uint32_t ea = smem_u32(e_base + es * kEStageBytes);
                    const uint64_t d0 = make_kmajor_sw128_desc(ea);
                    const uint64_t d1 = make_kmajor_sw128_desc(ea + kTileBytes);
                    const uint64_t db = make_kmajor_sw128_desc(smem_u32(v_base + vs * kTileBytes));
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        umma_f16_ss(d_tmem, d0 + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                        umma_f16_ss(d_tmem + BN, d1 + 2 * k, db + 2 * k, idesc, !(first && k == 0));
                    }
                    first = false;
                    umma_commit(&e_empty[es]);
                    umma_commit(&v_empty[vs]);
                    if (++es == kEStages) {
                        es = 0;
                        ephase ^= 1;
                    }
                    if (++vs == kVStages) {
                        vs = 0;
                        vphase ^= 1;
                    }
                }
                umma_commit(&tfull[acc]);
                ++local;
            }
        }
    } else {                                                   // ---- epilogue (warps 3-6)
        const int quad = warp & 3;
        int local = 0;
        long long pos = w_begin;
        while (pos < w_end) {
            const long long pt = pos / KB;
            const long long seg_end = min(w_end, (pt + 1) * KB);
            const int pb = static_cast<int>(pt / p.pair_tiles);
            const int mp = static_cast<int>(pt - static_cast<long long>(pb) * p.pair_tiles);
            const int acc = local & 1;
            mbar_wait(&tfull[acc], (local >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * (2 * BN);
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int n = (2 * mp + half) * BM + quad * 32 + lane;
                // acc is channel-major [P, 128, N] (the layout of the NCHW result): for a fixed channel the 32
                // lanes of a warp hit 32 consecutive floats, so every warp-level red is one coalesced 128 B line
                float* dst = p.acc + static_cast<long long>(pb) * BN * p.N + n;
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ++ch) {
                    uint32_t v[32];
                    tmem_ld_32x32(t_acc + half * BN + ch * 32, v);
                    tmem_ld_wait();
                    if (half == 1 && ch == BN / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[acc]);
                    }
                    if (n < p.N) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + static_cast<long long>(ch * 32 + j) * p.N),
                                         "f"(__uint_as_float(v[j]))
                                         : "memory");
                    }
                }
            }
            pos = seg_end;
            ++local;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

// out[p, c, n] = fmap[p, c, n] + acc[p, c, n] * (gamma / rowsum[p, n]);  acc <- 0.
// grid = (ceil(N / 1024), P * C): one float4 per thread, no index arithmetic, every load in flight at once.
template <typename T>
__global__ void __launch_bounds__(256) gma_finalize_kernel(const __grid_constant__ GmaAggParams p) {
    pdl_launch();
    pdl_wait();
    const int row = blockIdx.y;                          // p * C + c
    const int pb = row / p.C;
    const long long base = static_cast<long long>(row) * p.N;
    const T* fm = reinterpret_cast<const T*>(p.fmap) + base;
    float* acc = p.acc + base;
    float* out = p.out + base;
    const float* rs = p.rscale + static_cast<long long>(pb) * p.N;
    const int n = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (n >= p.N) return;
    if ((p.N & 3) == 0) {
        const float4 a = *reinterpret_cast<const float4*>(acc + n);
        const float4 r = __ldg(reinterpret_cast<const float4*>(rs + n));
        const float f0 = static_cast<float>(fm[n]), f1 = static_cast<float>(fm[n + 1]);
        const float f2 = static_cast<float>(fm[n + 2]), f3 = static_cast<float>(fm[n + 3]);
        *reinterpret_cast<float4*>(acc + n) = make_float4(0.f, 0.f, 0.f, 0.f);
        __stcs(reinterpret_cast<float4*>(out + n),
               make_float4(fmaf(a.x, r.x, f0), fmaf(a.y, r.y, f1), fmaf(a.z, r.z, f2), fmaf(a.w, r.w, f3)));
    } else {
        for (int e = n; e < min(n + 4, p.N); ++e) {
            const float a = acc[e];
            acc[e] = 0.f;
            out[e] = fmaf(a, __ldg(rs + e), static_cast<float>(fm[e]));
        }
    }
}

}  // namespace

int launch_gma_aggregate(const GmaAggParams& p, const CUtensorMap& tm_e, const CUtensorMap& tm_v, int num_sms,
                         cudaStream_t s) {
    GmaAggArgs args;
    args.tm_e = tm_e;
    args.tm_v = tm_v;
    args.p = p;
    const long long work = static_cast<long long>(p.P) * p.pair_tiles * p.k_blocks;
    const int grid = static_cast<int>(std::min<long long>(work, num_sms));
    SF_CUDA_CHECK(cudaFuncSetAttribute(gma_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    prof_before(SF_KERNEL_GMA_AGGREGATE, s);
    SF_CUDA_CHECK(launch_kernel(gma_aggregate_kernel, dim3(grid), dim3(kThreads), kSmemBytes, s, args));
    prof_after(SF_KERNEL_GMA_AGGREGATE, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

int launch_gma_finalize(const GmaAggParams& p, cudaStream_t s) {
    dim3 grid((p.N + 1023) / 1024, p.P * p.C);
    prof_before(SF_KERNEL_GMA_FINALIZE, s);
    switch (p.fmap_dtype) {
        case SF_DT_F32: SF_CUDA_CHECK(launch_kernel(gma_finalize_kernel<float>, grid, dim3(256), 0, s, p)); break;
        case SF_DT_F16: SF_CUDA_CHECK(launch_kernel(gma_finalize_kernel<__half>, grid, dim3(256), 0, s, p)); break;
        case SF_DT_BF16:
            SF_CUDA_CHECK(launch_kernel(gma_finalize_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, p));
            break;
        default: set_error("gma_finalize: unsupported dtype %d", p.fmap_dtype); return SF_ERR_INVALID;
    }
    prof_after(SF_KERNEL_GMA_FINALIZE, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
