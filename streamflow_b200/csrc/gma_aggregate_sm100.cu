// G3b: per-iteration GMA aggregation (core/gma.py:91-104) as ONE streaming tcgen05 GEMM with a fused epilogue:
//
//     out[p, c, n] = fmap[p, c, n] + (gamma / rowsum[p, n]) * sum_j E[p, n, j] * V[p, c, j]
//
// HBM-bound on the fp16 softmax numerators E (297 MB per Sintel clip and iteration).  Design points:
//   * one CTA owns TWO 128-query tiles (M = 256) per key block, so each 16 KB V tile fetched from L2 feeds two MMAs:
//     SM ingest is 1.5 B per E byte instead of 2 -- with M = 128 the kernel sat on the L2->SM bandwidth cap
//     (measured: 10.4 TB/s of L2 reads for 4.9 TB/s of HBM), not on HBM;
//   * stream-K: the linear (map, tile-pair, key-block) space is cut into gridDim equal contiguous ranges, so every SM
//     streams the same number of bytes (165 tiles over 148 SMs would otherwise quantise to 2 waves);
//   * E tiles are 16 KB contiguous blocks (tile-major layout written by gma_stats_kernel) on a deep mbarrier ring
//     (5 x 32 KB in flight per SM), V on a shallow one (3 x 16 KB);
//   * split tiles are reduced with the stream-K "last arriver" fix-up: every contributor parks its fp32 partial
//     (coalesced, channel-major) in a per-CTA slot, bumps the tile's counter, and the CTA that arrives last sums the
//     slots and runs the real epilogue -- residual add, gamma / rowsum scale and the NCHW store happen in-kernel,
//     so there is no accumulator zeroing, no atomics on the data and no separate finalize pass.
//
// warps: 0 = E producer (TMA), 1 = TMEM alloc + MMA issuer, 2 = V producer (TMA), 3-6 = epilogue (one TMEM lane
// quadrant each; thread = query row, loop over channels -> every global access is a coalesced 128 B line).
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
constexpr int kPartialFloats = 2 * BM * BN;                      // one parked partial: [128 ch][256 rows] fp32

struct GmaAggArgs {
    CUtensorMap tm_e, tm_v;
    GmaAggParams p;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ float to_float(T v) {
    return static_cast<float>(v);
}
template <>
__device__ __forceinline__ float to_float<__half>(__half v) {
    return __half2float(v);
}
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) {
    return __bfloat162float(v);
}

template <typename T>
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
    int* s_last = reinterpret_cast<int*>(tmem_slot + 1);

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
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

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
        const int et = threadIdx.x - 96;                       // 0..127 among the epilogue threads
        const T* fmap = reinterpret_cast<const T*>(p.fmap);
        int local = 0;
        long long pos = w_begin;
        auto owner = [&](long long u) { return ((u + 1) * G - 1) / work; };   // CTA whose range holds unit u
        while (pos < w_end) {
            const long long pt = pos / KB;
            const long long seg_end = min(w_end, (pt + 1) * KB);
            const int pb = static_cast<int>(pt / p.pair_tiles);
            const int mp = static_cast<int>(pt - static_cast<long long>(pb) * p.pair_tiles);
            const long long first_cta = owner(pt * KB), last_cta = owner((pt + 1) * KB - 1);
            const int contributors = static_cast<int>(last_cta - first_cta + 1);
            const int acc = local & 1;
            mbar_wait(&tfull[acc], (local >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * (2 * BN);
            const float* chan_base_dummy = nullptr;
            (void)chan_base_dummy;

            if (contributors == 1) {
                // whole tile pair accumulated here: epilogue straight from TMEM
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int n = (2 * mp + half) * BM + quad * 32 + lane;
                    const bool ok = n < p.N;
                    const float rs = ok ? __ldg(p.rscale + static_cast<long long>(pb) * p.N + n) : 0.f;
                    const long long base = static_cast<long long>(pb) * BN * p.N + n;
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
                        if (ok) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const long long idx = base + static_cast<long long>(ch * 32 + j) * p.N;
                                __stcs(p.out + idx, fmaf(__uint_as_float(v[j]), rs, to_float<T>(fmap[idx])));
                            }
                        }
                    }
                }
            } else {
                // park this CTA's partial: slot 0 if this is the first tile pair the CTA touches, else slot 1
                const int which = (w_begin / KB == pt) ? 0 : 1;
                float* mine = p.partials + (static_cast<long long>(blockIdx.x) * 2 + which) * kPartialFloats;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    float* dst = mine + half * BM + quad * 32 + lane;      // [channel][256 rows]: lanes contiguous
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
#pragma unroll
                        for (int j = 0; j < 32; ++j) __stcg(dst + (ch * 32 + j) * (2 * BM), __uint_as_float(v[j]));
                    }
                }
                // publish, and find out whether every other contributor has already published
                epi_bar_sync();
                if (et == 0) {
                    __threadfence();
                    const int prev = atomicAdd(p.counters + pt, 1);
                    const int last = (prev == contributors - 1);
                    if (last) p.counters[pt] = 0;                         // ready for the next launch
                    __threadfence();
                    *s_last = last;
                }
                epi_bar_sync();
                const int is_last = *s_last;
                epi_bar_sync();                                            // s_last may be rewritten next segment
                if (is_last) {
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        const int n = (2 * mp + half) * BM + quad * 32 + lane;
                        if (n >= p.N) continue;
                        const float rs = __ldg(p.rscale + static_cast<long long>(pb) * p.N + n);
                        const long long base = static_cast<long long>(pb) * BN * p.N + n;
                        const int roff = half * BM + quad * 32 + lane;
#pragma unroll 1
                        for (int c0 = 0; c0 < BN; c0 += 16) {
                            float sum[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) sum[j] = 0.f;
                            for (long long cta = first_cta; cta <= last_cta; ++cta) {
                                const int wh = ((work * cta / G) / KB == pt) ? 0 : 1;
                                const float* src = p.partials + (cta * 2 + wh) * kPartialFloats + roff;
#pragma unroll
                                for (int j = 0; j < 16; ++j) sum[j] += __ldcg(src + (c0 + j) * (2 * BM));
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const long long idx = base + static_cast<long long>(c0 + j) * p.N;
                                __stcs(p.out + idx, fmaf(sum[j], rs, to_float<T>(fmap[idx])));
                            }
                        }
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

}  // namespace

int gma_aggregate_max_ctas() { return 160; }
long long gma_aggregate_partial_bytes() {
    return static_cast<long long>(gma_aggregate_max_ctas()) * 2 * kPartialFloats * sizeof(float);
}

int launch_gma_aggregate(const GmaAggParams& p, const CUtensorMap& tm_e, const CUtensorMap& tm_v, int num_sms,
                         cudaStream_t s) {
    GmaAggArgs args;
    args.tm_e = tm_e;
    args.tm_v = tm_v;
    args.p = p;
    const long long work = static_cast<long long>(p.P) * p.pair_tiles * p.k_blocks;
    const int grid = static_cast<int>(std::min<long long>(work, std::min(num_sms, gma_aggregate_max_ctas())));
    auto launch = [&](auto kernel) -> int {
        SF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        prof_before(SF_KERNEL_GMA_AGGREGATE, s);
        kernel<<<grid, kThreads, kSmemBytes, s>>>(args);
        prof_after(SF_KERNEL_GMA_AGGREGATE, s);
        SF_CUDA_CHECK(cudaGetLastError());
        return SF_OK;
    };
    switch (p.fmap_dtype) {
        case SF_DT_F32: return launch(gma_aggregate_kernel<float>);
        case SF_DT_F16: return launch(gma_aggregate_kernel<__half>);
        case SF_DT_BF16: return launch(gma_aggregate_kernel<__nv_bfloat16>);
        default: set_error("gma_aggregate: unsupported dtype %d", p.fmap_dtype); return SF_ERR_INVALID;
    }
}

}  // namespace sf
