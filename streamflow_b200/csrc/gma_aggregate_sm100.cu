// G3b: per-iteration GMA aggregation (core/gma.py:91-104), one kernel, no split-K:
//
//     out[p, c, n] = fmap[p, c, n] + (gamma / rowsum[p, n]) * sum_j V[p, c, j] * E[p, n, j]
//
// HBM-bound on the fp16 softmax numerators E (297 MB per Sintel clip and iteration).  The GEMM is issued
// TRANSPOSED -- D[channel, query] = V[channel, key] . E[query, key]^T, M = 128 channels, N = queries -- because the
// UMMA N extent is any multiple of 16 up to 256: the query rows of all maps are cut into 16-row units and every CTA
// owns one contiguous range of units (the same count +-1 everywhere), for ALL key blocks.  Consequences:
//   * every SM streams the same number of E bytes without split-K: no fp32 scratch accumulator, no atomics, no
//     memset, no separate finalize pass, and the result is deterministic (one fp32 accumulator per output);
//   * the accumulator comes out of TMEM channel-major (lane = channel, column = query), which is the NCHW layout
//     of the result: the epilogue fuses the residual add and the softmax normalisation and writes `out` directly;
//   * E is tile-major: 16 KB blocks of 128 queries x 64 keys, each block stored by gma_stats_kernel as the
//     128B-swizzled K-major shared-memory image the UMMA descriptor expects, so a run of queries inside a block is
//     contiguous in HBM and lands with ONE non-tensor bulk copy (cp.async.bulk); a row range touches at most 3
//     blocks.  (Fetching the same rows as power-of-two TMA boxes cost 4-9 TMA instructions per key block and ran
//     at half the speed: per-instruction TMA cost, not bytes, was the limit.)
// Ring sizing: bulk copies of >= 16 KB saturate HBM with as few as 3 stages (scripts/probes/stream_probe.cu), so the
// E ring simply takes the shared memory that the 3-stage V ring (16 KB per key block, L2 resident, evict_last) and the
// prefetched fmap tile of the staged epilogue leave: 5 stages x 18 KB at Sintel size.  What does matter is the work
// split: a CTA that straddles two maps runs the key loop twice (twice the V traffic and iterations for the same E
// bytes) and was the tail of the whole kernel (98 us) until the split became per-map.
//
// warps: 0 = E producer (bulk copies), 1 = TMEM alloc + MMA issuer, 2 = V producer (TMA), 3-10 = epilogue.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "sf_internal.h"
#include "sm100_ptx.cuh"

namespace sf {

namespace {

constexpr int BK = 64, kCh = 128;
constexpr int kUnit = 16;                                   // query rows per work unit (UMMA N granularity at M=128)
constexpr int kMaxUnits = 16;                               // 256 rows = the UMMA N limit
constexpr int kMaxEStages = 12, kVStages = 3;
constexpr int kVBytes = kCh * BK * 2;                       // 16 KB
constexpr int kSmemBytes = 227 * 1024;                      // everything the SM has
constexpr int kBarBytes = 512;
constexpr int kRing = (kSmemBytes - 1024 /*align slack*/ - kBarBytes) / 1024 * 1024;
constexpr int kTmemCols = 512;                              // 2 accumulators x 256 columns
constexpr int kEpiWarps = 8;
constexpr int kThreads = (3 + kEpiWarps) * 32;

struct GmaAggArgs {
    CUtensorMap tm_v;
    GmaAggParams p;
    int units_per_map;              // ceil(N / 16)
    int e_stages, e_stage_bytes;    // E stage = the rows of the longest segment
    int fbuf_pitch;                 // > 0: the fmap tile of the CTA's single segment is prefetched into shared memory
    int fbuf_off, rbuf_off;         // byte offsets of that tile and of its rscale row inside the ring
};

struct Seg {
    int pb, row0, rows;
};

// next run of units inside one map, at most 256 rows
__device__ __forceinline__ bool next_seg(long long& u, long long u_end, int upm, Seg& s) {
    if (u >= u_end) return false;
    const int pb = static_cast<int>(u / upm);
    const int ul = static_cast<int>(u - static_cast<long long>(pb) * upm);
    const int n = min(min(kMaxUnits, upm - ul), static_cast<int>(u_end - u));
    s.pb = pb;
    s.row0 = ul * kUnit;
    s.rows = n * kUnit;
    u += n;
    return true;
}

template <typename T>
__device__ __forceinline__ void load_chunk(const T* fm, const float* rs, int n0, int N, bool vec, float (&f)[16],
                                           float (&r)[16]) {
    if constexpr (sizeof(T) == 4) {
        if (vec) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (n0 + 4 * j < N) {
                    a = *reinterpret_cast<const float4*>(fm + 4 * j);
                    b = __ldg(reinterpret_cast<const float4*>(rs + 4 * j));
                }
                f[4 * j] = a.x; f[4 * j + 1] = a.y; f[4 * j + 2] = a.z; f[4 * j + 3] = a.w;
                r[4 * j] = b.x; r[4 * j + 1] = b.y; r[4 * j + 2] = b.z; r[4 * j + 3] = b.w;
            }
            return;
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const bool ok = n0 + j < N;
        f[j] = ok ? static_cast<float>(fm[j]) : 0.f;
        r[j] = ok ? __ldg(rs + j) : 0.f;
    }
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// out[n0 .. n0+16) = fmap + acc * rscale for one channel; `vec`: N % 4 == 0, so float4 stores never straddle N
__device__ __forceinline__ void store_chunk(float* out, int n0, int N, bool vec, const uint32_t (&v)[16],
                                            const float (&f)[16], const float (&r)[16]) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (n0 + 4 * j < N)
                __stcs(reinterpret_cast<float4*>(out + n0 + 4 * j),
                       make_float4(fmaf(__uint_as_float(v[4 * j]), r[4 * j], f[4 * j]),
                                   fmaf(__uint_as_float(v[4 * j + 1]), r[4 * j + 1], f[4 * j + 1]),
                                   fmaf(__uint_as_float(v[4 * j + 2]), r[4 * j + 2], f[4 * j + 2]),
                                   fmaf(__uint_as_float(v[4 * j + 3]), r[4 * j + 3], f[4 * j + 3])));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (n0 + j < N) out[n0 + j] = fmaf(__uint_as_float(v[j]), r[j], f[j]);
    }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) gma_aggregate_kernel(const __grid_constant__ GmaAggArgs args) {
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic (not through an integer) so the compiler keeps the shared address space: STS / LDS
    // instead of generic ST / LD in the epilogue
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* e_base = smem;
    uint8_t* v_base = smem + kRing - kVStages * kVBytes;     // [E ring | fmap tile + rscale (optional) | V ring | barriers]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRing);
    uint64_t* e_full = bars;
    uint64_t* e_empty = e_full + kMaxEStages;
    uint64_t* v_full = e_empty + kMaxEStages;
    uint64_t* v_empty = v_full + kVStages;
    uint64_t* tfull = v_empty + kVStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const GmaAggParams& p = args.p;
    // warp index through a shuffle so the compiler knows the role dispatch is warp-uniform (see gma_sm100.cu)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int KB = p.k_blocks;
    const int upm = args.units_per_map;
    const int e_stages = args.e_stages;
    // Work split.  With at least one CTA per map, every map gets floor(G / P) or one more CTAs and each CTA a
    // contiguous run of that map's units: no CTA straddles two maps (a straddler would run the key loop twice --
    // twice the V traffic and iterations for the same E bytes -- and become the tail of the kernel).
    long long u_begin, u_end;
    {
        const int G = gridDim.x, bid = blockIdx.x;
        if (p.P <= G) {
            const int base = G / p.P, extra = G % p.P;
            int pb, j, g;
            if (bid < extra * (base + 1)) {
                pb = bid / (base + 1);
                j = bid - pb * (base + 1);
                g = base + 1;
            } else {
                const int b2 = bid - extra * (base + 1);
                pb = extra + b2 / base;
                j = b2 - (b2 / base) * base;
                g = base;
            }
            const long long m0 = static_cast<long long>(pb) * upm;
            u_begin = m0 + static_cast<long long>(upm) * j / g;
            u_end = m0 + static_cast<long long>(upm) * (j + 1) / g;
        } else {
            const long long U = static_cast<long long>(p.P) * upm;
            u_begin = U * bid / G;
            u_end = U * (bid + 1) / G;
        }
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tm_v);
        for (int i = 0; i < kMaxEStages; ++i) {
            mbar_init(&e_full[i], 1);
            mbar_init(&e_empty[i], 1);
        }
        for (int i = 0; i < kVStages; ++i) {
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], kEpiWarps);
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
        {                                                      // ---- E producer (warp-uniform loop, elected issue)
            int stage = 0;
            uint32_t phase = 0;
            long long u = u_begin;
            Seg s;
            while (next_seg(u, u_end, upm, s)) {
                // contiguous runs of the row range: one per 128-row block it touches (at most 3)
                const int end = s.row0 + s.rows;
                const int cut1 = min(end, (s.row0 | 127) + 1);
                const int cut2 = min(end, cut1 + 128);
                const __half* eb = p.e_ptr + static_cast<long long>(s.pb) * p.e_map_stride;
                const __half* src0 = eb + (static_cast<long long>(s.row0 >> 7) * KB * 128 + (s.row0 & 127)) * 64;
                const __half* src1 = eb + static_cast<long long>(cut1 >> 7) * KB * 128 * 64;
                const __half* src2 = eb + static_cast<long long>(cut2 >> 7) * KB * 128 * 64;
                const uint32_t bytes = static_cast<uint32_t>(s.rows) * 128u;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&e_empty[stage], phase ^ 1);
                    uint8_t* dst = e_base + stage * args.e_stage_bytes;
                    const long long blk = static_cast<long long>(kb) * 128 * 64;
                    if (elect_one()) {
                        mbar_expect_tx(&e_full[stage], bytes);
                        bulk_load_hint(dst, src0 + blk, (cut1 - s.row0) * 128, &e_full[stage], kEvictFirst);
                        if (cut1 < end)
                            bulk_load_hint(dst + (cut1 - s.row0) * 128, src1 + blk, (cut2 - cut1) * 128, &e_full[stage],
                                           kEvictFirst);
                        if (cut2 < end)
                            bulk_load_hint(dst + (cut2 - s.row0) * 128, src2 + blk, (end - cut2) * 128, &e_full[stage],
                                           kEvictFirst);
                    }
                    __syncwarp();
                    if (++stage == e_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 2) {
        {                                                      // ---- V producer
            int stage = 0;
            uint32_t phase = 0;
            long long u = u_begin;
            Seg s;
            while (next_seg(u, u_end, upm, s)) {
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&v_empty[stage], phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&v_full[stage], kVBytes);
                        tma_load_3d_hint(&args.tm_v, &v_full[stage], v_base + stage * kVBytes, kb * BK, 0, s.pb,
                                         kEvictLast);
                    }
                    __syncwarp();
                    if (++stage == kVStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        {                                                      // ---- MMA issuer: warp-uniform loop, one elected lane issues
            int es = 0, vs = 0, local = 0;
            uint32_t ephase = 0, vphase = 0;
            long long u = u_begin;
            Seg s;
            while (next_seg(u, u_end, upm, s)) {
                const int acc = local & 1;
                mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t idesc = make_idesc_f16_f32(kCh, s.rows);
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&v_full[vs], vphase);
                    mbar_wait(&e_full[es], ephase);
                    tc_fence_after();
                    const uint64_t da = make_kmajor_sw128_desc(smem_u32(v_base + vs * kVBytes));
                    const uint64_t db = make_kmajor_sw128_desc(smem_u32(e_base + es * args.e_stage_bytes));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        umma_commit(&e_empty[es]);
                        umma_commit(&v_empty[vs]);
                        if (kb == KB - 1) umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++es == e_stages) {
                        es = 0;
                        ephase ^= 1;
                    }
                    if (++vs == kVStages) {
                        vs = 0;
                        vphase ^= 1;
                    }
                }
                ++local;
            }
        }
    } else {                                                   // ---- epilogue (warps 3-10)
        const int quad = warp & 3;                             // TMEM lane quadrant this warp may read
        const int sub = (warp - 3) >> 2;                       // which half of the 16-column chunks
        const int ch = quad * 32 + lane;                       // channel = TMEM lane
        int local = 0;
        long long u = u_begin;
        Seg s;
        if (args.fbuf_pitch > 0) {
            // Single-segment CTA: the epilogue is the un-overlapped tail of the kernel, so everything it reads from
            // global memory (the fmap tile, its rscale row) is fetched into shared memory while the key loop runs.
            if (next_seg(u, u_end, upm, s)) {
                uint8_t* fbuf = smem + args.fbuf_off;
                float* rbuf = reinterpret_cast<float*>(smem + args.rbuf_off);
                const int t = threadIdx.x - 96;
                const int valid = min(s.rows, p.N - s.row0);
                const int cpr = valid * static_cast<int>(sizeof(T)) / 16;         // 16-byte chunks per channel row
                const T* fm0 = reinterpret_cast<const T*>(p.fmap) + static_cast<long long>(s.pb) * kCh * p.N + s.row0;
                for (int i = t; i < kCh * cpr; i += kEpiWarps * 32) {
                    const int c = i / cpr, k = i - c * cpr;
                    cp_async16(fbuf + c * args.fbuf_pitch + k * 16,
                               reinterpret_cast<const uint8_t*>(fm0 + static_cast<long long>(c) * p.N) + k * 16);
                }
                const float* rs = p.rscale + static_cast<long long>(s.pb) * p.N + s.row0;
                for (int i = t; i < valid / 4; i += kEpiWarps * 32) cp_async16(rbuf + 4 * i, rs + 4 * i);
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");

                float* out = p.out + (static_cast<long long>(s.pb) * kCh + ch) * p.N;
                const uint8_t* frow = fbuf + ch * args.fbuf_pitch;
                const int chunks = s.rows / 16;
                mbar_wait(&tfull[0], 0);
                tc_fence_after();
                const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
                for (int c = sub; c < chunks; c += 2) {
                    uint32_t v[16];
                    tmem_ld_32x16(t_acc + c * 16, v);
                    float f[16], r[16];
                    if (c * 16 < valid) {      // chunks past N are never stored
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b = *reinterpret_cast<const float4*>(rbuf + c * 16 + 4 * j);
                            r[4 * j] = b.x; r[4 * j + 1] = b.y; r[4 * j + 2] = b.z; r[4 * j + 3] = b.w;
                        }
                        if constexpr (sizeof(T) == 4) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 a = *reinterpret_cast<const float4*>(frow + c * 64 + 16 * j);
                                f[4 * j] = a.x; f[4 * j + 1] = a.y; f[4 * j + 2] = a.z; f[4 * j + 3] = a.w;
                            }
                        } else {
                            alignas(16) T h[16];
                            *reinterpret_cast<uint4*>(h) = *reinterpret_cast<const uint4*>(frow + c * 32);
                            *reinterpret_cast<uint4*>(h + 8) = *reinterpret_cast<const uint4*>(frow + c * 32 + 16);
#pragma unroll
                            for (int j = 0; j < 16; ++j) f[j] = static_cast<float>(h[j]);
                        }
                    }
                    tmem_ld_wait();
                    store_chunk(out, s.row0 + c * 16, p.N, true, v, f, r);
                }
            }
        } else {
        const bool vec = (p.N & 3) == 0;
        while (next_seg(u, u_end, upm, s)) {
            const int acc = local & 1;
            const int chunks = s.rows / 16;
            const long long base = (static_cast<long long>(s.pb) * kCh + ch) * p.N;
            const T* fm = reinterpret_cast<const T*>(p.fmap) + base;
            const float* rs = p.rscale + static_cast<long long>(s.pb) * p.N;
            float* out = p.out + base;
            float f[16], r[16];
            if (sub < chunks) load_chunk<T>(fm + s.row0 + sub * 16, rs + s.row0 + sub * 16, s.row0 + sub * 16, p.N, vec, f, r);
            mbar_wait(&tfull[acc], (local >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256;
#pragma unroll 1
            for (int c = sub; c < chunks; c += 2) {
                uint32_t v[16];
                tmem_ld_32x16(t_acc + c * 16, v);
                float fn[16], rn[16];
                const int n0 = s.row0 + c * 16;
                if (c + 2 < chunks) load_chunk<T>(fm + n0 + 32, rs + n0 + 32, n0 + 32, p.N, vec, fn, rn);
                tmem_ld_wait();
                store_chunk(out, n0, p.N, vec, v, f, r);
                if (c + 2 < chunks) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        f[j] = fn[j];
                        r[j] = rn[j];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            ++local;
        }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

template <typename T>
int launch_typed(const GmaAggArgs& args, int grid, cudaStream_t s) {
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gma_aggregate_kernel<T>), kSmemBytes)) return rc;
    prof_before(SF_KERNEL_GMA_AGGREGATE, s);
    SF_CUDA_CHECK(launch_kernel(gma_aggregate_kernel<T>, dim3(grid), dim3(kThreads), kSmemBytes, s, args));
    prof_after(SF_KERNEL_GMA_AGGREGATE, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace

int launch_gma_aggregate(const GmaAggParams& p, const CUtensorMap& tm_v, int num_sms, cudaStream_t s) {
    GmaAggArgs args;
    args.tm_v = tm_v;
    args.p = p;
    args.units_per_map = (p.N + kUnit - 1) / kUnit;
    const long long U = static_cast<long long>(p.P) * args.units_per_map;
    const int grid = static_cast<int>(std::min<long long>(U, num_sms));
    const long long per_cta = (p.P <= grid) ? (args.units_per_map + grid / p.P - 1) / (grid / p.P) : (U + grid - 1) / grid;
    const int seg_units = static_cast<int>(std::min<long long>(kMaxUnits, per_cta));
    args.e_stage_bytes = seg_units * kUnit * 128;
    const int elt = (p.fmap_dtype == SF_DT_F32) ? 4 : 2;
    const int rows_cap = seg_units * kUnit;
    int avail = kRing - kVStages * kVBytes;
    // staged epilogue: one segment per CTA, 16-byte aligned fmap rows, and room for at least 4 E stages next to the tile
    args.fbuf_pitch = 0;
    args.fbuf_off = args.rbuf_off = 0;
    const int fbuf_bytes = (kCh * (rows_cap * elt + 16) + rows_cap * 4 + 1023) / 1024 * 1024;
    // (one segment per CTA needs the per-map split, i.e. P <= grid: a linear split may straddle two maps)
    if (p.P <= grid && per_cta <= kMaxUnits && (static_cast<long long>(p.N) * elt) % 16 == 0 &&
        (reinterpret_cast<uintptr_t>(p.fmap) & 15) == 0 && avail - fbuf_bytes >= 4 * args.e_stage_bytes) {
        args.fbuf_pitch = rows_cap * elt + 16;
        avail -= fbuf_bytes;
        args.fbuf_off = avail;
        args.rbuf_off = avail + kCh * args.fbuf_pitch;
    }
    args.e_stages = std::min(kMaxEStages, avail / args.e_stage_bytes);
    switch (p.fmap_dtype) {
        case SF_DT_F32: return launch_typed<float>(args, grid, s);
        case SF_DT_F16: return launch_typed<__half>(args, grid, s);
        case SF_DT_BF16: return launch_typed<__nv_bfloat16>(args, grid, s);
        default: set_error("gma_aggregate: unsupported dtype %d", p.fmap_dtype); return SF_ERR_INVALID;
    }
}

}  // namespace sf
