// G3b: per-iteration GMA aggregation (core/gma.py:91-104), one streaming kernel, no split-K, no separate projection:
//
//     out[p, c, n] = fmap[p, c, n] + gamma * sum_c' W_v[c, c'] * ( sum_j X[p, c', j] * E[p, n, j] / rowsum[p, n] )
//
// i.e. the 1x1 `to_v` convolution is applied AFTER the attention-weighted sum (sum_j (W_v X)[c, j] E[n, j] =
// (W_v (X E^T))[c, n]): the kernel streams the fp16 motion features X themselves (NCHW = K-major over keys, written
// by gma_cast_kernel) through the big GEMM and applies W_v to the 128 x rows result with one extra small MMA per
// CTA.  That removes the per-iteration v-projection launch (10.8-14.7 us of mma.sync work and a grid-wide
// dependency in front of this kernel, 9 % of the round-1 step) and the fp16 rounding of V.
//
// HBM-bound on the fp16 softmax numerators E (297 MB per Sintel clip and iteration).  The GEMM is issued
// TRANSPOSED -- Y[channel, query] = X[channel, key] . E[query, key]^T, M = 128 channels, N = queries -- because the
// UMMA N extent is any multiple of 16 up to 256: the query rows of all maps are cut into 16-row units and every CTA
// owns one contiguous range of units (the same count +-1 everywhere), for ALL key blocks.  Consequences:
//   * every SM streams the same number of E bytes without split-K: no fp32 scratch accumulator, no atomics, no
//     memset, no separate finalize pass, and the result is deterministic (one fp32 accumulator per output);
//   * E is tile-major: 16 KB blocks of 128 queries x 64 keys, each block stored by gma_stats_kernel as the
//     128B-swizzled K-major shared-memory image the UMMA descriptor expects, so a run of queries inside a block is
//     contiguous in HBM and lands with ONE non-tensor bulk copy (cp.async.bulk); a row range touches at most 3
//     blocks.  (Fetching the same rows as power-of-two TMA boxes cost 4-9 TMA instructions per key block and ran
//     at half the speed: per-instruction TMA cost, not bytes, was the limit.)
// Epilogue (8 warps), per segment of <= 256 queries, in passes of `stage_rows` queries:
//   1. Y leaves TMEM channel-major (lane = c', column = query); each value is normalised by 1 / rowsum[query]
//      (a convex combination of X values: back in fp16 range), split into fp16 hi + lo (~22 bits) and written
//      TRANSPOSED into a 128B-swizzled K-major staging tile [query][c'] -- the B operand of the second GEMM;
//   2. one elected thread issues D2[c, query] = W_v[c, c'] . (hi + lo)[query, c']^T (16 tcgen05.mma of K = 16;
//      W_v fp16, resident in shared memory since kernel start) INTO THE SAME TMEM COLUMNS the pass just drained;
//   3. D2 leaves TMEM channel-major = NCHW, and `out = fmap + gamma * D2` is written directly.
// A CTA with a single segment (Sintel / KITTI: 8-10 units per CTA) stages whole segments in the idle E / X rings and
// prefetches its fmap tile there with cp.async while the second GEMM runs; multi-segment CTAs (Spring, batched
// KITTI) keep a dedicated 32 KB staging tile so the epilogue of segment i overlaps the key loop of segment i + 1.
// Ring sizing: bulk copies of >= 16 KB saturate HBM with as few as 3 stages (scripts/probes/stream_probe.cu).  What
// does matter is the work split: a CTA that straddles two maps runs the key loop twice (twice the X traffic and
// iterations for the same E bytes) and was the tail of the whole kernel (98 us) until the split became per-map.
//
// warps: 0 = E producer (bulk copies), 1 = TMEM alloc + MMA issuer, 2 = W_v + X producer (TMA), 3-10 = epilogue.
#include <cstdlib>

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
constexpr int kVBytes = kCh * BK * 2;                       // 16 KB: one key block of X (128 channels x 64 keys)
constexpr int kWBytes = 2 * kVBytes;                        // 32 KB: W_v [c][c'] fp16 as two 64-wide k-blocks
constexpr int kSmemBytes = 227 * 1024;                      // everything the SM has
constexpr int kBarBytes = 2048;                            // barriers + the 1 / rowsum row of the current segment
constexpr int kRing = (kSmemBytes - 1024 /*align slack*/ - kBarBytes) / 1024 * 1024;
constexpr int kTmemCols = 512;                              // 2 accumulators x 256 columns
constexpr int kEpiWarps = 8;
constexpr int kThreads = (3 + kEpiWarps) * 32;

struct GmaAggArgs {
    CUtensorMap tm_v;               // X16 [P, 128, Npad] fp16
    CUtensorMap tm_w;               // W_v [128, 128] fp16
    GmaAggParams p;
    int units_per_map;              // ceil(N / 16)
    int e_stages, e_stage_bytes;    // E stage = the rows of the longest segment
    int x_off, w_off;               // byte offsets of the X ring and of W_v inside the ring area
    int stage_off, stage_rows;      // staging tile of the second GEMM: 4 sub-tiles [hi|lo][k-block] of stage_rows x 128 B
    int fbuf_pitch;                 // > 0: single-segment CTAs; the fmap tile is prefetched into shared memory
    int fbuf_off;                   // byte offset of that tile inside the ring area
    int dbg;                        // STREAMCORR_AGG_DEBUG ablation bits (measurement only; results are wrong when set)
    // Early E stream.  E is written once per clip by the attention kernels and then only read, so from the SECOND
    // aggregate call on a handle it cannot depend on the kernel launched right before this one (the operand cast): the E
    // producer then skips griddepcontrol.wait and fills its ring (147 KB per SM) while the predecessor is still running.
    // `settled` (a word in the GMA workspace) is cleared by sf_gma_attention* and set to kSettledMagic by the end of
    // every aggregate launch; a launch that reads anything else -- the first after an attention call, whose stats kernels
    // may still be running under programmatic dependent launch, or a workspace the attention call never saw -- waits
    // like every other warp.
    unsigned* settled;
};
constexpr unsigned kSettledMagic = 0x5E771ED1u;

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

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

// 16 fmap values of one channel (zero past N)
template <typename T>
__device__ __forceinline__ void load_chunk(const T* fm, int n0, int N, bool vec, float (&f)[16]) {
    if constexpr (sizeof(T) == 4) {
        if (vec) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + 4 * j < N) a = *reinterpret_cast<const float4*>(fm + 4 * j);
                f[4 * j] = a.x; f[4 * j + 1] = a.y; f[4 * j + 2] = a.z; f[4 * j + 3] = a.w;
            }
            return;
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = (n0 + j < N) ? static_cast<float>(fm[j]) : 0.f;
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// out[n0 .. n0+16) = fmap + gamma * acc for one channel; `vec`: N % 4 == 0, so float4 stores never straddle N
__device__ __forceinline__ void store_chunk(float* out, int n0, int N, bool vec, const uint32_t (&v)[16],
                                            const float (&f)[16], float g) {
    if (vec) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (n0 + 4 * j < N)
                __stcs(reinterpret_cast<float4*>(out + n0 + 4 * j),
                       make_float4(fmaf(__uint_as_float(v[4 * j]), g, f[4 * j]),
                                   fmaf(__uint_as_float(v[4 * j + 1]), g, f[4 * j + 1]),
                                   fmaf(__uint_as_float(v[4 * j + 2]), g, f[4 * j + 2]),
                                   fmaf(__uint_as_float(v[4 * j + 3]), g, f[4 * j + 3])));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (n0 + j < N) out[n0 + j] = fmaf(__uint_as_float(v[j]), g, f[j]);
    }
}

__device__ __forceinline__ void epi_bar_sync() {       // the 8 epilogue warps only
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) gma_aggregate_kernel(const __grid_constant__ GmaAggArgs args) {
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic (not through an integer) so the compiler keeps the shared address space: STS / LDS
    // instead of generic ST / LD in the epilogue
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* e_base = smem;                                  // [E ring | (staging tile) | X ring | W_v | barriers]
    uint8_t* v_base = smem + args.x_off;
    uint8_t* w_base = smem + args.w_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRing);
    uint64_t* e_full = bars;
    uint64_t* e_empty = e_full + kMaxEStages;
    uint64_t* v_full = e_empty + kMaxEStages;
    uint64_t* v_empty = v_full + kVStages;
    uint64_t* tfull = v_empty + kVStages;
    uint64_t* tempty = tfull + 2;
    uint64_t* w_full = tempty + 2;
    uint64_t* d2_full = w_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_full + 1);
    float* gamma_s = reinterpret_cast<float*>(tmem_slot + 1);
    float* rinv_s = reinterpret_cast<float*>(smem + kRing + 1024);     // [256] 1 / rowsum of the current segment

    const GmaAggParams& p = args.p;
    // warp index through a shuffle so the compiler knows the role dispatch is warp-uniform (see gma_sm100.cu)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int KB = p.k_blocks;
    const int upm = args.units_per_map;
    const int e_stages = args.e_stages;
    // Work split.  With at least one CTA per map, every map gets floor(G / P) or one more CTAs and each CTA a
    // contiguous run of that map's units: no CTA straddles two maps (a straddler would run the key loop twice --
    // twice the X traffic and iterations for the same E bytes -- and become the tail of the kernel).
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
        tma_prefetch_desc(&args.tm_w);
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
        mbar_init(w_full, 1);
        mbar_init(d2_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    pdl_launch();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // every role but (possibly) the E producer depends on the preceding kernel (the fp16 cast of X, the fmap itself)
    bool e_early = false;
    if (warp == 0) {
        unsigned st = 0;
        if (lane == 0 && args.settled != nullptr) st = ld_acquire_gpu_u32(args.settled);
        e_early = __shfl_sync(0xffffffffu, st, 0) == kSettledMagic;
    }
    if (!e_early) pdl_wait();

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
        {                                                      // ---- W_v (once) + X producer
            if (elect_one()) {
                mbar_expect_tx(w_full, kWBytes);
                tma_load_3d_hint(&args.tm_w, w_full, w_base, 0, 0, 0, kEvictLast);
                tma_load_3d_hint(&args.tm_w, w_full, w_base + kVBytes, BK, 0, 0, kEvictLast);
            }
            __syncwarp();
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
        const int ch = quad * 32 + lane;                       // channel = TMEM lane (c' of Y, then c of D2)
        const int et = static_cast<int>(threadIdx.x) - 96;     // 0 .. 255
        if (et == 0) *gamma_s = __ldg(p.gamma);                // ONE read of the scalar per CTA (visible after epi_bar_sync)
        const bool single = args.fbuf_pitch > 0;
        const bool vec = (p.N & 3) == 0;
        uint8_t* stage = smem + args.stage_off;
        const int SR = args.stage_rows;
        const uint32_t sub_bytes = static_cast<uint32_t>(SR) * 128u;      // one [query][64 x c'] sub-tile
        // staging address of (row, ch): 128-byte rows, 16-byte chunk (ch & 63) >> 3 of row r stored at chunk ^ (r & 7)
        const uint32_t st_kb = static_cast<uint32_t>(ch >> 6) * sub_bytes;
        const uint32_t st_chunk = static_cast<uint32_t>((ch & 63) >> 3), st_in = static_cast<uint32_t>(ch & 7) * 2u;
        uint32_t d2_phase = 0;
        bool w_ready = false;
        int local = 0;
        long long u = u_begin;
        Seg s;
        while (next_seg(u, u_end, upm, s)) {
            const int acc = local & 1;
            const int chunks = s.rows / 16;
            const long long base = (static_cast<long long>(s.pb) * kCh + ch) * p.N;
            const T* fm = reinterpret_cast<const T*>(p.fmap) + base;
            const float* rsum = p.rowsum + static_cast<long long>(s.pb) * p.N;
            float* out = p.out + base;
            uint8_t* fbuf = smem + args.fbuf_off;
            // While the key loop of this segment runs (the epilogue warps have nothing else to do): 1 / rowsum of its
            // queries into shared memory, and the lines of its fmap tile into L2, so that neither is a DRAM round trip
            // in the un-overlapped tail.  (rinv_s was last read before the previous segment's final epi_bar_sync.)
            if (et < s.rows) {
                const int n = s.row0 + et;
                rinv_s[et] = (n < p.N && !(args.dbg & 1)) ? __frcp_rn(__ldg(rsum + n)) : ((args.dbg & 1) ? 1.f : 0.f);
            }
            if (!(args.dbg & 8)) {
                const int valid = min(s.rows, p.N - s.row0);
                const int lines = (valid * static_cast<int>(sizeof(T)) + 127) / 128;      // 128-byte lines per channel row
                const uint8_t* fm0 = reinterpret_cast<const uint8_t*>(
                    reinterpret_cast<const T*>(p.fmap) + static_cast<long long>(s.pb) * kCh * p.N + s.row0);
                for (int i = et; i < kCh * lines; i += kEpiWarps * 32) {
                    const int c = i / lines, k = i - c * lines;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(fm0 + static_cast<long long>(c) * p.N * sizeof(T) + k * 128));
                }
            }
            epi_bar_sync();
            mbar_wait(&tfull[acc], (local >> 1) & 1);          // every MMA of the key loop is done: the rings are idle
            tc_fence_after();
            if (single && !(args.dbg & 4)) {
                // the un-overlapped tail of the kernel: fetch the fmap tile into the idle ring while the second GEMM runs
                const int valid = min(s.rows, p.N - s.row0);
                const int cpr = valid * static_cast<int>(sizeof(T)) / 16;         // 16-byte chunks per channel row
                const uint8_t* fm0 = reinterpret_cast<const uint8_t*>(
                    reinterpret_cast<const T*>(p.fmap) + static_cast<long long>(s.pb) * kCh * p.N + s.row0);
                const long long row_bytes = static_cast<long long>(p.N) * sizeof(T);
                // chunk i = et, et + 256, ... of the [128 channels][cpr] tile, (c, k) advanced without divisions
                const int dc = (kEpiWarps * 32) / cpr, dk = (kEpiWarps * 32) - dc * cpr;
                int c = et / cpr, k = et - c * cpr;
                while (c < kCh) {
                    cp_async16(fbuf + c * args.fbuf_pitch + k * 16, fm0 + c * row_bytes + k * 16);
                    c += dc;
                    k += dk;
                    if (k >= cpr) {
                        k -= cpr;
                        ++c;
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 256;
#pragma unroll 1
            for (int q0 = 0; q0 < s.rows && !(args.dbg & 2); q0 += SR) {      // ---- second GEMM, `SR` queries per pass
                const int w = min(SR, s.rows - q0);
#pragma unroll 1
                for (int cc = sub; cc < w / 16; cc += 2) {
                    uint32_t v[16];
                    tmem_ld_32x16(t_acc + q0 + cc * 16, v);
                    float ri[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 r4 = *reinterpret_cast<const float4*>(rinv_s + q0 + cc * 16 + 4 * j);
                        ri[4 * j] = r4.x; ri[4 * j + 1] = r4.y; ri[4 * j + 2] = r4.z; ri[4 * j + 3] = r4.w;
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // rows past N carry rinv = 0; their E rows may be uninitialised, so select instead of multiply
                        const float y = (ri[j] != 0.f) ? __uint_as_float(v[j]) * ri[j] : 0.f;
                        const __half hi = __float2half_rn(y);
                        const __half lo = __float2half_rn(y - __half2float(hi));
                        const uint32_t r = static_cast<uint32_t>(cc * 16 + j);
                        const uint32_t off = st_kb + r * 128u + ((st_chunk ^ (r & 7u)) << 4) + st_in;
                        *reinterpret_cast<__half*>(stage + off) = hi;
                        *reinterpret_cast<__half*>(stage + 2u * sub_bytes + off) = lo;
                    }
                }
                fence_proxy_async_smem();                       // staging writes -> visible to the tensor core (async proxy)
                tc_fence_before();                              // the TMEM reads of this pass precede the MMA that overwrites them
                epi_bar_sync();
                if (warp == 3) {
                    if (!w_ready) {
                        mbar_wait(w_full, 0);
                        w_ready = true;
                    }
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t idesc = make_idesc_f16_f32(kCh, w);
                        const uint32_t d_tmem = tmem_base + acc * 256 + q0;
#pragma unroll
                        for (int part = 0; part < 2; ++part)            // hi, lo
#pragma unroll
                            for (int kb2 = 0; kb2 < 2; ++kb2) {         // c' 0-63, 64-127
                                const uint64_t da = make_kmajor_sw128_desc(smem_u32(w_base + kb2 * kVBytes));
                                const uint64_t db =
                                    make_kmajor_sw128_desc(smem_u32(stage + (part * 2 + kb2) * sub_bytes));
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (part | kb2 | k) != 0);
                            }
                        umma_commit(d2_full);
                    }
                    __syncwarp();
                }
                mbar_wait(d2_full, d2_phase);                   // D2 columns written, staging tile free again
                d2_phase ^= 1;
                tc_fence_after();
            }
            const float g = *gamma_s;
            if (single) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                epi_bar_sync();                                 // every thread's part of the fmap tile has landed
                const uint8_t* frow = fbuf + ch * args.fbuf_pitch;
                const int valid = min(s.rows, p.N - s.row0);
#pragma unroll 1
                for (int c = sub; c < chunks; c += 2) {
                    uint32_t v[16];
                    tmem_ld_32x16(t_acc + c * 16, v);
                    float f[16];
                    if (c * 16 < valid) {      // chunks past N are never stored
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
                    store_chunk(out, s.row0 + c * 16, p.N, true, v, f, g);
                }
            } else {
                float f[16];
                if (sub < chunks) load_chunk<T>(fm + s.row0 + sub * 16, s.row0 + sub * 16, p.N, vec, f);
#pragma unroll 1
                for (int c = sub; c < chunks; c += 2) {
                    uint32_t v[16];
                    tmem_ld_32x16(t_acc + c * 16, v);
                    float fn[16];
                    const int n0 = s.row0 + c * 16;
                    if (c + 2 < chunks) load_chunk<T>(fm + n0 + 32, n0 + 32, p.N, vec, fn);
                    tmem_ld_wait();
                    store_chunk(out, n0, p.N, vec, v, f, g);
                    if (c + 2 < chunks) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) f[j] = fn[j];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            ++local;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && args.settled != nullptr) *args.settled = kSettledMagic;
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

int launch_gma_aggregate(const GmaAggParams& p, const CUtensorMap& tm_x, const CUtensorMap& tm_w, int num_sms,
                         unsigned* settled, cudaStream_t s) {
    GmaAggArgs args;
    args.tm_v = tm_x;
    args.tm_w = tm_w;
    args.p = p;
    static const bool want_early = [] {
        const char* e = getenv("STREAMCORR_AGG_EARLY_E");
        return !(e && e[0] == '0');
    }();
    args.settled = want_early ? settled : nullptr;
    args.units_per_map = (p.N + kUnit - 1) / kUnit;
    const long long U = static_cast<long long>(p.P) * args.units_per_map;
    const int grid = static_cast<int>(std::min<long long>(U, num_sms));
    const long long per_cta = (p.P <= grid) ? (args.units_per_map + grid / p.P - 1) / (grid / p.P) : (U + grid - 1) / grid;
    const int seg_units = static_cast<int>(std::min<long long>(kMaxUnits, per_cta));
    args.e_stage_bytes = seg_units * kUnit * 128;
    const int elt = (p.fmap_dtype == SF_DT_F32) ? 4 : 2;
    const int rows_cap = seg_units * kUnit;
    args.w_off = kRing - kWBytes;
    args.x_off = args.w_off - kVStages * kVBytes;
    // Single-segment CTAs (one segment per CTA needs the per-map split, i.e. P <= grid, and 16-byte aligned fmap rows):
    // after the key loop the E and X rings are idle, so the staging tile of the second GEMM (whole segment if it fits)
    // and the prefetched fmap tile live there and the E ring keeps every byte during the loop.
    args.fbuf_pitch = 0;
    args.fbuf_off = 0;
    bool single = p.P <= grid && per_cta <= kMaxUnits && (static_cast<long long>(p.N) * elt) % 16 == 0 &&
                  (reinterpret_cast<uintptr_t>(p.fmap) & 15) == 0;
    int e_avail = 0;
    if (single) {
        const int pitch = rows_cap * elt + 16;
        const int fbuf_bytes = kCh * pitch;
        int sr = rows_cap;
        while (sr > 16 && 512 * sr + fbuf_bytes > args.w_off) sr -= 16;
        if (512 * sr + fbuf_bytes > args.w_off) {
            single = false;
        } else {
            args.stage_rows = sr;
            args.stage_off = 0;
            args.fbuf_pitch = pitch;
            args.fbuf_off = 512 * sr;
            e_avail = args.x_off;
        }
    }
    if (!single) {      // dedicated staging tile: the epilogue of one segment overlaps the key loop of the next
        args.stage_rows = 32;
        args.stage_off = args.x_off - 512 * args.stage_rows;
        e_avail = args.stage_off;
    }
    static const int dbg = [] {
        const char* e = getenv("STREAMCORR_AGG_DEBUG");
        return e ? atoi(e) : 0;
    }();
    args.dbg = dbg;
    args.e_stages = std::min(kMaxEStages, e_avail / args.e_stage_bytes);
    SF_REQUIRE(args.e_stages >= 2, "gma_aggregate: internal error, E ring of %d stages", args.e_stages);
    switch (p.fmap_dtype) {
        case SF_DT_F32: return launch_typed<float>(args, grid, s);
        case SF_DT_F16: return launch_typed<__half>(args, grid, s);
        case SF_DT_BF16: return launch_typed<__nv_bfloat16>(args, grid, s);
        default: set_error("gma_aggregate: unsupported dtype %d", p.fmap_dtype); return SF_ERR_INVALID;
    }
}

}  // namespace sf
