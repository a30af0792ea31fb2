// G3: GMA softmax attention (core/gma.py:53-65) and motion-feature aggregation (core/gma.py:91-104) on
// tcgen05 tensor cores.
//
// Q and K do not change across refinement iterations, so the softmax numerator is computed ONCE per clip and
// kept in HBM as fp16 (E = 2^12 * exp(s - rowmax), tile-major 16 KB blocks, 99 MB per Sintel map -- trivial against 180 GB),
// exactly the matrix the reference's autocast path re-casts to fp16 every iteration (core/gma.py:95-97).
// Every iteration is then one streaming GEMM  acc = E . V^T  bound by reading E from HBM:
//
//   gma_stats_kernel      S = Q K^T tiles (128 x 256, K = d or 3d for hi/lo-split operands) in TMEM;
//                         pass 1 reduces the row max, pass 2 writes E with TMA stores and the row sums.
//   gma_aggregate_kernel  (gma_aggregate_sm100.cu) the per-iteration streaming GEMM with the fused epilogue.
#include <cuda_bf16.h>

#include "sf_internal.h"
#include "sm100_ptx.cuh"

namespace sf {

namespace {

constexpr float kLog2e = 1.4426950408889634f;
// E = 2^12 * exp(.): the scale keeps small weights out of the fp16 subnormals (folded into the exponent)

// 2^x with one MUFU.EX2 (rel. error 2^-22; exp2f() adds a range check and two rescaling multiplies per element,
// which made the pass-2 epilogue issue-bound at 13 instructions per logit)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ unsigned enc_ordered(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// =====================================================================================================
// stats: S = Q K^T
// =====================================================================================================
namespace st {
constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kMaxStages = 6, kMaxQBlocks = 4;          // K ring depth / resident Q k-blocks (hi and lo halves of d = 128)
constexpr int kABytes = BM * BK * 2, kBBytes = BN * BK * 2;
constexpr int kEpiWarps = 8;                            // two warps per TMEM lane quadrant, 128 columns each
constexpr int kEpiBuf = 32 * 128;                       // 32 rows x 64 fp16, one staging buffer per warp
constexpr int kEpiBytes = kEpiWarps * kEpiBuf;
constexpr int kSmemBytes = 227 * 1024;
constexpr int kBarBytes = 256;
constexpr int kRing = (kSmemBytes - 1024 - kBarBytes - kEpiBytes) / 1024 * 1024;     // Q blocks + K stages
constexpr int kTmemCols = 512;
constexpr int kThreads = 64 + 32 * kEpiWarps;
}  // namespace st

struct GmaStatsArgs {
    CUtensorMap tm_q, tm_k, tm_e;
    GmaStatsParams p;
    int stages;                     // K ring depth: whatever fits next to the resident Q blocks
};

__global__ void __launch_bounds__(st::kThreads, 1) gma_stats_kernel(const __grid_constant__ GmaStatsArgs args) {
    using namespace st;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // The 128-query Q tile of a work unit stays resident (one 16 KB block per 64 columns of K-depth, each with its own
    // full/empty barrier so the next unit's blocks stream in behind the last key tile); only K tiles go through the ring.
    uint8_t* q_base = smem;
    const int kStages = args.stages;
    uint8_t* epi_base = smem + kRing;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + kEpiBytes);
    uint64_t* full = bars;
    uint64_t* empty = full + kMaxStages;
    uint64_t* q_full = empty + kMaxStages;
    uint64_t* q_empty = q_full + kMaxQBlocks;
    uint64_t* tfull = q_empty + kMaxQBlocks;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const GmaStatsParams& p = args.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Operand schedule.  d-wide hi parts are `dblocks` 64-column blocks; with split operands (q = [hi | lo],
    // k = [hi | lo], Kp = 2d) the logit is hi.hi + lo.hi + hi.lo: the K tile streams as hi_0, lo_0, hi_1, lo_1, ... and
    // every K block is used while it sits in its stage -- hi_j against the resident Q_hi_j AND Q_lo_j, lo_j against
    // Q_hi_j -- so each K byte is fetched once per tile (the earlier [hi | hi' | lo] x [hi | lo | hi'] packing
    // streamed 3d columns for the same three products).  lo parts are stored unscaled: for |x| < 0.25 they are fp16
    // subnormals with 2^-25 absolute error, far below the 2^-12 relative error of an unsplit operand.
    // Pass 1 only needs an approximate row max (any m within a few units of the true max keeps exp() in range and
    // cancels in E / rowsum): it uses the hi parts alone.
    const bool split2 = (p.pass == 2) && p.split;
    const int dblocks = (p.split ? p.Kp / 2 : p.Kp) / BK;
    const int qblocks = split2 ? 2 * dblocks : dblocks;       // resident Q blocks; Q/K column of block b is b * BK
    const int ksteps = qblocks;                               // K blocks streamed per key tile
    uint8_t* stage_base = q_base + qblocks * kABytes;
    const int per_chunk = (p.n_tiles + p.chunks - 1) / p.chunks;
    const long long units = static_cast<long long>(p.P) * p.m_tiles * p.chunks;
    const long long u_begin = units * blockIdx.x / gridDim.x;
    const long long u_end = units * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tm_q);
        tma_prefetch_desc(&args.tm_k);
        tma_prefetch_desc(&args.tm_e);
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < kMaxQBlocks; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
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

    auto unit_coords = [&](long long u, int& pb, int& mt, int& nt0, int& nt1) {
        const int per_p = p.m_tiles * p.chunks;
        pb = static_cast<int>(u / per_p);
        const int r = static_cast<int>(u - static_cast<long long>(pb) * per_p);
        mt = r / p.chunks;
        const int ck = r - mt * p.chunks;
        nt0 = ck * per_chunk;
        nt1 = min(p.n_tiles, nt0 + per_chunk);
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, qphase = 0;
            for (long long u = u_begin; u < u_end; ++u) {
                int pb, mt, nt0, nt1;
                unit_coords(u, pb, mt, nt0, nt1);
                if (nt0 >= nt1) continue;               // empty key chunk: no Q load, no phase flip
                for (int nt = nt0; nt < nt1; ++nt)
                    for (int t = 0; t < ksteps; ++t) {
                        const int j = split2 ? (t >> 1) : t;
                        const bool is_lo = split2 && (t & 1);
                        if (nt == nt0 && !is_lo) {          // Q blocks first used by this step: free once the previous
                            for (int b = j; b < qblocks; b += dblocks) {       // unit's last key tile has consumed them
                                mbar_wait(&q_empty[b], qphase ^ 1);
                                mbar_expect_tx(&q_full[b], kABytes);
                                tma_load_3d(&args.tm_q, &q_full[b], q_base + b * kABytes, b * BK, mt * BM, pb);
                            }
                        }
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], kBBytes);
                        tma_load_3d(&args.tm_k, &full[stage], stage_base + stage * kBBytes,
                                    (is_lo ? dblocks + j : j) * BK, nt * BN, pb);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                qphase ^= 1;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16_f32(BM, BN);
            int stage = 0, local = 0;
            uint32_t phase = 0, qphase = 0;
            for (long long u = u_begin; u < u_end; ++u) {
                int pb, mt, nt0, nt1;
                unit_coords(u, pb, mt, nt0, nt1);
                if (nt0 >= nt1) continue;               // empty key chunk: no Q load, no phase flip
                for (int nt = nt0; nt < nt1; ++nt, ++local) {
                    const int acc = local & 1;
                    mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    for (int t = 0; t < ksteps; ++t) {
                        const int j = split2 ? (t >> 1) : t;
                        const bool is_lo = split2 && (t & 1);
                        const int n_a = (split2 && !is_lo) ? 2 : 1;        // K_hi_j meets Q_hi_j and Q_lo_j
                        if (nt == nt0 && !is_lo)
                            for (int b = j; b < qblocks; b += dblocks) mbar_wait(&q_full[b], qphase);
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint64_t db = make_kmajor_sw128_desc(smem_u32(stage_base + stage * kBBytes));
                        for (int a = 0; a < n_a; ++a) {
                            const uint64_t da = make_kmajor_sw128_desc(smem_u32(q_base + (j + a * dblocks) * kABytes));
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (t | a | k) != 0);
                        }
                        umma_commit(&empty[stage]);
                        if (nt == nt1 - 1) {                // last use of a resident Q block in this unit
                            if (!split2) umma_commit(&q_empty[j]);
                            else umma_commit(&q_empty[is_lo ? j : dblocks + j]);
                        }
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma_commit(&tfull[acc]);
                }
                qphase ^= 1;
            }
        }
    } else {
        const int e = warp - 2, quad = warp & 3, half = e >> 2;     // half: which 128 of the 256 key columns
        uint8_t* buf = epi_base + e * kEpiBuf;
        const int kbk = p.Npad / 64;                                // 64-key blocks per row of E
        int local = 0;
        for (long long u = u_begin; u < u_end; ++u) {
            int pb, mt, nt0, nt1;
            unit_coords(u, pb, mt, nt0, nt1);
            const int row = mt * BM + quad * 32 + lane;
            const bool row_ok = row < p.N;
            const long long ridx = static_cast<long long>(pb) * p.N + row;
            float run_max = -INFINITY, run_sum = 0.f, mrow = 0.f;
            // E = 2^(s*log2e - (rowmax*log2e - 12)): the 2^12 scale rides in the exponent
            if (p.pass == 2 && row_ok) mrow = dec_ordered(p.rowmax_bits[ridx]) * kLog2e - 12.0f;
            for (int nt = nt0; nt < nt1; ++nt, ++local) {
                const int acc = local & 1;
                mbar_wait(&tfull[acc], (local >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int cq = 0; cq < 2; ++cq) {            // this warp's two 64-key column groups
                    const int cp = half * 2 + cq;
                    uint32_t v0[32], v1[32];
                    const uint32_t taddr =
                        tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + cp * 64;
                    tmem_ld_32x32(taddr, v0);
                    tmem_ld_32x32(taddr + 32, v1);
                    tmem_ld_wait();
                    if (cq == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[acc]);
                    }
                    const int col0 = nt * BN + cp * 64;
                    if (col0 >= p.Npad) continue;
                    const bool full = col0 + 64 <= p.N;     // warp-uniform: no per-element key masking needed
                    if (p.pass == 1) {
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                run_max = fmaxf(run_max, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (col0 + j < p.N) run_max = fmaxf(run_max, __uint_as_float(v0[j]));
                                if (col0 + 32 + j < p.N) run_max = fmaxf(run_max, __uint_as_float(v1[j]));
                            }
                        }
                    } else {
                        // one staging buffer per warp: the previous chunk's store has had the whole tcgen05.ld + exp
                        // phase of this chunk to finish reading it
                        if (lane == 0) tma_store_wait_read<0>();
                        __syncwarp();
                        __half2 h[32];
                        float sum0 = 0.f, sum1 = 0.f;
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float a0 = ex2_approx(fmaf(__uint_as_float(v0[j]), kLog2e, -mrow));
                                const float a1 = ex2_approx(fmaf(__uint_as_float(v0[j + 1]), kLog2e, -mrow));
                                const float b0 = ex2_approx(fmaf(__uint_as_float(v1[j]), kLog2e, -mrow));
                                const float b1 = ex2_approx(fmaf(__uint_as_float(v1[j + 1]), kLog2e, -mrow));
                                h[j >> 1] = __floats2half2_rn(a0, a1);
                                h[16 + (j >> 1)] = __floats2half2_rn(b0, b1);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float a0 = (col0 + j < p.N)
                                    ? ex2_approx(fmaf(__uint_as_float(v0[j]), kLog2e, -mrow)) : 0.f;
                                const float a1 = (col0 + j + 1 < p.N)
                                    ? ex2_approx(fmaf(__uint_as_float(v0[j + 1]), kLog2e, -mrow)) : 0.f;
                                const float b0 = (col0 + 32 + j < p.N)
                                    ? ex2_approx(fmaf(__uint_as_float(v1[j]), kLog2e, -mrow)) : 0.f;
                                const float b1 = (col0 + 32 + j + 1 < p.N)
                                    ? ex2_approx(fmaf(__uint_as_float(v1[j + 1]), kLog2e, -mrow)) : 0.f;
                                h[j >> 1] = __floats2half2_rn(a0, a1);
                                h[16 + (j >> 1)] = __floats2half2_rn(b0, b1);
                            }
                        }
                        // row sum of the ROUNDED numerators (fp32 adds, two chains): sum_j E / rowsum == 1 for what is
                        // stored (summing the un-rounded values leaves up to 2^-11 of normalisation error on peaked rows)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float2 f = __half22float2(h[j]);
                            sum0 += f.x;
                            sum1 += f.y;
                        }
                        run_sum += sum0 + sum1;
#pragma unroll
                        for (int c16 = 0; c16 < 8; ++c16) {   // 8 x 16-byte chunks (8 halfs) per 128 B row
                            int4 o;
                            o.x = *reinterpret_cast<int*>(&h[c16 * 4 + 0]);
                            o.y = *reinterpret_cast<int*>(&h[c16 * 4 + 1]);
                            o.z = *reinterpret_cast<int*>(&h[c16 * 4 + 2]);
                            o.w = *reinterpret_cast<int*>(&h[c16 * 4 + 3]);
                            *reinterpret_cast<int4*>(buf + lane * 128 + ((c16 ^ (lane & 7)) << 4)) = o;
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            // E is tile-major: [P][m-tile][64-key block][128 rows][64 keys], 16 KB per tile
                            tma_store_3d(&args.tm_e, buf, 0, (mt * kbk + (col0 >> 6)) * BM + quad * 32, pb);
                            tma_store_commit();
                        }
                    }
                }
            }
            if (row_ok) {
                if (p.pass == 1)
                    atomicMax(p.rowmax_bits + ridx, enc_ordered(run_max));
                else   // fp32 sum of fp16 values: always a multiple of 2^-24, so the conversion is exact
                    atomicAdd(p.rowsum_fx + ridx, __float2ull_rn(run_sum * 16777216.0f));
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

__global__ void gma_rowsum_finish_kernel(const unsigned long long* fx, float* rowsum, long long n) {
    pdl_launch();
    pdl_wait();
    const long long i = blockIdx.x * 256ll + threadIdx.x;
    if (i < n) rowsum[i] = __ull2float_rn(fx[i]) * (1.0f / 16777216.0f);
}

}  // namespace

int launch_gma_rowsum_finish(const unsigned long long* fx, float* rowsum, long long n, cudaStream_t s) {
    SF_CUDA_CHECK(launch_kernel(gma_rowsum_finish_kernel, dim3(static_cast<unsigned>((n + 255) / 256)), dim3(256), 0, s,
                                fx, rowsum, n));
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

int launch_gma_stats(const GmaStatsParams& p, const CUtensorMap& tm_q, const CUtensorMap& tm_k,
                     const CUtensorMap& tm_e, int num_sms, cudaStream_t s) {
    GmaStatsArgs args;
    args.tm_q = tm_q;
    args.tm_k = tm_k;
    args.tm_e = tm_e;
    args.p = p;
    const int d = p.split ? p.Kp / 2 : p.Kp;
    SF_REQUIRE(d % st::BK == 0 && 2 * (d / st::BK) <= st::kMaxQBlocks, "gma_stats: head dimension %d not supported", d);
    const int qblocks = (p.pass == 2 && p.split) ? 2 * (d / st::BK) : d / st::BK;
    args.stages = std::min(st::kMaxStages, (st::kRing - qblocks * st::kABytes) / st::kBBytes);
    SF_CUDA_CHECK(cudaFuncSetAttribute(gma_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, st::kSmemBytes));
    const long long units = static_cast<long long>(p.P) * p.m_tiles * p.chunks;
    const int grid = static_cast<int>(std::min<long long>(units, num_sms));
    prof_before(SF_KERNEL_GMA_STATS, s);
    SF_CUDA_CHECK(launch_kernel(gma_stats_kernel, dim3(grid), dim3(st::kThreads), st::kSmemBytes, s, args));
    prof_after(SF_KERNEL_GMA_STATS, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
