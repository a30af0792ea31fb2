// G3a: GMA softmax attention statistics (core/gma.py:53-65) on tcgen05 tensor cores, once per clip.
//
// Q and K do not change across refinement iterations, so the softmax numerator is computed ONCE per clip and
// kept in HBM as fp16 (E = 2^12 * exp(s - rowmax), tile-major 16 KB blocks, 99 MB per Sintel map -- trivial against
// 180 GB), exactly the matrix the reference's autocast path re-casts to fp16 every iteration (core/gma.py:95-97).
// Every iteration is then one streaming GEMM acc = E . V^T (gma_aggregate_sm100.cu).
//
// gma_stats_kernel computes S = Q K^T with hi/lo-split fp16 operands (K = 3d, fp32-faithful logits):
//   pass 1  row max (hi parts only: any m within a few units of the true max keeps exp() in range and cancels in
//           E / rowsum);
//   pass 2  E = fp16(2^(s*log2e - m*log2e + 12)) written straight from registers (a thread's 64 logits are exactly
//           one 128-byte row of a tile-major E tile) and the row sums of the rounded values.
//
// Tiling: a CTA keeps TWO 128-query tiles of Q resident in shared memory (2 x 96 KB for K = 384) and streams 64-key
// K blocks (8 KB per k-block, 4-stage mbarrier ring), so every K byte fetched from L2 feeds two MMAs.  The first
// version streamed Q and K per 128 x 256 tile: 1.33 GB of L2 reads per pass = the L2->SM bandwidth cap (133 us
// measured, tensor pipe 50 % idle).  Accumulators: 2 tiles x 64 columns, double buffered in TMEM.
// warps: 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2-9 = epilogue (TMEM lane quadrant x query tile).
#include <cuda_bf16.h>

#include "sf_internal.h"
#include "sm100_ptx.cuh"

namespace sf {

namespace {

constexpr float kLog2e = 1.4426950408889634f;

// 2^x with one MUFU.EX2 (rel. error 2^-22; exp2f() adds a range check and two rescaling multiplies per element)
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

constexpr int BM = 128, BN = 64, BK = 64;
constexpr int kMaxKBlocks = 6;                          // Kp <= 384
constexpr int kStages = 4;
constexpr int kQTileBytes = BM * BK * 2;                // 16 KB: one query tile, one k-block
constexpr int kKTileBytes = BN * BK * 2;                // 8 KB
constexpr int kQBytes = 2 * kMaxKBlocks * kQTileBytes;  // 192 KB resident
constexpr int kSmemBytes = kQBytes + kStages * kKTileBytes + 1024 + 256;
constexpr int kTmemCols = 256;                          // 2 buffers x (2 query tiles x 64 columns)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;

struct GmaStatsArgs {
    CUtensorMap tm_q, tm_k;
    GmaStatsParams p;
};

__global__ void __launch_bounds__(kThreads, 1) gma_stats_kernel(const __grid_constant__ GmaStatsArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_base = smem;                               // [2 tiles][kblocks][128 x 64] fp16, 128B-swizzled
    uint8_t* k_base = smem + kQBytes;                     // ring of [64 x 64] fp16
    uint64_t* bars = reinterpret_cast<uint64_t*>(k_base + kStages * kKTileBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tfull = bars + 2 * kStages;
    uint64_t* tempty = bars + 2 * kStages + 2;
    uint64_t* qfull = bars + 2 * kStages + 4;
    uint64_t* qempty = bars + 2 * kStages + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 6);

    const GmaStatsParams& p = args.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pass 1 only needs an approximate row max: hi parts alone (first d columns of the split operands)
    const int kblocks = (p.pass == 1 ? (p.Kp / 3 + BK - 1) / BK : (p.Kp + BK - 1) / BK);
    const int per_chunk = (p.n_tiles + p.chunks - 1) / p.chunks;       // 64-key tiles per unit
    const long long units = static_cast<long long>(p.P) * p.pair_tiles * p.chunks;
    const long long u_begin = units * blockIdx.x / gridDim.x;
    const long long u_end = units * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tm_q);
        tma_prefetch_desc(&args.tm_k);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], kEpiWarps);
        }
        mbar_init(qfull, 1);
        mbar_init(qempty, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
    pdl_launch();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // unit -> (map, query-tile pair, range of 64-key tiles); units of one pair are adjacent
    auto unit_coords = [&](long long u, int& pb, int& mp, int& nt0, int& nt1) {
        const int per_p = p.pair_tiles * p.chunks;
        pb = static_cast<int>(u / per_p);
        const int r = static_cast<int>(u - static_cast<long long>(pb) * per_p);
        mp = r / p.chunks;
        const int ck = r - mp * p.chunks;
        nt0 = ck * per_chunk;
        nt1 = min(p.n_tiles, nt0 + per_chunk);
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, qphase = 0;
            long long cur_pair = -1;
            for (long long u = u_begin; u < u_end; ++u) {
                int pb, mp, nt0, nt1;
                unit_coords(u, pb, mp, nt0, nt1);
                const long long pair_id = static_cast<long long>(pb) * p.pair_tiles + mp;
                if (pair_id != cur_pair) {                 // (re)load the two resident Q tiles
                    mbar_wait(qempty, qphase ^ 1);         // all MMAs that read the previous Q have completed
                    mbar_expect_tx(qfull, 2 * kblocks * kQTileBytes);
                    for (int half = 0; half < 2; ++half)
                        for (int kb = 0; kb < kblocks; ++kb)   // rows past N (and a tile past the last) are zero-filled
                            tma_load_3d(&args.tm_q, qfull, q_base + (half * kMaxKBlocks + kb) * kQTileBytes, kb * BK,
                                        (2 * mp + half) * BM, pb);
                    cur_pair = pair_id;
                    qphase ^= 1;
                }
                for (int nt = nt0; nt < nt1; ++nt)
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], kKTileBytes);
                        tma_load_3d(&args.tm_k, &full[stage], k_base + stage * kKTileBytes, kb * BK, nt * BN, pb);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16_f32(BM, BN);
            int stage = 0, local = 0;
            uint32_t phase = 0, qphase = 0;
            long long cur_pair = -1;
            for (long long u = u_begin; u < u_end; ++u) {
                int pb, mp, nt0, nt1;
                unit_coords(u, pb, mp, nt0, nt1);
                const long long pair_id = static_cast<long long>(pb) * p.pair_tiles + mp;
                if (pair_id != cur_pair) {
                    if (cur_pair >= 0) umma_commit(qempty);    // previous Q tiles are free once issued MMAs finish
                    mbar_wait(qfull, qphase);
                    tc_fence_after();
                    cur_pair = pair_id;
                    qphase ^= 1;
                }
                for (int nt = nt0; nt < nt1; ++nt, ++local) {
                    const int acc = local & 1;
                    mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * (2 * BN);
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint64_t db = make_kmajor_sw128_desc(smem_u32(k_base + stage * kKTileBytes));
                        const uint64_t d0 = make_kmajor_sw128_desc(smem_u32(q_base + kb * kQTileBytes));
                        const uint64_t d1 = make_kmajor_sw128_desc(smem_u32(q_base + (kMaxKBlocks + kb) * kQTileBytes));
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            umma_f16_ss(d_tmem, d0 + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                            umma_f16_ss(d_tmem + BN, d1 + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit(&empty[stage]);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    umma_commit(&tfull[acc]);
                }
            }
        }
    } else {
        const int quad = warp & 3, half = (warp - 2) >> 2;           // TMEM lane quadrant, query tile of the pair
        const int kbk = p.Npad / 64;                                 // 64-key blocks per E row
        int local = 0;
        for (long long u = u_begin; u < u_end; ++u) {
            int pb, mp, nt0, nt1;
            unit_coords(u, pb, mp, nt0, nt1);
            const int mt = 2 * mp + half;
            const int row = mt * BM + quad * 32 + lane;
            const bool row_ok = row < p.N;
            const bool tile_ok = mt < p.m_tiles;                     // second tile of an odd last pair does not exist
            const long long ridx = static_cast<long long>(pb) * p.N + row;
            float run_max = -INFINITY, sum0 = 0.f, sum1 = 0.f, mrow = 0.f;
            // E = 2^(s*log2e - (rowmax*log2e - 12)): the 2^12 scale (keeps small weights out of the fp16
            // subnormals) rides in the exponent
            if (p.pass == 2 && row_ok) mrow = dec_ordered(p.rowmax_bits[ridx]) * kLog2e - 12.0f;
            for (int nt = nt0; nt < nt1; ++nt, ++local) {
                const int acc = local & 1;
                mbar_wait(&tfull[acc], (local >> 1) & 1);
                tc_fence_after();
                uint32_t v0[32], v1[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * (2 * BN) + half * BN;
                tmem_ld_32x32(taddr, v0);
                tmem_ld_32x32(taddr + 32, v1);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                const int col0 = nt * BN;
                if (!tile_ok || col0 >= p.Npad) continue;
                const bool fullcols = col0 + BN <= p.N;              // warp-uniform: no per-element key masking
                if (p.pass == 1) {
                    if (fullcols) {
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
                    __half2 h[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float a0 = ex2_approx(fmaf(__uint_as_float(v0[j]), kLog2e, -mrow));
                        float a1 = ex2_approx(fmaf(__uint_as_float(v0[j + 1]), kLog2e, -mrow));
                        float b0 = ex2_approx(fmaf(__uint_as_float(v1[j]), kLog2e, -mrow));
                        float b1 = ex2_approx(fmaf(__uint_as_float(v1[j + 1]), kLog2e, -mrow));
                        if (!fullcols) {                             // pad keys (zero-filled K rows) must store 0
                            a0 = (col0 + j < p.N) ? a0 : 0.f;
                            a1 = (col0 + j + 1 < p.N) ? a1 : 0.f;
                            b0 = (col0 + 32 + j < p.N) ? b0 : 0.f;
                            b1 = (col0 + 32 + j + 1 < p.N) ? b1 : 0.f;
                        }
                        h[j >> 1] = __floats2half2_rn(a0, a1);
                        h[16 + (j >> 1)] = __floats2half2_rn(b0, b1);
                    }
                    // row sum of the ROUNDED numerators (fp32 adds, two chains): sum_j E / rowsum == 1 for what is
                    // stored (summing the un-rounded values leaves up to 2^-11 of normalisation error on peaked rows)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float2 f = __half22float2(h[j]);
                        sum0 += f.x;
                        sum1 += f.y;
                    }
                    // tile-major E: [P][m-tile][64-key block][128 rows][64 keys]; this thread owns one 128 B row
                    int4* dst = reinterpret_cast<int4*>(
                        p.E + ((static_cast<long long>(pb) * p.m_tiles + mt) * kbk + nt) * (BM * 64) +
                        (quad * 32 + lane) * 64);
#pragma unroll
                    for (int c16 = 0; c16 < 8; ++c16) {
                        int4 o;
                        o.x = *reinterpret_cast<int*>(&h[c16 * 4 + 0]);
                        o.y = *reinterpret_cast<int*>(&h[c16 * 4 + 1]);
                        o.z = *reinterpret_cast<int*>(&h[c16 * 4 + 2]);
                        o.w = *reinterpret_cast<int*>(&h[c16 * 4 + 3]);
                        __stcs(dst + c16, o);
                    }
                }
            }
            if (row_ok && tile_ok) {
                if (p.pass == 1)
                    atomicMax(p.rowmax_bits + ridx, enc_ordered(run_max));
                else
                    atomicAdd(p.rowsum + ridx, sum0 + sum1);
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

}  // namespace

int launch_gma_stats(const GmaStatsParams& p, const CUtensorMap& tm_q, const CUtensorMap& tm_k, int num_sms,
                     cudaStream_t s) {
    SF_REQUIRE(p.Kp <= kMaxKBlocks * BK, "gma_stats: Kp %d exceeds the resident-Q capacity", p.Kp);
    GmaStatsArgs args;
    args.tm_q = tm_q;
    args.tm_k = tm_k;
    args.p = p;
    SF_CUDA_CHECK(cudaFuncSetAttribute(gma_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const long long units = static_cast<long long>(p.P) * p.pair_tiles * p.chunks;
    const int grid = static_cast<int>(std::min<long long>(units, num_sms));
    prof_before(SF_KERNEL_GMA_STATS, s);
    SF_CUDA_CHECK(launch_kernel(gma_stats_kernel, dim3(grid), dim3(kThreads), kSmemBytes, s, args));
    prof_after(SF_KERNEL_GMA_STATS, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
