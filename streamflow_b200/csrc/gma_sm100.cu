// G3: GMA softmax attention (core/gma.py:53-65) and motion-feature aggregation (core/gma.py:91-104) on
// tcgen05 tensor cores.
//
// Q and K do not change across refinement iterations, so the softmax numerator is computed ONCE per clip and
// kept in HBM as fp16 (E = 2^12 * exp(s - rowmax), tile-major 16 KB blocks, 99 MB per Sintel map -- trivial against 180 GB),
// exactly the matrix the reference's autocast path re-casts to fp16 every iteration (core/gma.py:95-97).
// Every iteration is then one streaming GEMM  acc = E . V^T  bound by reading E from HBM:
//
//   gma_stats_kernel      S = Q K^T tiles (CTA pairs, 256 x 256; K = d, or [hi | lo] operands for three products) in TMEM;
//                         pass 1 reduces the row max, pass 2 writes E (4 KB bulk stores) and the row sums.
//   gma_aggregate_kernel  (gma_aggregate_sm100.cu) the per-iteration streaming GEMM with the fused epilogue.
#include <type_traits>

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
constexpr int kMaxStages = 8, kMaxQBlocks = 4;          // K ring depth / resident Q k-blocks (hi and lo halves of d = 128)
constexpr int kABytes = BM * BK * 2;                    // this CTA's 128 query rows of one k-block
constexpr int kBBytes = (BN / 2) * BK * 2;              // this CTA's half (128 key rows) of one K k-block
constexpr int kEpiWarps = 16;                           // four warps per TMEM lane quadrant, 64 key columns each
constexpr int kEpiBuf = 32 * 128;                       // 32 rows x 64 fp16, one staging buffer per warp
constexpr int kEpiBytes = kEpiWarps * kEpiBuf;
constexpr int kSmemBytes = 227 * 1024;
constexpr int kBarBytes = 512;
constexpr int kRing = (kSmemBytes - 1024 - kBarBytes - kEpiBytes) / 1024 * 1024;     // Q blocks + K stages
constexpr int kTmemCols = 512;
constexpr int kThreads = 64 + 32 * kEpiWarps;
}  // namespace st

struct GmaStatsArgs {
    CUtensorMap tm_q, tm_k;
    GmaStatsParams p;
    int stages;                     // K ring depth: whatever fits next to the resident Q blocks
};

// CTA pairs (cluster of 2, cta_group::2): the pair owns two 128-query tiles; each CTA keeps its own Q tile resident,
// loads its own half (128 key rows) of every K block, and the leader CTA issues M = 256 x N = 256 MMAs that write
// each CTA's 128 x 256 logits into that CTA's TMEM.  Per SM an MMA step then reads 8 KB of operands from shared
// memory instead of 12 KB and the TMA fill halves -- the single-CTA kernel sat on shared-memory bandwidth.
// Barriers: full / q_full live in the leader (one expect_tx arrival from the leader, one plain arrival from the peer,
// bytes from both CTAs' TMA loads); empty / q_empty / tfull are signalled in both CTAs by multicast tcgen05.commit;
// tempty lives in the leader and collects the epilogue warps of both CTAs.
__global__ void __launch_bounds__(st::kThreads, 1) gma_stats_kernel(const __grid_constant__ GmaStatsArgs args) {
    using namespace st;
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic (not through an integer) so the compiler keeps the shared address space: STS / LDS
    // instead of generic ST / LD in the epilogue
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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
    // warp index through a shuffle: the compiler then knows the role dispatch below is warp-uniform and keeps MMA
    // descriptors in uniform registers (a per-thread `threadIdx.x >> 5` made every tcgen05.mma a 15-instruction
    // ELECT / R2UR.BROADCAST waterfall loop)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                  // 0 = leader of the pair
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
    const long long units = static_cast<long long>(p.P) * p.pair_tiles * p.chunks;
    const long long n_pairs = gridDim.x / 2, pair_id = blockIdx.x / 2;
    const long long u_begin = units * pair_id / n_pairs;
    const long long u_end = units * (pair_id + 1) / n_pairs;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tm_q);
        tma_prefetch_desc(&args.tm_k);
        for (int i = 0; i < kMaxStages; ++i) {
            mbar_init(&full[i], 2);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < kMaxQBlocks; ++i) {
            mbar_init(&q_full[i], 2);
            mbar_init(&q_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 2 * kEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm<kTmemCols>(tmem_slot);
    pdl_launch();
    tc_fence_before();
    __syncthreads();                       // CTA-local ordering of the TMEM-address write (tcgen05.alloc -> tmem_slot) that
                                           // compute-sanitizer's racecheck can see; the cluster barrier below subsumes it
    cluster_sync_all();                    // both CTAs' barriers are initialised before any remote arrive / TMA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    auto unit_coords = [&](long long u, int& pb, int& mt, int& nt0, int& nt1) {
        const int per_p = p.pair_tiles * p.chunks;
        pb = static_cast<int>(u / per_p);
        const int r = static_cast<int>(u - static_cast<long long>(pb) * per_p);
        const int mp = r / p.chunks;
        mt = 2 * mp + static_cast<int>(rank);      // this CTA's query tile (may lie past the last one: all rows masked)
        const int ck = r - mp * p.chunks;
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
                                const uint32_t lead = map_shared_rank(&q_full[b], 0);
                                if (rank == 0) mbar_expect_tx(&q_full[b], 2 * kABytes);
                                else mbar_arrive_cluster(lead);
                                tma_load_3d_2sm(&args.tm_q, lead, q_base + b * kABytes, b * BK, mt * BM, pb);
                            }
                        }
                        mbar_wait(&empty[stage], phase ^ 1);
                        const uint32_t lead = map_shared_rank(&full[stage], 0);
                        if (rank == 0) mbar_expect_tx(&full[stage], 2 * kBBytes);
                        else mbar_arrive_cluster(lead);
                        tma_load_3d_2sm(&args.tm_k, lead, stage_base + stage * kBBytes, (is_lo ? dblocks + j : j) * BK,
                                        nt * BN + static_cast<int>(rank) * (BN / 2), pb);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                qphase ^= 1;
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // One elected thread feeds the tensor pipe; the loop itself runs warp-uniformly.  Measured with clock64 on an earlier, loop-heavy version of this
            // branch: ~107 clk of single-thread instruction latency per MMA issued (descriptor arithmetic, predicates,
            // loop control at one dependent instruction per ~4.5 clk) against 128 clk of execution -- the issue thread,
            // not shared memory or L2, set the pace (4870 clk per tile for 3072 clk of MMA work).  Hence: descriptors
            // are built once, the K-step loop is unrolled at compile time, and only one 32-bit add per step remains.
            constexpr uint32_t idesc = make_idesc_f16_f32(2 * BM, BN);
            uint64_t qd[kMaxQBlocks];
#pragma unroll
            for (int b = 0; b < kMaxQBlocks; ++b) qd[b] = make_kmajor_sw128_desc(smem_u32(q_base + b * kABytes));
            const uint64_t kd0 = make_kmajor_sw128_desc(smem_u32(stage_base));
            int stage = 0, local = 0;
            uint32_t phase = 0, qphase = 0;
            auto tile = [&](auto split_tag, uint32_t d_tmem, uint64_t* tfull_bar, bool first_tile, bool last_tile) {
                constexpr bool kSplit = decltype(split_tag)::value;
                constexpr int kSteps = kSplit ? 4 : 2;          // d = 128: two 64-column blocks per operand half
#pragma unroll
                for (int t = 0; t < kSteps; ++t) {
                    const int j = kSplit ? (t >> 1) : t;
                    const bool is_lo = kSplit && (t & 1);
                    if (first_tile && !is_lo) {
                        mbar_wait(&q_full[j], qphase);
                        if (kSplit) mbar_wait(&q_full[2 + j], qphase);
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t db = kd0 + static_cast<uint32_t>(stage * (kBBytes >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16_ss_2sm(d_tmem, qd[j] + 2 * k, db + 2 * k, idesc, (t | k) != 0);
                        if (kSplit && !is_lo) {                 // K_hi_j also meets Q_lo_j
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_f16_ss_2sm(d_tmem, qd[2 + j] + 2 * k, db + 2 * k, idesc, 1u);
                        }
                        umma_commit_2sm(&empty[stage]);
                        if (last_tile) umma_commit_2sm(&q_empty[kSplit ? (is_lo ? j : 2 + j) : j]);   // last use of that Q block
                        if (t == kSteps - 1) umma_commit_2sm(tfull_bar);
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            };
            for (long long u = u_begin; u < u_end; ++u) {
                int pb, mt, nt0, nt1;
                unit_coords(u, pb, mt, nt0, nt1);
                if (nt0 >= nt1) continue;               // empty key chunk: no Q load, no phase flip
                for (int nt = nt0; nt < nt1; ++nt, ++local) {
                    const int acc = local & 1;
                    mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    if (split2) tile(std::true_type{}, d_tmem, &tfull[acc], nt == nt0, nt == nt1 - 1);
                    else tile(std::false_type{}, d_tmem, &tfull[acc], nt == nt0, nt == nt1 - 1);
                }
                qphase ^= 1;
            }
        }
    } else {
        // 16 epilogue warps: warp e owns TMEM lane quadrant (warp & 3) and the 64-key column group cp = e >> 2 of the
        // 256-column tile, processed as two 32-column halves (the exp / convert / row-sum phase is MUFU- and
        // latency-bound: with 8 warps it took ~3900 clk per tile against 3072 clk of MMA work)
        const int e = warp - 2, quad = warp & 3, cp = e >> 2;
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
            if (p.pass == 2 && row_ok && nt0 < nt1) mrow = dec_ordered(p.rowmax_bits[ridx]) * kLog2e - 12.0f;
            for (int nt = nt0; nt < nt1; ++nt, ++local) {
                const int acc = local & 1;
                mbar_wait(&tfull[acc], (local >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + cp * 64;
                const int col0 = nt * BN + cp * 64;
                const bool live = col0 < p.Npad;            // warp-uniform: key group inside the padded row
                const bool full = col0 + 64 <= p.N;         // warp-uniform: no per-element key masking needed
                if (p.pass == 2 && live) {
                    // one staging buffer per warp: the previous tile's store has had a whole tile period to be read
                    if (elect_one()) tma_store_wait_read<0>();
                    __syncwarp();
                }
                float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {            // 32 columns at a time keeps the register footprint small
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + hf * 32, v);
                    tmem_ld_wait();
                    if (hf == 1) {                          // accumulator drained: tell the leader's MMA thread
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (rank == 0) mbar_arrive(&tempty[acc]);
                            else mbar_arrive_cluster(map_shared_rank(&tempty[acc], 0));
                        }
                    }
                    if (!live) continue;
                    const int c0 = col0 + hf * 32;
                    if (p.pass == 1) {
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) run_max = fmaxf(run_max, __uint_as_float(v[j]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < p.N) run_max = fmaxf(run_max, __uint_as_float(v[j]));
                        }
                    } else {
                        __half2 h[16];
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float a0 = ex2_approx(fmaf(__uint_as_float(v[j]), kLog2e, -mrow));
                                const float a1 = ex2_approx(fmaf(__uint_as_float(v[j + 1]), kLog2e, -mrow));
                                h[j >> 1] = __floats2half2_rn(a0, a1);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float a0 = (c0 + j < p.N) ? ex2_approx(fmaf(__uint_as_float(v[j]), kLog2e, -mrow)) : 0.f;
                                const float a1 =
                                    (c0 + j + 1 < p.N) ? ex2_approx(fmaf(__uint_as_float(v[j + 1]), kLog2e, -mrow)) : 0.f;
                                h[j >> 1] = __floats2half2_rn(a0, a1);
                            }
                        }
                        // row sum of the ROUNDED numerators (fp32 adds, two chains): sum_j E / rowsum == 1 for what is
                        // stored (summing the un-rounded values leaves up to 2^-11 of normalisation error on peaked rows)
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float2 f = __half22float2(h[j]);
                            sum0 += f.x;
                            sum1 += f.y;
                        }
#pragma unroll
                        for (int c16 = 0; c16 < 4; ++c16) {   // 4 x 16-byte chunks (8 halfs) of this half's 64 B per row
                            int4 o;
                            o.x = *reinterpret_cast<int*>(&h[c16 * 4 + 0]);
                            o.y = *reinterpret_cast<int*>(&h[c16 * 4 + 1]);
                            o.z = *reinterpret_cast<int*>(&h[c16 * 4 + 2]);
                            o.w = *reinterpret_cast<int*>(&h[c16 * 4 + 3]);
                            *reinterpret_cast<int4*>(buf + lane * 128 + (((hf * 4 + c16) ^ (lane & 7)) << 4)) = o;
                        }
                    }
                }
                if (p.pass == 2 && live) {
                    run_sum += sum0 + sum1;
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (elect_one()) {
                        // E is tile-major: [P][m-tile][64-key block][128 rows][64 keys]; the staging buffer already is
                        // the swizzled image of its 32 rows, which are 4 KB contiguous in HBM: one bulk copy (a pair's
                        // second CTA may hold a tile past the last one: nothing to store)
                        if (mt < p.m_tiles) {
                            __half* dst = p.E + (static_cast<long long>(pb) * p.m_tiles * kbk +
                                                 static_cast<long long>(mt) * kbk + (col0 >> 6)) * (BM * 64) + quad * 32 * 64;
                            bulk_store(dst, buf, kEpiBuf);
                        }
                        tma_store_commit();
                    }
                }
            }
            if (row_ok && nt0 < nt1) {
                if (p.pass == 1)
                    atomicMax(p.rowmax_bits + ridx, enc_ordered(run_max));
                else   // fp32 sum of fp16 values: always a multiple of 2^-24, so the conversion is exact
                    atomicAdd(p.rowsum_fx + ridx, __float2ull_rn(run_sum * 16777216.0f));
            }
        }
        if (elect_one()) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();                    // the peer's shared memory and barriers stay valid until both are done
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<kTmemCols>(tmem_base);
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

int launch_gma_stats(const GmaStatsParams& p, const CUtensorMap& tm_q, const CUtensorMap& tm_k, int num_sms,
                     cudaStream_t s) {
    GmaStatsArgs args;
    args.tm_q = tm_q;
    args.tm_k = tm_k;
    args.p = p;
    const int d = p.split ? p.Kp / 2 : p.Kp;
    SF_REQUIRE(d == 2 * st::BK, "gma_stats: head dimension %d not supported (the issue loop is unrolled for d = 128)", d);
    const int qblocks = (p.pass == 2 && p.split) ? 2 * (d / st::BK) : d / st::BK;
    args.stages = std::min(st::kMaxStages, (st::kRing - qblocks * st::kABytes) / st::kBBytes);
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gma_stats_kernel), st::kSmemBytes)) return rc;
    const long long units = static_cast<long long>(p.P) * p.pair_tiles * p.chunks;
    const int grid = 2 * static_cast<int>(std::min<long long>(units, num_sms / 2));       // CTA pairs
    prof_before(SF_KERNEL_GMA_STATS, s);
    SF_CUDA_CHECK(launch_kernel_cluster(gma_stats_kernel, dim3(grid), dim3(st::kThreads), st::kSmemBytes, s, 2u, args));
    prof_after(SF_KERNEL_GMA_STATS, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
