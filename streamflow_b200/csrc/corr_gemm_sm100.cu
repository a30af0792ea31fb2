// G1: all-pairs correlation volume + pooled pyramid as ONE persistent tcgen05 GEMM on CTA pairs
// (replaces CorrBlock.corr + the three avg_pool2d passes, core/corr.py:13-21,46-54).
//
//   level_l[b, n, m] = alpha * sum_k A[b, n, k] * B_l[b, m, k]        A = packed fmap1 [B, N, Kp] fp16
//                                                                     B_l = packed pooled fmap2 [B, rows_l, Kp]
// The four levels are extra N-tiles of the same GEMM (pooling is linear, see corr_pack.cu), so every pyramid
// level is written exactly once, straight from the accumulator, and the 1/sqrt(D) scale is folded into the
// epilogue.  In the single-product modes the kernel is output-store bound (261 MB per Sintel pair vs 33 GFLOP; it runs
// at 86 % of the measured HBM write ceiling): the stores of tile i overlap the MMAs of tile i+1.
//
// Cluster of 2 CTAs x 192 threads, 1 CTA / SM, persistent over a contiguous range of (batch, m-tile PAIR, n-tile) tiles:
//   warp 0      TMA producer: this CTA's A block (128 x 64) and its HALF of the B block (128 of 256 rows x 64) per
//               k-block, fp16, 128B swizzle, 5-stage mbarrier ring of 32 KB
//   warp 1      TMEM allocator; in the leader CTA the elected thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256,
//               K = 16, kind::f16, fp32 accumulate) for both CTAs; two 256-column accumulators per CTA (double buffered
//               against the epilogue); the number of k-blocks per tile depends on the precision mode and, in mode `auto`,
//               on the exactness flag the absmax pass left on the device
//   warps 2-5   epilogue of the CTA's own 128 rows: tcgen05.ld 32 lanes x 64 columns -> scale -> swizzled smem ->
//               transposed read-back -> st.global.cs: every store instruction writes four full 128-byte lines
//
// Why pairs: the single-CTA kernel of round 1 (M = 128, one CTA loads A and all of B) was bound by L2 bandwidth, not by
// the tensor pipe or by DRAM: per 128 x 256 tile a CTA pulled 192 KB of operands through L2 (K = 256) and pushed 128 KB
// of results, and the measured cost per streamed k-block was the same in all precision modes (0.70-0.93 us, i.e.
// ~12 TB/s of aggregate L2 traffic; profiles/r2_probes.txt has the A/B numbers: GEMM of a 3-pair group build
// 162 -> 155 us single-product, 359 -> 289 us three-product, 58 % tensor pipe in ncu).  With cta_group::2 a CTA loads
// A (16 KB) + half of B (16 KB) per k-block instead of 48 KB.  Barrier protocol as in gma_stats_kernel: `full` lives in
// the leader (its expect_tx arrival + the peer's remote arrival, bytes of both CTAs' TMA loads), `empty` / `tfull` are
// signalled in both CTAs by multicast tcgen05.commit, `tempty` lives in the leader and collects the epilogue warps of
// both CTAs.
#include "sf_internal.h"
#include "sm100_ptx.cuh"

namespace sf {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kEpiBufs = 4;                           // 4 KB staging buffers per epilogue warp (two pairs, ping-pong)
constexpr int kABytes = BM * BK * 2;
constexpr int kEpiBuf = 32 * 32 * 4;                 // 32 rows x 128 B
constexpr int kEpiBytes = 4 * kEpiBufs * kEpiBuf;
constexpr int kTmemCols = 512;

struct CorrGemmArgs {
    CUtensorMap tm_a;
    CUtensorMap tm_b[SF_NUM_LEVELS];
    CorrGemmParams p;
    int n_cols[SF_NUM_LEVELS];     // valid output columns per level (rows_l) = row pitch of the level in floats
    float* out[SF_NUM_LEVELS];     // level buffers [B * N, n_cols]
};

struct TileCoord {
    int b, mt, level, ntl;
};

namespace pair {
constexpr int kPStages = 5;
constexpr int kBHalfBytes = (BN / 2) * BK * 2;        // this CTA's 128 rows of the 256-row B tile
constexpr int kPStageBytes = kABytes + kBHalfBytes;    // 32 KB
constexpr int kPSmemBytes = kPStages * kPStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
}  // namespace pair

__device__ __forceinline__ TileCoord decode_pair_tile(const CorrGemmParams& p, int pair_tiles, long long t) {
    TileCoord c;
    const int per_b = pair_tiles * p.n_tiles_total;
    c.b = static_cast<int>(t / per_b);
    const int r = static_cast<int>(t - static_cast<long long>(c.b) * per_b);
    c.mt = r / p.n_tiles_total;                        // index of the m-tile PAIR
    int nt = r - c.mt * p.n_tiles_total;
    c.level = 0;
#pragma unroll
    for (int l = 0; l < SF_NUM_LEVELS - 1; ++l) {
        if (c.level == l && nt >= p.n_tiles[l]) {
            nt -= p.n_tiles[l];
            c.level = l + 1;
        }
    }
    c.ntl = nt;
    return c;
}

__global__ void __launch_bounds__(192, 1) corr_gemm_pair_kernel(const __grid_constant__ CorrGemmArgs args) {
    using namespace pair;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_base = smem;
    uint8_t* epi_base = smem + kPStages * kPStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + kEpiBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kPStages;
    uint64_t* tfull = bars + 2 * kPStages;
    uint64_t* tempty = bars + 2 * kPStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPStages + 4);

    const CorrGemmParams& p = args.p;
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                  // 0 = leader of the pair
    int kb_l0 = p.kb_single, kb_lx = p.kb_single;
    if (p.mode == 1 || (p.mode == 2 && p.amax_bits[2] != 0u)) {
        kb_l0 = kb_lx = p.kb_split;
    } else if (p.mode == 2) {
        kb_lx = p.kb_pool;
    }
    const int pair_tiles = (p.m_tiles + 1) / 2;
    const long long total = static_cast<long long>(p.B) * pair_tiles * p.n_tiles_total;
    const long long n_pairs = gridDim.x / 2, pair_id = blockIdx.x / 2;
    const long long t_begin = total * pair_id / n_pairs;
    const long long t_end = total * (pair_id + 1) / n_pairs;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tm_a);
        for (int l = 0; l < SF_NUM_LEVELS; ++l) tma_prefetch_desc(&args.tm_b[l]);
        for (int i = 0; i < kPStages; ++i) {
            mbar_init(&full[i], 2);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 2 * 4);
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

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (long long t = t_begin; t < t_end; ++t) {
            const TileCoord c = decode_pair_tile(p, pair_tiles, t);
            const int mt = 2 * c.mt + static_cast<int>(rank);   // may lie past the last m-tile: TMA zero-fills, nothing stored
            const int kblocks = c.level == 0 ? kb_l0 : kb_lx;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = stage_base + stage * kPStageBytes;
                if (elect_one()) {
                    const uint32_t lead = map_shared_rank(&full[stage], 0);
                    if (rank == 0) mbar_expect_tx(&full[stage], 2 * kPStageBytes);
                    else mbar_arrive_cluster(lead);
                    tma_load_3d_2sm(&args.tm_a, lead, sa, kb * BK, mt * BM, c.b);
                    tma_load_3d_2sm(&args.tm_b[c.level], lead, sa + kABytes, kb * BK,
                                    c.ntl * BN + static_cast<int>(rank) * (BN / 2), c.b);
                }
                __syncwarp();
                if (++stage == kPStages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {                   // the leader's elected thread issues the M = 256 MMAs for both CTAs
            constexpr uint32_t idesc = make_idesc_f16_f32(2 * BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (long long t = t_begin; t < t_end; ++t, ++local) {
                const int acc = local & 1;
                const uint32_t acc_phase = (local >> 1) & 1;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                const int kblocks = decode_pair_tile(p, pair_tiles, t).level == 0 ? kb_l0 : kb_lx;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage_base + stage * kPStageBytes);
                    const uint64_t da = make_kmajor_sw128_desc(sa);
                    const uint64_t db = make_kmajor_sw128_desc(sa + kABytes);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            umma_f16_ss_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        umma_commit_2sm(&empty[stage]);
                        if (kb == kblocks - 1) umma_commit_2sm(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++stage == kPStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else {
        const int e = warp - 2;            // staging buffers of this warp
        const int quad = warp & 3;         // TMEM lane quadrant this warp may read
        uint8_t* bufs = epi_base + e * kEpiBufs * kEpiBuf;
        const int e1 = scale_exponent_from_bits(p.amax_bits[0]);
        const int e2 = scale_exponent_from_bits(p.amax_bits[1]);
        const float alpha = p.inv_sqrt_d * exp2f(static_cast<float>(-(e1 + e2)));
        int local = 0;
        int buf_sel = 0;
        for (long long t = t_begin; t < t_end; ++t, ++local) {
            const TileCoord c = decode_pair_tile(p, pair_tiles, t);
            const int mt = 2 * c.mt + static_cast<int>(rank);
            const int acc = local & 1;
            const uint32_t acc_phase = (local >> 1) & 1;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int row0 = mt * BM + quad * 32;
            const int ncols = args.n_cols[c.level];
            const long long pitch = ncols;
            float* out_base = args.out[c.level] + static_cast<long long>(c.b) * p.N * pitch;
#pragma unroll 1
            for (int cp = 0; cp < BN / 64; ++cp) {      // 64 columns per round trip through the staging buffers
                uint32_t v0[32], v1[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + cp * 64;
                tmem_ld_32x32(taddr, v0);
                tmem_ld_32x32(taddr + 32, v1);
                tmem_ld_wait();
                if (cp == BN / 64 - 1) {   // accumulator fully drained into registers: hand TMEM back to the leader
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (rank == 0) mbar_arrive(&tempty[acc]);
                        else mbar_arrive_cluster(map_shared_rank(&tempty[acc], 0));
                    }
                }
                const int col0 = c.ntl * BN + cp * 64;
                if (col0 >= ncols || row0 >= p.N) continue;          // warp-uniform
                uint8_t* buf = bufs + buf_sel * (2 * kEpiBuf);
                __syncwarp();               // the read-back of this buffer pair two iterations ago is complete
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 o;
                    o.x = __uint_as_float(v0[4 * j + 0]) * alpha;
                    o.y = __uint_as_float(v0[4 * j + 1]) * alpha;
                    o.z = __uint_as_float(v0[4 * j + 2]) * alpha;
                    o.w = __uint_as_float(v0[4 * j + 3]) * alpha;
                    *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                    o.x = __uint_as_float(v1[4 * j + 0]) * alpha;
                    o.y = __uint_as_float(v1[4 * j + 1]) * alpha;
                    o.z = __uint_as_float(v1[4 * j + 2]) * alpha;
                    o.w = __uint_as_float(v1[4 * j + 3]) * alpha;
                    *reinterpret_cast<float4*>(buf + kEpiBuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                }
                __syncwarp();
                const int sub = lane >> 3, chunk = lane & 7;
                float* dst = out_base + static_cast<long long>(row0 + sub) * pitch + col0 + chunk * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = 4 * i + sub;
                    const float4 a = *reinterpret_cast<const float4*>(buf + rr * 128 + ((chunk ^ (rr & 7)) << 4));
                    const float4 bq = *reinterpret_cast<const float4*>(buf + kEpiBuf + rr * 128 + ((chunk ^ (rr & 7)) << 4));
                    if (row0 + rr < p.N) {        // level images are multiples of 16 floats wide: clip per 16-byte chunk
                        float* d = dst + static_cast<long long>(4 * i) * pitch;
                        if (col0 + chunk * 4 < ncols) __stcs(reinterpret_cast<float4*>(d), a);
                        if (col0 + 32 + chunk * 4 < ncols) __stcs(reinterpret_cast<float4*>(d + 32), bq);
                    }
                }
                buf_sel ^= 1;
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();                    // the peer's shared memory and barriers stay valid until both are done
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<kTmemCols>(tmem_base);
    }
}

}  // namespace

int launch_corr_gemm(const CorrGemmParams& p, const CUtensorMap& tm_a, const CUtensorMap tm_b[SF_NUM_LEVELS],
                     const int n_cols[SF_NUM_LEVELS], float* const levels[SF_NUM_LEVELS], int num_sms,
                     cudaStream_t s) {
    CorrGemmArgs args;
    args.tm_a = tm_a;
    for (int l = 0; l < SF_NUM_LEVELS; ++l) {
        args.tm_b[l] = tm_b[l];             // 64 x 128 boxes: each CTA of a pair loads half of a 256-row B tile
        args.n_cols[l] = n_cols[l];
        args.out[l] = levels[l];
    }
    args.p = p;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(corr_gemm_pair_kernel), pair::kPSmemBytes)) return rc;
    const long long total = static_cast<long long>(p.B) * ((p.m_tiles + 1) / 2) * p.n_tiles_total;
    const int grid = 2 * static_cast<int>(std::min<long long>(total, num_sms / 2));
    prof_before(SF_KERNEL_CORR_GEMM, s);
    SF_CUDA_CHECK(launch_kernel_cluster(corr_gemm_pair_kernel, dim3(grid), dim3(192), pair::kPSmemBytes, s, 2u, args));
    prof_after(SF_KERNEL_CORR_GEMM, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
