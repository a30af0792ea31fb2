// G1: all-pairs correlation volume + pooled pyramid as ONE persistent tcgen05 GEMM
// (replaces CorrBlock.corr + the three avg_pool2d passes, core/corr.py:13-21,46-54).
//
//   level_l[b, n, m] = alpha * sum_k A[b, n, k] * B_l[b, m, k]        A = packed fmap1 [B, N, Kp] fp16
//                                                                     B_l = packed pooled fmap2 [B, rows_l, Kp]
// The four levels are extra N-tiles of the same GEMM (pooling is linear, see corr_pack.cu), so every pyramid
// level is written exactly once, straight from the accumulator, and the 1/sqrt(D) scale is folded into the
// epilogue.  The kernel is output-store bound (261 MB per Sintel pair vs 33 GFLOP): the design goal is that
// the stores of tile i overlap the MMAs of tile i+1.
//
// CTA = 192 threads, 1 CTA / SM, persistent over a contiguous range of (batch, m-tile, n-tile) tiles:
//   warp 0      TMA producer: A (128 x 64) and B (256 x 64) fp16 k-blocks, 128B swizzle, 3-stage mbarrier ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M=128, N=256, K=16, kind::f16, fp32 accum),
//               two 256-column accumulators in TMEM (double buffered against the epilogue)
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 64 columns -> scale -> swizzled smem -> transposed read-back ->
//               st.global.cs: every store instruction writes four full 128-byte lines of the level image
#include "sf_internal.h"
#include "sm100_ptx.cuh"

namespace sf {

namespace {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kStages = 3;
constexpr int kEpiBufs = 4;                           // 4 KB staging buffers per epilogue warp (two pairs, ping-pong)
constexpr int kABytes = BM * BK * 2;
constexpr int kBBytes = BN * BK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kEpiBuf = 32 * 32 * 4;                 // 32 rows x 128 B
constexpr int kEpiBytes = 4 * kEpiBufs * kEpiBuf;
constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/;
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

__device__ __forceinline__ TileCoord decode_tile(const CorrGemmParams& p, long long t) {
    TileCoord c;
    const int per_b = p.m_tiles * p.n_tiles_total;
    c.b = static_cast<int>(t / per_b);
    const int r = static_cast<int>(t - static_cast<long long>(c.b) * per_b);
    c.mt = r / p.n_tiles_total;
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

__global__ void __launch_bounds__(192, 1) corr_gemm_kernel(const __grid_constant__ CorrGemmArgs args) {
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic (not through an integer) so the compiler keeps the shared address space: STS / LDS
    // instead of generic ST / LD in the epilogue
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* stage_base = smem;
    uint8_t* epi_base = smem + kStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + kEpiBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tfull = bars + 2 * kStages;
    uint64_t* tempty = bars + 2 * kStages + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const CorrGemmParams& p = args.p;
    // warp index through a shuffle so the compiler knows the role dispatch is warp-uniform (see gma_sm100.cu)
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int kblocks = (p.Kp + BK - 1) / BK;

    const long long total = static_cast<long long>(p.B) * p.m_tiles * p.n_tiles_total;
    const long long t_begin = total * blockIdx.x / gridDim.x;
    const long long t_end = total * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tm_a);
        for (int l = 0; l < SF_NUM_LEVELS; ++l) {
            tma_prefetch_desc(&args.tm_b[l]);
        }
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
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
    pdl_wait();        // everything above overlapped the previous kernel; its results are visible from here on

    if (warp == 0) {
        {   // warp-uniform loop, one elected lane issues
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = t_begin; t < t_end; ++t) {
                const TileCoord c = decode_tile(p, t);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = stage_base + stage * kStageBytes;
                    if (elect_one()) {
                        mbar_expect_tx(&full[stage], kStageBytes);
                        tma_load_3d(&args.tm_a, &full[stage], sa, kb * BK, c.mt * BM, c.b);
                        tma_load_3d(&args.tm_b[c.level], &full[stage], sa + kABytes, kb * BK, c.ntl * BN, c.b);
                    }
                    __syncwarp();
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        {   // whole warp runs the loop; one elected lane issues
            constexpr uint32_t idesc = make_idesc_f16_f32(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            for (long long t = t_begin; t < t_end; ++t, ++local) {
                const int acc = local & 1;
                const uint32_t acc_phase = (local >> 1) & 1;
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
                    const uint64_t da = make_kmajor_sw128_desc(sa);
                    const uint64_t db = make_kmajor_sw128_desc(sa + kABytes);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // advance 16 fp16 = 32 B along K inside the 128 B swizzle row: +2 in (addr >> 4) units
                            umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit(&empty[stage]);
                        if (kb == kblocks - 1) umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++stage == kStages) {
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
            const TileCoord c = decode_tile(p, t);
            const int acc = local & 1;
            const uint32_t acc_phase = (local >> 1) & 1;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int row0 = c.mt * BM + quad * 32;
            const int ncols = args.n_cols[c.level];
            const long long pitch = ncols;
            float* out_base = args.out[c.level] + static_cast<long long>(c.b) * p.N * pitch;
#pragma unroll 1
            for (int cp = 0; cp < BN / 64; ++cp) {      // 64 columns (two 32-column TMA boxes) per fence
                uint32_t v0[32], v1[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + cp * 64;
                tmem_ld_32x32(taddr, v0);
                tmem_ld_32x32(taddr + 32, v1);
                tmem_ld_wait();
                if (cp == BN / 64 - 1) {   // accumulator fully drained into registers: hand TMEM back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
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
                // Transposed read-back: 8 lanes cover the 128 B of one row segment, so every store instruction writes
                // four full 128-byte lines.  (TMA stores of the same 32-row x 128-byte boxes were the bottleneck of this
                // kernel: ~250 clk of TMA-engine time per 4 KB box = 16 B/clk per SM = the 4.4 TB/s it was stuck at.)
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
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<kTmemCols>(tmem_base);
    }
}

}  // namespace

int launch_corr_gemm(const CorrGemmParams& p, const CUtensorMap& tm_a, const CUtensorMap tm_b[SF_NUM_LEVELS],
                     const int n_cols[SF_NUM_LEVELS], float* const levels[SF_NUM_LEVELS], int num_sms,
                     cudaStream_t s) {
    CorrGemmArgs args;
    args.tm_a = tm_a;
    for (int l = 0; l < SF_NUM_LEVELS; ++l) {
        args.tm_b[l] = tm_b[l];
        args.n_cols[l] = n_cols[l];
        args.out[l] = levels[l];
    }
    args.p = p;
    if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(corr_gemm_kernel), kSmemBytes)) return rc;
    const long long total = static_cast<long long>(p.B) * p.m_tiles * p.n_tiles_total;
    const int grid = static_cast<int>(std::min<long long>(total, num_sms));
    prof_before(SF_KERNEL_CORR_GEMM, s);
    SF_CUDA_CHECK(launch_kernel(corr_gemm_kernel, dim3(grid), dim3(192), kSmemBytes, s, args));
    prof_after(SF_KERNEL_CORR_GEMM, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
