// G1, resident-A variant of the correlation GEMM (used when the packed K fits: Kp <= 256, i.e. the model's D = 256
// in SF_PREC_F16 mode; the streaming kernel in corr_gemm_sm100.cu handles larger Kp).
//
// The streaming kernel re-reads a 64 KB A tile and a 128 KB B tile from L2 for every 128 x 256 output tile: with the
// 268 MB of stores that is ~660 MB of L2 traffic per Sintel pair, i.e. the kernel sits on the L2->SM bandwidth cap
// (57 us, tensor pipe 32 %), not on HBM (42 us of stores) or the MMA (20 us).  Here a CTA keeps TWO 128-query tiles of
// A resident in shared memory for its whole run of N-tiles (128 KB, loaded ~1.5 times per CTA) and streams 128-column
// B tiles (16 KB per k-block), each of which feeds two MMAs: B traffic 134 MB, A 28 MB -> ~430 MB through L2.
//
// CTA = 192 threads, persistent over a contiguous range of (batch, query-tile pair, 128-column n-tile) units:
//   warp 0      TMA producer: resident A (one mbarrier, reloaded when the pair changes) + 3-stage B ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer: 2 x (M=128, N=128, K=16) per k-step,
//               accumulators 2 tiles x 128 columns, double buffered (512 TMEM columns)
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns -> scale -> swizzled smem -> 3-D TMA store,
//               three 4 KB staging buffers per warp
#include "sf_internal.h"
#include "sm100_ptx.cuh"

namespace sf {

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int kMaxKBlocks = 4;                       // Kp <= 256
constexpr int kStages = 3;
constexpr int kATileBytes = BM * BK * 2;             // 16 KB: one query tile, one k-block
constexpr int kBTileBytes = BN * BK * 2;             // 16 KB
constexpr int kABytes = 2 * kMaxKBlocks * kATileBytes;   // 128 KB resident
constexpr int kEpiBufs = 3;
constexpr int kEpiBuf = 32 * 32 * 4;                 // 32 rows x 128 B
constexpr int kEpiBytes = 4 * kEpiBufs * kEpiBuf;    // 48 KB
constexpr int kSmemBytes = kABytes + kStages * kBTileBytes + kEpiBytes + 1024 + 256;
constexpr int kTmemCols = 512;

struct CorrGemmRaArgs {
    CUtensorMap tm_a;
    CUtensorMap tm_b[SF_NUM_LEVELS];
    CUtensorMap tm_out[SF_NUM_LEVELS];
    CorrGemmParams p;
    int n_cols[SF_NUM_LEVELS];     // valid output columns per level
    int n_tiles[SF_NUM_LEVELS];    // 128-column tiles per level
    int n_tiles_total;
    int pair_tiles;                // ceil(m_tiles / 2)
};

struct Unit {
    int b, mp, level, ntl;
};

__device__ __forceinline__ Unit decode_unit(const CorrGemmRaArgs& a, long long u) {
    Unit c;
    const int per_b = a.pair_tiles * a.n_tiles_total;
    c.b = static_cast<int>(u / per_b);
    const int r = static_cast<int>(u - static_cast<long long>(c.b) * per_b);
    c.mp = r / a.n_tiles_total;
    int nt = r - c.mp * a.n_tiles_total;
    c.level = 0;
#pragma unroll
    for (int l = 0; l < SF_NUM_LEVELS - 1; ++l) {
        if (c.level == l && nt >= a.n_tiles[l]) {
            nt -= a.n_tiles[l];
            c.level = l + 1;
        }
    }
    c.ntl = nt;
    return c;
}

__global__ void __launch_bounds__(192, 1) corr_gemm_ra_kernel(const __grid_constant__ CorrGemmRaArgs args) {
    extern __shared__ uint8_t smem_raw[];
    // align by pointer arithmetic (not through an integer) so the compiler keeps the shared address space: STS / LDS
    // instead of generic ST / LD in the epilogue
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_base = smem;                                // [2 tiles][kblocks][128 x 64] fp16, 128B-swizzled
    uint8_t* b_base = smem + kABytes;                      // ring of [128 x 64] fp16
    uint8_t* epi_base = b_base + kStages * kBTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + kEpiBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* tfull = bars + 2 * kStages;
    uint64_t* tempty = bars + 2 * kStages + 2;
    uint64_t* afull = bars + 2 * kStages + 4;
    uint64_t* aempty = bars + 2 * kStages + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 6);

    const CorrGemmParams& p = args.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = (p.Kp + BK - 1) / BK;
    const long long total = static_cast<long long>(p.B) * args.pair_tiles * args.n_tiles_total;
    const long long u_begin = total * blockIdx.x / gridDim.x;
    const long long u_end = total * (blockIdx.x + 1) / gridDim.x;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&args.tm_a);
        for (int l = 0; l < SF_NUM_LEVELS; ++l) {
            tma_prefetch_desc(&args.tm_b[l]);
            tma_prefetch_desc(&args.tm_out[l]);
        }
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);
        }
        mbar_init(afull, 1);
        mbar_init(aempty, 1);
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
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, aphase = 0;
            long long cur_pair = -1;
            for (long long u = u_begin; u < u_end; ++u) {
                const Unit c = decode_unit(args, u);
                const long long pair_id = static_cast<long long>(c.b) * args.pair_tiles + c.mp;
                if (pair_id != cur_pair) {                  // (re)load the two resident A tiles
                    mbar_wait(aempty, aphase ^ 1);          // every MMA that read the previous A has completed
                    mbar_expect_tx(afull, 2 * kblocks * kATileBytes);
                    for (int half = 0; half < 2; ++half)
                        for (int kb = 0; kb < kblocks; ++kb)    // rows past N / a tile past the last: zero-filled
                            tma_load_3d(&args.tm_a, afull, a_base + (half * kMaxKBlocks + kb) * kATileBytes, kb * BK,
                                        (2 * c.mp + half) * BM, c.b);
                    cur_pair = pair_id;
                    aphase ^= 1;
                }
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], kBTileBytes);
                    tma_load_3d(&args.tm_b[c.level], &full[stage], b_base + stage * kBTileBytes, kb * BK, c.ntl * BN,
                                c.b);
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
            uint32_t phase = 0, aphase = 0;
            long long cur_pair = -1;
            for (long long u = u_begin; u < u_end; ++u, ++local) {
                const Unit c = decode_unit(args, u);
                const long long pair_id = static_cast<long long>(c.b) * args.pair_tiles + c.mp;
                if (pair_id != cur_pair) {
                    if (cur_pair >= 0) umma_commit(aempty);     // previous A is free once the issued MMAs finish
                    mbar_wait(afull, aphase);
                    tc_fence_after();
                    cur_pair = pair_id;
                    aphase ^= 1;
                }
                const int acc = local & 1;
                mbar_wait(&tempty[acc], ((local >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (2 * BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t db = make_kmajor_sw128_desc(smem_u32(b_base + stage * kBTileBytes));
                    const uint64_t d0 = make_kmajor_sw128_desc(smem_u32(a_base + kb * kATileBytes));
                    const uint64_t d1 = make_kmajor_sw128_desc(smem_u32(a_base + (kMaxKBlocks + kb) * kATileBytes));
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
    } else {
        const int e = warp - 2;            // staging buffers of this warp
        const int quad = warp & 3;         // TMEM lane quadrant this warp may read
        uint8_t* bufs = epi_base + e * kEpiBufs * kEpiBuf;
        const int e1 = scale_exponent_from_bits(p.amax_bits[0]);
        const int e2 = scale_exponent_from_bits(p.amax_bits[1]);
        const float alpha = p.inv_sqrt_d * exp2f(static_cast<float>(-(e1 + e2)));
        int local = 0, buf_sel = 0;
        for (long long u = u_begin; u < u_end; ++u, ++local) {
            const Unit c = decode_unit(args, u);
            const int acc = local & 1;
            mbar_wait(&tfull[acc], (local >> 1) & 1);
            tc_fence_after();
            const int ncols = args.n_cols[c.level];
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int row0 = (2 * c.mp + half) * BM + quad * 32;
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ++ch) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * (2 * BN) + half * BN +
                                      ch * 32,
                                  v);
                    tmem_ld_wait();
                    if (half == 1 && ch == BN / 32 - 1) {   // accumulators fully drained: hand TMEM back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[acc]);
                    }
                    const int col0 = c.ntl * BN + ch * 32;
                    if (col0 >= ncols || row0 >= p.N) continue;          // warp-uniform
                    uint8_t* buf = bufs + buf_sel * kEpiBuf;
                    if (lane == 0) tma_store_wait_read<kEpiBufs - 1>();  // the store that last read `buf` is done
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o;
                        o.x = __uint_as_float(v[4 * j + 0]) * alpha;
                        o.y = __uint_as_float(v[4 * j + 1]) * alpha;
                        o.z = __uint_as_float(v[4 * j + 2]) * alpha;
                        o.w = __uint_as_float(v[4 * j + 3]) * alpha;
                        *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&args.tm_out[c.level], buf, col0, row0, c.b);
                        tma_store_commit();
                    }
                    buf_sel = (buf_sel + 1) % kEpiBufs;
                }
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

}  // namespace

bool corr_gemm_ra_supported(int Kp) { return Kp <= kMaxKBlocks * BK; }

int launch_corr_gemm_ra(const CorrGemmParams& p, const CUtensorMap& tm_a, const CUtensorMap tm_b[SF_NUM_LEVELS],
                        const CUtensorMap tm_out[SF_NUM_LEVELS], const int n_cols[SF_NUM_LEVELS], int num_sms,
                        cudaStream_t s) {
    CorrGemmRaArgs args;
    args.tm_a = tm_a;
    args.n_tiles_total = 0;
    for (int l = 0; l < SF_NUM_LEVELS; ++l) {
        args.tm_b[l] = tm_b[l];
        args.tm_out[l] = tm_out[l];
        args.n_cols[l] = n_cols[l];
        args.n_tiles[l] = (n_cols[l] + BN - 1) / BN;
        args.n_tiles_total += args.n_tiles[l];
    }
    args.p = p;
    args.pair_tiles = (p.m_tiles + 1) / 2;
    SF_CUDA_CHECK(cudaFuncSetAttribute(corr_gemm_ra_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const long long total = static_cast<long long>(p.B) * args.pair_tiles * args.n_tiles_total;
    const int grid = static_cast<int>(std::min<long long>(total, num_sms));
    prof_before(SF_KERNEL_CORR_GEMM, s);
    SF_CUDA_CHECK(launch_kernel(corr_gemm_ra_kernel, dim3(grid), dim3(192), kSmemBytes, s, args));
    prof_after(SF_KERNEL_CORR_GEMM, s);
    SF_CUDA_CHECK(cudaGetLastError());
    return SF_OK;
}

}  // namespace sf
