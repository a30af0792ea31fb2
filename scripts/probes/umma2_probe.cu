// Probe: minimal cta_group::2 (2-SM) tcgen05 GEMM.  D[256 x 256] = A[256 x K] . B[256 x K]^T, one cluster of two CTAs:
// CTA r loads A rows [128r, +128) and B rows [128r, +128) per 64-wide k-block; the leader issues M=256 N=256 MMAs.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../streamflow_b200/csrc/sm100_ptx.cuh"
using namespace sf;

constexpr int BK = 64, kStages = 2, kHalf = 128 * BK * 2;      // 16 KB
constexpr int kStageBytes = 2 * kHalf;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}

struct Args { CUtensorMap tm_a, tm_b; float* out; int kblocks; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) probe(const __grid_constant__ Args args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full = bars;                 // used in the leader CTA only
    uint64_t* empty = bars + kStages;      // per CTA
    uint64_t* tfull = bars + 2 * kStages;  // per CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], 1); }
        mbar_init(tfull, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < args.kblocks; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * kStageBytes;
                const uint32_t lead_full = map_to_rank(smem_u32(&full[stage]), 0);
                if (rank == 0) mbar_expect_tx(&full[stage], 2 * kStageBytes);
                else asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(lead_full) : "memory");
                tma_load_2d_2sm(&args.tm_a, lead_full, sa, kb * BK, rank * 128);
                tma_load_2d_2sm(&args.tm_b, lead_full, sa + kHalf, kb * BK, rank * 128);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = make_idesc_f16_f32(256, 256);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < args.kblocks; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                const uint64_t da = make_kmajor_sw128_desc(sa), db = make_kmajor_sw128_desc(sa + kHalf);
                for (int k = 0; k < BK / 16; ++k) umma_f16_ss_2sm(tmem_base, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                umma_commit_2sm(&empty[stage]);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit_2sm(tfull);
        }
    } else {
        const int quad = warp & 3;
        mbar_wait(tfull, 0);
        tc_fence_after();
        const int row = rank * 128 + quad * 32 + lane;
        for (int ch = 0; ch < 8; ++ch) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + ch * 32, v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) args.out[row * 256 + ch * 32 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int M = 256, N = 256, K = 256;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeFn enc = reinterpret_cast<EncodeFn>(p);
    std::vector<__half> ha(M * K), hb(N * K);
    std::vector<float> fa(M * K), fb(N * K);
    srand(1);
    for (int i = 0; i < M * K; ++i) { fa[i] = (rand() % 17 - 8) / 8.0f; ha[i] = __float2half(fa[i]); }
    for (int i = 0; i < N * K; ++i) { fb[i] = (rand() % 13 - 6) / 4.0f; hb[i] = __float2half(fb[i]); }
    __half *da, *db; float* dout;
    cudaMalloc(&da, M * K * 2); cudaMalloc(&db, N * K * 2); cudaMalloc(&dout, M * N * 4);
    cudaMemcpy(da, ha.data(), M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), N * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, M * N * 4);
    Args args; args.out = dout; args.kblocks = K / BK;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {64, 128}; const cuuint32_t es[2] = {1, 1};
    CUresult r1 = enc(&args.tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, da, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&args.tm_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d %d\n", (int)r1, (int)r2);
    const int smem = kStages * kStageBytes + 1024 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<2, 192, smem>>>(args);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> out(M * N);
    cudaMemcpy(out.data(), dout, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)fa[i * K + k] * fb[j * K + k];
        const double err = fabs(ref - out[i * N + j]); if (err > maxerr) maxerr = err; if (err > 1e-3) ++bad;
    }
    printf("max abs err %.3g, bad %d of %d; out[0]=%g out[128*256+5]=%g out[255*256+255]=%g\n", maxerr, bad, M * N, out[0], out[128 * 256 + 5], out[255 * 256 + 255]);
    return 0;
}
