// Probe: how fast can 148 persistent CTAs stream a contiguous HBM buffer into shared memory with bulk copies,
// as a function of bytes per stage and ring depth?   nvcc -arch=sm_100a -O3 -o stream_probe stream_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../streamflow_b200/csrc/sm100_ptx.cuh"
using namespace sf;

__global__ void __launch_bounds__(96, 1) probe(const uint8_t* src, long long bytes_per_cta, int stage_bytes, int stages,
                                               int pieces, int hint) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[16], empty[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const uint8_t* base = src + bytes_per_cta * blockIdx.x;
    const int iters = static_cast<int>(bytes_per_cta / stage_bytes);
    if (warp == 0 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], stage_bytes);
            const int pb = stage_bytes / pieces;
            for (int q = 0; q < pieces; ++q) {
                if (hint) bulk_load_hint(smem + stage * stage_bytes + q * pb, base + (long long)it * stage_bytes + q * pb, pb, &full[stage], kEvictFirst);
                else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem + stage * stage_bytes + q * pb)), "l"(base + (long long)it * stage_bytes + q * pb), "r"(pb), "r"(smem_u32(&full[stage])) : "memory");
            }
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            mbar_wait(&full[stage], phase);
            mbar_arrive(&empty[stage]);
            if (++stage == stages) { stage = 0; phase ^= 1; }
        }
    }
}

int main() {
    const long long total = 296ll << 20;
    uint8_t* buf; cudaMalloc(&buf, total + (1 << 20)); cudaMemset(buf, 1, total);
    uint8_t* flush; cudaMalloc(&flush, 256 << 20);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int cfgs[][4] = {{16384, 5, 1, 1}, {16384, 10, 1, 1}, {32768, 5, 1, 1}, {32768, 6, 2, 1}, {18432, 3, 1, 1}, {18432, 6, 1, 1},
                           {18432, 9, 1, 1}, {18432, 11, 1, 1}, {18432, 9, 1, 0}, {18432, 9, 9, 1}, {36864, 5, 1, 1}, {8192, 16, 1, 1}, {8192, 8, 1, 1}, {65536, 3, 1, 1}};
    for (auto& c : cfgs) {
        const int sb = c[0], st = c[1], pieces = c[2], hint = c[3];
        const long long per_cta = (total / 148) / sb * sb;
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemsetAsync(flush, rep, 256 << 20);
            cudaEventRecord(e0);
            probe<<<148, 96, sb * st + 1024>>>(buf, per_cta, sb, st, pieces, hint);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        printf("stage %6d B x %2d stages, %d pieces, hint %d: %7.1f us  %6.0f GB/s  (%s)\n", sb, st, pieces, hint, best * 1e3,
               per_cta * 148 / (best * 1e-3) / 1e9, cudaGetErrorString(err));
    }
    return 0;
}
