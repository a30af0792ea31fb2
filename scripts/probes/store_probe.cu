// Probe: HBM write throughput for a 264 MB buffer: plain 16-byte stores vs bulk smem->global copies of S bytes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../streamflow_b200/csrc/sm100_ptx.cuh"
using namespace sf;

__global__ void __launch_bounds__(256) plain_store(float4* dst, long long n16) {
    const long long stride = (long long)gridDim.x * 256;
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n16; i += stride) __stcs(dst + i, v);
}

// each CTA owns a contiguous slice and writes it as bulk copies of `sz` bytes from one smem buffer, `inflight` groups
__global__ void __launch_bounds__(128, 1) bulk_store(uint8_t* dst, long long bytes_per_cta, int sz, int inflight) {
    extern __shared__ __align__(128) uint8_t smem[];
    for (int i = threadIdx.x; i < sz / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint8_t* base = dst + bytes_per_cta * blockIdx.x;
        const int iters = static_cast<int>(bytes_per_cta / sz);
        for (int it = 0; it < iters; ++it) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + (long long)it * sz),
                         "r"(smem_u32(smem)), "r"(sz) : "memory");
            tma_store_commit();
            if (inflight == 1) tma_store_wait_read<0>();
            else if (inflight == 4) tma_store_wait_read<3>();
            else tma_store_wait_read<7>();
        }
        tma_store_wait_all<0>();
    }
}

int main() {
    const long long total = 264ll << 20;
    uint8_t* buf; cudaMalloc(&buf, total); 
    uint8_t* flush; cudaMalloc(&flush, 256 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(bulk_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    auto report = [&](const char* name, float ms) {
        printf("%-44s %7.1f us  %6.0f GB/s (%s)\n", name, ms * 1e3, total / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    for (int ctas : {148 * 4, 148 * 8, 148 * 16}) {
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemsetAsync(flush, rep, 256 << 20);
            cudaEventRecord(e0);
            plain_store<<<ctas, 256>>>(reinterpret_cast<float4*>(buf), total / 16);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        char name[64]; snprintf(name, 64, "plain st.cs 16 B, %d CTAs", ctas); report(name, best);
    }
    const int cfg[][2] = {{4096, 4}, {4096, 8}, {8192, 8}, {16384, 4}, {16384, 8}, {32768, 4}, {65536, 4}};
    for (auto& c : cfg) {
        const long long per = (total / 148) / c[0] * c[0];
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemsetAsync(flush, rep, 256 << 20);
            cudaEventRecord(e0);
            bulk_store<<<148, 128, c[0]>>>(buf, per, c[0], c[1]);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        char name[64]; snprintf(name, 64, "bulk store %d B x %d in flight, 148 CTAs", c[0], c[1]); report(name, best);
    }
    // memset as the library reference
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0); cudaMemsetAsync(buf, 0, total); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    report("cudaMemsetAsync", best);
    return 0;
}
