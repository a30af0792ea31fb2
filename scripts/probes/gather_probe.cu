// Probe: HBM throughput for random aligned reads of G bytes (G = 32, 64, 128, 256) out of an 800 MB buffer.
// Each group of G/16 consecutive threads reads one random G-byte block as 16-byte loads; sums keep the loads alive.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int G>
__global__ void __launch_bounds__(256) gather(const float4* __restrict__ src, unsigned long long nblocks,
                                              int per_thread, float* sink) {
    constexpr int T = G / 16;
    const unsigned long long gtid = blockIdx.x * 256ull + threadIdx.x;
    const unsigned long long grp = gtid / T;
    const int sub = gtid % T;
    float acc = 0.f;
    unsigned long long x = grp * 0x9E3779B97F4A7C15ull + 12345;
#pragma unroll 8
    for (int i = 0; i < per_thread; ++i) {
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        const unsigned long long blk = x % nblocks;
        const float4 v = __ldcs(src + blk * T + sub);
        acc += v.x + v.w;
    }
    if (acc == 123.456f) *sink = acc;
}

template <int G>
void run(const float4* buf, size_t bytes, float* sink) {
    const unsigned long long nblocks = bytes / G;
    const int per_thread = 32;
    const int ctas = 148 * 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        gather<G><<<ctas, 256>>>(buf, nblocks, per_thread, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double total = double(ctas) * 256 * per_thread * 16;
    printf("random %3d-byte blocks: %8.1f us for %.0f MB -> %6.0f GB/s (%s)\n", G, best * 1e3, total / 1e6,
           total / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
    const size_t bytes = (argc > 1 ? static_cast<size_t>(atoi(argv[1])) : 800ull) << 20;
    printf("buffer %zu MB\n", bytes >> 20);
    float4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    float* sink; cudaMalloc(&sink, 4);
    run<32>(buf, bytes, sink); run<64>(buf, bytes, sink); run<128>(buf, bytes, sink); run<256>(buf, bytes, sink);
    run<512>(buf, bytes, sink);
    return 0;
}
