// Probe: can TMA (cp.async.bulk.tensor, rank 4, one small box per query) do the lookup's window gather?
//  part 1  smem image of a (16 floats, ncol tiles, nrow tile-rows, 1 query) box under SWIZZLE_NONE / 64B / 128B, with
//          out-of-bounds coordinates, destination at a 1024-aligned address and at +128 (is the swizzle address-based?)
//  part 2  throughput: 84480 boxes (= 21120 queries x 4 levels at Sintel size) of 3-4 x 3-4 tiles at random positions of
//          a 605 MB tensor, issued by W producer warps per CTA with S stages of 32 boxes in flight per warp.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_gather_probe tma_gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../streamflow_b200/csrc/sm100_ptx.cuh"

using namespace sf;

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_fn() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return reinterpret_cast<EncodeFn>(p);
}
static CUtensorMap make_map(const float* base, int tw, int th, long long Q, int ncol, int nrow, CUtensorMapSwizzle sw) {
    CUtensorMap m;
    const cuuint64_t dims[4] = {16, (cuuint64_t)tw, (cuuint64_t)th, (cuuint64_t)Q};
    const cuuint64_t strides[3] = {64, (cuuint64_t)tw * 64, (cuuint64_t)tw * th * 64};
    const cuuint32_t box[4] = {16, (cuuint32_t)ncol, (cuuint32_t)nrow, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = get_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

__device__ __forceinline__ void tma_load_4d_hint(const CUtensorMap* m, uint64_t* bar, uint32_t dst, int c0, int c1, int c2,
                                                 int c3, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(hint)
        : "memory");
}

// ------------------------------------------------------------------ part 1
__global__ void dump_kernel(const __grid_constant__ CUtensorMap m, int tx0, int ty0, int q, int dst_off, float* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 8192);
    float* buf = reinterpret_cast<float*>(sm);
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = -1.f;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 3 * 4 * 64);
        tma_load_4d_hint(&m, bar, smem_u32(sm + dst_off), 0, tx0, ty0, q, kEvictLast);
    }
    mbar_wait(bar, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = buf[i];
}

// ------------------------------------------------------------------ part 2
constexpr int kSlot = 1152;
constexpr int kStageBytes = 32 * kSlot;

struct GatherParams {
    CUtensorMap map[4];     // shape id = (nrow - 3) * 2 + (ncol - 3)
    int Q, items, w, h, seed, stages;
    float* sink;
};

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <bool kShflIssue>
__global__ void __launch_bounds__(128) tma_gather_kernel(const __grid_constant__ GatherParams p) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
    const int S = p.stages;                                  // stages per warp
    unsigned char* ring = sm + warp * S * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + W * S * kStageBytes) + warp * 8;
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    float acc = 0.f;
    const int gw = blockIdx.x * W + warp, nw = gridDim.x * W;
    int k = 0;
    for (int it = gw; it < p.items; it += nw, ++k) {
        const int s = k % S;
        if (k >= S) {                                        // previous use of this stage has landed
            mbar_wait(&bars[s], ((k / S) - 1) & 1);
            acc += *reinterpret_cast<const float*>(ring + s * kStageBytes + lane * kSlot);
        }
        const unsigned hsh = hash32((it * 32 + lane) * 2654435761u + p.seed);
        const int x0 = static_cast<int>(hsh % (p.w + 2)) - 4, y0 = static_cast<int>((hsh >> 12) % (p.h + 2)) - 4;
        const int ox = x0 & 3, oy = y0 & 3, tx0 = x0 >> 2, ty0 = y0 >> 2;
        const int ncol = ox == 3 ? 4 : 3, nrow = oy == 3 ? 4 : 3;
        const int shape = (nrow - 3) * 2 + (ncol - 3);
        const int q = (it * 32 + lane) % p.Q;
        const unsigned bytes = nrow * ncol * 64;
        const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == 0) mbar_expect_tx(&bars[s], total);
        __syncwarp();
        const uint32_t dst = smem_u32(ring + s * kStageBytes + lane * kSlot);
        if constexpr (kShflIssue) {
#pragma unroll 1
            for (int i = 0; i < 32; ++i) {
                const int sh = __shfl_sync(0xffffffffu, shape, i), a = __shfl_sync(0xffffffffu, tx0, i),
                          b = __shfl_sync(0xffffffffu, ty0, i), c = __shfl_sync(0xffffffffu, q, i);
                const uint32_t d = __shfl_sync(0xffffffffu, dst, i);
                if (lane == 0) tma_load_4d_hint(&p.map[sh], &bars[s], d, 0, a, b, c, kEvictLast);
            }
        } else {
            tma_load_4d_hint(&p.map[shape], &bars[s], dst, 0, tx0, ty0, q, kEvictLast);
        }
    }
    const int n = k;
    for (int j = (n > S ? n - S : 0); j < n; ++j) mbar_wait(&bars[j % S], (j / S) & 1);
    if (acc == 123.456f) *p.sink = acc;
}

int main() {
    // ---- part 2
    {
        const int tw = 32, th = 14, Q = 21120, items = 2640;
        float* d;
        const size_t bytes = (size_t)Q * th * tw * 64;
        cudaMalloc(&d, bytes);
        cudaMemset(d, 0, bytes);
        GatherParams p{};
        for (int nr = 3; nr <= 4; ++nr)
            for (int nc = 3; nc <= 4; ++nc) p.map[(nr - 3) * 2 + (nc - 3)] = make_map(d, tw, th, Q, nc, nr, CU_TENSOR_MAP_SWIZZLE_64B);
        p.Q = Q; p.items = items; p.w = 128; p.h = 55;
        cudaMalloc(&p.sink, 4);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        struct Cfg { int ctas_per_sm, W, S; bool shfl; };
        const Cfg cfgs[] = {{1, 1, 5, false}, {1, 2, 3, false}, {1, 4, 1, false}, {2, 1, 3, false}, {2, 2, 1, false},
                            {1, 1, 5, true},  {1, 2, 3, true},  {2, 1, 3, true},  {1, 3, 2, false}, {1, 6, 1, false}};
        for (const Cfg& c : cfgs) {
            const int smem = c.W * c.S * kStageBytes + c.W * 64;
            auto kern = c.shfl ? tma_gather_kernel<true> : tma_gather_kernel<false>;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
                printf("cfg skipped (smem %d)\n", smem);
                cudaGetLastError();
                continue;
            }
            p.stages = c.S;
            for (int same = 0; same < 2; ++same) {
                float best = 1e9f, sum = 0.f;
                const int reps = 20;
                for (int rep = 0; rep < reps + 3; ++rep) {
                    p.seed = same ? 7 : 1000 + rep;
                    cudaEventRecord(e0);
                    kern<<<148 * c.ctas_per_sm, 32 * c.W, smem>>>(p);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (rep >= 3) { sum += ms; if (ms < best) best = ms; }
                }
                const double tiles = items * 32.0 * 3.25 * 3.25;
                printf("ctas/SM %d  warps %d  stages/warp %d  %s  %s coords: best %6.2f us  mean %6.2f us  -> %5.1f G tiles/s  (%s)\n",
                       c.ctas_per_sm, c.W, c.S, c.shfl ? "shfl-issue" : "lane-issue", same ? "same " : "fresh", best * 1e3,
                       sum / reps * 1e3, tiles / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
            }
        }
    }
    // ---- part 1
    {
        const int tw = 6, th = 5, Q = 4;
        std::vector<float> h(Q * th * tw * 16);
        for (int q = 0; q < Q; ++q)
            for (int ty = 0; ty < th; ++ty)
                for (int tx = 0; tx < tw; ++tx)
                    for (int i = 0; i < 16; ++i) h[((q * th + ty) * tw + tx) * 16 + i] = q * 10000 + ty * 1000 + tx * 100 + i;
        float *d, *out;
        cudaMalloc(&d, h.size() * 4);
        cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        cudaMalloc(&out, 8192);
        std::vector<float> o(2048);
        const CUtensorMapSwizzle sws[3] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
        const char* names[3] = {"none", "64B", "128B"};
        cudaFuncSetAttribute(dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
        for (int v = 0; v < 2; ++v)
            for (int off : {64}) {
                CUtensorMap m = make_map(d, tw, th, Q, 3, 4, sws[v]);
                dump_kernel<<<1, 128, 16384>>>(m, -1, 2, 1, off, out);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(o.data(), out, 8192, cudaMemcpyDeviceToHost);
                printf("swizzle %s dst+%d (%s): box (16, 3 cols from tx=-1, 4 rows from ty=2, q=1); 16-byte chunks (value of first float, -1 = untouched):\n",
                       names[v], off, cudaGetErrorString(e));
                for (int c = (off / 128) * 8; c < (off / 128) * 8 + 64; ++c) {
                    if (c % 8 == 0) printf("  byte %4d:", c * 16);
                    printf(" %6.0f", o[c * 4]);
                    if (c % 8 == 7) printf("\n");
                }
            }
    }
    return 0;
}
