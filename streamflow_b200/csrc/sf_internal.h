// Internal declarations shared by the kernel translation units and the C-ABI layer (api.cu).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/streamcorr.h"

namespace sf {

void set_error(const char* fmt, ...);

#define SF_CUDA_CHECK(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::sf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                            __LINE__);                                                        \
            return SF_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

#define SF_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            ::sf::set_error(__VA_ARGS__); \
            return SF_ERR_INVALID;       \
        }                                \
    } while (0)

// launch accounting + optional event bracketing of one kernel kind (see sf_profile_kernel)
int debug_gma_mask();
int debug_corr_mask();
void prof_before(int kind, cudaStream_t s);
void prof_after(int kind, cudaStream_t s);

// Per-query correlation images are stored as 4x4 tiles (64 B each, tile-row-major): the lookup's 10x10 window then
// touches ~3.25 x 3.25 sixty-four-byte blocks instead of 10 rows x 1.6 blocks of a row-major image.
// Programmatic dependent launch: every kernel is launched with the stream-serialisation attribute, releases its
// dependents right away (pdl_launch) and waits for its predecessor's memory (pdl_wait) only after its own
// prologue (barrier init, TMEM allocation, descriptor prefetch, index setup), so launch latency and prologues of
// the ~60 kernels of a step overlap the tail of the previous kernel.  STREAMCORR_PDL=0 disables the attribute.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                 Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// Same for a kernel whose CTAs form clusters of `cluster_x` along x (grid.x must be a multiple of it).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                         unsigned cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

struct LevelGeom {
    int h[SF_NUM_LEVELS], w[SF_NUM_LEVELS];       // valid cells
    int th[SF_NUM_LEVELS], tw[SF_NUM_LEVELS];     // tiles: ceil(h_l / 4), ceil(w_l / 4)
    long long img[SF_NUM_LEVELS];                 // floats per query image = th * tw * 16 (cells past h_l, w_l are 0)
};
// offset of cell (v, u) inside a tiled image with `tw` tiles per row
__host__ __device__ inline int tiled_offset(int v, int u, int tw) {
    return (((v >> 2) * tw + (u >> 2)) << 4) + ((v & 3) << 2) + (u & 3);
}
LevelGeom make_level_geom(int64_t h, int64_t w);

struct DeviceInfo {
    int ok = -1;      // -1 unknown, 0 unusable, 1 usable
    int sms = 0;
    int dev = -1;
};
// Opt the kernel `func` in to `bytes` of dynamic shared memory on the CURRENT device (remembered per device).
int ensure_dynamic_smem(const void* func, int bytes);
// SF_OK and fills `out` when the current device is sm_100; SF_ERR_NODEVICE (with message) otherwise.
int query_device(DeviceInfo* out);
// rank-3 tiled TMA map (128-byte swizzle by default; 64 or 0 = none): dims / box innermost first; strides (bytes) of dims 1 and 2; box[2] = 1.
int make_tmap3(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
               uint64_t stride1, uint64_t stride2, uint32_t b0, uint32_t b1, const char* what, int swizzle_bytes = 128);
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Power-of-two operand scale derived from a tensor's absmax (bits of a non-negative float):
// amax * 2^e lands in [2^13, 2^14) so fp16 never overflows; e is clamped to [-40, 40].
__host__ __device__ inline int scale_exponent_from_bits(unsigned bits) {
    const int biased = static_cast<int>((bits >> 23) & 0xFF);
    if (bits == 0u || biased == 0 || biased == 0xFF) return 0;   // zero / denormal / inf-nan: no scaling
    int e = 13 - (biased - 127);
    return e < -40 ? -40 : (e > 40 ? 40 : e);
}

// ------------------------------------------------------------------ lookup (corr_lookup.cu)
struct LookupParams {
    const float* lvl[SF_MAX_GROUPS][SF_NUM_LEVELS];
    const float* coords[SF_MAX_GROUPS];
    void* out[SF_MAX_GROUPS];
    int hl[SF_NUM_LEVELS], wl[SF_NUM_LEVELS], th[SF_NUM_LEVELS], tw[SF_NUM_LEVELS];
    long long img[SF_NUM_LEVELS];
    int N;            // h * w
    long long BN;     // queries per group
    int out_f16;
    long long tiles, items;   // filled by the launcher: 32-query tiles per group, tiles * groups * levels
};
int launch_corr_lookup(const LookupParams& p, int groups, int num_sms, cudaStream_t s);

// ------------------------------------------------------------- operand packing (corr_pack.cu)
struct PackParams {
    const float* src[2];                 // fmap1 (-> A), fmap2 (-> B levels): [B, D, h, w], element strides below
    long long sb[2], sk[2], sy[2], sx[2];
    __half* dst_a;                       // [B, N, Kp] fp16, K contiguous
    __half* dst_b[SF_NUM_LEVELS];        // [B, th_l*tw_l*16, Kp]; row m = tiled_offset(v, u, tw_l)
    int h, w, D, split;                  // split: 0 = [hi] only; 1 = Kp = 3 * D with A = [hi | hi*2^-11 | lo*2^11],
                                         //        B = [hi | lo*2^11 | hi*2^-11]; 2 = decided by amax_bits[2] (auto):
                                         //        inexact inputs -> as 1, fp16-exact inputs -> A = [hi | hi*2^-11],
                                         //        B_0 = [hi], B_l>0 = [hi | lo*2^11]
    int kp;                              // row pitch of every packed operand in elements (D or 3 * D)
    int hl[SF_NUM_LEVELS], wl[SF_NUM_LEVELS], th[SF_NUM_LEVELS], tw[SF_NUM_LEVELS], rows[SF_NUM_LEVELS];
    int bx, by;                          // 8x8 source-pixel blocks per image
    const unsigned* amax_bits;           // [3] absmax bits of fmap1, fmap2; OR of the low 13 mantissa bits of all values
};
int launch_absmax2(const float* f1, const float* f2, int64_t B, int64_t D, int64_t h, int64_t w,
                   const int64_t s1[4], const int64_t s2[4], unsigned* amax_bits, cudaStream_t s);
int launch_corr_pack(const PackParams& p, int64_t B, cudaStream_t s);

// --------------------------------------------------------- strict fp32 path (corr_simt.cu)
int launch_corr_simt(const float* f1, const float* f2, int64_t B, int64_t D, int64_t h, int64_t w,
                     const int64_t s1[4], const int64_t s2[4], float* const levels[SF_NUM_LEVELS],
                     float* ws_a, float* ws_b, cudaStream_t s);

// ----------------------------------------------------- tcgen05 GEMM (corr_gemm_sm100.cu)
struct CorrGemmParams {
    int B, N, Kp;                          // batch (pairs), queries per pair, packed K pitch
    int mode;                              // 0: kb_single k-blocks everywhere; 1: kb_split; 2 (auto): amax_bits[2] != 0 ->
                                           // kb_split, else level 0 kb_single and pooled levels kb_pool
    int kb_single, kb_split, kb_pool;
    int m_tiles;                           // ceil(N / 128)
    int n_tiles[SF_NUM_LEVELS];            // ceil(rows_l / 256)
    int n_tiles_total;
    const unsigned* amax_bits;             // [2] operand scales (device)
    float inv_sqrt_d;
};
int launch_corr_gemm(const CorrGemmParams& p, const CUtensorMap& tm_a, const CUtensorMap tm_b[SF_NUM_LEVELS],
                     const int n_cols[SF_NUM_LEVELS], float* const levels[SF_NUM_LEVELS], int num_sms,
                     cudaStream_t s);

// ----------------------------------------------------------------------- GMA (gma_sm100.cu)
struct GmaProjParams {
    const void* x;        // [P, C, N]
    int x_dtype;
    const float* w;       // [O, C] rows o0 .. o0+O-1 used
    int P, C, N, O;
    float scale;          // multiplies the result (q gets d^-1/2)
    __half* out;          // token-major: [P, Nrows, ldo] (ldo >= O) or channel-major: [P, O, ldn]
    long long out_batch_stride;
    int ld;               // row pitch of out in elements
    int token_major;
    int split;            // token-major only: also write lo/hi split parts at column offsets O and 2*O
    int is_b;
    // optional second projection in the same launch (blockIdx.z == 1): k next to q
    const void* x2;
    const float* w2;
    float scale2;
    __half* out2;
    // optional: clear per-token accumulators of the following stats passes ([P, N] each) instead of two memsets
    unsigned* zero_u32;
    unsigned long long* zero_u64;
};
int launch_gma_proj(const GmaProjParams& p, cudaStream_t s);
// Per-iteration operand preparation of the aggregate: x [P, C, N] (any float dtype) -> x16 [P, C, Npad] fp16 (K-major
// over keys = NCHW, pad columns zero) and w_v [d, C] (fp32 or fp16) -> w16 [d, C] fp16.
int launch_gma_cast(const void* x, int x_dtype, __half* x16, int64_t rows, int64_t N, int64_t Npad, const void* w,
                    int w_dtype, __half* w16, int64_t w_elems, cudaStream_t s);

struct GmaStatsParams {
    int P, N, Npad, Kp;
    int split;                  // Kp = 2d, q = [hi | lo], k = [hi | lo] (pass 1 uses the hi halves only)
    int m_tiles, n_tiles;       // ceil(N/128), ceil(N/256)
    int pair_tiles;             // ceil(m_tiles / 2): a CTA pair owns two query tiles
    int chunks;                 // key-chunks per m-tile (work split)
    unsigned* rowmax_bits;      // [P, N] ordered-int encoded running max (pass 1 out / pass 2 in)
    unsigned long long* rowsum_fx;   // [P, N] sum of stored E in units of 2^-24 (pass 2 out): every fp16 value is a
                                     // multiple of 2^-24, so integer atomics make the sum exact and order-independent
    __half* E;                  // [P, N, Npad]
    int pass;
};
int launch_gma_stats(const GmaStatsParams& p, const CUtensorMap& tm_q, const CUtensorMap& tm_k, int num_sms,
                     cudaStream_t s);
int launch_gma_rowsum_finish(const unsigned long long* fx, float* rowsum, long long n, cudaStream_t s);

struct GmaAggParams {
    int P, N, Npad, C;          // C == d == 128
    int k_blocks;               // Npad / 64
    const float* rowsum;        // [P, N] softmax denominators (sum of the stored numerators)
    const float* gamma;         // device scalar
    const void* fmap;           // [P, C, N]
    int fmap_dtype;
    float* out;                 // [P, C, N]
    const __half* e_ptr;        // tile-major E (swizzled 16 KB blocks) and its per-map stride in elements
    long long e_map_stride;
};
// `settled`: one word of the GMA workspace, 0 after sf_gma_attention*, set by every aggregate launch (see the kernel)
int launch_gma_aggregate(const GmaAggParams& p, const CUtensorMap& tm_x, const CUtensorMap& tm_w, int num_sms,
                         unsigned* settled, cudaStream_t s);
int launch_gma_identity(float* dst, int d, cudaStream_t s);      // dst[d, d] <- I

// y = gelu(x + W2 . gelu(W1 . x + b1) + b2) per pixel (pcblock_ffn1_sm100.cu); w1p [Hp, Kp], w2p [ceil16(C), Hp] fp16 padded (Hp = ceil128(hidden), Kp = ceil64(C))
int launch_pcblock_ffn1(const void* x, int x_dtype, const void* w1p, const float* b1p, const void* w2p, const float* b2,
                        void* out, int out_dtype, int64_t P, int64_t C, int64_t Hd, int64_t N, cudaStream_t s);

void set_ffn1_trace(void* dev_ptr);     // debug: 32 x u64 of %globaltimer stamps (see pcblock_ffn1_sm100.cu)

int launch_upsample_flow(const float* flow, const void* mask, int mask_dtype, float* out, int64_t N, int64_t H,
                         int64_t W, cudaStream_t s);

}  // namespace sf
