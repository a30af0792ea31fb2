// C ABI of libstreamcorr.so (declared in include/streamcorr.h): argument validation, workspace carving,
// TMA tensor-map construction and kernel launches.  No torch types, no allocation, no host synchronisation.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "sf_internal.h"

namespace sf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static thread_local int64_t g_launches = 0;
static thread_local int g_prof_kind = 0;
static thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;

static thread_local int g_gma_mask = 7, g_corr_mask = 3;
int debug_gma_mask() { return g_gma_mask; }
int debug_corr_mask() { return g_corr_mask; }

void prof_before(int kind, cudaStream_t s) {
    ++g_launches;
    if (kind == g_prof_kind && g_prof_start) cudaEventRecord(g_prof_start, s);
}
void prof_after(int kind, cudaStream_t s) {
    if (kind == g_prof_kind && g_prof_stop) cudaEventRecord(g_prof_stop, s);
}

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("STREAMCORR_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

LevelGeom make_level_geom(int64_t h, int64_t w) {
    LevelGeom g;
    for (int l = 0; l < SF_NUM_LEVELS; ++l) {
        g.h[l] = static_cast<int>(h >> l);
        g.w[l] = static_cast<int>(w >> l);
        g.th[l] = (g.h[l] + 3) >> 2;
        g.tw[l] = (g.w[l] + 3) >> 2;
        g.img[l] = static_cast<long long>(g.th[l]) * g.tw[l] * 16;
    }
    return g;
}

int query_device(DeviceInfo* out) {
    static thread_local DeviceInfo cache;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        set_error("no CUDA device available: libstreamcorr has no CPU fallback");
        return SF_ERR_NODEVICE;
    }
    if (cache.ok < 0 || cache.dev != dev) {
        int major = 0, sms = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            cudaGetLastError();
            set_error("cannot query CUDA device %d", dev);
            return SF_ERR_NODEVICE;
        }
        cache.dev = dev;
        cache.sms = sms;
        cache.ok = (major == 10) ? 1 : 0;
        // STREAMCORR_L2_FETCH=32|64|128 (opt-in, measurement only) changes cudaLimitMaxL2FetchGranularity for the
        // WHOLE device/process -- it also applies to every other kernel (cuDNN / cuBLAS of the update block), so the
        // library never touches it unless asked; it made no measurable difference to the lookup.
        if (cache.ok == 1) {
            const char* env = getenv("STREAMCORR_L2_FETCH");
            const int gran = env ? atoi(env) : 0;
            if (gran == 32 || gran == 64 || gran == 128) {
                if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, static_cast<size_t>(gran)) != cudaSuccess)
                    cudaGetLastError();
            }
        }
    }
    if (cache.ok != 1) {
        set_error("device %d is not sm_100 (Blackwell B200): kernels are built for sm_100a only, no fallback", dev);
        return SF_ERR_NODEVICE;
    }
    *out = cache;
    return SF_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): remember what each device already has so a
// launch costs no driver call, and a second device in the same process (nn.DataParallel, evaluate_mf.py:1207) gets
// its own opt-in instead of an invalid-value launch failure.
int ensure_dynamic_smem(const void* func, int bytes) {
    int dev = 0;
    SF_CUDA_CHECK(cudaGetDevice(&dev));
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> configured;
    std::lock_guard<std::mutex> lock(mu);
    int& cur = configured[std::make_pair(func, dev)];
    if (bytes > cur) {
        SF_CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        cur = bytes;
    }
    return SF_OK;
}

namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

}  // namespace

int make_tmap3(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
               uint64_t stride1, uint64_t stride2, uint32_t b0, uint32_t b1, const char* what, int swizzle_bytes) {
    // A tensor map is a pure function of its arguments (the driver call costs ~1-2 us, and a step re-encodes the
    // same ~10 maps for every clip of a stream): small per-thread direct-mapped cache keyed by all of them.
    struct Entry {
        uint64_t key[9];
        bool valid = false;
        CUtensorMap map;
    };
    constexpr int kSlots = 64;
    static thread_local Entry cache[kSlots];
    const uint64_t key[9] = {static_cast<uint64_t>(dt), reinterpret_cast<uint64_t>(base), d0, d1, d2, stride1, stride2,
                             (static_cast<uint64_t>(b0) << 32) | b1, static_cast<uint64_t>(swizzle_bytes)};
    uint64_t hsh = 1469598103934665603ull;
    for (uint64_t k : key) hsh = (hsh ^ k) * 1099511628211ull;
    Entry& e = cache[(hsh >> 17) % kSlots];
    if (e.valid && memcmp(e.key, key, sizeof(key)) == 0) {
        *m = e.map;
        return SF_OK;
    }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return SF_ERR_CUDA;
    }
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {stride1, stride2};
    const cuuint32_t box[3] = {b0, b1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(m, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d (dims %llu,%llu,%llu strides %llu,%llu)", what,
                  static_cast<int>(r), (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                  (unsigned long long)stride1, (unsigned long long)stride2);
        return SF_ERR_CUDA;
    }
    memcpy(e.key, key, sizeof(key));
    e.map = *m;
    e.valid = true;
    return SF_OK;
}

namespace {

int check_corr_shape(int64_t B, int64_t D, int64_t h, int64_t w) {
    SF_REQUIRE(B >= 1 && D >= 1 && h >= 1 && w >= 1, "corr: non-positive shape B=%lld D=%lld h=%lld w=%lld",
               (long long)B, (long long)D, (long long)h, (long long)w);
    // the reference divides by (h_l - 1) and (w_l - 1) when normalising (core/utils/utils.py:69-70)
    SF_REQUIRE((h >> (SF_NUM_LEVELS - 1)) >= 2 && (w >> (SF_NUM_LEVELS - 1)) >= 2,
               "corr: %lldx%lld is too small for a %d-level pyramid (coarsest level must be at least 2x2)",
               (long long)h, (long long)w, SF_NUM_LEVELS);
    SF_REQUIRE(B * h * w < (1ll << 31) && h * w * 2 < (1ll << 31), "corr: shape too large");
    return SF_OK;
}

struct CorrWs {
    int64_t amax_off, a_off, b_off[SF_NUM_LEVELS], total;
    int Kp;
};

CorrWs corr_ws_layout(int64_t B, int64_t D, int64_t h, int64_t w, int precision) {
    CorrWs ws{};
    const LevelGeom g = make_level_geom(h, w);
    const int64_t N = h * w;
    if (precision == SF_PREC_FP32_SIMT) {
        ws.a_off = 0;
        ws.b_off[0] = align_up(B * D * N * 4, 256);
        ws.total = ws.b_off[0] + align_up(B * D * N * 4, 256);
        ws.Kp = static_cast<int>(D);
        return ws;
    }
    ws.Kp = static_cast<int>((precision == SF_PREC_F16X2 || precision == SF_PREC_AUTO) ? 3 * D : D);
    int64_t off = 0;
    ws.amax_off = off;
    off += 256;
    ws.a_off = off;
    off += align_up(B * N * ws.Kp * 2, 1024);
    for (int l = 0; l < SF_NUM_LEVELS; ++l) {
        ws.b_off[l] = off;
        off += align_up(B * g.img[l] * ws.Kp * 2, 1024);
    }
    ws.total = off;
    return ws;
}

}  // namespace
}  // namespace sf

using namespace sf;

extern "C" {

int sf_version(void) { return SF_VERSION; }

const char* sf_last_error(void) { return g_err; }

int64_t sf_launch_count(void) { return g_launches; }

void sf_profile_kernel(int which, void* start, void* stop) {
    g_prof_kind = (start && stop) ? which : 0;
    g_prof_start = static_cast<cudaEvent_t>(start);
    g_prof_stop = static_cast<cudaEvent_t>(stop);
}

void sf_debug_select_kernels(int gma_aggregate_mask, int corr_build_mask) {
    g_gma_mask = gma_aggregate_mask;
    g_corr_mask = corr_build_mask;
}

int sf_device_ok(void) {
    DeviceInfo di;
    return query_device(&di);
}

void sf_corr_level_dims(int64_t h, int64_t w, int level, int64_t* h_l, int64_t* w_l, int64_t* tiles_y,
                        int64_t* tiles_x) {
    const int64_t hl = h >> level, wl = w >> level;
    if (h_l) *h_l = hl;
    if (w_l) *w_l = wl;
    if (tiles_y) *tiles_y = (hl + 3) >> 2;
    if (tiles_x) *tiles_x = (wl + 3) >> 2;
}

int64_t sf_corr_workspace_bytes(int64_t B, int64_t D, int64_t h, int64_t w, int precision) {
    if (B < 1 || D < 1 || h < 1 || w < 1) return 0;
    return corr_ws_layout(B, D, h, w, precision).total;
}

int sf_corr_build(const float* fmap1, const float* fmap2, int64_t B, int64_t D, int64_t h, int64_t w,
                  const int64_t f1_strides[4], const int64_t f2_strides[4], float* const levels[SF_NUM_LEVELS],
                  void* workspace, int64_t workspace_bytes, int precision, void* stream) {
    DeviceInfo di;
    if (int rc = query_device(&di)) return rc;
    if (int rc = check_corr_shape(B, D, h, w)) return rc;
    SF_REQUIRE(fmap1 && fmap2 && levels && f1_strides && f2_strides, "corr_build: null pointer argument");
    SF_REQUIRE(precision == SF_PREC_F16 || precision == SF_PREC_F16X2 || precision == SF_PREC_FP32_SIMT ||
                   precision == SF_PREC_AUTO,
               "corr_build: unknown precision mode %d", precision);
    // the exact-input fast path of SF_PREC_AUTO addresses k-blocks of 64: other D run the three-product mode
    if (precision == SF_PREC_AUTO && D % 64 != 0) precision = SF_PREC_F16X2;
    for (int l = 0; l < SF_NUM_LEVELS; ++l)
        SF_REQUIRE(levels[l] && (reinterpret_cast<uintptr_t>(levels[l]) & 15) == 0,
                   "corr_build: level %d buffer is null or not 16-byte aligned", l);
    const CorrWs ws = corr_ws_layout(B, D, h, w, precision);
    if (!workspace || workspace_bytes < ws.total) {
        set_error("corr_build: workspace of %lld bytes needed, %lld given", (long long)ws.total,
                  (long long)workspace_bytes);
        return SF_ERR_WORKSPACE;
    }
    SF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "corr_build: workspace must be 1024-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    const LevelGeom g = make_level_geom(h, w);
    const int64_t N = h * w;

    if (precision == SF_PREC_FP32_SIMT) {
        return launch_corr_simt(fmap1, fmap2, B, D, h, w, f1_strides, f2_strides, levels,
                                reinterpret_cast<float*>(wsb + ws.a_off), reinterpret_cast<float*>(wsb + ws.b_off[0]),
                                s);
    }

    SF_REQUIRE(D % 8 == 0, "corr_build: tensor-core modes need D %% 8 == 0 (got %lld); use SF_PREC_FP32_SIMT",
               (long long)D);
    unsigned* amax = reinterpret_cast<unsigned*>(wsb + ws.amax_off);
    const int parts = debug_corr_mask();
    if (parts & 1)
        if (int rc = launch_absmax2(fmap1, fmap2, B, D, h, w, f1_strides, f2_strides, amax, s)) return rc;

    PackParams pp{};
    pp.src[0] = fmap1; pp.src[1] = fmap2;
    pp.sb[0] = f1_strides[0]; pp.sk[0] = f1_strides[1]; pp.sy[0] = f1_strides[2]; pp.sx[0] = f1_strides[3];
    pp.sb[1] = f2_strides[0]; pp.sk[1] = f2_strides[1]; pp.sy[1] = f2_strides[2]; pp.sx[1] = f2_strides[3];
    pp.dst_a = reinterpret_cast<__half*>(wsb + ws.a_off);
    pp.h = (int)h; pp.w = (int)w; pp.D = (int)D;
    pp.split = (precision == SF_PREC_F16X2) ? 1 : (precision == SF_PREC_AUTO ? 2 : 0);
    pp.kp = ws.Kp;
    for (int l = 0; l < SF_NUM_LEVELS; ++l) {
        pp.dst_b[l] = reinterpret_cast<__half*>(wsb + ws.b_off[l]);
        pp.hl[l] = g.h[l]; pp.wl[l] = g.w[l]; pp.th[l] = g.th[l]; pp.tw[l] = g.tw[l]; pp.rows[l] = (int)g.img[l];
    }
    pp.bx = (int)((w + 7) / 8); pp.by = (int)((h + 7) / 8);
    pp.amax_bits = amax;
    // tile-grid cells of levels 2-3 can lie outside every 8x8 source block: clear those (small) operands first
    if (parts & 1) {
        SF_CUDA_CHECK(cudaMemsetAsync(wsb + ws.b_off[2], 0, ws.total - ws.b_off[2], s));
        if (int rc = launch_corr_pack(pp, B, s)) return rc;
    }
    if (!(parts & 2)) return SF_OK;

    CorrGemmParams gp{};
    gp.B = (int)B; gp.N = (int)N; gp.Kp = ws.Kp;
    gp.mode = pp.split;
    gp.kb_single = (int)((D + 63) / 64);
    gp.kb_split = (int)((3 * D + 63) / 64);
    gp.kb_pool = (int)((2 * D + 63) / 64);
    gp.m_tiles = (int)((N + 127) / 128);
    gp.n_tiles_total = 0;
    int n_cols[SF_NUM_LEVELS];
    for (int l = 0; l < SF_NUM_LEVELS; ++l) {
        gp.n_tiles[l] = (int)((g.img[l] + 255) / 256);
        gp.n_tiles_total += gp.n_tiles[l];
        n_cols[l] = (int)g.img[l];
    }
    gp.amax_bits = amax;
    gp.inv_sqrt_d = 1.0f / sqrtf(static_cast<float>(D));

    CUtensorMap tm_a, tm_b[SF_NUM_LEVELS];
    const uint64_t kp = static_cast<uint64_t>(ws.Kp);
    if (int rc = make_tmap3(&tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, wsb + ws.a_off, kp, N, B, kp * 2, N * kp * 2, 64,
                            128, "A"))
        return rc;
    for (int l = 0; l < SF_NUM_LEVELS; ++l) {
        const uint64_t rows = static_cast<uint64_t>(g.img[l]);
        if (int rc = make_tmap3(&tm_b[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, wsb + ws.b_off[l], kp, rows, B, kp * 2,
                                rows * kp * 2, 64, 128, "B"))
            return rc;
    }
    return launch_corr_gemm(gp, tm_a, tm_b, n_cols, levels, di.sms, s);
}

static int lookup_common(int G, const float* const* levels, const float* const* coords, void* const* out,
                         int out_dtype, int64_t B, int64_t h, int64_t w, int radius, int num_levels, void* stream) {
    DeviceInfo di;
    if (int rc = query_device(&di)) return rc;
    if (int rc = check_corr_shape(B, 1, h, w)) return rc;
    SF_REQUIRE(radius == SF_RADIUS && num_levels == SF_NUM_LEVELS,
               "corr_lookup: specialised for radius=%d, num_levels=%d (got %d, %d); no generic fallback", SF_RADIUS,
               SF_NUM_LEVELS, radius, num_levels);
    SF_REQUIRE(G >= 1 && G <= SF_MAX_GROUPS, "corr_lookup: group count %d outside [1, %d]", G, SF_MAX_GROUPS);
    SF_REQUIRE(out_dtype == SF_DT_F32 || out_dtype == SF_DT_F16, "corr_lookup: unsupported output dtype %d", out_dtype);
    SF_REQUIRE(levels && coords && out, "corr_lookup: null pointer argument");
    const LevelGeom g = make_level_geom(h, w);
    LookupParams p{};
    for (int gi = 0; gi < G; ++gi) {
        for (int l = 0; l < SF_NUM_LEVELS; ++l) {
            const float* lp = levels[gi * SF_NUM_LEVELS + l];
            SF_REQUIRE(lp && (reinterpret_cast<uintptr_t>(lp) & 15) == 0,
                       "corr_lookup: level buffer (group %d, level %d) null or not 16-byte aligned", gi, l);
            p.lvl[gi][l] = lp;
        }
        SF_REQUIRE(coords[gi] && out[gi], "corr_lookup: null coords/out for group %d", gi);
        p.coords[gi] = coords[gi];
        p.out[gi] = out[gi];
    }
    for (int l = 0; l < SF_NUM_LEVELS; ++l) {
        p.hl[l] = g.h[l]; p.wl[l] = g.w[l]; p.th[l] = g.th[l]; p.tw[l] = g.tw[l]; p.img[l] = g.img[l];
    }
    p.N = (int)(h * w);
    p.BN = B * h * w;
    p.out_f16 = (out_dtype == SF_DT_F16);
    return launch_corr_lookup(p, G, di.sms, static_cast<cudaStream_t>(stream));
}

int sf_corr_lookup(const float* const levels[SF_NUM_LEVELS], const float* coords, float* out, int64_t B, int64_t h,
                   int64_t w, int radius, int num_levels, void* stream) {
    void* o = out;
    return lookup_common(1, levels, &coords, &o, SF_DT_F32, B, h, w, radius, num_levels, stream);
}

int sf_corr_lookup_group(int G, const float* const* levels, const float* const* coords, void* const* out,
                         int out_dtype, int64_t B, int64_t h, int64_t w, int radius, int num_levels, void* stream) {
    return lookup_common(G, levels, coords, out, out_dtype, B, h, w, radius, num_levels, stream);
}

void sf_debug_ffn1_trace(void* dev_ptr) { set_ffn1_trace(dev_ptr); }

int sf_pcblock_ffn1(const void* x, int x_dtype, const void* w1p, const float* b1p, const void* w2p, const float* b2,
                    void* out, int out_dtype, int64_t P, int64_t C, int64_t hidden, int64_t N, void* stream) {
    DeviceInfo di;
    if (int rc = query_device(&di)) return rc;
    SF_REQUIRE(x && w1p && b1p && w2p && b2 && out, "pcblock_ffn1: null pointer argument");
    SF_REQUIRE((x_dtype == SF_DT_F32 || x_dtype == SF_DT_F16) && (out_dtype == SF_DT_F32 || out_dtype == SF_DT_F16),
               "pcblock_ffn1: x / out must be fp32 or fp16 (got dtype codes %d, %d)", x_dtype, out_dtype);
    SF_REQUIRE(((reinterpret_cast<uintptr_t>(w1p) | reinterpret_cast<uintptr_t>(w2p) | reinterpret_cast<uintptr_t>(b1p)) & 15) == 0,
               "pcblock_ffn1: packed weights must be 16-byte aligned");
    return launch_pcblock_ffn1(x, x_dtype, w1p, b1p, w2p, b2, out, out_dtype, P, C, hidden, N,
                               static_cast<cudaStream_t>(stream));
}

int sf_upsample_flow(const float* flow, const void* mask, int mask_dtype, float* out, int64_t N, int64_t H, int64_t W,
                     int ratio, void* stream) {
    DeviceInfo di;
    if (int rc = query_device(&di)) return rc;
    SF_REQUIRE(flow && mask && out, "upsample_flow: null pointer argument");
    SF_REQUIRE(ratio == 8, "upsample_flow: specialised for ratio 8 (got %d)", ratio);
    SF_REQUIRE(N >= 1 && H >= 1 && W >= 1 && N < 65536 && H < 65536, "upsample_flow: bad shape");
    SF_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "upsample_flow: out must be 16-byte aligned");
    return launch_upsample_flow(flow, mask, mask_dtype, out, N, H, W, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
