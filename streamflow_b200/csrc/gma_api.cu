// C ABI for the GMA path: sf_gma_attention (once per clip) and sf_gma_aggregate (every refinement iteration).
#include "sf_internal.h"

namespace sf {
namespace {

struct GmaWs {
    int64_t q_off, k_off, rowmax_off, rowsum_fx_off, x16_off, w16_off, eye_off, flag_off, total;
    int Kp;
    int64_t Npad;
};

GmaWs gma_ws_layout(int64_t P, int64_t N, int64_t d) {
    GmaWs ws{};
    ws.Kp = static_cast<int>(2 * d);            // room for [hi | lo]-split q, k (SF_PREC_F16X2); SF_PREC_F16 uses d
    ws.Npad = align_up(N, 64);
    int64_t off = 0;
    ws.q_off = off;       off += align_up(P * N * ws.Kp * 2, 1024);
    ws.k_off = off;       off += align_up(P * N * ws.Kp * 2, 1024);
    ws.rowmax_off = off;  off += align_up(P * N * 4, 1024);
    ws.rowsum_fx_off = off;  off += align_up(P * N * 8, 1024);
    ws.x16_off = off;     off += align_up(P * d * ws.Npad * 2, 1024);
    ws.w16_off = off;     off += align_up(d * d * 2, 1024);
    ws.eye_off = off;     off += align_up(d * d * 4, 1024);
    ws.flag_off = off;    off += 1024;          // "E is settled" word of the aggregate's early E stream
    ws.total = off;
    return ws;
}

int check_gma(int64_t P, int64_t C, int64_t N, int64_t d, const void* ws, int64_t ws_bytes, const GmaWs& lay) {
    SF_REQUIRE(P >= 1 && N >= 1, "gma: non-positive shape P=%lld N=%lld", (long long)P, (long long)N);
    SF_REQUIRE(d == 128 && C == 128,
               "gma: kernels are specialised for heads=1, dim=dim_head=128 (the shipped model); got dim=%lld "
               "dim_head=%lld -- no generic fallback",
               (long long)C, (long long)d);
    SF_REQUIRE(P * N < (1ll << 31), "gma: shape too large");
    if (!ws || ws_bytes < lay.total) {
        set_error("gma: workspace of %lld bytes needed, %lld given", (long long)lay.total, (long long)ws_bytes);
        return SF_ERR_WORKSPACE;
    }
    SF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "gma: workspace must be 1024-byte aligned");
    return SF_OK;
}

}  // namespace
}  // namespace sf

namespace sf {
namespace {

// E, rowsum from  q = scale * W_q . xq,  k = W_k . xk  (xq == xk == fmap for the reference's Attention)
int attention_common(const void* xq, const void* xk, int x_dtype, const float* w_q, const float* w_k, int64_t P,
                     int64_t C, int64_t N, int64_t d, float scale, int precision, void* E, float* rowsum,
                     uint8_t* wsb, const GmaWs& ws, const DeviceInfo& di, cudaStream_t s) {
    const int split = (precision == SF_PREC_F16X2);
    const int Kp = split ? ws.Kp : static_cast<int>(d);
    const int64_t Npad = ws.Npad;
    // E is about to be rewritten: the next aggregate launch on this workspace must wait for its predecessors
    SF_CUDA_CHECK(cudaMemsetAsync(wsb + ws.flag_off, 0, 1024, s));

    GmaProjParams pq{};
    pq.x = xq; pq.x_dtype = x_dtype; pq.w = w_q;
    pq.P = (int)P; pq.C = (int)C; pq.N = (int)N; pq.O = 128;
    pq.scale = scale;
    pq.out = reinterpret_cast<__half*>(wsb + ws.q_off);
    pq.out_batch_stride = N * Kp; pq.ld = Kp; pq.token_major = 1; pq.split = split; pq.is_b = 0;
    pq.zero_u32 = reinterpret_cast<unsigned*>(wsb + ws.rowmax_off);          // row max / row sum accumulators of the
    pq.zero_u64 = reinterpret_cast<unsigned long long*>(wsb + ws.rowsum_fx_off);   // stats passes start from zero
    pq.x2 = xk; pq.w2 = w_k; pq.scale2 = 1.0f;                       // k rides in the same launch (blockIdx.z = 1)
    pq.out2 = reinterpret_cast<__half*>(wsb + ws.k_off);
    SF_REQUIRE((reinterpret_cast<uintptr_t>(w_k) & 15) == 0, "gma_attention: weight pointer must be 16-byte aligned");
    if (int rc = launch_gma_proj(pq, s)) return rc;


    CUtensorMap tm_q, tm_k;
    const uint64_t kp = static_cast<uint64_t>(Kp);
    if (int rc = make_tmap3(&tm_q, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, wsb + ws.q_off, kp, N, P, kp * 2, N * kp * 2, 64,
                            128, "Q"))
        return rc;
    if (int rc = make_tmap3(&tm_k, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, wsb + ws.k_off, kp, N, P, kp * 2, N * kp * 2, 64,
                            128, "K"))      // each CTA of a pair loads half of a 256-key tile
        return rc;
    GmaStatsParams sp{};
    sp.P = (int)P; sp.N = (int)N; sp.Npad = (int)Npad; sp.Kp = Kp; sp.split = split;
    sp.m_tiles = (int)((N + 127) / 128);
    sp.n_tiles = (int)((N + 255) / 256);
    sp.pair_tiles = (sp.m_tiles + 1) / 2;
    const int64_t base_units = P * sp.pair_tiles;
    int chunks = (int)((4ll * (di.sms / 2) + base_units - 1) / base_units);
    sp.chunks = std::max(1, std::min(chunks, sp.n_tiles));
    sp.rowmax_bits = reinterpret_cast<unsigned*>(wsb + ws.rowmax_off);
    sp.rowsum_fx = reinterpret_cast<unsigned long long*>(wsb + ws.rowsum_fx_off);
    sp.E = static_cast<__half*>(E);
    sp.pass = 1;
    if (int rc = launch_gma_stats(sp, tm_q, tm_k, di.sms, s)) return rc;
    sp.pass = 2;
    if (int rc = launch_gma_stats(sp, tm_q, tm_k, di.sms, s)) return rc;
    return launch_gma_rowsum_finish(sp.rowsum_fx, rowsum, P * N, s);
}

}  // namespace
}  // namespace sf

using namespace sf;

extern "C" {

int64_t sf_gma_npad(int64_t N) { return align_up(N, 64); }

int64_t sf_gma_e_elems(int64_t P, int64_t N) { return P * align_up(N, 128) * align_up(N, 64); }

int64_t sf_gma_workspace_bytes(int64_t P, int64_t C, int64_t N, int64_t d) {
    (void)C;
    if (P < 1 || N < 1 || d < 1) return 0;
    return gma_ws_layout(P, N, d).total;
}

int sf_gma_attention(const void* fmap, int fmap_dtype, const float* w_qk, int64_t P, int64_t C, int64_t N, int64_t d,
                     float scale, int precision, void* E, float* rowsum, void* workspace, int64_t workspace_bytes,
                     void* stream) {
    DeviceInfo di;
    if (int rc = query_device(&di)) return rc;
    const GmaWs ws = gma_ws_layout(P, N, d);
    if (int rc = check_gma(P, C, N, d, workspace, workspace_bytes, ws)) return rc;
    SF_REQUIRE(fmap && w_qk && E && rowsum, "gma_attention: null pointer argument");
    SF_REQUIRE(precision == SF_PREC_F16 || precision == SF_PREC_F16X2, "gma_attention: unknown precision mode %d",
               precision);
    SF_REQUIRE((reinterpret_cast<uintptr_t>(E) & 15) == 0, "gma_attention: E must be 16-byte aligned");
    return attention_common(fmap, fmap, fmap_dtype, w_qk, w_qk + d * C, P, C, N, d, scale, precision, E, rowsum,
                            static_cast<uint8_t*>(workspace), ws, di, static_cast<cudaStream_t>(stream));
}

int sf_gma_attention_qk(const void* q, const void* k, int qk_dtype, int64_t P, int64_t N, int64_t d, float scale,
                        int precision, void* E, float* rowsum, void* workspace, int64_t workspace_bytes,
                        void* stream) {
    DeviceInfo di;
    if (int rc = query_device(&di)) return rc;
    const GmaWs ws = gma_ws_layout(P, N, d);
    if (int rc = check_gma(P, d, N, d, workspace, workspace_bytes, ws)) return rc;
    SF_REQUIRE(q && k && E && rowsum, "gma_attention_qk: null pointer argument");
    SF_REQUIRE(precision == SF_PREC_F16 || precision == SF_PREC_F16X2, "gma_attention_qk: unknown precision mode %d",
               precision);
    SF_REQUIRE((reinterpret_cast<uintptr_t>(E) & 15) == 0, "gma_attention_qk: E must be 16-byte aligned");
    // q and k arrive already projected: run them through the same packing kernel with identity weights
    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    float* eye = reinterpret_cast<float*>(wsb + ws.eye_off);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = launch_gma_identity(eye, static_cast<int>(d), s)) return rc;
    return attention_common(q, k, qk_dtype, eye, eye, P, d, N, d, scale, precision, E, rowsum, wsb, ws, di, s);
}

int sf_gma_aggregate(const void* E, const float* rowsum, const void* fmap, int fmap_dtype, const void* w_v,
                     int w_dtype, const float* gamma, float* out, int64_t P, int64_t C, int64_t N, int64_t d, void* workspace,
                     int64_t workspace_bytes, void* stream) {
    DeviceInfo di;
    if (int rc = query_device(&di)) return rc;
    const GmaWs ws = gma_ws_layout(P, N, d);
    if (int rc = check_gma(P, C, N, d, workspace, workspace_bytes, ws)) return rc;
    SF_REQUIRE(E && rowsum && fmap && w_v && gamma && out, "gma_aggregate: null pointer argument");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t* wsb = static_cast<uint8_t*>(workspace);
    const int64_t Npad = ws.Npad;

    SF_REQUIRE(w_dtype == SF_DT_F32 || w_dtype == SF_DT_F16, "gma_aggregate: w_v must be fp32 or fp16");
    __half* x16 = reinterpret_cast<__half*>(wsb + ws.x16_off);
    __half* w16 = reinterpret_cast<__half*>(wsb + ws.w16_off);
    const int parts = debug_gma_mask();
    if (parts & 1)
        if (int rc = launch_gma_cast(fmap, fmap_dtype, x16, P * C, N, Npad, w_v, w_dtype, w16, d * C, s)) return rc;

    CUtensorMap tm_x, tm_w;
    const int64_t e_rows = (N + 127) / 128 * (Npad / 64) * 128;       // 128-byte rows of tile-major E per map
    if (int rc = make_tmap3(&tm_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, x16, Npad, C, P, Npad * 2, C * Npad * 2, 64, 128,
                            "X16"))
        return rc;
    if (int rc = make_tmap3(&tm_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, w16, C, d, 1, C * 2, d * C * 2, 64, 128, "Wv"))
        return rc;

    GmaAggParams ap{};
    ap.P = (int)P; ap.N = (int)N; ap.Npad = (int)Npad; ap.C = (int)C;
    ap.k_blocks = (int)(Npad / 64);
    ap.rowsum = rowsum;
    ap.gamma = gamma;
    ap.fmap = fmap; ap.fmap_dtype = fmap_dtype;
    ap.out = out;
    ap.e_ptr = static_cast<const __half*>(E);
    ap.e_map_stride = e_rows * 64;
    return (parts & 2) ? launch_gma_aggregate(ap, tm_x, tm_w, di.sms, reinterpret_cast<unsigned*>(wsb + ws.flag_off), s)
                       : SF_OK;
}

}  // extern "C"
