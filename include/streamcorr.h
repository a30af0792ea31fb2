/* streamcorr.h -- C ABI of libstreamcorr.so, the B200 (sm_100a) implementation of StreamFlow's
 * per-iteration correlation hot path.
 *
 * The reference (littlespray/StreamFlow) has no FFI of its own: the boundary of this path is its Python
 * operator API.  Each entry point below replaces the body of one reference operator; the Python mirror
 * in streamflow_b200/{corr,gma}.py binds them with ctypes:
 *
 *   sf_corr_build           <- CorrBlock.__init__ + CorrBlock.corr     core/corr.py:7-21, 46-54
 *   sf_corr_lookup[_group]  <- CorrBlock.__call__ + bilinear_sampler   core/corr.py:23-44, core/utils/utils.py:65-79
 *   sf_gma_attention        <- gma.Attention.forward                   core/gma.py:53-65
 *   sf_gma_aggregate        <- gma.Aggregate.forward                   core/gma.py:91-104
 *   sf_upsample_flow        <- SKFlow_MF8.upsample_flow                core/models/streamflow.py:82-93   (8(f) row 3)
 *   sf_pcblock_ffn1         <- first line of PCBlock4_Deep_nopool_res.forward    core/update.py:18-22, 31   (8(f) row 2)
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated; the library never
 *     allocates or frees caller-visible memory (workspaces are sized by the *_workspace_bytes queries);
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns without synchronising,
 *     so all calls are CUDA-graph capturable;
 *   - return value 0 = success, negative = error; sf_last_error() returns a thread-local message;
 *   - there is no CPU fallback: on a machine without an sm_100 device every compute entry point fails.
 */
#ifndef STREAMCORR_H_
#define STREAMCORR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SF_API __attribute__((visibility("default")))
#else
#define SF_API
#endif

#define SF_VERSION 210          /* major*100 + minor */
#define SF_NUM_LEVELS 4         /* corr_levels fixed by the model: core/models/streamflow.py:38 */
#define SF_RADIUS 4             /* corr_radius fixed by the model: core/models/streamflow.py:39 */
#define SF_MAX_GROUPS 8         /* CorrBlocks batched into one lookup launch */

/* error codes */
#define SF_OK 0
#define SF_ERR_INVALID (-1)     /* bad argument (shape, alignment, unsupported specialisation) */
#define SF_ERR_CUDA (-2)        /* CUDA runtime / driver error, message has the detail */
#define SF_ERR_NODEVICE (-3)    /* no sm_100 device: the library has no fallback path */
#define SF_ERR_WORKSPACE (-4)   /* workspace too small */

/* precision of the correlation GEMM (stated per north_star: fp32 accumulate in every mode) */
#define SF_PREC_F16 0           /* operands rounded to fp16 after a per-tensor power-of-two scale; EXACT
                                   products when the feature maps are fp16-representable (the model's
                                   mixed-precision path, core/models/streamflow.py:107); 2^-11 operand
                                   rounding otherwise (same mantissa as TF32)                                */
#define SF_PREC_F16X2 1         /* hi/lo split fp16 operands, 3 tensor-core products: ~2^-21, fp32-faithful  */
#define SF_PREC_FP32_SIMT 2     /* plain fp32 FFMA tiles + hierarchical pooling: the reference's arithmetic  */
#define SF_PREC_AUTO 3          /* decided ON THE DEVICE by the pass that already reads every element for the
                                   operand scale (no host sync): if every value of both feature maps is
                                   fp16-representable (11 significant bits -- the model's mixed-precision path)
                                   level 0 runs single-product (exact products) and the pooled levels, whose
                                   averaged operands are not fp16-exact, run hi*hi + hi*lo; otherwise the whole
                                   build runs as SF_PREC_F16X2.  Either way the result is fp32-faithful
                                   (~1e-6 norm-wise vs the reference's SGEMM).  Needs D % 64 == 0 for the fast
                                   path (else it behaves as SF_PREC_F16X2).                                   */

/* dtype codes for tensors whose element type may vary */
#define SF_DT_F32 0
#define SF_DT_F16 1
#define SF_DT_BF16 2

SF_API int sf_version(void);
SF_API const char* sf_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), SF_ERR_NODEVICE otherwise */
SF_API int sf_device_ok(void);

/* Measurement hooks (used by bench.py; no effect on results).
 * sf_launch_count: number of kernels this library has launched in the calling thread since load.
 * sf_profile_kernel: record CUDA events `start` / `stop` (cudaEvent_t as void*) on the call's stream
 * immediately before / after the NEXT launches of kernel `which`; pass NULL events to stop recording.  */
#define SF_KERNEL_LOOKUP 1
#define SF_KERNEL_CORR_GEMM 2
#define SF_KERNEL_GMA_AGGREGATE 3
#define SF_KERNEL_GMA_STATS 4
#define SF_KERNEL_CORR_PACK 5
#define SF_KERNEL_GMA_PROJ 6     /* q/k projection (attention) and the per-iteration fp16 operand cast (aggregate) */
#define SF_KERNEL_CORR_SIMT 8
#define SF_KERNEL_UPSAMPLE 9
#define SF_KERNEL_PCBLOCK_FFN1 10
SF_API int64_t sf_launch_count(void);
SF_API void sf_profile_kernel(int which, void* start, void* stop);
/* Measurement only: restrict the calling thread's sf_gma_aggregate to a subset of its kernels (bit 0 = fp16 operand cast,
 * bit 1 = streaming GEMM) and sf_corr_build to (bit 0 = absmax + pack, bit 1 = GEMM), so bench.py
 * can time one kernel back-to-back inside a CUDA graph.  Results are meaningless unless all bits are set (default). */
SF_API void sf_debug_select_kernels(int gma_aggregate_mask, int corr_build_mask);

/* ---- correlation pyramid -------------------------------------------------------------------------
 * Level l of the pyramid is a dense matrix [B*N, tiles_y * tiles_x * 16] fp32 (N = h*w): row b*N + y*w + x is the
 * h_l x w_l correlation image of query (b,y,x), stored as 4x4 tiles (64 bytes each, tile-row-major; cell (v,u) at
 * ((v/4)*tiles_x + u/4)*16 + (v%4)*4 + u%4; cells past h_l / w_l are zero) so that the lookup's 10x10 window
 * touches few 64-byte blocks.  h_l = h >> l, w_l = w >> l (avg_pool2d(2,2) floor mode, core/corr.py:19-21),
 * tiles = ceil(./4).                                                                                      */
SF_API void sf_corr_level_dims(int64_t h, int64_t w, int level, int64_t* h_l, int64_t* w_l, int64_t* tiles_y,
                        int64_t* tiles_x);
SF_API int64_t sf_corr_workspace_bytes(int64_t B, int64_t D, int64_t h, int64_t w, int precision);

/* fmap1/fmap2: [B, D, h, w] fp32 with arbitrary element strides {sB, sD, sh, sw} (the model passes
 * channels-last views).  levels[l]: output buffers as described above, 16-byte aligned.               */
SF_API int sf_corr_build(const float* fmap1, const float* fmap2, int64_t B, int64_t D, int64_t h, int64_t w,
                  const int64_t f1_strides[4], const int64_t f2_strides[4], float* const levels[SF_NUM_LEVELS],
                  void* workspace, int64_t workspace_bytes, int precision, void* stream);

/* coords [B, 2, h, w] fp32 contiguous (channel 0 = x, 1 = y) -> out [B, 4*81, h, w] fp32 contiguous;
 * channel l*81 + i*9 + j samples level l bilinearly (zeros outside, align_corners) at
 * (x, y) = (cx / 2^l + i - 4, cy / 2^l + j - 4).                                                       */
SF_API int sf_corr_lookup(const float* const levels[SF_NUM_LEVELS], const float* coords, float* out, int64_t B,
                   int64_t h, int64_t w, int radius, int num_levels, void* stream);

/* G CorrBlocks of identical shape in ONE launch (the T-1 frame pairs of a clip, streamflow.py:132).
 * levels: G*4 pointers (group-major); coords, out: G pointers.  out[g] may point into one
 * [G*B, 324, h, w] tensor so the stack + rearrange of the caller disappears.
 * out_dtype: SF_DT_F32 (reference layout) or SF_DT_F16 (for the autocast consumer).                    */
SF_API int sf_corr_lookup_group(int G, const float* const* levels, const float* const* coords, void* const* out,
                         int out_dtype, int64_t B, int64_t h, int64_t w, int radius, int num_levels, void* stream);

/* ---- GMA ------------------------------------------------------------------------------------------
 * heads = 1, dim = dim_head = d = 128 (the shipped model, core/models/streamflow.py:49).
 * Q and K are constant over the refinement iterations, so sf_gma_attention computes ONCE per clip
 *     E[p,i,j]   = fp16(2^12 * exp(s_ij - max_j s_ij)),   s = (scale * q) . k,   [q; k] = W_qk . fmap
 *     rowsum[p,i] = sum_j E[p,i,j]                         (softmax = E / rowsum)
 * with E stored tile-major as [P][ceil(N/128)][Npad/64][128][64] fp16 (Npad = sf_gma_npad(N) = round_up(N, 64);
 * every 128-query x 64-key tile is one contiguous 16 KB block holding the 128-byte-swizzled K-major image the
 * tensor-core descriptor reads -- 16-byte chunk c of query row r sits at chunk c ^ (r & 7) of its 128-byte row --
 * so any run of queries is one contiguous bulk copy; pad key columns are zero; sf_gma_e_elems(P, N) is the
 * element count to allocate), and
 * sf_gma_aggregate computes every iteration
 *     out = fmap + gamma * W_v . ((E / rowsum) . fmap^T)^T          (= fmap + gamma * attn . to_v(fmap), core/gma.py:94-102)
 * i.e. the 1x1 `to_v` convolution is applied AFTER the attention-weighted sum, inside the streaming kernel: the
 * motion features are cast to fp16 (one small launch) and streamed against E; the 128 x rows result is normalised,
 * split into fp16 hi + lo and multiplied by W_v (fp16) with a second tensor-core GEMM in the epilogue.
 * `workspace` (sf_gma_workspace_bytes, 1024-byte aligned) is scratch for the projections / fp16 operands; one buffer
 * serves the attention call and all aggregate calls that use its E on the same stream (it also carries one word of
 * state: "E is settled", which lets the second and later aggregate launches start streaming E before the kernel
 * launched in front of them has finished; with any other workspace the aggregate simply waits).  The aggregate is
 * deterministic (one fp32 accumulator per output element, no atomics). */
SF_API int64_t sf_gma_npad(int64_t N);
SF_API int64_t sf_gma_e_elems(int64_t P, int64_t N);
SF_API int64_t sf_gma_workspace_bytes(int64_t P, int64_t C, int64_t N, int64_t d);

/* fmap: [P, C, N] (NCHW flattened, contiguous) of dtype fmap_dtype; w_qk: [2*d, C] fp32 contiguous
 * (rows [0,d) -> q, rows [d,2d) -> k, as nn.Conv2d(dim, 2*inner, 1).weight chunked, core/gma.py:56).   */
/* precision: SF_PREC_F16 = q, k rounded to fp16 for the logit GEMM (the operand precision of the reference's own
 * autocast path, core/models/streamflow.py:118-124); SF_PREC_F16X2 = hi/lo-split projections and logits
 * (fp32-faithful, 3x the tensor-core work of this once-per-clip kernel).                                    */
SF_API int sf_gma_attention(const void* fmap, int fmap_dtype, const float* w_qk, int64_t P, int64_t C, int64_t N,
                     int64_t d, float scale, int precision, void* E, float* rowsum, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* The authors' memory-saving call convention (demo.py:235-282, test_memory.py:240-282): Attention.forward returns
 * the projected (q, k) and Aggregate.forward(q, k, fmap) redoes the attention every iteration with
 * flash_attn_func(q, k, v, softmax_scale = dim_head^-0.5).  q, k: [P, d, N] (NCHW flattened, contiguous) of dtype
 * qk_dtype, NOT pre-scaled.  Produces the same E / rowsum as sf_gma_attention would from the fmap they were
 * projected from; the caller caches them for as long as q and k are unchanged.                              */
SF_API int sf_gma_attention_qk(const void* q, const void* k, int qk_dtype, int64_t P, int64_t N, int64_t d, float scale,
                        int precision, void* E, float* rowsum, void* workspace, int64_t workspace_bytes,
                        void* stream);

/* fmap: [P, C, N] of dtype fmap_dtype; w_v: [d, C] of dtype w_dtype (SF_DT_F32 or SF_DT_F16; the tensor core
 * consumes fp16 weights either way, as the reference's autocast does); gamma: DEVICE pointer to
 * 1 float (no host sync); out: [P, C, N] fp32 (requires C == d: the reference's `project` is None,
 * core/gma.py:86-89).                                                                                    */
SF_API int sf_gma_aggregate(const void* E, const float* rowsum, const void* fmap, int fmap_dtype, const void* w_v,
                     int w_dtype, const float* gamma, float* out, int64_t P, int64_t C, int64_t N, int64_t d,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* ---- convex flow upsampling -------------------------------------------------------
 * Replaces SKFlow_MF8.upsample_flow (core/models/streamflow.py:82-93), ratio 8:
 *   out[n, c, 8y+i, 8x+j] = sum_k softmax_k(mask[n, k*64 + i*8 + j, y, x]) * 8 * flow[n, c, y + k/3 - 1, x + k%3 - 1]
 * flow: [N, 2, H, W] fp32 contiguous; mask: [N, 576, H, W] contiguous of dtype mask_dtype; out: [N, 2, 8H, 8W] fp32. */
SF_API int sf_upsample_flow(const float* flow, const void* mask, int mask_dtype, float* out, int64_t N, int64_t H,
                     int64_t W, int ratio, void* stream);

/* Motion-encoder entry, SURVEY 8(f) row 2.  Replaces the first line of PCBlock4_Deep_nopool_res.forward
 * (core/update.py:31, modules built at core/update.py:18-22): `x = F.gelu(x + self.ffn1(x))` with
 * ffn1 = Conv2d(C, 1.5 C, 1) -> GELU -> Conv2d(1.5 C, C, 1), i.e. per pixel
 *   out[p, :, n] = gelu(x[p, :, n] + W2 . gelu(W1 . x[p, :, n] + b1) + b2)         (exact-erf GELU)
 * x, out: [P, C, N] contiguous (N = h * w), dtype fp32 or fp16 (SF_DT_*); fp16 operands, fp32 accumulate.
 * w1p: fp16 [ceil128(hidden), ceil64(C)] = W1 zero-padded;   b1p: fp32 [ceil128(hidden)] zero-padded;
 * w2p: fp16 [ceil16(C), ceil128(hidden)] = W2 zero-padded;   b2:  fp32 [C].
 * Specialised for C <= 384 and hidden <= 512 (convc1: 324 / 486, convc2 and conv: 256 / 384, convf2: 128 / 192). */
SF_API int sf_pcblock_ffn1(const void* x, int x_dtype, const void* w1p, const float* b1p, const void* w2p, const float* b2,
                    void* out, int out_dtype, int64_t P, int64_t C, int64_t hidden, int64_t N, void* stream);

/* Debug only: device buffer of 32 x uint64 receiving %globaltimer stamps of one CTA of sf_pcblock_ffn1 (NULL = off). */
SF_API void sf_debug_ffn1_trace(void* dev_ptr);

#ifdef __cplusplus
}
#endif
#endif /* STREAMCORR_H_ */
