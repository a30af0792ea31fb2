// G2: radius-4 bilinear pyramid lookup (replaces CorrBlock.__call__, core/corr.py:23-44 and
// bilinear_sampler / grid_sample, core/utils/utils.py:65-79).
//
// HBM-bound gather.  Algorithmic bytes per query: 4 levels x 10x10 fp32 window (1600 B) + 8 B coords
// + 4 x 81 fp32 outputs (1296 B) = 2904 B.
//
// Each query's correlation image is stored as 4x4 tiles of 64 B (see sf_internal.h): a 10x10 window at an
// arbitrary offset touches on average 3.25 x 3.25 tiles = 676 B of 64-byte DRAM fetches, against 1024 B for a
// row-major image (10 rows x 1.6 blocks; measured 87.7 MB of DRAM reads for 33.8 MB of window bytes).
//
// Work item = 32 consecutive queries x ONE pyramid level; persistent CTAs (4 per SM) pipeline three stages:
//   A  warp 0, lane = query: coordinates of item k+2 into registers;
//   B  warp 0: coords -> integer window origin (x0, y0) and the single fractional pair (ax, ay) shared by all 81
//      taps of the level (window offsets are integers); then all 128 threads, thread = (query, tile column): the
//      up-to 4x4 tiles under the window of item k+1 as 16-byte cp.async.cg chunks (zero-filled outside the image);
//   C  item k, lane = query: conflict-free LDS.128 of its own rows, horizontal then vertical lerp in registers, and
//      one 128-byte coalesced store per output channel straight into the NCHW result
//      (channel = l*81 + i*9 + j, i moves x, j moves y).  Warps split the 9 y-offsets.
#include <algorithm>

#include "sf_internal.h"

namespace sf {

namespace {

constexpr int kQ = 32;              // queries per CTA
constexpr int kRowFloats = 16;      // 4 tile columns of 4 floats per staged row
constexpr int kRows = 2 * SF_RADIUS + 2;                 // 10 window rows / columns
constexpr int kStageRows = kRows + 3;                    // window may start at row 0..3 of its first tile
constexpr int kWinStride = kStageRows * kRowFloats + 4;  // 212 floats: 8 consecutive queries -> distinct bank quads
constexpr int kSide = 2 * SF_RADIUS + 1;                 // 9

// 16-byte async copy, zero-filled when !pred.  The L2 evict_last policy keeps the window tiles resident for the
// next refinement iteration (flow moves by ~1 px, so ~80 % of the tiles are touched again) while the 300 MB
// softmax stream of the aggregation passes through L2 as evict_first.
__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, bool pred, unsigned long long policy) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst));
    const int bytes = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(d), "l"(src), "r"(bytes),
                 "l"(policy)
                 : "memory");
}

// Per-item metadata, written once by warp 0 (lane = query) so that neither the 128 loader threads nor the 128
// interpolation threads repeat divisions or 64-bit address arithmetic.
struct Meta {
    float ax[kQ], ay[kQ];
    int x0[kQ], y0[kQ];
    const float* img[kQ];        // this query's correlation image at the item's level (nullptr: query out of range)
    unsigned out_off[kQ];        // element offset of out[b, lvl*81, n] inside the group's output tensor
    int grp, lvl;
};

// Persistent, software-pipelined: a CTA walks work items (32 queries x 1 level) with stride gridDim.x and keeps
// three items in flight -- coordinates of item k+2 (registers), window tiles of item k+1 (cp.async into the other
// shared-memory buffer) and the interpolation + stores of item k -- so the DRAM latency of the gather is hidden
// behind the previous item's compute instead of being paid once per CTA wave.
template <bool kHalfOut>
__global__ void __launch_bounds__(128) corr_lookup_kernel(const __grid_constant__ LookupParams p) {
    extern __shared__ __align__(16) float smem_f[];
    float* win0 = smem_f;                                   // [2][kQ * kWinStride]
    Meta* meta = reinterpret_cast<Meta*>(smem_f + 2 * kQ * kWinStride);   // [2]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int items = static_cast<int>(p.items), tiles = static_cast<int>(p.tiles), BN = static_cast<int>(p.BN);

    // item -> (group, query tile, level); levels of one tile are adjacent items.  Only warp 0 decodes (32-bit math).
    auto decode = [&](int it, int& grp, int& q0, int& lvl) {
        lvl = it & 3;
        const unsigned t = static_cast<unsigned>(it) >> 2;
        grp = static_cast<int>(t / static_cast<unsigned>(tiles));
        q0 = static_cast<int>(t - static_cast<unsigned>(grp) * tiles) * kQ;
    };
    // stage A: warp 0 fetches the coordinates of an item into registers (and remembers where its outputs go)
    struct Pending {
        float cx, cy;
        int grp, lvl, qid, b, n;
    };
    auto load_coords = [&](int it, Pending& pd) {
        pd.cx = pd.cy = -1e30f;
        pd.qid = -1;
        pd.grp = pd.lvl = pd.b = pd.n = 0;
        if (it < items) {
            int q0;
            decode(it, pd.grp, q0, pd.lvl);
            const int qid = q0 + lane;
            if (qid < BN) {
                pd.qid = qid;
                pd.b = (p.N >= BN) ? 0 : static_cast<int>(static_cast<unsigned>(qid) / static_cast<unsigned>(p.N));
                pd.n = qid - pd.b * p.N;
                const float* c = p.coords[pd.grp] + static_cast<long long>(pd.b) * 2 * p.N + pd.n;
                pd.cx = __ldg(c);
                pd.cy = __ldg(c + p.N);
            }
        }
    };
    // stage B1: warp 0 turns coordinates into the window origin + fractions of the item's level
    auto write_meta = [&](int it, int buf, const Pending& pd) {
        if (it >= items) return;
        const int lvl = pd.lvl;
        const float inv = 1.0f / static_cast<float>(1 << lvl);
        // Far outside the image every tap is zero; clamping keeps the int conversion defined
        // (NaN / missing coordinates clamp to the lower bound and yield zeros).
        const float X0 = fminf(fmaxf(pd.cx * inv - static_cast<float>(SF_RADIUS), -16.f), static_cast<float>(p.wl[lvl] + 8));
        const float Y0 = fminf(fmaxf(pd.cy * inv - static_cast<float>(SF_RADIUS), -16.f), static_cast<float>(p.hl[lvl] + 8));
        const float xf = floorf(X0), yf = floorf(Y0);
        Meta& m = meta[buf];
        m.ax[lane] = X0 - xf;
        m.ay[lane] = Y0 - yf;
        m.x0[lane] = static_cast<int>(xf);
        m.y0[lane] = static_cast<int>(yf);
        m.img[lane] = pd.qid >= 0 ? p.lvl[pd.grp][lvl] + static_cast<long long>(pd.qid) * p.img[lvl] : nullptr;
        m.out_off[lane] = static_cast<unsigned>((pd.b * (SF_NUM_LEVELS * kSide * kSide) + lvl * (kSide * kSide)) * p.N + pd.n);
        if (lane == 0) {
            m.grp = pd.grp;
            m.lvl = lvl;
        }
    };
    unsigned long long policy;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    // stage B2: all threads, thread = (query, row inside a tile): issue the window as 16-byte chunks
    auto issue_window = [&](int it, int buf) {
        if (it < items) {
            const Meta& m = meta[buf];
            const int lvl = m.lvl;
            const int th = p.th[lvl], tw = p.tw[lvl];
            const int q = tid >> 2, rr = tid & 3;           // lane quad = the four 16-byte rows of ONE 64-byte tile:
            const float* base = m.img[q];                   // every request carries two full sectors, fetched once
            const int x0 = m.x0[q], y0 = m.y0[q];
            const int ox = x0 & 3, oy = y0 & 3;             // window origin inside its first tile
            const int tx0 = x0 >> 2, ty0 = y0 >> 2;
            if (base != nullptr) {
                const unsigned rowmask = ((1u << kRows) - 1u) << oy;      // staged rows covered by the window
                const int ncol = (ox + kRows + 3) >> 2;                   // tile columns overlapping window columns ox .. ox+9
                float* dst = win0 + buf * (kQ * kWinStride) + q * kWinStride;
#pragma unroll
                for (int tr = 0; tr < 4; ++tr) {            // tile row
                    const int R = tr * 4 + rr;
                    if (R >= kStageRows || !(rowmask & (1u << R))) continue;      // row outside the window
                    const int ty = ty0 + tr;
                    const bool rowok = (ty >= 0) && (ty < th);
                    const float* row = base + (rowok ? (ty * tw + tx0) * 16 + rr * 4 : 0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {           // tile column: constant 64-byte steps from here on
                        if (j >= ncol) continue;
                        const bool ok = rowok && (tx0 + j >= 0) && (tx0 + j < tw);
                        cp_async16_zfill(dst + R * kRowFloats + j * 4, ok ? row + j * 16 : base, ok, policy);
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    const int first = blockIdx.x, step = gridDim.x;
    Pending pd;
    pdl_launch();
    pdl_wait();
    // prologue: item 0 fully staged, coordinates of item 1 in flight
    if (warp == 0) {
        load_coords(first, pd);
        write_meta(first, 0, pd);
        load_coords(first + step, pd);
    }
    __syncthreads();
    issue_window(first, 0);

    int buf = 0;
    for (int it = first; it < items; it += step, buf ^= 1) {
        if (warp == 0) {
            write_meta(it + step, buf ^ 1, pd);
            load_coords(it + 2 * step, pd);
        }
        __syncthreads();                                    // meta[buf^1] visible; win[buf^1] free (read 2 items ago)
        issue_window(it + step, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // this thread's chunks of item `it` have landed
        __syncthreads();

        // stage C: lane = query; warp w owns y-offsets j in [jb, je]
        const Meta& m = meta[buf];
        if (m.img[lane] != nullptr) {
            const int jb = (warp == 0) ? 0 : (2 * warp + 1);       // 0,3,5,7
            const int je = (warp == 0) ? 2 : (2 * warp + 2);       // 2,4,6,8
            const float ax = m.ax[lane], ay = m.ay[lane];
            const int o = m.x0[lane] & 3;
            const float4* wq = reinterpret_cast<const float4*>(win0 + buf * (kQ * kWinStride) + lane * kWinStride) +
                               (m.y0[lane] & 3) * 4;
            // uniform 64-bit base (kernel parameter) + 32-bit per-lane element offset: one IMAD.WIDE per store
            float* const outf = reinterpret_cast<float*>(p.out[m.grp]);
            __half* const outh = reinterpret_cast<__half*>(p.out[m.grp]);
            const unsigned ooff = m.out_off[lane];
            const unsigned nine_n = static_cast<unsigned>(kSide * p.N), un = static_cast<unsigned>(p.N);
            float hprev[kSide];
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                if (r < jb || r > je + 1) continue;
                float t[16];
                {
                    const float4 c0 = wq[r * 4 + 0], c1 = wq[r * 4 + 1], c2 = wq[r * 4 + 2], c3 = wq[r * 4 + 3];
                    t[0] = c0.x; t[1] = c0.y; t[2] = c0.z; t[3] = c0.w;
                    t[4] = c1.x; t[5] = c1.y; t[6] = c1.z; t[7] = c1.w;
                    t[8] = c2.x; t[9] = c2.y; t[10] = c2.z; t[11] = c2.w;
                    t[12] = c3.x; t[13] = c3.y; t[14] = c3.z; t[15] = c3.w;
                }
                float s1[12], u[10], hcur[kSide];
#pragma unroll
                for (int i = 0; i < 12; ++i) s1[i] = (o & 1) ? t[i + 1] : t[i];
#pragma unroll
                for (int i = 0; i < 10; ++i) u[i] = (o & 2) ? s1[i + 2] : s1[i];
#pragma unroll
                for (int i = 0; i < kSide; ++i) hcur[i] = fmaf(ax, u[i + 1] - u[i], u[i]);
                if (r > jb) {
                    unsigned off = ooff + static_cast<unsigned>(r - 1) * un;      // channel i*9 + j, j = r - 1
#pragma unroll
                    for (int i = 0; i < kSide; ++i) {
                        const float v = fmaf(ay, hcur[i] - hprev[i], hprev[i]);
                        if (kHalfOut) {
                            outh[off] = __float2half_rn(v);
                        } else {
                            __stcs(outf + off, v);
                        }
                        off += nine_n;
                    }
                }
#pragma unroll
                for (int i = 0; i < kSide; ++i) hprev[i] = hcur[i];
            }
        }
        __syncthreads();                                    // meta[buf] / win[buf] are rewritten two stages later
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

constexpr int kLookupSmem = 2 * kQ * kWinStride * 4 + 2 * static_cast<int>(sizeof(Meta));

}  // namespace

int launch_corr_lookup(const LookupParams& p_in, int groups, int num_sms, cudaStream_t s) {
    LookupParams p = p_in;
    p.tiles = (p.BN + kQ - 1) / kQ;
    SF_REQUIRE(p.tiles > 0 && p.BN * (SF_NUM_LEVELS * kSide * kSide) < (1ll << 31) &&
                   p.tiles * groups * SF_NUM_LEVELS < (1ll << 31),
               "corr_lookup: %lld queries per group exceed the 32-bit index range of the kernel", p.BN);
    p.items = p.tiles * groups * SF_NUM_LEVELS;
    // 4 CTAs of 55 KB per SM, persistent over the work items
    const int grid = static_cast<int>(std::min<long long>(p.items, 4ll * num_sms));
    auto launch = [&](auto kernel) -> int {
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), kLookupSmem)) return rc;
        prof_before(SF_KERNEL_LOOKUP, s);
        SF_CUDA_CHECK(launch_kernel(kernel, dim3(grid), dim3(128), static_cast<size_t>(kLookupSmem), s, p));
        prof_after(SF_KERNEL_LOOKUP, s);
        SF_CUDA_CHECK(cudaGetLastError());
        return SF_OK;
    };
    return p.out_f16 ? launch(corr_lookup_kernel<true>) : launch(corr_lookup_kernel<false>);
}

}  // namespace sf
