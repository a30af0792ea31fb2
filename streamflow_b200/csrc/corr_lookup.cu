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
// Work item = 32 consecutive queries x ONE pyramid level; persistent CTAs (5 per SM) run a software pipeline with TWO
// block-wide barriers per item and no single-warp phase (an ncu capture of the previous three-barrier version, where
// warp 0 alone prepared the item's metadata, showed `barrier` as the top stall reason at 41 % issue utilisation):
//   A  every thread, thread = (query, row inside a tile): coordinates of item k+2 into registers;
//   B  the same thread turns the coordinates of item k+1 into the integer window origin (x0, y0) and the single
//      fractional pair (ax, ay) shared by all 81 taps of the level (window offsets are integers) -- redundantly in the 4
//      lanes of a quad, one of which leaves them in shared memory for stage C -- and issues the up-to 4x4 tiles under
//      the window as 16-byte cp.async chunks (zero-filled outside the image); the lane quad fetches the four rows of
//      ONE 64-byte tile, and only the 10 window rows are staged (row 0 = first window row);
//   C  item k, lane = query: conflict-free LDS.128 of its own rows, horizontal then vertical lerp in registers, and
//      one 128-byte coalesced store per output channel straight into the NCHW result
//      (channel = l*81 + i*9 + j, i moves x, j moves y).  Warps split the 9 y-offsets.
#include <algorithm>
#include <cstdlib>

#include "sf_internal.h"

namespace sf {

namespace {

constexpr int kQ = 32;              // queries per CTA
constexpr int kRowFloats = 16;      // 4 tile columns of 4 floats per staged row
constexpr int kRows = 2 * SF_RADIUS + 2;                 // 10 window rows / columns
constexpr int kWinStride = kRows * kRowFloats + 4;       // 164 floats: 8 consecutive queries -> distinct bank quads
constexpr int kSide = 2 * SF_RADIUS + 1;                 // 9

// 16-byte async copy, zero-filled when !pred.  The L2 evict_last policy keeps the window tiles resident for the
// next refinement iteration (flow moves by ~1 px; a warm-cache ncu capture shows 27 % fewer DRAM bytes than cold)
// while the 300 MB softmax stream of the aggregation passes through L2 as evict_first.
template <bool kViaL1>
__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, bool pred, unsigned long long policy) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(dst));
    const int bytes = pred ? 16 : 0;
    if constexpr (kViaL1)       // through L1: the four 16-byte requests of a lane quad merge into one 64-byte L2 request
        asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(d), "l"(src), "r"(bytes),
                     "l"(policy)
                     : "memory");
    else
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(d), "l"(src), "r"(bytes),
                     "l"(policy)
                     : "memory");
}

// Per-item metadata for stage C, written by one lane of each loader quad.
struct Meta {
    float ax[kQ], ay[kQ];
    int o[kQ];                   // x0 & 3: window origin inside its first tile column; -1: query out of range
    unsigned out_off[kQ];        // element offset of out[b, lvl*81, n] inside the group's output tensor
    int grp, lvl;
};

template <bool kHalfOut, bool kViaL1>
__global__ void __launch_bounds__(128) corr_lookup_kernel(const __grid_constant__ LookupParams p) {
    extern __shared__ __align__(16) float smem_f[];
    float* win0 = smem_f;                                   // [2][kQ * kWinStride]
    Meta* meta = reinterpret_cast<Meta*>(smem_f + 2 * kQ * kWinStride);   // [2]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int lq = tid >> 2, rr = tid & 3;                  // loader role: query of the item, row inside a tile
    const int items = static_cast<int>(p.items), tiles = static_cast<int>(p.tiles), BN = static_cast<int>(p.BN);

    // item -> (group, query tile, level); levels of one tile are adjacent items (block-uniform, 32-bit math)
    auto decode = [&](int it, int& grp, int& q0, int& lvl) {
        lvl = it & 3;
        const unsigned t = static_cast<unsigned>(it) >> 2;
        grp = static_cast<int>(t / static_cast<unsigned>(tiles));
        q0 = static_cast<int>(t - static_cast<unsigned>(grp) * tiles) * kQ;
    };
    // stage A: coordinates of this thread's query of an item into registers
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
            const int qid = q0 + lq;
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
    unsigned long long policy;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
    // stage B: window origin + fractions of the item's level, then the window as 16-byte chunks
    auto stage_item = [&](int it, int buf, const Pending& pd) {
        if (it < items) {
            const int lvl = pd.lvl;
            const float inv = 1.0f / static_cast<float>(1 << lvl);
            // Far outside the image every tap is zero; clamping keeps the int conversion defined
            // (NaN / missing coordinates clamp to the lower bound and yield zeros).
            const float X0 = fminf(fmaxf(pd.cx * inv - static_cast<float>(SF_RADIUS), -16.f), static_cast<float>(p.wl[lvl] + 8));
            const float Y0 = fminf(fmaxf(pd.cy * inv - static_cast<float>(SF_RADIUS), -16.f), static_cast<float>(p.hl[lvl] + 8));
            const float xf = floorf(X0), yf = floorf(Y0);
            const int x0 = static_cast<int>(xf), y0 = static_cast<int>(yf);
            const int ox = x0 & 3, oy = y0 & 3;             // window origin inside its first tile
            if (rr == 0) {
                Meta& m = meta[buf];
                m.ax[lq] = X0 - xf;
                m.ay[lq] = Y0 - yf;
                m.o[lq] = pd.qid >= 0 ? ox : -1;
                m.out_off[lq] = static_cast<unsigned>((pd.b * (SF_NUM_LEVELS * kSide * kSide) + lvl * (kSide * kSide)) * p.N + pd.n);
                if (tid == 0) {
                    m.grp = pd.grp;
                    m.lvl = lvl;
                }
            }
            if (pd.qid >= 0) {
                const int th = p.th[lvl], tw = p.tw[lvl];
                const float* base = p.lvl[pd.grp][lvl] + static_cast<long long>(pd.qid) * p.img[lvl];
                const int tx0 = x0 >> 2, ty0 = y0 >> 2;
                const int ncol = (ox + kRows + 3) >> 2;     // tile columns overlapping window columns ox .. ox+9
                float* dst = win0 + buf * (kQ * kWinStride) + lq * kWinStride;
#pragma unroll
                for (int tr = 0; tr < 4; ++tr) {            // tile row; this lane owns row rr of every tile
                    const int wr = tr * 4 + rr - oy;        // window row staged by this (tile row, lane)
                    if (wr < 0 || wr >= kRows) continue;
                    const int ty = ty0 + tr;
                    const bool rowok = (ty >= 0) && (ty < th);
                    const float* row = base + (rowok ? (ty * tw + tx0) * 16 + rr * 4 : 0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {           // tile column: constant 64-byte steps from here on
                        if (j >= ncol) continue;
                        const bool ok = rowok && (tx0 + j >= 0) && (tx0 + j < tw);
                        cp_async16_zfill<kViaL1>(dst + wr * kRowFloats + j * 4, ok ? row + j * 16 : base, ok, policy);
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
    // prologue: item 0 fully issued, coordinates of item 1 in flight
    load_coords(first, pd);
    stage_item(first, 0, pd);
    load_coords(first + step, pd);

    int buf = 0;
    for (int it = first; it < items; it += step, buf ^= 1) {
        stage_item(it + step, buf ^ 1, pd);                 // win / meta[buf^1] were last read before the previous barrier
        load_coords(it + 2 * step, pd);
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // this thread's chunks of item `it` have landed
        __syncthreads();                                    // ... and everybody else's, and meta[buf]

        // stage C: lane = query; warp w owns y-offsets j in [jb, je]
        // (8-byte loads + one select per value instead of 16-byte loads + two selects were measured: same time --
        // 17.7 vs 17.4 us -- the kernel is bound by the shared-memory / LSU pipe and the gather, not by issue slots)
        const Meta& m = meta[buf];
        const int o = m.o[lane];
        if (o >= 0) {
            const int jb = (warp == 0) ? 0 : (2 * warp + 1);       // 0,3,5,7
            const int je = (warp == 0) ? 2 : (2 * warp + 2);       // 2,4,6,8
            const float ax = m.ax[lane], ay = m.ay[lane];
            const float4* wq = reinterpret_cast<const float4*>(win0 + buf * (kQ * kWinStride) + lane * kWinStride);
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
        __syncthreads();                                    // win[buf] / meta[buf] are rewritten by the next iteration
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
    // 5 CTAs of 43 KB per SM, persistent over the work items
    const int grid = static_cast<int>(std::min<long long>(p.items, 5ll * num_sms));
    // cp.async.cg (L2 only) by default: measured 17.4 us vs 21.0 us through L1 for the 3-pair launch (the L1 path
    // wins only when the whole footprint is L2-resident, e.g. a single pair: 7.3 vs 8.1 us); STREAMCORR_LOOKUP_CA=1
    static const bool via_l1 = [] {
        const char* e = getenv("STREAMCORR_LOOKUP_CA");
        return e && e[0] == '1';
    }();
    auto launch = [&](auto kernel) -> int {
        if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), kLookupSmem)) return rc;
        prof_before(SF_KERNEL_LOOKUP, s);
        SF_CUDA_CHECK(launch_kernel(kernel, dim3(grid), dim3(128), static_cast<size_t>(kLookupSmem), s, p));
        prof_after(SF_KERNEL_LOOKUP, s);
        SF_CUDA_CHECK(cudaGetLastError());
        return SF_OK;
    };
    if (via_l1) return p.out_f16 ? launch(corr_lookup_kernel<true, true>) : launch(corr_lookup_kernel<false, true>);
    return p.out_f16 ? launch(corr_lookup_kernel<true, false>) : launch(corr_lookup_kernel<false, false>);
}

}  // namespace sf
