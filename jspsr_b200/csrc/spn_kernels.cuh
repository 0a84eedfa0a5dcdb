// Per-tile context, the tap gather and the launch descriptor shared by the forward,
// backward and iterate kernels.
#pragma once
#include "spn_common.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

constexpr int FWD_MIN_BLOCKS = 4;  // 1024 threads/SM, <= 64 registers/thread
constexpr int BWD_MIN_BLOCKS = 3;       // fp32: 80 registers beat 64 + spills (measured)
constexpr int BWD_MIN_BLOCKS_BF16 = 4;  // bf16 is issue-bound: the extra resident CTA is worth 5 %
constexpr int pixels_per_thread(int th) { return th * TILE_W / THREADS; }  // 8 for TH = 16

// Pixel `it` of a thread inside the TH x 128 block.  A warp always covers 32 consecutive x of
// one row (coalesced 128-byte lines).  Two mappings, chosen per kernel from measurements
// (B200, 4096 tiles of 128x128, fp32):
//   LINEAR = false: warp w owns rows w, w+8, ... and walks the four 32-column segments of a row
//                   (forward: 1.27 ms vs 1.37 ms linear)
//   LINEAR = true : pixel index it*256 + tid, two full rows per pass
//                   (backward at TH = 8: 2.47 ms vs 2.57 ms for rows/TH = 16)
template <int TH, bool LINEAR>
__device__ __forceinline__ int pix_row(int it) {
    if (!LINEAR && TH >= 2 * WARPS) return (int)(threadIdx.x >> 5) + WARPS * (it / (TILE_W / 32));
    return (it * THREADS + (int)threadIdx.x) >> 7;
}
template <int TH, bool LINEAR>
__device__ __forceinline__ int pix_col(int it) {
    if (!LINEAR && TH >= 2 * WARPS) return (int)(threadIdx.x & 31) + 32 * (it % (TILE_W / 32));
    return (it * THREADS + (int)threadIdx.x) & (TILE_W - 1);
}

struct TileCtx {
    int b;        // sample
    int x0, y0;   // first output column / strip-local row of the tile
    int ox, oy;   // GLOBAL column / row of staged-tile element [0][0]
    int r_lo;     // staged rows [r_lo, r_lo + r_span] hold correct data for a (r, r+1) pair
    unsigned r_span;
    unsigned oy_lo;  // (unsigned)(oy + r_lo): global row of the first trusted staged row
};

// edge_top / edge_bot (fused halo exchange, B = 1): that many tile rows at the top / bottom of the strip are scheduled
// FIRST (CTAs are dispatched in blockIdx order).  They are the CTAs that feed the neighbouring ranks and that wait for
// them: run first, their rows reach the neighbours a whole kernel before the neighbours' next application needs them,
// and what they wait for was produced at the start of the neighbours' previous kernel.
template <int TH>
__device__ __forceinline__ TileCtx make_tile_ctx(const Geom& g, int edge_top = 0, int edge_bot = 0) {
    constexpr int SH = staged_rows(TH);
    TileCtx c;
    unsigned t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    int ty = t % g.tiles_y;
    if (edge_bot > 0 && edge_top + edge_bot <= g.tiles_y) {
        if (ty >= edge_top) ty = ty < edge_top + edge_bot ? g.tiles_y - edge_bot + (ty - edge_top) : ty - edge_bot;
    }
    c.b = t / g.tiles_y;
    c.x0 = tx * TILE_W;
    c.y0 = ty * TH;
    c.ox = c.x0 - HALO_L;
    c.oy = g.row0 + c.y0 - HALO_T;
    // A staged row is trustworthy when it lies outside the image (zero is the right
    // value) or inside the init buffer.  Only row strips have untrustworthy rows.
    int lo = 0, hi = SH;
    if (g.init_row0 > 0) lo = max(0, g.init_row0 - c.oy);
    if (g.init_row0 + g.init_rows < g.H_img) hi = min(SH, g.init_row0 + g.init_rows - c.oy);
    c.r_lo = lo;
    c.r_span = (hi - 1 > lo) ? (unsigned)(hi - 1 - lo) : 0u;
    c.oy_lo = (unsigned)(c.oy + lo);
    return c;
}

// a2 of SURVEY section 8: residual -> subtract the tap mean; sum -> divide by the tap sum.
// `mode` is warp-uniform (a kernel argument).  Returns the raw sum (needed by the
// sum-mode Jacobian).  Sum mode multiplies by one correctly rounded reciprocal instead of
// nine divisions (<= 1 ulp from torch's a / s).
__device__ __forceinline__ float normalise9(float (&a)[9], int mode) {
    if (mode == NORM_NONE) return 0.f;
    float s = a[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) s += a[k];
    if (mode == NORM_RESIDUAL) {
        const float mean = __fdiv_rn(s, 9.f);  // torch.mean: sum / n
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] -= mean;
    } else {
        const float inv = __fdiv_rn(1.f, s);
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] *= inv;
    }
    return s;
}

// ---------------------------------------------------------------------------
// Branch-free tap (the hot loop).  Every tap loads its four neighbours from the
// staged tile through ONE address; a tap whose footprint is not wholly inside the
// trusted part of the tile reads element 0 instead and is flagged, and the flagged
// taps of a pixel (rare: offsets beyond the halo, non-finite offsets, strip edges)
// are redone afterwards through slow_tap().  Keeping branches out of the 9-tap body
// lets the 36 shared loads of a pixel overlap instead of serialising per tap.
// ---------------------------------------------------------------------------
struct FastTap {
    float v1, v2, v3, v4, lh, lw;
    int h0, w0;
    bool ok;
};

// MANTISSA_FLOOR: floor without the conversion pipe (F2I / FRND run at a quarter of the FP32 rate): adding 1.5 * 2^23
// with round-down leaves floor(x) in the low mantissa bits for |x| < 2^22, and subtracting it back gives floor(x) as a
// float exactly, so h0 / lh are bit-identical to __float2int_rd / x - floorf(x) there.  Anything further out (or NaN)
// yields an index far outside the staged rows / columns (host: H, W < 2^22), fails the range test below and goes
// through slow_tap, which keeps the saturating conversions.  Same-box A/B (tools/ab_hot.py, 2048 tiles): one more
// instruction per coordinate costs the issue-bound propagation kernels 1-6 % (forward 627 -> 633 us, bf16 forward
// 592 -> 627 us), while the NLSPN affinity backward - 8 gathers whose conversions sit on its critical path - gains
// 7.6 % (2880 -> 2662 us); so only that kernel turns it on.
template <typename T, bool MANTISSA_FLOOR = false>
__device__ __forceinline__ FastTap fast_tap(const T* __restrict__ tile_lo, const TileCtx& c, float h, float w) {
    FastTap t;
    if (MANTISSA_FLOOR) {
        constexpr float MAGIC = 12582912.f;
        const float th = __fadd_rd(h, MAGIC), tw = __fadd_rd(w, MAGIC);
        t.h0 = __float_as_int(th) - 0x4B400000;
        t.w0 = __float_as_int(tw) - 0x4B400000;
        t.lh = h - (th - MAGIC);
        t.lw = w - (tw - MAGIC);
    } else {
        t.h0 = __float2int_rd(h);  // saturating; NaN -> 0
        t.w0 = __float2int_rd(w);
        t.lh = h - floorf(h);
        t.lw = w - floorf(w);
    }
    const unsigned r = (unsigned)t.h0 - c.oy_lo;
    const unsigned q = (unsigned)t.w0 - (unsigned)c.ox;
    t.ok = (r < c.r_span) && (q < (unsigned)(SW - 1));
    const T* s = tile_lo + (t.ok ? r * SW + q : 0u);
    t.v1 = to_f32(s[0]);
    t.v2 = to_f32(s[1]);
    t.v3 = to_f32(s[SW]);
    t.v4 = to_f32(s[SW + 1]);
    return t;
}

// bilinear value in lerp form (6 flops; within 1 ulp-ish of torchvision's 4-product form)
__device__ __forceinline__ float bilerp(float v1, float v2, float v3, float v4, float lh, float lw) {
    const float top = fmaf(lw, v2 - v1, v1);
    const float bot = fmaf(lw, v4 - v3, v3);
    return fmaf(lh, bot - top, top);
}

// Slow tap: per-corner bounds-checked global loads (exactly torchvision's corner rule).
struct SlowTap {
    float v1, v2, v3, v4, lh, lw;
    int h0, w0;
    bool finite;
};
template <typename T>
__device__ __noinline__ SlowTap slow_tap(const T* __restrict__ init_b, const Geom g, float h, float w, int* status) {
    SlowTap t;
    t.h0 = __float2int_rd(h);
    t.w0 = __float2int_rd(w);
    t.lh = h - floorf(h);
    t.lw = w - floorf(w);
    t.v1 = t.v2 = t.v3 = t.v4 = 0.f;
    t.finite = fabsf(h) < 1.0e9f && fabsf(w) < 1.0e9f;
    if (t.finite) {
        t.v1 = fetch_corner_global(init_b, g, t.h0, t.w0, status);
        t.v2 = fetch_corner_global(init_b, g, t.h0, t.w0 + 1, status);
        t.v3 = fetch_corner_global(init_b, g, t.h0 + 1, t.w0, status);
        t.v4 = fetch_corner_global(init_b, g, t.h0 + 1, t.w0 + 1, status);
    } else if (h == h && w == w) {
        t.lh = t.lw = 0.f;  // +-inf / absurdly far: torchvision returns 0; keep inf - inf = NaN out of it
    }                       // NaN positions keep lh/lw = NaN so the result is NaN like the reference's
    return t;
}

// Block-floating-point accumulation tile for the gradient scatters (grad_init, grad_confidence): see spn_backward.cu
struct GiScale {
    float scale, inv;
    int poison;
};
__device__ __forceinline__ GiScale gi_scale_from_sum(float S) {
    GiScale r;
    r.poison = !(S < 3.0e38f);  // inf or NaN somewhere in this CTA's gradients
    int ex = 0;
    if (!r.poison && S > 0.f) frexpf(S, &ex);  // S < 2^ex
    int e = 30 - ex;  // S * 2^e < 2^30; the per-add rounding slack (<= 0.5 each) keeps every cell below 2^31
    e = max(-120, min(120, e));
    r.scale = ldexpf(1.f, e);
    r.inv = ldexpf(1.f, -e);
    return r;
}
__device__ __forceinline__ void gi_add(int* cell, float v_scaled) { atomicAdd(cell, __float2int_rn(v_scaled)); }
// Two-tile form for scales that come from a loose bound: the rounding residual of the coarse add goes, in units
// of 2^-16 of the coarse quantum, into a second integer tile (|residual| <= 2^15 per add, so 2^15 adds cannot
// overflow it): 16 more bits of precision for one more native atomic.
__device__ __forceinline__ void gi_add2(int* cell, int* cell_lo, float v_scaled) {
    const int hi = __float2int_rn(v_scaled);
    atomicAdd(cell, hi);
    atomicAdd(cell_lo, __float2int_rn((v_scaled - (float)hi) * 65536.f));
}

// per-variant entry points used by abi.cu
cudaError_t launch_spn_forward(const LaunchArgs& la);
cudaError_t launch_spn_backward(const LaunchArgs& la);
// backward of the fixed-affinity loop (spn_iterate_backward.cu): 8 rows per CTA, fp32
cudaError_t launch_iter_carry(const float* g_a, const float* g_b, const float* aff, const float* offset,
                              const float* asum_in, float* asum_out, float* carry_out, const Geom& g, cudaStream_t stream);
cudaError_t launch_iter_grad(const float* grad_list, const float* carry, const float* feat_init, const float* list_out,
                             const float* aff, const float* offset, float* grad_aff, float* grad_offset, const Geom& g, int T,
                             int tile_h, bool use_tma, const CUtensorMap& tmap_init, const CUtensorMap& tmap_list,
                             cudaStream_t stream);
int iter_grad_tile_h(int T);
int stage_box_cols();        // extents of the staged DEM box = the TMA box
int stage_box_rows(int th);


}  // namespace JSPSR_VARIANT
}  // namespace jspsr
