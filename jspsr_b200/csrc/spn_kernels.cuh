// Per-tile context, the tap gather and the launch descriptor shared by the forward,
// backward and iterate kernels.
#pragma once
#include "spn_common.cuh"

namespace jspsr {

constexpr int FWD_MIN_BLOCKS = 4;  // 1024 threads/SM, <= 64 registers/thread
constexpr int BWD_MIN_BLOCKS = 3;
constexpr int PIX_PER_THREAD = (TILE_H / WARPS) * (TILE_W / 32);  // 8

struct TileCtx {
    int b;        // sample
    int x0, y0;   // first output column / strip-local row of the tile
    int ox, oy;   // GLOBAL column / row of staged-tile element [0][0]
    int r_lo;     // staged rows [r_lo, r_lo + r_span] hold correct data for a (r, r+1) pair
    unsigned r_span;
};

__device__ __forceinline__ TileCtx make_tile_ctx(const Geom& g) {
    TileCtx c;
    unsigned t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    c.b = t / g.tiles_y;
    c.x0 = tx * TILE_W;
    c.y0 = ty * TILE_H;
    c.ox = c.x0 - HALO_L;
    c.oy = g.row0 + c.y0 - HALO_T;
    // A staged row is trustworthy when it lies outside the image (zero is the right
    // value) or inside the init buffer.  Only row strips have untrustworthy rows.
    int lo = 0, hi = SH;
    if (g.init_row0 > 0) lo = max(0, g.init_row0 - c.oy);
    if (g.init_row0 + g.init_rows < g.H_img) hi = min(SH, g.init_row0 + g.init_rows - c.oy);
    c.r_lo = lo;
    c.r_span = (hi - 1 > lo) ? (unsigned)(hi - 1 - lo) : 0u;
    return c;
}

template <int MODE>
__device__ __forceinline__ void normalise9(float (&a)[9]) {
    if (MODE == NORM_NONE) return;
    float s = a[0];
#pragma unroll
    for (int k = 1; k < 9; ++k) s += a[k];
    if (MODE == NORM_RESIDUAL) {
        const float mean = __fdiv_rn(s, 9.f);  // torch.mean: sum / n
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] -= mean;
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] = __fdiv_rn(a[k], s);
    }
}

struct Tap {
    float v1, v2, v3, v4;  // (h0,w0) (h0,w1) (h1,w0) (h1,w1), zero outside the image
    float lh, lw;          // fractional parts
    int h0, w0;            // GLOBAL integer corner
    bool in_tile;
};

// The four bilinear neighbours of a sample at GLOBAL position (h, w).
// Fast path: both rows and both columns are inside the staged tile -> 4 shared loads
// off one address (the tile is zero outside the image, so no bounds logic).
// Slow path: per-corner bounds-checked global loads.
template <typename T>
__device__ __forceinline__ Tap gather_tap(const T* __restrict__ tile, const T* __restrict__ init_b, const Geom& g,
                                          const TileCtx& c, float h, float w, int* status) {
    Tap t;
    const float hf = floorf(h), wf = floorf(w);
    t.lh = h - hf;
    t.lw = w - wf;
    t.h0 = __float2int_rd(h);  // saturating; NaN -> 0
    t.w0 = __float2int_rd(w);
    const unsigned r = (unsigned)t.h0 - (unsigned)c.oy;
    const unsigned q = (unsigned)t.w0 - (unsigned)c.ox;
    t.in_tile = (r - (unsigned)c.r_lo < c.r_span) && (q < (unsigned)(SW - 1));
    if (t.in_tile) {
        const T* s = tile + r * SW + q;
        t.v1 = to_f32(s[0]);
        t.v2 = to_f32(s[1]);
        t.v3 = to_f32(s[SW]);
        t.v4 = to_f32(s[SW + 1]);
    } else {
        t.v1 = t.v2 = t.v3 = t.v4 = 0.f;
        if (fabsf(h) < 1.0e9f && fabsf(w) < 1.0e9f) {
            // finite position: per-corner validity only.  This equals torchvision's
            // forward (its whole-sample test changes nothing when corners are checked)
            // and is exactly its backward (get_coordinate_weight has no such test).
            t.v1 = fetch_corner_global(init_b, g, t.h0, t.w0, status);
            t.v2 = fetch_corner_global(init_b, g, t.h0, t.w0 + 1, status);
            t.v3 = fetch_corner_global(init_b, g, t.h0 + 1, t.w0, status);
            t.v4 = fetch_corner_global(init_b, g, t.h0 + 1, t.w0 + 1, status);
        } else if (h == h && w == w) {
            // +-inf / absurdly far: torchvision returns 0; keep inf - inf = NaN out of it.
            t.lh = t.lw = 0.f;
        }  // NaN positions keep lh/lw = NaN so the result is NaN like the reference's
    }
    return t;
}

// Host-side launch descriptor (filled by abi.cu)
struct LaunchArgs {
    const void* init = nullptr;
    const void* weight = nullptr;
    const void* offset = nullptr;
    const float* w9 = nullptr;
    const float* b1 = nullptr;
    void* out = nullptr;
    // backward only
    const void* grad_out = nullptr;
    float* grad_init = nullptr;
    void* grad_weight = nullptr;
    void* grad_offset = nullptr;
    float* grad_w9 = nullptr;
    float* grad_b1 = nullptr;
    void* workspace = nullptr;
    bool accumulate = false;
    Geom g{};
    int mode = NORM_RESIDUAL;
    float scale = 1.f;
    bool bf16 = false;
    bool use_tma = false;
    int* status = nullptr;
    cudaStream_t stream = nullptr;
    CUtensorMap tmap{};
};

cudaError_t launch_spn_forward(const LaunchArgs& la);
cudaError_t launch_spn_backward(const LaunchArgs& la);

// reduction workspace layout (caller-owned, zero on entry, zero on exit)
struct alignas(16) ReduceWs {
    double sums[12];          // grad_w[0..8], grad_b, spare
    unsigned int ticket;      // CTAs that have contributed
    unsigned int pad[3];
};

}  // namespace jspsr
