// Backward of the fixed-affinity T-step loop (autograd of NLSPN.forward's loop, models/components/nlspn.py:222-235:
// w = 1, b = 0, mask = aff, no normalisation), as TWO kinds of launches instead of T full backward applications.
//
// Step t of the loop is  out[t] = P(out[t-1])  with the same affinities and offsets for every t, so with
//   g_t = grad_list[t] + (gradient flowing back from step t + 1)
// the three gradients separate:
//   (A) the carry chain      g_{t-1} += P^T g_t : a scatter of g_t * a_k through the four bilinear coefficients of each
//       tap - coefficients that depend on (aff, offset) only.  No DEM is read, nothing but the carry is written:
//       iter_carry_kernel, T launches of ~128 B/pixel (the full backward application moves 116 B in, 108 B of
//       read-modify-write REDs and the carry: ~340 B/pixel, and is HBM-bound on exactly that: tools/nlspn_step_probe.py).
//   (B) grad_aff[k]    = sum_t g_t * v_k(t),   grad_offset[2k (+1)] = sum_t g_t * a_k * dv_k/dh (dw) (t):
//       the tap geometry is computed once per pixel, the sums over t run in registers against the T staged features
//       (all T tiles of a CTA are resident in shared memory, one TMA box each), and each of the 27 gradients is written
//       once: iter_grad_kernel, one launch of ~270 B/pixel bound by its 36 T shared-memory gathers.
// The per-step arithmetic is that of spn_backward_kernel (same bilinear derivatives, same block-floating-point scatter
// tile); the sums over t are associated in the same order (t = T - 1 first), so results agree with the T-application
// path to fp32 rounding (tests hold them to a relative 1e-6).
#include "spn_kernels.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

constexpr int ITER_BWD_TMAX = 8;  // steps held in registers / shared memory by iter_grad_kernel
constexpr int iter_tile_stride(int th) { return (staged_rows(th) * SW + 31) / 32 * 32; }

// tap geometry without touching the tile: the arithmetic of fast_tap (spn_kernels.cuh)
struct TapGeo {
    float lh, lw;
    int h0, w0;
    unsigned idx;  // element index of corner (h0, w0) relative to the first trusted staged row
    bool ok;
};
__device__ __forceinline__ TapGeo tap_geo(const TileCtx& c, float h, float w) {
    TapGeo t;
    t.h0 = __float2int_rd(h);
    t.w0 = __float2int_rd(w);
    t.lh = h - floorf(h);
    t.lw = w - floorf(w);
    const unsigned r = (unsigned)t.h0 - c.oy_lo;
    const unsigned q = (unsigned)t.w0 - (unsigned)c.ox;
    t.ok = (r < c.r_span) && (q < (unsigned)(SW - 1));
    t.idx = t.ok ? r * SW + q : 0u;
    return t;
}

__device__ __forceinline__ void carry_corner_global(float* __restrict__ dst_b, const Geom& g, int hi, int wi, float v) {
    if ((unsigned)hi >= (unsigned)g.H || (unsigned)wi >= (unsigned)g.W) return;
    atomicAdd(dst_b + (size_t)hi * g.W + wi, v);
}

// (A) one step of the carry chain.  g = g_a (+ g_b); carry_out (zero on entry) += P^T g.
// asum = sum_k |a_k| per pixel, the iteration-invariant factor of the tile's scale bound: the first launch of a chain
// (asum_in == nullptr) forms it from the nine affinities and stores it (asum_out), the others read one value per
// pixel instead of nine in their pre-pass.
// resident CTAs per SM: 4 (55 registers); 5 (48 registers, no spills) measured the same on the same box (T = 6, 2048 tiles:
// 8.45 vs 8.46 ms for the whole backward) - the kernel sits at 70 % of both the issue and the LSU wavefront rate
#ifndef JSPSR_CARRY_MIN_BLOCKS
#define JSPSR_CARRY_MIN_BLOCKS 4
#endif
template <int CS, int TH>
__global__ void __launch_bounds__(THREADS, JSPSR_CARRY_MIN_BLOCKS)
iter_carry_kernel(const float* __restrict__ g_a, const float* __restrict__ g_b, const float* __restrict__ aff,
                  const float* __restrict__ offset, const float* __restrict__ asum_in, float* __restrict__ asum_out,
                  float* __restrict__ carry_out, const Geom g) {
    constexpr int SH = staged_rows(TH);
    constexpr int PPT = pixels_per_thread(TH);
    __shared__ __align__(16) int gtile[SH * SW];  // block-floating-point accumulation tile (spn_backward.cu)
    __shared__ float s_gi[WARPS];
    __shared__ GiScale s_gis;

    const TileCtx c = make_tile_ctx<TH>(g);
    for (int i = threadIdx.x; i < SH * SW / 4; i += THREADS) reinterpret_cast<int4*>(gtile)[i] = make_int4(0, 0, 0, 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t cs = CS ? (size_t)CS : (size_t)g.H * g.W;
    const size_t plane = (size_t)c.b * cs;
    const float* aff_b = aff + 9 * plane;
    const float* off_b = offset + 18 * plane;
    float* out_b = carry_out + plane;
    int* gtile_lo = gtile + c.r_lo * SW;

    // pre-pass: S = sum over the CTA of |g| * sum_k |a_k| bounds every cell of the tile
    float part = 0.f;
#pragma unroll
    for (int it = 0; it < PPT; ++it) {
        const int y = c.y0 + pix_row<TH, true>(it), x = c.x0 + pix_col<TH, true>(it);
        if (y < g.H && x < g.W) {
            const size_t q = plane + (size_t)y * g.W + x;
            const float gq = g_b ? g_a[q] + g_b[q] : g_a[q];
            float sa;
            if (asum_in) {
                sa = asum_in[q];
            } else {
                sa = 0.f;
#pragma unroll
                for (int k = 0; k < 9; ++k) sa += fabsf(aff_b[q - plane + k * cs]);
                if (asum_out) asum_out[q] = sa;
            }
            part = fmaf(fabsf(gq), sa, part);
        }
    }
    part = warp_sum(part);
    if (lane == 0) s_gi[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float S = 0.f;
#pragma unroll
        for (int wi = 0; wi < WARPS; ++wi) S += s_gi[wi];
        s_gis = gi_scale_from_sum(S);
    }
    __syncthreads();
    const float gscale = s_gis.scale;

#pragma unroll 1
    for (int it = 0; it < PPT; ++it) {
        __syncwarp();
        const int y = c.y0 + pix_row<TH, true>(it), x = c.x0 + pix_col<TH, true>(it);
        if (!(y < g.H && x < g.W)) continue;
        const size_t p = (size_t)y * g.W + x;
        float a[9], oh[9], ow[9];
        const float* pw = aff_b + p;
        const float* po = off_b + p;
        float go = ld_stream(g_a + plane + p);
        if (g_b) go += ld_stream(g_b + plane + p);
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] = ld_stream(pw + k * cs);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            oh[k] = ld_stream(po + (2 * k) * cs);
            ow[k] = ld_stream(po + (2 * k + 1) * cs);
        }
        const float fy = (float)(g.row0 + y), fx = (float)x;
        const float hk[3] = {fy - 1.f, fy, fy + 1.f};
        const float wk[3] = {fx - 1.f, fx, fx + 1.f};
        go *= gscale;  // exact: a power of two
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float h = hk[k / 3] + oh[k], w = wk[k % 3] + ow[k];
            const TapGeo t = tap_geo(c, h, w);
            const float gks = go * a[k];
            const float ch = gks * t.lh, cl = gks - ch;
            const float c2 = cl * t.lw, c4 = ch * t.lw;
            if (t.ok) {
                int* gt = gtile_lo + t.idx;
                gi_add(gt, cl - c2);
                if (c2 != 0.f) gi_add(gt + 1, c2);
                if (ch != c4) gi_add(gt + SW, ch - c4);
                if (c4 != 0.f) gi_add(gt + SW + 1, c4);
            } else if (fabsf(h) < 1.0e9f && fabsf(w) < 1.0e9f) {
                // rare: footprint outside the accumulation tile -> global atomics with torchvision's corner rule
                // (handled in place: nothing of the pixel's 27 inputs has to stay live for a second pass)
                const float ginv = s_gis.inv;
                carry_corner_global(out_b, g, t.h0, t.w0, (cl - c2) * ginv);
                carry_corner_global(out_b, g, t.h0, t.w0 + 1, c2 * ginv);
                carry_corner_global(out_b, g, t.h0 + 1, t.w0, (ch - c4) * ginv);
                carry_corner_global(out_b, g, t.h0 + 1, t.w0 + 1, c4 * ginv);
            }
        }
    }

    // flush: contributions to cells outside the image are dropped (zero padding has no gradient)
    __syncthreads();
    const bool vec_ok = (g.W & 3) == 0 && ((reinterpret_cast<uintptr_t>(carry_out) & 15) == 0);
    const float ginv = s_gis.poison ? __int_as_float(0x7fc00000) : s_gis.inv;
    for (int i = threadIdx.x; i < SH * (SW / 4); i += THREADS) {
        const int r = i / (SW / 4), q = (i - r * (SW / 4)) * 4;
        const int4 iv = *reinterpret_cast<const int4*>(gtile + r * SW + q);
        if ((iv.x | iv.y | iv.z | iv.w) == 0 && !s_gis.poison) continue;
        const float4 v = make_float4((float)iv.x * ginv, (float)iv.y * ginv, (float)iv.z * ginv, (float)iv.w * ginv);
        const int gy = c.oy + r, gx = c.ox + q;
        if ((unsigned)gy >= (unsigned)g.H) continue;
        if (vec_ok && gx >= 0 && gx + 3 < g.W) {
            float* dst = out_b + (size_t)gy * g.W + gx;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
        } else {
            carry_corner_global(out_b, g, gy, gx, v.x);
            carry_corner_global(out_b, g, gy, gx + 1, v.y);
            carry_corner_global(out_b, g, gy, gx + 2, v.z);
            carry_corner_global(out_b, g, gy, gx + 3, v.w);
        }
    }
}

// (B) grad_aff / grad_offset over all T steps.  src(t) = feat_init for t = 0, list_out[t - 1] otherwise;
// g_t = grad_list[t] + carry[t] (t < T - 1; carry[t] = what step t + 1 sent back, written by (A)).
// Two shapes: 8 rows x 256 threads (three CTAs per SM up to T = 6, two for T = 7, 8) and 16 rows x 512 threads (two CTAs
// per SM = 32 warps instead of 24, half the halo rows per output row; T <= 6: 6 x 16.4 KB of tiles per CTA).
template <bool TMA, int CS, int TH, int NT>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 3)
iter_grad_kernel(const float* __restrict__ grad_list, const float* __restrict__ carry, const float* __restrict__ feat_init,
                 const float* __restrict__ list_out, const float* __restrict__ aff, const float* __restrict__ offset,
                 float* __restrict__ grad_aff, float* __restrict__ grad_offset, const Geom g, const int T,
                 const __grid_constant__ CUtensorMap tmap_init, const __grid_constant__ CUtensorMap tmap_list) {
    constexpr int SH = staged_rows(TH);
    constexpr int PPT = TH * TILE_W / NT;
    constexpr int TILE_ELEMS = iter_tile_stride(TH);  // SH * SW rounded up to 128 bytes: every TMA box lands 128-byte aligned
    extern __shared__ __align__(128) float tiles[];  // [T][TILE_ELEMS]
    __shared__ __align__(8) uint64_t bar;

    const TileCtx c = make_tile_ctx<TH>(g);
    const size_t cs = CS ? (size_t)CS : (size_t)g.H * g.W;
    const size_t step = (size_t)g.B * cs;  // elements between consecutive steps of grad_list / carry / list_out
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
            mbar_arrive_expect_tx(&bar, (uint32_t)(T * SH * SW * sizeof(float)));
            tma_load_3d(tiles, &tmap_init, &bar, c.ox, c.oy, c.b);
            for (int t = 1; t < T; ++t) tma_load_3d(tiles + t * TILE_ELEMS, &tmap_list, &bar, c.ox, c.oy, (t - 1) * g.B + c.b);
        }
    } else {
        for (int t = 0; t < T; ++t) {
            const float* src = (t == 0 ? feat_init : list_out + (size_t)(t - 1) * step) + (size_t)c.b * cs;
            for (int i = threadIdx.x; i < SH * SW; i += NT) {
                const int r = i / SW, q = i - r * SW;
                const int gy = c.oy + r, gx = c.ox + q;
                float v = 0.f;
                if ((unsigned)gy < (unsigned)g.H && (unsigned)gx < (unsigned)g.W) v = src[(size_t)gy * g.W + gx];
                tiles[t * TILE_ELEMS + i] = v;
            }
        }
    }
    const float* aff_b = aff + (size_t)c.b * 9 * cs;
    const float* off_b = offset + (size_t)c.b * 18 * cs;
    float* gaff_b = grad_aff + (size_t)c.b * 9 * cs;
    float* goff_b = grad_offset + (size_t)c.b * 18 * cs;
    const float* tile_lo = tiles + c.r_lo * SW;
    __syncthreads();
    if (TMA) mbar_wait(&bar, 0);

#pragma unroll 1
    for (int it = 0; it < PPT; ++it) {
        __syncwarp();
        const int pi = it * NT + (int)threadIdx.x;  // a warp = 32 consecutive columns of one row
        const int y = c.y0 + (pi >> 7), x = c.x0 + (pi & (TILE_W - 1));
        if (!(y < g.H && x < g.W)) continue;
        const size_t p = (size_t)y * g.W + x;
        // everything this pixel needs from global memory is requested up front (27 + 2 T - 1 loads in flight per lane)
        float a[9], oh[9], ow[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) a[k] = ld_stream(aff_b + p + k * cs);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            oh[k] = ld_stream(off_b + p + (2 * k) * cs);
            ow[k] = ld_stream(off_b + p + (2 * k + 1) * cs);
        }
        float gt[ITER_BWD_TMAX];
#pragma unroll
        for (int t = 0; t < ITER_BWD_TMAX; ++t) {
            gt[t] = 0.f;
            if (t < T) {
                gt[t] = ld_stream(grad_list + (size_t)t * step + (size_t)c.b * cs + p);
                if (t < T - 1) gt[t] += ld_stream(carry + (size_t)t * step + (size_t)c.b * cs + p);
            }
        }
        const float fy = (float)(g.row0 + y), fx = (float)x;
        const float hk[3] = {fy - 1.f, fy, fy + 1.f};
        const float wk[3] = {fx - 1.f, fx, fx + 1.f};
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float h = hk[k / 3] + oh[k], w = wk[k % 3] + ow[k];
            const TapGeo tg = tap_geo(c, h, w);
            float s_val = 0.f, s_dh = 0.f, s_dw = 0.f;
            // branch-free over the lanes: a tap outside the staged tile reads element 0 and contributes through the
            // (rare) global-corner pass below instead
            const float* s = tile_lo + tg.idx;
#pragma unroll
            for (int t = ITER_BWD_TMAX - 1; t >= 0; --t) {
                if (t < T) {
                    const float v1 = s[t * TILE_ELEMS], v2 = s[t * TILE_ELEMS + 1];
                    const float v3 = s[t * TILE_ELEMS + SW], v4 = s[t * TILE_ELEMS + SW + 1];
                    const float d21 = v2 - v1, d43 = v4 - v3;
                    const float top = fmaf(tg.lw, d21, v1), bot = fmaf(tg.lw, d43, v3);
                    const float dh = bot - top;
                    const float val = fmaf(tg.lh, dh, top);
                    const float dw = fmaf(tg.lh, d43 - d21, d21);
                    const float gkm = gt[t] * a[k];
                    s_val += gt[t] * val;
                    s_dh += gkm * dh;
                    s_dw += gkm * dw;
                }
            }
            if (!tg.ok) {  // rare: bounds-checked global corners, step by step
                s_val = s_dh = s_dw = 0.f;
                Geom gs = g;
                gs.H_img = g.H; gs.init_row0 = 0; gs.init_rows = g.H;
#pragma unroll  // (static indices: a rolled loop would put gt[] in local memory for every pixel)
                for (int t = ITER_BWD_TMAX - 1; t >= 0; --t) {
                    if (t < T) {
                        const float* src = (t == 0 ? feat_init : list_out + (size_t)(t - 1) * step) + (size_t)c.b * cs;
                        const SlowTap st = slow_tap<float>(src, gs, h, w, nullptr);
                        const float d21 = st.v2 - st.v1, d43 = st.v4 - st.v3;
                        const float top = fmaf(st.lw, d21, st.v1), bot = fmaf(st.lw, d43, st.v3);
                        const float dh = bot - top, val = fmaf(st.lh, dh, top), dw = fmaf(st.lh, d43 - d21, d21);
                        const float gkm = gt[t] * a[k];
                        s_val += gt[t] * val;
                        s_dh += gkm * dh;
                        s_dw += gkm * dw;
                    }
                }
            }
            st_stream(gaff_b + p + k * cs, s_val);
            st_stream(goff_b + p + (2 * k) * cs, s_dh);
            st_stream(goff_b + p + (2 * k + 1) * cs, s_dw);
        }
    }
}

// rows per CTA of iter_grad_kernel for T steps: 16 (512 threads) while its T tiles leave room for two CTAs per SM
int iter_grad_tile_h(int T) {
    if (const char* e = getenv("JSPSR_ITER_GRAD_TH")) {
        if (atoi(e) == 8) return 8;
        if (atoi(e) == 16 && T <= 6) return 16;
    }
    return T <= 6 ? 16 : 8;
}

cudaError_t launch_iter_carry(const float* g_a, const float* g_b, const float* aff, const float* offset,
                              const float* asum_in, float* asum_out, float* carry_out, const Geom& g, cudaStream_t stream) {
    // geometry: 8 rows per CTA (the caller filled tiles_y for that)
    const dim3 grid((unsigned)((size_t)g.tiles_x * g.tiles_y * g.B));
    if ((size_t)g.H * g.W == 16384)
        iter_carry_kernel<16384, 8><<<grid, THREADS, 0, stream>>>(g_a, g_b, aff, offset, asum_in, asum_out, carry_out, g);
    else
        iter_carry_kernel<0, 8><<<grid, THREADS, 0, stream>>>(g_a, g_b, aff, offset, asum_in, asum_out, carry_out, g);
    return cudaGetLastError();
}

// g.tiles_y and the tensor maps' boxes are the caller's, for `tile_h` rows per CTA
cudaError_t launch_iter_grad(const float* grad_list, const float* carry, const float* feat_init, const float* list_out,
                             const float* aff, const float* offset, float* grad_aff, float* grad_offset, const Geom& g, int T,
                             int tile_h, bool use_tma, const CUtensorMap& tmap_init, const CUtensorMap& tmap_list,
                             cudaStream_t stream) {
    const dim3 grid((unsigned)((size_t)g.tiles_x * g.tiles_y * g.B));
    const bool cs128 = (size_t)g.H * g.W == 16384;
    // the shared-memory opt-in is made once per kernel: ask for the largest T the shape is used with
#define JSPSR_LAUNCH_ITER_GRAD(TMA_, CS_, TH_, NT_, TMAX_)                                                             \
    do {                                                                                                               \
        const size_t smem = (size_t)T * iter_tile_stride(TH_) * sizeof(float);                                         \
        cudaError_t e = ensure_dynamic_smem((const void*)iter_grad_kernel<TMA_, CS_, TH_, NT_>,                        \
                                            (size_t)(TMAX_) * iter_tile_stride(TH_) * sizeof(float));                  \
        if (e != cudaSuccess) return e;                                                                                \
        iter_grad_kernel<TMA_, CS_, TH_, NT_><<<grid, NT_, smem, stream>>>(grad_list, carry, feat_init, list_out, aff, \
                                                                          offset, grad_aff, grad_offset, g, T,         \
                                                                          tmap_init, tmap_list);                       \
    } while (0)
#define JSPSR_LAUNCH_ITER_GRAD_SHAPE(TMA_, CS_)                                                                         \
    do {                                                                                                               \
        if (tile_h == 16) JSPSR_LAUNCH_ITER_GRAD(TMA_, CS_, 16, 512, 6);                                               \
        else JSPSR_LAUNCH_ITER_GRAD(TMA_, CS_, 8, 256, ITER_BWD_TMAX);                                                 \
    } while (0)
    if (tile_h == 16 && T > 6) return cudaErrorInvalidValue;
    if (use_tma) {
        if (cs128) JSPSR_LAUNCH_ITER_GRAD_SHAPE(true, 16384); else JSPSR_LAUNCH_ITER_GRAD_SHAPE(true, 0);
    } else {
        if (cs128) JSPSR_LAUNCH_ITER_GRAD_SHAPE(false, 16384); else JSPSR_LAUNCH_ITER_GRAD_SHAPE(false, 0);
    }
#undef JSPSR_LAUNCH_ITER_GRAD_SHAPE
#undef JSPSR_LAUNCH_ITER_GRAD
    return cudaGetLastError();
}

}  // namespace JSPSR_VARIANT
}  // namespace jspsr
