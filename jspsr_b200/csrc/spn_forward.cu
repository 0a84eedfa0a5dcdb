// Fused forward of one propagation application:
//   normalise 9 affinities -> 9-tap deformable bilinear gather of the DEM -> weighted
//   sum + bias (+ scale * centre in residual mode)
// Replaces PostProcessor.forward (models/components/spn.py:99-118),
// Post_process_deconv.forward (models/LRRU.py:267-298) and NLSPN._propagate_once
// (models/components/nlspn.py:177-187), i.e. 2 elementwise kernels + torchvision's
// im2col + GEMM, with one pass over HBM: 27 streamed channels in, 1 out, the DEM tile
// staged in shared memory by TMA.
#include "spn_kernels.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

// ---------------------------------------------------------------------------
// Fused halo exchange of a row strip (PEER instantiations; include/jspsr_peer.h).  Both halves are separate,
// non-inlined functions that recompute everything they need from blockIdx and the kernel parameters: nothing of the
// exchange is live across the per-pixel loop, whose code and register allocation stay those of the plain kernel
// (the first version tested `y < halo` per pixel: 1560 vs 1352 SASS instructions, 3x the spill traffic, -9 %).
// ---------------------------------------------------------------------------
struct EdgeRows {
    int n_top, n_bot;  // tile rows within `halo` of a strip edge that has a neighbour
};
template <int TH>
__device__ __forceinline__ EdgeRows edge_rows(const StripPeerDev& sp, const Geom& g) {
    EdgeRows e;
    e.n_top = sp.wait_up != nullptr ? (sp.halo + TH - 1) / TH : 0;
    e.n_bot = sp.wait_dn != nullptr ? g.tiles_y - max(0, g.H - sp.halo) / TH : 0;
    return e;
}

// before the DEM tile is staged: the edge CTAs wait until the neighbour's rows of this generation have landed
template <int TH>
__device__ __noinline__ void strip_peer_wait(const StripPeerDev& sp, const Geom& g, const bool all_threads_read) {
    const EdgeRows e = edge_rows<TH>(sp, g);
    const TileCtx c = make_tile_ctx<TH>(g, (sp.debug & 1) ? 0 : e.n_top, (sp.debug & 1) ? 0 : e.n_bot);
    const bool top_edge = e.n_top > 0 && c.y0 < sp.halo;
    const bool bot_edge = e.n_bot > 0 && c.y0 + TH > g.H - sp.halo;
    if (!(top_edge || bot_edge) || (sp.debug & 4)) return;
    if (threadIdx.x == 0) {
        if (top_edge) wait_stamp(sp.wait_up, sp.stamp);
        if (bot_edge) wait_stamp(sp.wait_dn, sp.stamp);
        asm volatile("fence.proxy.async;" ::: "memory");  // the TMA box copy reads what the peers wrote
    }
    if (all_threads_read) __syncthreads();  // the cooperative loader reads the halo rows from every thread
}

// after the CTA's pixels are stored: copy its edge rows into the neighbours' next DEM buffer (peer memory over NVLink)
// and, as the last edge CTA of that side, raise the neighbour's flag: "your halo rows of generation stamp + 1 are
// there, and I have finished reading mine of generation stamp"
template <typename TI, int TH>
__device__ __noinline__ void strip_peer_push(const StripPeerDev& sp, const Geom& g, const TI* __restrict__ out) {
    const EdgeRows e = edge_rows<TH>(sp, g);
    const TileCtx c = make_tile_ctx<TH>(g, (sp.debug & 1) ? 0 : e.n_top, (sp.debug & 1) ? 0 : e.n_bot);
    const bool top_edge = e.n_top > 0 && c.y0 < sp.halo;
    const bool bot_edge = e.n_bot > 0 && c.y0 + TH > g.H - sp.halo;
    if (!(top_edge || bot_edge) || (sp.debug & 8)) return;
    __syncthreads();  // every thread's st.global.cs of this CTA is performed at L2 before the rows are read back
    for (int i = threadIdx.x; i < ((sp.debug & 2) ? 0 : TH * TILE_W); i += THREADS) {
        const int y = c.y0 + i / TILE_W, x = c.x0 + (i & (TILE_W - 1));
        if (y >= g.H || x >= g.W) continue;
        const bool up = top_edge && sp.up_dst != nullptr && y < sp.halo;
        const bool dn = bot_edge && sp.dn_dst != nullptr && y >= g.H - sp.halo;
        if (!(up || dn)) continue;
        const TI v = __ldcg(out + (size_t)y * g.W + x);
        if (up) static_cast<TI*>(sp.up_dst)[(size_t)y * g.W + x] = v;
        if (dn) static_cast<TI*>(sp.dn_dst)[(size_t)(y - (g.H - sp.halo)) * g.W + x] = v;
    }
    // one system-scope fence per CTA, not per thread: the CTA barrier orders every thread's peer stores before thread 0's
    // fence (cumulativity), and 256 MEMBAR.SYS per edge CTA on SMs that are streaming at the HBM rate are not free
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned per_row = (unsigned)g.tiles_x;
        if (top_edge && atomicAdd(sp.tickets, 1u) == per_row * (unsigned)e.n_top - 1u) {
            sp.tickets[0] = 0u;
            __threadfence_system();
            st_release_sys(sp.up_flag, sp.stamp + 1u);
        }
        if (bot_edge && atomicAdd(sp.tickets + 1, 1u) == per_row * (unsigned)e.n_bot - 1u) {
            sp.tickets[1] = 0u;
            __threadfence_system();
            st_release_sys(sp.dn_flag, sp.stamp + 1u);
        }
    }
}

// CS: compile-time channel stride H*W (0 = runtime).  With CS known (the reference's
// 128x128 tiles: 16384) the 27 channel loads of a pixel share one address register with
// immediate offsets instead of a 64-bit add per channel.
// TH: rows per CTA (see spn_common.cuh).  `mode` (normalisation) is a runtime, warp-uniform switch.
// LINEAR: pixel-to-thread mapping (spn_kernels.cuh).  The row mapping is fastest when image rows are
// 128-byte aligned; when they are not (W*sizeof(T) % 128 != 0) a warp's 128-byte request straddles
// DRAM granules and only the linear mapping, where the four warps of a row issue together, lets L2
// merge them (measured at W = 2004: 11.1 GB read from DRAM for 7.5 GB of data with the row mapping).
// T: element type of weight / offset; TI: element type of init / out (TI = float with T = bf16 is what
// torch.autocast produces: bf16 Generator outputs, fp32 DEM; torchvision's operator promotes to fp32 there).
// PEER: row strip of a raster sharded over GPUs with the halo exchange fused in (include/jspsr_peer.h): the CTAs
// whose rows lie within `halo` of a strip edge - the only ones whose taps can reach a neighbour's rows - wait for
// that neighbour's flag before they stage their DEM tile, store their edge rows a second time into the neighbour's
// next DEM buffer (peer memory over NVLink), and the last of them to finish raises the neighbour's flag.  All other
// CTAs run exactly the non-PEER code, so the bulk of the band is computed while the boundary rows travel.
// PAIR (bf16 weight / offset on 128-byte aligned rows): a thread owns two horizontally adjacent pixels and streams
// their 27 channels as 27 bf16x2 words instead of 54 scalar loads, then runs the same per-pixel arithmetic on the
// low and the high halves one after the other (identical bits).  The kernels with bf16 inputs are bound by the
// load/store pipe (27 LDG + 36 LDS of a random gather per pixel-warp), not by HBM: halving the LDG count and
// amortising the address arithmetic over two pixels is what this mapping buys.
template <typename T, typename TI, bool TMA, int CS, int TH, bool LINEAR, bool PEER = false, bool PAIR = false>
__global__ void __launch_bounds__(THREADS, FWD_MIN_BLOCKS)
spn_forward_kernel(const TI* __restrict__ init, const T* __restrict__ weight, const T* __restrict__ offset,
                   const float* __restrict__ w9, const float* __restrict__ b1, TI* __restrict__ out,
                   const __grid_constant__ Geom g, const int mode, const float scale, int* __restrict__ status,
                   const __grid_constant__ CUtensorMap tmap, const __grid_constant__ StripPeerDev sp) {
    constexpr int SH = staged_rows(TH);
    constexpr int PPT = pixels_per_thread(TH);
    __shared__ __align__(128) TI tile[SH * SW];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) float s_w[12];  // w[0..8], b, 2 x padding (read as three 16-byte broadcasts)

    // PEER: the edge tile rows are scheduled first (make_tile_ctx) and wait for the neighbours' flag before staging
    const EdgeRows er = PEER ? edge_rows<TH>(sp, g) : EdgeRows{0, 0};
    const TileCtx c = make_tile_ctx<TH>(g, (PEER && (sp.debug & 1)) ? 0 : er.n_top, (PEER && (sp.debug & 1)) ? 0 : er.n_bot);
    if (PEER) strip_peer_wait<TH>(sp, g, !TMA);
    stage_tile_begin<TI, TMA, TH>(tile, &bar, &tmap, init, g, c.b, c.ox, c.oy - g.init_row0);
    // w9 == nullptr: frozen unit weight / zero bias (NLSPN, nlspn.py:61-65)
    if (threadIdx.x < 9) s_w[threadIdx.x] = w9 ? w9[threadIdx.x] : 1.f;
    if (threadIdx.x == 9) s_w[9] = b1 ? b1[0] : 0.f;

    const size_t cs = CS ? (size_t)CS : (size_t)g.H * g.W;
    const size_t csb = cs * sizeof(T);  // channel stride in bytes  // channel stride
    const T* wgt_b = weight + (size_t)c.b * 9 * cs;
    const T* off_b = offset + (size_t)c.b * 18 * cs;
    const TI* init_b = init + (size_t)c.b * g.init_rows * g.W;
    TI* out_b = out + (size_t)c.b * cs;
    const TI* tile_lo = tile + c.r_lo * SW;

    // pixel `it` of this thread: linear index it*256 + tid in the TH x 128 block (a warp = 32 consecutive x)
    struct PixelIn {
        float a[9], oh[9], ow[9];
        bool active;
        size_t p;
    };
    auto load_inputs = [&](int it, PixelIn& in) {
        float (&a)[9] = in.a;
        float (&oh)[9] = in.oh;
        float (&ow)[9] = in.ow;
        bool& active = in.active;
        size_t& p = in.p;
        const int y = c.y0 + pix_row<TH, LINEAR>(it), x = c.x0 + pix_col<TH, LINEAR>(it);
        active = (y < g.H) && (x < g.W);
        p = (size_t)y * g.W + x;
        if (active) {
            const T* pw = wgt_b + p;
            const T* po = off_b + p;
            if (CS) {  // compile-time stride: immediate offsets off one address register
#pragma unroll
                for (int k = 0; k < 9; ++k) a[k] = ld_stream(pw + k * cs);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    oh[k] = ld_stream(po + (2 * k) * cs);
                    ow[k] = ld_stream(po + (2 * k + 1) * cs);
                }
            } else {   // runtime stride: one opaque 64-bit add per channel off a few bases (short chains,
                       // so all 27 loads still issue back to back)
                const T* pw3 = step_ptr(pw, 3 * csb);
                const T* pw6 = step_ptr(pw, 6 * csb);
                a[0] = ld_stream(pw);
                a[1] = ld_stream(step_ptr(pw, csb));
                a[2] = ld_stream(step_ptr(pw, 2 * csb));
                a[3] = ld_stream(pw3);
                a[4] = ld_stream(step_ptr(pw3, csb));
                a[5] = ld_stream(step_ptr(pw3, 2 * csb));
                a[6] = ld_stream(pw6);
                a[7] = ld_stream(step_ptr(pw6, csb));
                a[8] = ld_stream(step_ptr(pw6, 2 * csb));
#pragma unroll
                for (int k3 = 0; k3 < 3; ++k3) {  // 6 offset channels per base
                    const T* pb = k3 == 0 ? po : step_ptr(po, (size_t)(6 * k3) * csb);
                    oh[3 * k3] = ld_stream(pb);
                    ow[3 * k3] = ld_stream(step_ptr(pb, csb));
                    oh[3 * k3 + 1] = ld_stream(step_ptr(pb, 2 * csb));
                    ow[3 * k3 + 1] = ld_stream(step_ptr(pb, 3 * csb));
                    oh[3 * k3 + 2] = ld_stream(step_ptr(pb, 4 * csb));
                    ow[3 * k3 + 2] = ld_stream(step_ptr(pb, 5 * csb));
                }
            }
        }
    };

    // one pixel at tile-local (ry, cx): consumes a[] / oh[] / ow[], returns the output value
    auto compute = [&](const int ry, const int cx, float (&a)[9], float (&oh)[9], float (&ow)[9]) -> float {
        normalise9(a, mode);
        {   // w_k * m_k up front: three 16-byte broadcast loads instead of one load per tap (the product is formed first
            // in either order of writing it, so the bits do not change; the load/store pipe is what bounds the gather)
            const float4 wa = *reinterpret_cast<const float4*>(s_w), wb = *reinterpret_cast<const float4*>(s_w + 4);
            const float w8 = s_w[8];
            a[0] *= wa.x; a[1] *= wa.y; a[2] *= wa.z; a[3] *= wa.w;
            a[4] *= wb.x; a[5] *= wb.y; a[6] *= wb.z; a[7] *= wb.w;
            a[8] *= w8;
        }

        // torchvision: (out_y - pad + i*dil) formed as an integer, converted, + offset
        const float fy = (float)(g.row0 + c.y0 + ry), fx = (float)(c.x0 + cx);
        const float hk[3] = {fy - 1.f, fy, fy + 1.f};
        const float wk[3] = {fx - 1.f, fx, fx + 1.f};
        // Each tap's contribution replaces its (normalised) affinity in a[]; a flagged tap keeps the
        // affinity until the slow pass has its value.  The final sum runs in tap order whatever the
        // mix of fast and slow taps, so the result does not depend on tiling or strip layout.
        unsigned slow = 0u;
        // Centre tap with a zero offset pair - what every producer in the reference emits (spn.py:72, LRRU.py, nlspn.py
        // insert a zero pair at the reference index): its position is the pixel itself, so the footprint address is known
        // without the position / floor / range arithmetic and the four loads are lane-consecutive (no bank conflicts).
        // Taken only when the whole warp agrees (one uniform branch); same loads, same bilerp with lh = lw = +0, hence
        // the same bits, including NaN / inf neighbours (0 * inf = NaN is kept: no fast-math).
        const TI* ctr = tile + (ry + HALO_T) * SW + (cx + HALO_L);
        const bool centre_fast =
            __all_sync(__activemask(), ((__float_as_uint(oh[4]) | __float_as_uint(ow[4])) << 1) == 0u &&
                                           (unsigned)(ry + HALO_T - c.r_lo) < c.r_span);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (k == 4 && centre_fast) {
                a[4] = a[4] * bilerp(to_f32(ctr[0]), to_f32(ctr[1]), to_f32(ctr[SW]), to_f32(ctr[SW + 1]), 0.f, 0.f);
                continue;
            }
            const FastTap t = fast_tap<TI>(tile_lo, c, hk[k / 3] + oh[k], wk[k % 3] + ow[k]);
            const float val = bilerp(t.v1, t.v2, t.v3, t.v4, t.lh, t.lw);
            a[k] = t.ok ? a[k] * val : a[k];
            slow |= t.ok ? 0u : (1u << k);
        }
        if (slow) {  // rare: redo the flagged taps through the bounds-checked global path
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (slow & (1u << k)) {
                    const SlowTap t = slow_tap<TI>(init_b, g, hk[k / 3] + oh[k], wk[k % 3] + ow[k], status);
                    a[k] = a[k] * bilerp(t.v1, t.v2, t.v3, t.v4, t.lh, t.lw);
                }
            }
        }
        float acc = a[0];
#pragma unroll
        for (int k = 1; k < 9; ++k) acc += a[k];
        acc += s_w[9];
        if (mode == NORM_RESIDUAL) acc = fmaf(scale, to_f32(ctr[0]), acc);
        return acc;
    };

    if constexpr (PAIR) {
        static_assert(sizeof(T) == 2 && CS != 0 && !LINEAR && PPT >= 2, "PAIR: bf16 weight / offset, compile-time stride, 128-byte aligned rows, >= 2 pixels per thread");
        // pair `jt` of this thread: a warp covers 64 consecutive x of one row (lane l: x = 2 l, 2 l + 1)
        struct PairIn {
            uint32_t a[9], oh[9], ow[9];
            bool active;
            size_t p;
        };
        auto pair_row = [](int jt) {
            if (TH >= 2 * WARPS) return (int)(threadIdx.x >> 5) + WARPS * (jt / (TILE_W / 64));
            return (jt * THREADS + (int)threadIdx.x) >> 6;
        };
        auto pair_col = [](int jt) {
            if (TH >= 2 * WARPS) return 2 * (int)(threadIdx.x & 31) + 64 * (jt % (TILE_W / 64));
            return 2 * ((jt * THREADS + (int)threadIdx.x) & 63);
        };
        auto load_pair = [&](int jt, PairIn& in) {
            const int y = c.y0 + pair_row(jt), x = c.x0 + pair_col(jt);
            in.active = (y < g.H) && (x < g.W);  // W is even here: both pixels or neither
            in.p = (size_t)y * g.W + x;
            if (in.active) {
                const T* pw = wgt_b + in.p;
                const T* po = off_b + in.p;
#pragma unroll
                for (int k = 0; k < 9; ++k) in.a[k] = ld_stream_x2(pw + k * cs);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    in.oh[k] = ld_stream_x2(po + (2 * k) * cs);
                    in.ow[k] = ld_stream_x2(po + (2 * k + 1) * cs);
                }
            }
        };
        // One pixel of the pair: half `sub` of the 27 words.  Same
        // arithmetic, in the same order, as compute() above, arranged so that nothing but the 27 words, two normalisation
        // constants and the running sum stays live: the affinity of tap k is widened and normalised when the tap needs it
        // (one FFMA covers the three modes exactly: a - mean = fma(a, 1, -mean), a * inv = fma(a, inv, -0),
        // a = fma(a, 1, -0)), and the tap contributions are summed as they come.  A pixel with an out-of-tile tap (rare) is
        // summed again from scratch in the same tap order with those taps taken from global memory, so the result does not
        // depend on which taps were on chip.
        auto compute_half = [&](const int ry, const int cx, const PairIn& in, const int sub) -> float {
            float nmul = 1.f, nadd = -0.f;
            if (mode != NORM_NONE) {
                float s = bf16x2_half(in.a[0], sub);
#pragma unroll
                for (int k = 1; k < 9; ++k) s += bf16x2_half(in.a[k], sub);
                if (mode == NORM_RESIDUAL) nadd = -__fdiv_rn(s, 9.f);
                else nmul = __fdiv_rn(1.f, s);
            }
            const float fy = (float)(g.row0 + c.y0 + ry), fx = (float)(c.x0 + cx);
            const float hk[3] = {fy - 1.f, fy, fy + 1.f};
            const float wk[3] = {fx - 1.f, fx, fx + 1.f};
            const TI* ctr = tile + (ry + HALO_T) * SW + (cx + HALO_L);
            const bool centre_fast =
                __all_sync(__activemask(), ((__float_as_uint(bf16x2_half(in.oh[4], sub)) | __float_as_uint(bf16x2_half(in.ow[4], sub))) << 1) == 0u &&
                                               (unsigned)(ry + HALO_T - c.r_lo) < c.r_span);
            bool slow = false;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float m = fmaf(bf16x2_half(in.a[k], sub), nmul, nadd);
                float val;
                if (k == 4 && centre_fast) {
                    val = bilerp(to_f32(ctr[0]), to_f32(ctr[1]), to_f32(ctr[SW]), to_f32(ctr[SW + 1]), 0.f, 0.f);
                } else {
                    const FastTap t = fast_tap<TI>(tile_lo, c, hk[k / 3] + bf16x2_half(in.oh[k], sub), wk[k % 3] + bf16x2_half(in.ow[k], sub));
                    val = bilerp(t.v1, t.v2, t.v3, t.v4, t.lh, t.lw);
                    slow |= !t.ok;
                }
                const float ck = __fmul_rn(s_w[k] * m, val);  // (no contraction with the running sum: compute() adds
                acc = k == 0 ? ck : __fadd_rn(acc, ck);        //  rounded products)
            }
            if (slow) {
#pragma unroll  // (a rolled loop would index the words dynamically and push them into local memory)
                for (int k = 0; k < 9; ++k) {
                    const float m = fmaf(bf16x2_half(in.a[k], sub), nmul, nadd);
                    const float h = hk[k / 3] + bf16x2_half(in.oh[k], sub), w = wk[k % 3] + bf16x2_half(in.ow[k], sub);
                    const FastTap t = fast_tap<TI>(tile_lo, c, h, w);
                    float val = bilerp(t.v1, t.v2, t.v3, t.v4, t.lh, t.lw);
                    if (!t.ok) {
                        const SlowTap u = slow_tap<TI>(init_b, g, h, w, status);
                        val = bilerp(u.v1, u.v2, u.v3, u.v4, u.lh, u.lw);
                    }
                    const float ck = __fmul_rn(s_w[k] * m, val);
                    acc = k == 0 ? ck : __fadd_rn(acc, ck);
                }
            }
            acc += s_w[9];
            if (mode == NORM_RESIDUAL) acc = fmaf(scale, to_f32(ctr[0]), acc);
            return acc;
        };
        // Lanes 0..15 take their left pixel first, lanes 16..31 their right one: in either pass the warp's 32 pixels sit
        // at x = 2 l + (l >= 16), i.e. on 32 different shared-memory banks when the offsets agree - with every lane on its
        // left pixel, lanes l and l + 16 are 32 columns apart and share a bank (ncu: 121 instead of 99 shared wavefronts
        // per pixel-warp, on a kernel whose load/store pipe is 89 % busy).  The upper lanes swap the halves of their words
        // once after the load (a per-lane byte selector in the unpacking itself made ptxas spill 18 of the 27 words).
        const int first = (int)((threadIdx.x >> 4) & 1u);
        const unsigned swap_sel = first ? 0x1032u : 0x3210u;
        PairIn cur;
        auto swap_halves = [&](PairIn& in) {
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                in.a[k] = __byte_perm(in.a[k], 0u, swap_sel);
                in.oh[k] = __byte_perm(in.oh[k], 0u, swap_sel);
                in.ow[k] = __byte_perm(in.ow[k], 0u, swap_sel);
            }
        };
        load_pair(0, cur);
        stage_tile_wait<TMA>(&bar);
#pragma unroll 1
        for (int jt = 0; jt < PPT / 2; ++jt) {
            __syncwarp();
            if (jt > 0) load_pair(jt, cur);
            const int ry = pair_row(jt), cx = pair_col(jt);
            float ra = 0.f, rb = 0.f;
            if (cur.active) {
                swap_halves(cur);
                ra = compute_half(ry, cx + first, cur, 0);
            }
            __syncwarp();  // the first pixel may have split the warp on its out-of-tile path
            if (cur.active) rb = compute_half(ry, cx + (first ^ 1), cur, 1);
            if (cur.active) st_stream_x2(out_b + cur.p, first ? rb : ra, first ? ra : rb);
        }
        if (PEER) strip_peer_push<TI, TH>(sp, g, out);
        return;
    }

    // the first pixel's 27 streamed loads are in flight while the tile lands
    PixelIn cur;
    load_inputs(0, cur);
    stage_tile_wait<TMA>(&bar);

    // (A two-stage software pipeline - next pixel's loads issued before the current pixel's taps - was measured
    //  and rejected: 1.50-1.61 ms vs 1.23 ms at 4096 tiles, 40-44 us vs 32 us at 70 tiles; the 27 extra live
    //  registers cost more than the overlap gains, warp-level parallelism already hides the latency.)
#pragma unroll 1
    for (int it = 0; it < PPT; ++it) {
        __syncwarp();  // lanes that took the out-of-tile path on the previous pixel rejoin here: without it a warp stays split for the
                       // rest of the loop (ncu: 5.8 active threads per instruction with far offsets), every later pixel issued per fragment
        if (it > 0) load_inputs(it, cur);
        if (cur.active) st_stream(out_b + cur.p, compute(pix_row<TH, LINEAR>(it), pix_col<TH, LINEAR>(it), cur.a, cur.oh, cur.ow));
    }

    if (PEER) strip_peer_push<TI, TH>(sp, g, out);
}

template <typename T, typename TI, bool TMA, int CS, int TH, bool LINEAR, bool PEER = false, bool PAIR = false>
static void launch_fwd_one(const LaunchArgs& la) {
    dim3 grid((unsigned)((size_t)la.g.tiles_x * la.g.tiles_y * la.g.B));
    spn_forward_kernel<T, TI, TMA, CS, TH, LINEAR, PEER, PAIR><<<grid, THREADS, 0, la.stream>>>(
        (const TI*)la.init, (const T*)la.weight, (const T*)la.offset, la.w9, la.b1, (TI*)la.out, la.g, la.mode, la.scale,
        la.status, la.tmap, PEER ? *la.strip_peer : StripPeerDev{});
}

// fused halo exchange: rasters only (B = 1, runtime channel stride), 16 rows per CTA, fp32 or bf16 throughout
template <typename T>
static cudaError_t launch_fwd_peer(const LaunchArgs& la) {
    if (la.tile_h != 16 || la.g.B != 1) return cudaErrorInvalidValue;
    const bool aligned = ((size_t)la.g.W * sizeof(T)) % 128 == 0;
    if (la.use_tma && aligned) launch_fwd_one<T, T, true, 0, 16, false, true>(la);
    else if (la.use_tma) launch_fwd_one<T, T, true, 0, 16, true, true>(la);
    else launch_fwd_one<T, T, false, 0, 16, true, true>(la);
    return cudaGetLastError();
}

// (TMA, CS, LINEAR) variants: the compile-time stride only exists for 128x128-pixel planes (always
// TMA-able and row-aligned); rows that are not 128-byte aligned use the linear mapping
template <typename T, typename TI, int TH>
static void launch_fwd_th(const LaunchArgs& la) {
    const size_t cs = (size_t)la.g.H * la.g.W;
    const bool aligned = ((size_t)la.g.W * sizeof(T)) % 128 == 0;
    if constexpr (sizeof(T) == 2 && TH >= 4) {  // bf16 weight / offset on 128 x 128 planes: two pixels per thread
        if (la.pair && la.use_tma && cs == 16384 && aligned) {
            launch_fwd_one<T, TI, true, 16384, TH, false, false, true>(la);
            return;
        }
    }
    if (la.use_tma && cs == 16384 && aligned) launch_fwd_one<T, TI, true, 16384, TH, false>(la);
    else if (la.use_tma && aligned) launch_fwd_one<T, TI, true, 0, TH, false>(la);
    else if (la.use_tma) launch_fwd_one<T, TI, true, 0, TH, true>(la);
    else launch_fwd_one<T, TI, false, 0, TH, true>(la);
}

template <typename T, typename TI>
static cudaError_t launch_fwd_dtype(const LaunchArgs& la) {
    switch (la.tile_h) {
        case 16: launch_fwd_th<T, TI, 16>(la); break;
        case 8: launch_fwd_th<T, TI, 8>(la); break;
        case 4: launch_fwd_th<T, TI, 4>(la); break;
        case 2: launch_fwd_th<T, TI, 2>(la); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int stage_box_cols() { return SW; }
int stage_box_rows(int th) { return staged_rows(th); }

cudaError_t launch_spn_forward(const LaunchArgs& la) {
    if (la.strip_peer != nullptr) {
        if (la.bf16 && la.init_f32) return cudaErrorNotSupported;
        return la.bf16 ? launch_fwd_peer<__nv_bfloat16>(la) : launch_fwd_peer<float>(la);
    }
    if (la.bf16 && la.init_f32) return launch_fwd_dtype<__nv_bfloat16, float>(la);
    return la.bf16 ? launch_fwd_dtype<__nv_bfloat16, __nv_bfloat16>(la) : launch_fwd_dtype<float, float>(la);
}

}  // namespace JSPSR_VARIANT
}  // namespace jspsr
