// Fused backward of one propagation application (autograd of spn.py:99-118 /
// LRRU.py:267-298 / nlspn.py:177-187; upstream: deformable_col2im,
// deformable_col2im_coord, the GEMM for grad_weight and the bias sum, plus the
// Jacobian of the affinity normalisation) in ONE pass:
//   reads  grad_out, DEM tile (TMA-staged), 9 affinities, 18 offsets
//   writes grad_affinity[9], grad_offset[18]            (dense, one owner per element)
//          grad_w[9], grad_b                            (warp shuffle -> CTA -> fp64 atomics,
//                                                        last CTA publishes and re-zeroes)
//          grad_init (optional)                         (scatter: block-floating-point shared-memory
//                                                        tile of native integer atomics, flushed
//                                                        with global vector REDs)
#include "spn_kernels.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

template <typename T>
__device__ __forceinline__ void scatter_corner_global(float* __restrict__ gi_b, const Geom& g, int hi, int wi, float v) {
    if ((unsigned)hi >= (unsigned)g.H_img || (unsigned)wi >= (unsigned)g.W) return;
    const int br = hi - g.init_row0;
    if ((unsigned)br >= (unsigned)g.init_rows) return;
    atomicAdd(gi_b + (size_t)br * g.W + wi, v);
}

// grad_init accumulation tile.  sm_100a has no native shared-memory fp32 add: atomicAdd(float*) on shared
// memory is an LDS + ATOMS.CAST.SPIN loop (measured: 36 of them per pixel double the kernel time, 0.46 of the
// HBM roofline; lane-private tile copies buy 9 %, a hand-written atomicCAS loop is 4.7x slower).  Native
// ATOMS.ADD exists for 32-bit integers, so the tile is block floating point: a pre-pass over the CTA's own
// pixels sums |grad_out * m_k * w_k| over all taps (S); every cell's final magnitude is <= S because the four
// bilinear coefficients of a tap sum to 1, so with scale = 2^e, S * 2^e <= 2^30, no cell can overflow, the
// scaling itself is exact, and each contribution is rounded once to S * 2^-30 (typically ~1e-5 of the mean
// contribution, far below fp32 atomics' own order-dependent rounding).  Sums inside a CTA are exact integers,
// i.e. independent of the order the atomics land in.  A non-finite S poisons the CTA's cells with NaN.
// ACC: grad_weight / grad_offset are added to (fixed-affinity T-step loop) instead of written.
// CS : compile-time channel stride H*W (0 = runtime), see spn_forward.cu.
// TH : rows per CTA.  `mode` is a runtime, warp-uniform switch.
// T: element type of weight / offset and their gradients; TI: element type of grad_out / init (see spn_forward.cu).
// GZ (generator-tail training, JSPSR_BWD_GEN_PREACT): `grad_weight` is a [B,25,H,W] tensor that receives the gradients
// w.r.t. the PRE-ACTIVATIONS of Generator.conv_weight / conv_offset (spn.py:41-52,66-73): channels 0..8 =
// dL/d(weight_k) * weight_k * (1 - weight_k) (sigmoid'), channels 9..24 = the 16 offset gradients without the centre
// pair - the operand of the two 1x1-convolution gradient GEMMs, written here instead of by four elementwise passes.
// PACK (bf16 weight / offset, compile-time stride, no grad_init): the 27 bf16 inputs of a pixel are kept two to a
// register (14 instead of 27) and widened where a tap needs them - same operations in the same order as the plain path.
// At four resident CTAs (64 registers) the plain bf16 kernel spills about ten registers, and those local-memory
// accesses go through the same load/store pipe that bounds the kernel (ncu: LSU wavefronts 77 %, issue 69 %).
template <typename T, typename TI, bool TMA, bool GRAD_INIT, bool ACC, int CS, int TH, bool GZ = false, bool PACK = false>
__global__ void __launch_bounds__(THREADS, sizeof(T) == 2 ? BWD_MIN_BLOCKS_BF16 : BWD_MIN_BLOCKS)
spn_backward_kernel(const TI* __restrict__ gout, const TI* __restrict__ init, const T* __restrict__ weight,
                    const T* __restrict__ offset, const float* __restrict__ w9, float* __restrict__ grad_init,
                    T* __restrict__ grad_weight, T* __restrict__ grad_offset, float* __restrict__ grad_w9,
                    float* __restrict__ grad_b1, ReduceWs* __restrict__ ws, const Geom g, const int mode,
                    const float scale, const __grid_constant__ CUtensorMap tmap,
                    const __grid_constant__ PeerReduceDev pr) {
    constexpr int SH = staged_rows(TH);
    constexpr int PPT = pixels_per_thread(TH);
    __shared__ __align__(128) TI tile[SH * SW];
    __shared__ __align__(16) int gtile[GRAD_INIT ? SH * SW : 4];  // fixed-point accumulation tile
    __shared__ float s_gi[WARPS];
    __shared__ GiScale s_gis;
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_w[9];
    __shared__ float s_red[WARPS][10];
    __shared__ bool s_last;

    const TileCtx c = make_tile_ctx<TH>(g);
    stage_tile_begin<TI, TMA, TH>(tile, &bar, &tmap, init, g, c.b, c.ox, c.oy - g.init_row0);
    if (threadIdx.x < 9) s_w[threadIdx.x] = w9 ? w9[threadIdx.x] : 1.f;
    if (GRAD_INIT) {
        for (int i = threadIdx.x; i < SH * SW / 4; i += THREADS) reinterpret_cast<int4*>(gtile)[i] = make_int4(0, 0, 0, 0);
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t cs = CS ? (size_t)CS : (size_t)g.H * g.W;
    const size_t csb = cs * sizeof(T);  // channel stride in bytes
    const TI* gout_b = gout + (size_t)c.b * cs;
    const T* wgt_b = weight + (size_t)c.b * 9 * cs;
    const T* off_b = offset + (size_t)c.b * 18 * cs;
    const TI* init_b = init + (size_t)c.b * g.init_rows * g.W;
    T* gwgt_b = grad_weight + (size_t)c.b * (GZ ? 25 : 9) * cs;
    T* goff_b = GZ ? gwgt_b + 9 * cs : grad_offset + (size_t)c.b * 18 * cs;
    // channel of tap k's row-offset gradient inside goff_b (GZ: the centre pair k = 4 has no source and is skipped)
    auto och = [](int k) { return GZ ? 2 * (k - (k > 4 ? 1 : 0)) : 2 * k; };
    float* gi_b = GRAD_INIT ? grad_init + (size_t)c.b * g.init_rows * g.W : nullptr;
    const TI* tile_lo = tile + c.r_lo * SW;
    int* gtile_lo = gtile + (GRAD_INIT ? c.r_lo * SW : 0);

    float a[9], oh[9], ow[9], go;
    auto load_inputs = [&](int it, bool& active, size_t& p) {
        const int y = c.y0 + pix_row<TH, true>(it), x = c.x0 + pix_col<TH, true>(it);
        active = (y < g.H) && (x < g.W);
        p = (size_t)y * g.W + x;
        if (active) {
            const T* pw = wgt_b + p;
            const T* po = off_b + p;
            go = ld_stream(gout_b + p);
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
    auto store = [&](T* ptr, float v) {
        if (ACC && sizeof(T) == 4) {
            // accumulate at L2 (RED.ADD.F32, no return): no dependent load in front of every store.  Each element has
            // exactly one owner thread, so the result does not depend on ordering.
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(ptr), "f"(v) : "memory");
            return;
        }
        if (ACC) v += to_f32(*ptr);
        st_stream(ptr, v);
    };

    float acc_w[9], acc_b = 0.f;  // this thread's share of grad_w / grad_b
#pragma unroll
    for (int k = 0; k < 9; ++k) acc_w[k] = 0.f;

    bool active;
    size_t p;
    float gscale = 0.f;
    if (GRAD_INIT) {
        // pre-pass: S = sum over this CTA's pixels and taps of |grad_out * m_k * w_k| (+ the residual term);
        // runs while the TMA box is in flight, its 10 channels are re-read from L2 by the main loop
        float wabs[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) wabs[k] = fabsf(w9 ? __ldg(w9 + k) : 1.f);
        float part = 0.f;
#pragma unroll 1
        for (int it = 0; it < PPT; ++it) {
            const int y = c.y0 + pix_row<TH, true>(it), x = c.x0 + pix_col<TH, true>(it);
            if (y < g.H && x < g.W) {
                const size_t q = (size_t)y * g.W + x;
                const float gq = to_f32(gout_b[q]);
                float m[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) m[k] = to_f32(wgt_b[q + k * cs]);
                normalise9(m, mode);
                float sa = (mode == NORM_RESIDUAL) ? fabsf(scale) : 0.f;
#pragma unroll
                for (int k = 0; k < 9; ++k) sa = fmaf(fabsf(m[k]), wabs[k], sa);
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
    }
    if constexpr (PACK) {
        static_assert(sizeof(T) == 2 && CS != 0 && !GRAD_INIT && !ACC, "PACK: bf16, compile-time stride, plain stores");
        // 27 bf16 inputs of a pixel in 14 registers: value i (a0..a8, then oh_k / ow_k interleaved) is half i & 1 of pk[i >> 1]
        uint32_t pk[14];
        auto val_of = [&](int i) { return bf16x2_half(pk[i >> 1], i & 1); };
        auto load_packed = [&](int it) {
            const int y = c.y0 + pix_row<TH, true>(it), x = c.x0 + pix_col<TH, true>(it);
            active = (y < g.H) && (x < g.W);
            p = (size_t)y * g.W + x;
            if (active) {
                const T* pw = wgt_b + p;
                const T* po = off_b + p;
                go = ld_stream(gout_b + p);
                unsigned u[28];
#pragma unroll
                for (int k = 0; k < 9; ++k) u[k] = ld_stream_u16(pw + k * cs);
#pragma unroll
                for (int k = 0; k < 18; ++k) u[9 + k] = ld_stream_u16(po + k * cs);
                u[27] = 0u;
#pragma unroll
                for (int j = 0; j < 14; ++j) pk[j] = u[2 * j] | (u[2 * j + 1] << 16);
            }
        };
        // one pixel; the affinity of tap k is widened and normalised where it is needed (one FFMA covers the three modes
        // exactly: a - mean = fma(a, 1, -mean), a * inv = fma(a, inv, -0), a = fma(a, 1, -0)), so the inputs stay packed
        auto pixel = [&](const int ry, const int cx) {
            float nmul = 1.f, nadd = -0.f, s = 0.f;
            if (mode != NORM_NONE) {
                s = val_of(0);
#pragma unroll
                for (int k = 1; k < 9; ++k) s += val_of(k);
                if (mode == NORM_RESIDUAL) nadd = -__fdiv_rn(s, 9.f);
                else nmul = __fdiv_rn(1.f, s);
            }
            auto m_of = [&](int k) { return fmaf(val_of(k), nmul, nadd); };
            const float fy = (float)(g.row0 + c.y0 + ry), fx = (float)(c.x0 + cx);
            const float hk[3] = {fy - 1.f, fy, fy + 1.f};
            const float wk[3] = {fx - 1.f, fx, fx + 1.f};
            T* po = goff_b + p;
            float gm[9];
            unsigned slow = 0u;
            acc_b += go;
            const TI* ctr = tile + (ry + HALO_T) * SW + (cx + HALO_L);
            // centre offsets = values 9 + 8 (high half of pk[8]) and 9 + 9 (low half of pk[9])
            const bool centre_fast =
                __all_sync(__activemask(), (((pk[8] >> 16) | pk[9]) & 0x7fffu) == 0u && (unsigned)(ry + HALO_T - c.r_lo) < c.r_span);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                FastTap t;
                if (k == 4 && centre_fast) {
                    t.v1 = to_f32(ctr[0]); t.v2 = to_f32(ctr[1]); t.v3 = to_f32(ctr[SW]); t.v4 = to_f32(ctr[SW + 1]);
                    t.lh = 0.f; t.lw = 0.f;
                    t.ok = true;
                } else {
                    t = fast_tap<TI>(tile_lo, c, hk[k / 3] + val_of(9 + 2 * k), wk[k % 3] + val_of(10 + 2 * k));
                }
                const float d21 = t.v2 - t.v1, d43 = t.v4 - t.v3;
                const float top = fmaf(t.lw, d21, t.v1), bot = fmaf(t.lw, d43, t.v3);
                float dh = bot - top;
                float val = fmaf(t.lh, dh, top);
                float dw = fmaf(t.lh, d43 - d21, d21);
                if (!t.ok) {
                    val = dh = dw = 0.f;
                    slow |= 1u << k;
                }
                const float ga = go * m_of(k);
                const float gkm = ga * s_w[k];
                acc_w[k] = fmaf(ga, val, acc_w[k]);
                gm[k] = (go * s_w[k]) * val;
                if (!(GZ && k == 4)) {
                    store(po + och(k) * cs, gkm * dh);
                    store(po + (och(k) + 1) * cs, gkm * dw);
                }
            }
            if (slow) {
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (slow & (1u << k)) {
                        const float h = hk[k / 3] + val_of(9 + 2 * k), w = wk[k % 3] + val_of(10 + 2 * k);
                        const SlowTap t = slow_tap<TI>(init_b, g, h, w, nullptr);
                        const float d21 = t.v2 - t.v1, d43 = t.v4 - t.v3;
                        const float top = fmaf(t.lw, d21, t.v1), bot = fmaf(t.lw, d43, t.v3);
                        const float dh = bot - top, val = fmaf(t.lh, dh, top), dw = fmaf(t.lh, d43 - d21, d21);
                        const float ga = go * m_of(k), gkm = ga * s_w[k];
                        acc_w[k] = fmaf(ga, val, acc_w[k]);
                        gm[k] += (go * s_w[k]) * val;
                        if (!(GZ && k == 4)) {
                            store(po + och(k) * cs, gkm * dh);
                            store(po + (och(k) + 1) * cs, gkm * dw);
                        }
                    }
                }
            }
            if (mode == NORM_RESIDUAL) {
                float sg = gm[0];
#pragma unroll
                for (int k = 1; k < 9; ++k) sg += gm[k];
                const float mean = __fdiv_rn(sg, 9.f);
#pragma unroll
                for (int k = 0; k < 9; ++k) gm[k] -= mean;
            } else if (mode == NORM_SUM) {
                float dot = 0.f;
#pragma unroll
                for (int k = 0; k < 9; ++k) dot = fmaf(gm[k], m_of(k), dot);
                const float inv = __fdiv_rn(1.f, s);
#pragma unroll
                for (int k = 0; k < 9; ++k) gm[k] = (gm[k] - dot) * inv;
            }
            if (GZ) {
                const float mean = __fdiv_rn(s, 9.f);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const float mk = m_of(k);
                    const float raw = mode == NORM_RESIDUAL ? mk + mean : (mode == NORM_SUM ? mk * s : mk);
                    gm[k] *= raw * (1.f - raw);
                }
            }
            T* pw = gwgt_b + p;
#pragma unroll
            for (int k = 0; k < 9; ++k) store(pw + k * cs, gm[k]);
        };
        load_packed(0);
        stage_tile_wait<TMA>(&bar);
#pragma unroll 1
        for (int it = 0; it < PPT; ++it) {
            __syncwarp();
            if (it > 0) load_packed(it);
            if (active) pixel(pix_row<TH, true>(it), pix_col<TH, true>(it));
        }
    } else {
    load_inputs(0, active, p);
    stage_tile_wait<TMA>(&bar);
    if (GRAD_INIT) gscale = s_gis.scale;  // published before the barrier inside stage_tile_wait

#pragma unroll 1
    for (int it = 0; it < PPT; ++it) {
        __syncwarp();  // reconverge the lanes that took the out-of-tile path on the previous pixel (see spn_forward.cu)
        if (it > 0) load_inputs(it, active, p);
        if (!active) continue;
        const int ry = pix_row<TH, true>(it), cx = pix_col<TH, true>(it);

        const float s = normalise9(a, mode);  // a[] now holds the modulation m_k; s = raw sum

        const float fy = (float)(g.row0 + c.y0 + ry), fx = (float)(c.x0 + cx);
        const float hk[3] = {fy - 1.f, fy, fy + 1.f};
        const float wk[3] = {fx - 1.f, fx, fx + 1.f};
        T* po = goff_b + p;   // slow-pass stores index from here
        T* pos = po;          // running pointer of the fast pass when the stride is a runtime value
        float gm[9];
        unsigned slow = 0u;
        acc_b += go;
        // centre tap with a zero offset pair (every producer of the reference): footprint address known, loads
        // lane-consecutive; same values into the same arithmetic as the general path (see spn_forward.cu)
        const TI* ctr = tile + (ry + HALO_T) * SW + (cx + HALO_L);
        const bool centre_fast =
            __all_sync(__activemask(), ((__float_as_uint(oh[4]) | __float_as_uint(ow[4])) << 1) == 0u &&
                                           (unsigned)(ry + HALO_T - c.r_lo) < c.r_span);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            FastTap t;
            if (k == 4 && centre_fast) {
                t.v1 = to_f32(ctr[0]); t.v2 = to_f32(ctr[1]); t.v3 = to_f32(ctr[SW]); t.v4 = to_f32(ctr[SW + 1]);
                t.lh = 0.f; t.lw = 0.f;
                t.h0 = g.row0 + c.y0 + ry; t.w0 = c.x0 + cx;
                t.ok = true;
            } else {
                t = fast_tap<TI>(tile_lo, c, hk[k / 3] + oh[k], wk[k % 3] + ow[k]);
            }
            // value and both derivatives (torchvision get_coordinate_weight) share the two row differences
            const float d21 = t.v2 - t.v1, d43 = t.v4 - t.v3;
            const float top = fmaf(t.lw, d21, t.v1), bot = fmaf(t.lw, d43, t.v3);
            float dh = bot - top;
            float val = fmaf(t.lh, dh, top);
            float dw = fmaf(t.lh, d43 - d21, d21);
            if (!t.ok) {
                val = dh = dw = 0.f;
                slow |= 1u << k;
            }
            const float ga = go * a[k];          // go * m_k
            const float gkm = ga * s_w[k];       // dL/d(sample_k)
            acc_w[k] = fmaf(ga, val, acc_w[k]);
            gm[k] = (go * s_w[k]) * val;
            if (!(GZ && k == 4)) {
                if (CS) {
                    store(po + och(k) * cs, gkm * dh);
                    store(po + (och(k) + 1) * cs, gkm * dw);
                } else {
                    store(pos, gkm * dh);
                    pos = step_ptr(pos, csb);
                    store(pos, gkm * dw);
                    pos = step_ptr(pos, csb);
                }
            }
            if (GRAD_INIT) {
                if (t.ok) {
                    const float gks = gkm * gscale;              // exact: the scale is a power of two
                    const float ch = gks * t.lh, cl = gks - ch;  // rows h0+1 / h0
                    const float c2 = cl * t.lw, c4 = ch * t.lw;
                    int* gt = gtile_lo + ((unsigned)t.h0 - c.oy_lo) * SW + ((unsigned)t.w0 - (unsigned)c.ox);
                    gi_add(gt, cl - c2);
                    if (c2 != 0.f) gi_add(gt + 1, c2);            // zero for integer columns (centre tap)
                    if (ch != c4) gi_add(gt + SW, ch - c4);       // zero for integer rows
                    if (c4 != 0.f) gi_add(gt + SW + 1, c4);
                }
            }
        }
        if (slow) {  // rare: flagged taps through the bounds-checked global path
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (slow & (1u << k)) {
                    const float h = hk[k / 3] + oh[k], w = wk[k % 3] + ow[k];
                    const SlowTap t = slow_tap<TI>(init_b, g, h, w, nullptr);
                    const float d21 = t.v2 - t.v1, d43 = t.v4 - t.v3;
                    const float top = fmaf(t.lw, d21, t.v1), bot = fmaf(t.lw, d43, t.v3);
                    const float dh = bot - top, val = fmaf(t.lh, dh, top), dw = fmaf(t.lh, d43 - d21, d21);
                    const float ga = go * a[k], gkm = ga * s_w[k];
                    acc_w[k] = fmaf(ga, val, acc_w[k]);
                    gm[k] += (go * s_w[k]) * val;
                    if (!(GZ && k == 4)) {
                        store(po + och(k) * cs, gkm * dh);      // the fast pass stored 0 here (ACC: added 0)
                        store(po + (och(k) + 1) * cs, gkm * dw);
                    }
                    if (GRAD_INIT && t.finite) {
                        const float ch = gkm * t.lh, cl = gkm - ch, c2 = cl * t.lw, c4 = ch * t.lw;
                        scatter_corner_global<T>(gi_b, g, t.h0, t.w0, cl - c2);
                        scatter_corner_global<T>(gi_b, g, t.h0, t.w0 + 1, c2);
                        scatter_corner_global<T>(gi_b, g, t.h0 + 1, t.w0, ch - c4);
                        scatter_corner_global<T>(gi_b, g, t.h0 + 1, t.w0 + 1, c4);
                    }
                }
            }
        }
        // Jacobian of the normalisation
        if (mode == NORM_RESIDUAL) {
            float sg = gm[0];
#pragma unroll
            for (int k = 1; k < 9; ++k) sg += gm[k];
            const float mean = __fdiv_rn(sg, 9.f);
#pragma unroll
            for (int k = 0; k < 9; ++k) gm[k] -= mean;
            if (GRAD_INIT) gi_add(gtile + (ry + HALO_T) * SW + (cx + HALO_L), (scale * go) * gscale);
        } else if (mode == NORM_SUM) {
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) dot = fmaf(gm[k], a[k], dot);
            const float inv = __fdiv_rn(1.f, s);
#pragma unroll
            for (int k = 0; k < 9; ++k) gm[k] = (gm[k] - dot) * inv;
        }
        if (GZ) {  // sigmoid' at the RAW affinity: a[] holds the normalised m_k, the raw value is m_k + mean / m_k * s
            const float mean = __fdiv_rn(s, 9.f);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float raw = mode == NORM_RESIDUAL ? a[k] + mean : (mode == NORM_SUM ? a[k] * s : a[k]);
                gm[k] *= raw * (1.f - raw);
            }
        }
        T* pw = gwgt_b + p;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (CS) {
                store(pw + k * cs, gm[k]);
            } else {
                store(pw, gm[k]);
                pw = step_ptr(pw, csb);
            }
        }
    }

    }  // !PAIR

    // ---- grad_init: flush the shared accumulation tile (one 16-byte vector RED per 4 columns) ----
    if (GRAD_INIT) {
        __syncthreads();
        const bool vec_ok = (g.W & 3) == 0 && ((reinterpret_cast<uintptr_t>(grad_init) & 15) == 0);
        const float ginv = s_gis.poison ? __int_as_float(0x7fc00000) : s_gis.inv;
        for (int i = threadIdx.x; i < SH * (SW / 4); i += THREADS) {
            const int r = i / (SW / 4), q = (i - r * (SW / 4)) * 4;
            const int4 iv = *reinterpret_cast<const int4*>(gtile + r * SW + q);
            if ((iv.x | iv.y | iv.z | iv.w) == 0 && !s_gis.poison) continue;
            const float4 v = make_float4((float)iv.x * ginv, (float)iv.y * ginv, (float)iv.z * ginv, (float)iv.w * ginv);
            const int gy = c.oy + r, gx = c.ox + q;
            const int br = gy - g.init_row0;
            if ((unsigned)gy >= (unsigned)g.H_img || (unsigned)br >= (unsigned)g.init_rows) continue;
            if (vec_ok && gx >= 0 && gx + 3 < g.W) {
                float* dst = gi_b + (size_t)br * g.W + gx;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z),
                             "f"(v.w)
                             : "memory");
            } else {
                scatter_corner_global<T>(gi_b, g, gy, gx, v.x);
                scatter_corner_global<T>(gi_b, g, gy, gx + 1, v.y);
                scatter_corner_global<T>(gi_b, g, gy, gx + 2, v.z);
                scatter_corner_global<T>(gi_b, g, gy, gx + 3, v.w);
            }
        }
    }

    // ---- grad_w[9], grad_b: thread -> warp -> CTA -> fp64 atomics; last CTA publishes ----
    if (grad_w9 != nullptr) {
#pragma unroll
        for (int k = 0; k < 9; ++k) acc_w[k] = warp_sum(acc_w[k]);
        acc_b = warp_sum(acc_b);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) s_red[warp][k] = acc_w[k];
            s_red[warp][9] = acc_b;
        }
        __syncthreads();
        if (threadIdx.x < 10) {
            float v = 0.f;
#pragma unroll
            for (int wi = 0; wi < WARPS; ++wi) v += s_red[wi][threadIdx.x];
            atomicAdd(&ws->sums[threadIdx.x], (double)v);
            __threadfence();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(&ws->ticket, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            __shared__ double s_sum[10];
            if (threadIdx.x < 10) {
                s_sum[threadIdx.x] = atomicAdd(&ws->sums[threadIdx.x], 0.0);  // coherent read
                ws->sums[threadIdx.x] = 0.0;  // leave the workspace clean for the next call
            }
            if (threadIdx.x == 0) ws->ticket = 0u;
            if (pr.world > 1) {
                // All-reduce over the ranks, fused here (include/jspsr_peer.h): thread p stores this rank's ten fp64 sums
                // into its slot of rank p's buffer (peer memory over NVLink) and releases the step stamp behind them, then
                // waits for rank p's stamp in the local buffer.  Every rank adds the world's slots in rank order, so the
                // result is bit-identical everywhere.  Slots alternate with the stamp's parity: a rank can be at most one
                // step ahead of the slowest reader (it cannot finish step s + 1 before everyone has published step s + 1,
                // which they do after reading step s).
                __shared__ double s_all[8][10];
                __shared__ unsigned s_stamp;
                double* mine = pr.slots[pr.rank];
                if (threadIdx.x == 0) {
                    unsigned* counter = reinterpret_cast<unsigned*>(mine + REDUCE_COUNTER_OFFSET);
                    s_stamp = *counter + 1u;
                    *counter = s_stamp;
                }
                __syncthreads();
                const unsigned stamp = s_stamp;
                const int par = (int)(stamp & 1u);
                if ((int)threadIdx.x < pr.world) {
                    const int peer = (int)threadIdx.x;
                    double* dst = pr.slots[peer] + (size_t)(par * 8 + pr.rank) * REDUCE_SLOT;
#pragma unroll
                    for (int k = 0; k < 10; ++k) dst[k] = s_sum[k];
                    __threadfence_system();
                    st_release_sys(reinterpret_cast<unsigned*>(dst + REDUCE_SLOT - 1), stamp);
                    const double* src = mine + (size_t)(par * 8 + peer) * REDUCE_SLOT;
                    wait_stamp(reinterpret_cast<const unsigned*>(src + REDUCE_SLOT - 1), stamp);
#pragma unroll
                    for (int k = 0; k < 10; ++k) {
                        double v;
                        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(src + k) : "memory");
                        s_all[peer][k] = v;
                    }
                }
                __syncthreads();
                if (threadIdx.x < 10) {
                    double v = 0.0;
                    for (int r = 0; r < pr.world; ++r) v += s_all[r][threadIdx.x];
                    s_sum[threadIdx.x] = v * (double)pr.mul;
                }
                __syncthreads();
            }
            if (threadIdx.x < 9) grad_w9[threadIdx.x] = (float)s_sum[threadIdx.x];
            else if (threadIdx.x == 9 && grad_b1 != nullptr) grad_b1[0] = (float)s_sum[9];
        }
    }
}

template <typename T, typename TI, bool TMA, bool GI, bool ACC, int CS, int TH, bool GZ = false, bool PACK = false>
static void launch_one(const LaunchArgs& la) {
    dim3 grid((unsigned)((size_t)la.g.tiles_x * la.g.tiles_y * la.g.B));
    spn_backward_kernel<T, TI, TMA, GI, ACC, CS, TH, GZ, PACK><<<grid, THREADS, 0, la.stream>>>(
        (const TI*)la.grad_out, (const TI*)la.init, (const T*)la.weight, (const T*)la.offset, la.w9, la.grad_init,
        (T*)la.grad_weight, (T*)la.grad_offset, la.grad_w9, la.grad_b1, (ReduceWs*)la.workspace, la.g, la.mode,
        la.scale, la.tmap, la.peer_reduce ? *la.peer_reduce : PeerReduceDev{});
}

// (TMA, CS) variants: the compile-time stride only exists for 128x128-pixel planes, which always qualify for TMA
template <typename T, typename TI, bool GI, bool ACC, int TH, bool GZ = false>
static void launch_variant(const LaunchArgs& la) {
    const size_t cs = (size_t)la.g.H * la.g.W;
    if constexpr (sizeof(T) == 2 && !GI && !ACC) {  // bf16 on 128 x 128 planes: inputs packed two to a register
        if (la.pair && la.use_tma && cs == 16384) {
            launch_one<T, TI, true, GI, ACC, 16384, TH, GZ, true>(la);
            return;
        }
    }
    if (la.use_tma && cs == 16384) launch_one<T, TI, true, GI, ACC, 16384, TH, GZ>(la);
    else if (la.use_tma) launch_one<T, TI, true, GI, ACC, 0, TH, GZ>(la);
    else launch_one<T, TI, false, GI, ACC, 0, TH, GZ>(la);
}

template <typename T, typename TI, int TH>
static cudaError_t launch_bwd_th(const LaunchArgs& la) {
    const bool gi = la.grad_init != nullptr;
    if (la.gen_preact) {  // generator-tail training: the DEM is detached there (models/JSPSR.py:372), nothing accumulates
        if (gi || la.accumulate) return cudaErrorNotSupported;
        launch_variant<T, TI, false, false, TH, true>(la);
        return cudaGetLastError();
    }
    if (la.accumulate) {
        // only the fixed-affinity loop accumulates (NLSPN backward: grad_init always needed)
        if (!gi) return cudaErrorNotSupported;
        launch_variant<T, TI, true, true, TH>(la);
    } else if (gi) {
        launch_variant<T, TI, true, false, TH>(la);
    } else {
        launch_variant<T, TI, false, false, TH>(la);
    }
    return cudaGetLastError();
}

template <typename T, typename TI>
static cudaError_t launch_bwd_dtype(const LaunchArgs& la) {
    switch (la.tile_h) {
        case 8: return launch_bwd_th<T, TI, 8>(la);
        case 4: return launch_bwd_th<T, TI, 4>(la);
        case 2: return launch_bwd_th<T, TI, 2>(la);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_spn_backward(const LaunchArgs& la) {
    if (la.bf16 && la.init_f32) return launch_bwd_dtype<__nv_bfloat16, float>(la);
    return la.bf16 ? launch_bwd_dtype<__nv_bfloat16, __nv_bfloat16>(la) : launch_bwd_dtype<float, float>(la);
}

}  // namespace JSPSR_VARIANT
}  // namespace jspsr
