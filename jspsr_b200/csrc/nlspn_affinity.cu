// NLSPN affinity front-end after conv_offset_aff, one kernel each way.
// Forward replaces models/components/nlspn.py:82-175: chunk/cat/view/insert of the
// offsets, tanh/gamma scaling, eight separate 1x1 deform_conv2d confidence gathers,
// the multiply, abs-sum normalisation with its floor at 1, and the centre weight
// (about 30 launches in the reference).  Backward is its autograd.
//
// Same machinery as the propagation kernels: one CTA per TH x 128 pixel block, the
// confidence map staged in shared memory by TMA (zero fill = torchvision's zero-outside rule),
// branch-free taps with a deferred bounds-checked slow pass, 24 streamed channels in and
// 27 out per pixel; the backward scatters grad_confidence into a shared accumulation tile
// that is flushed with vector REDs.
#include "spn_kernels.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

enum { AFF_AS = 0, AFF_ASS = 1, AFF_TC = 2, AFF_TGASS = 3 };

constexpr int AFF_MIN_BLOCKS = 3;
constexpr int AFF_BWD_MIN_BLOCKS = 2;  // 128 registers: at 80 the backward spilled (STL/LDL were its top stall)

// tanh(x/100) / gamma' with the reference's operation order; the two divisions become multiplications
// by correctly rounded reciprocals (<= 1 ulp each)
__device__ __forceinline__ float aff_tanh(float af) { return tanhf(af * 0.01f); }

// position of the confidence sample of neighbour n (1x1 deformable gather, pad 0): pixel + offset
// (nlspn.py:130-139); tap index idx = n with the centre skipped
__device__ __forceinline__ int tap_index(int n) { return n < 4 ? n : n + 1; }

template <typename T, bool CONF, bool TMA, int TH>
__global__ void __launch_bounds__(THREADS, AFF_MIN_BLOCKS)
nlspn_affinity_fwd_kernel(const T* __restrict__ conv_out, const T* __restrict__ conf, const float* __restrict__ gamma_p,
                          T* __restrict__ offset_out, T* __restrict__ aff_out, const Geom g, const int affinity,
                          const int legacy, const __grid_constant__ CUtensorMap tmap) {
    constexpr int SH = staged_rows(TH);
    constexpr int PPT = pixels_per_thread(TH);
    __shared__ __align__(128) T tile[CONF ? SH * SW : 8];
    __shared__ __align__(8) uint64_t bar;

    const TileCtx c = make_tile_ctx<TH>(g);
    if (CONF) stage_tile_begin<T, TMA, TH>(tile, &bar, &tmap, conf, g, c.b, c.ox, c.oy - g.init_row0);

    const size_t cs = (size_t)g.H * g.W;
    const T* cv_b = conv_out + (size_t)c.b * 24 * cs;
    T* oo_b = offset_out + (size_t)c.b * 18 * cs;
    T* ao_b = aff_out + (size_t)c.b * 9 * cs;
    const T* conf_b = CONF ? conf + (size_t)c.b * cs : nullptr;
    const T* tile_lo = tile + (CONF ? c.r_lo * SW : 0);
    const bool tanh_type = affinity == AFF_TC || affinity == AFF_TGASS;
    const float gamma = gamma_p[0];
    const float rgam = __fdiv_rn(1.f, affinity == AFF_TGASS ? gamma + 1e-8f : gamma);

    float oh[8], ow[8], u[8];
    auto load_inputs = [&](int it, bool& active, size_t& p) {
        const int y = c.y0 + pix_row<TH, true>(it), x = c.x0 + pix_col<TH, true>(it);
        active = (y < g.H) && (x < g.W);
        p = (size_t)y * g.W + x;
        if (active) {
            const T* cv = cv_b + p;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                // cat(o1,o2).view(B,8,2,H,W): pair n = channels (2n, 2n+1)  (nlspn.py:85)
                oh[n] = ld_stream(cv + (2 * n) * cs);
                ow[n] = ld_stream(cv + (2 * n + 1) * cs);
                u[n] = ld_stream(cv + (16 + n) * cs);
            }
        }
    };

    bool active;
    size_t p;
    load_inputs(0, active, p);
    if (CONF) stage_tile_wait<TMA>(&bar);

#pragma unroll 1
    for (int it = 0; it < PPT; ++it) {
        __syncwarp();  // reconverge the lanes that took the out-of-tile path on the previous pixel (see spn_forward.cu)
        if (it > 0) load_inputs(it, active, p);
        if (!active) continue;
        const float fy = (float)(c.y0 + pix_row<TH, true>(it)), fx = (float)(c.x0 + pix_col<TH, true>(it));
        T* oo = oo_b + p;
        unsigned slow = 0u;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int idx = tap_index(n);
            if (legacy) {  // nlspn.py:122-128 shifts the shared storage in place
                oh[n] += (float)(idx / 3 - 1);
                ow[n] += (float)(idx % 3 - 1);
            }
            st_stream(oo + (2 * idx) * cs, oh[n]);
            st_stream(oo + (2 * idx + 1) * cs, ow[n]);
            if (tanh_type) u[n] = aff_tanh(u[n]) * rgam;
            if (CONF) {
                const FastTap t = fast_tap<T>(tile_lo, c, fy + oh[n], fx + ow[n]);
                const float val = bilerp(t.v1, t.v2, t.v3, t.v4, t.lh, t.lw);
                u[n] = t.ok ? u[n] * val : u[n];
                slow |= t.ok ? 0u : (1u << n);
            }
        }
        st_stream(oo + 8 * cs, 0.f);  // zero reference offset at idx_ref = 4 (nlspn.py:86-90)
        st_stream(oo + 9 * cs, 0.f);
        if (CONF && slow) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                if (slow & (1u << n)) {
                    const SlowTap t = slow_tap<T>(conf_b, g, fy + oh[n], fx + ow[n], nullptr);
                    u[n] *= bilerp(t.v1, t.v2, t.v3, t.v4, t.lh, t.lw);
                }
            }
        }
        // abs-sum normalisation with its floor, centre weight (nlspn.py:159-173)
        float sabs = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) sabs += fabsf(u[n]);
        sabs += 1e-4f;
        if (affinity == AFF_ASS || affinity == AFF_TGASS) sabs = sabs < 1.f ? 1.f : sabs;
        const float inv = affinity == AFF_TC ? 1.f : __fdiv_rn(1.f, sabs);
        float sum = 0.f;
        T* ao = ao_b + p;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float v = u[n] * inv;
            sum += v;
            st_stream(ao + tap_index(n) * cs, v);
        }
        st_stream(ao + 4 * cs, 1.f - sum);
    }
}

// CS: compile-time channel stride (0 = runtime).  ITILE: grad_confidence is scattered into a block-floating-point
// tile of native integer atomics (spn_backward.cu); its scale comes from a pre-pass BOUND on the CTA's sum of
// |contributions|: |dL/d(sample_n)| <= 2 * max_m|g_m| * inv * |t_n| with inv <= 1 for the clamped (ASS, TGASS) and the
// unnormalised (TC) affinities; AS (unclamped) has no such bound without the gathers and keeps the fp32 tile.
// The bound can overestimate the true sum by orders of magnitude (large raw affinities), so the tile is the
// two-level form (gi_add2: coarse + residual tile in dynamic shared memory, 45 bits in total).
template <typename T, bool CONF, bool TMA, int TH, int CS, bool ITILE>
__global__ void __launch_bounds__(THREADS, AFF_BWD_MIN_BLOCKS)
nlspn_affinity_bwd_kernel(const T* __restrict__ grad_offset, const T* __restrict__ grad_aff,
                          const T* __restrict__ conv_out, const T* __restrict__ conf, const float* __restrict__ gamma_p,
                          T* __restrict__ grad_conv, float* __restrict__ grad_conf, float* __restrict__ grad_scale,
                          ReduceWs* __restrict__ ws, const Geom g, const int affinity,
                          const __grid_constant__ CUtensorMap tmap) {
    constexpr int SH = staged_rows(TH);
    constexpr int PPT = pixels_per_thread(TH);
    __shared__ __align__(128) T tile[CONF ? SH * SW : 8];
    __shared__ __align__(16) float gtile[CONF ? SH * SW : 4];  // fp32, or int32 bit patterns when ITILE
    extern __shared__ __align__(16) int gtile_res[];           // ITILE: residual tile [SH * SW]
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_red[WARPS];
    __shared__ GiScale s_gis;
    __shared__ bool s_last;

    const TileCtx c = make_tile_ctx<TH>(g);
    if (CONF) {
        stage_tile_begin<T, TMA, TH>(tile, &bar, &tmap, conf, g, c.b, c.ox, c.oy - g.init_row0);
        for (int i = threadIdx.x; i < SH * SW; i += THREADS) {
            gtile[i] = 0.f;  // all-zero bits in both formats
            if (ITILE) gtile_res[i] = 0;
        }
    }
    const bool scatter = CONF && grad_conf != nullptr;

    const size_t cs = CS ? (size_t)CS : (size_t)g.H * g.W;
    const T* cv_b = conv_out + (size_t)c.b * 24 * cs;
    const T* go_b = grad_offset + (size_t)c.b * 18 * cs;
    const T* ga_b = grad_aff + (size_t)c.b * 9 * cs;
    T* gc_b = grad_conv + (size_t)c.b * 24 * cs;
    const T* conf_b = CONF ? conf + (size_t)c.b * cs : nullptr;
    float* gcf_b = scatter ? grad_conf + (size_t)c.b * cs : nullptr;
    const T* tile_lo = tile + (CONF ? c.r_lo * SW : 0);
    float* gtile_lo = gtile + (CONF ? c.r_lo * SW : 0);
    const bool tanh_type = affinity == AFF_TC || affinity == AFF_TGASS;
    const bool normalised = affinity != AFF_TC;
    const float gamma = gamma_p[0];
    const float rgam = __fdiv_rn(1.f, affinity == AFF_TGASS ? gamma + 1e-8f : gamma);
    float acc_gamma = 0.f;

    float oh[8], ow[8], af[8], gn[8];
    auto load_inputs = [&](int it, bool& active, size_t& p) {
        const int y = c.y0 + pix_row<TH, true>(it), x = c.x0 + pix_col<TH, true>(it);
        active = (y < g.H) && (x < g.W);
        p = (size_t)y * g.W + x;
        if (active) {
            const T* cv = cv_b + p;
            const T* ga = ga_b + p;
            const float gcen = ld_stream(ga + 4 * cs);
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                oh[n] = ld_stream(cv + (2 * n) * cs);
                ow[n] = ld_stream(cv + (2 * n + 1) * cs);
                af[n] = ld_stream(cv + (16 + n) * cs);
                gn[n] = ld_stream(ga + tap_index(n) * cs) - gcen;  // centre = 1 - sum
            }
        }
    };

    bool active;
    size_t p;
    float gscale = 0.f;
    if (CONF && ITILE) {
        // pre-pass: S >= sum over this CTA of |dL/d(confidence sample)| from the affinity gradients alone
        float part = 0.f;
        if (scatter) {
#pragma unroll 1
            for (int it = 0; it < PPT; ++it) {
                const int y = c.y0 + pix_row<TH, true>(it), x = c.x0 + pix_col<TH, true>(it);
                if (y < g.H && x < g.W) {
                    const size_t q = (size_t)y * g.W + x;
                    const float gcen = to_f32(ga_b[q + 4 * cs]);
                    float gmax = 0.f, tsum = 0.f;
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        gmax = fmaxf(gmax, fabsf(to_f32(ga_b[q + tap_index(n) * cs]) - gcen));
                        // |t_n|: tanh types are bounded by 1 / gamma', the raw types are the convolution output
                        tsum += tanh_type ? fabsf(rgam) : fabsf(to_f32(cv_b[q + (16 + n) * cs]));
                    }
                    part = fmaf(2.f * gmax, tsum, part);
                }
            }
        }
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
        __syncthreads();
        if (threadIdx.x == 0) {
            float S = 0.f;
#pragma unroll
            for (int i = 0; i < WARPS; ++i) S += s_red[i];
            s_gis = gi_scale_from_sum(S);
        }
    }
    load_inputs(0, active, p);
    if (CONF) stage_tile_wait<TMA>(&bar);
    else __syncthreads();
    if (CONF && ITILE) gscale = s_gis.scale;  // published before the barrier above

#pragma unroll 1
    for (int it = 0; it < PPT; ++it) {
        __syncwarp();  // reconverge the lanes that took the out-of-tile path on the previous pixel (see spn_forward.cu)
        if (it > 0) load_inputs(it, active, p);
        if (!active) continue;
        const float fy = (float)(c.y0 + pix_row<TH, true>(it)), fx = (float)(c.x0 + pix_col<TH, true>(it));
        const T* go = go_b + p;
        T* gc = gc_b + p;
        // offsets feed the propagation directly; the confidence gather detaches them (nlspn.py:118)
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int idx = tap_index(n);
            st_stream(gc + (2 * n) * cs, ld_stream(go + (2 * idx) * cs));
            st_stream(gc + (2 * n + 1) * cs, ld_stream(go + (2 * idx + 1) * cs));
        }
        // recompute t_n = scaled affinity, cval_n = gathered confidence
        float th[8], cval[8];
        unsigned slow = 0u;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            th[n] = tanh_type ? aff_tanh(af[n]) : 0.f;
            af[n] = tanh_type ? th[n] * rgam : af[n];  // af[] now holds t_n
            cval[n] = 1.f;
            if (CONF) {
                const FastTap t = fast_tap<T, true>(tile_lo, c, fy + oh[n], fx + ow[n]);
                cval[n] = t.ok ? bilerp(t.v1, t.v2, t.v3, t.v4, t.lh, t.lw) : 0.f;
                slow |= t.ok ? 0u : (1u << n);
            }
        }
        if (CONF && slow) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                if (slow & (1u << n)) {
                    const SlowTap t = slow_tap<T>(conf_b, g, fy + oh[n], fx + ow[n], nullptr);
                    cval[n] = bilerp(t.v1, t.v2, t.v3, t.v4, t.lh, t.lw);
                }
            }
        }
        float sabs = 0.f, dot = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float u = af[n] * cval[n];
            sabs += fabsf(u);
            dot = fmaf(gn[n], u, dot);
        }
        sabs += 1e-4f;
        bool clamped = false;
        if (affinity == AFF_ASS || affinity == AFF_TGASS) {
            clamped = sabs < 1.f;
            if (clamped) sabs = 1.f;
        }
        const float inv = normalised ? __fdiv_rn(1.f, sabs) : 1.f;
        const float corr = (normalised && !clamped) ? dot * inv * inv : 0.f;  // d(sum|u|)/du term
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float u = af[n] * cval[n];
            const float sgn = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f);
            const float gu = gn[n] * inv - sgn * corr;   // dL/du_n
            const float gt = gu * cval[n];               // dL/dt_n
            if (scatter) {
                const float gv = gu * af[n];             // dL/d(confidence sample)
                const float h = fy + oh[n], w = fx + ow[n];
                if (slow & (1u << n)) {
                    const SlowTap t = slow_tap<T>(conf_b, g, h, w, nullptr);
                    // torchvision's forward returns 0 for a sample wholly outside (-1,H)x(-1,W): no gradient there
                    const bool inside = t.finite && h > -1.f && h < (float)g.H_img && w > -1.f && w < (float)g.W;
                    if (inside && gv != 0.f) {
                        const float ch = gv * t.lh, cl = gv - ch, c2 = cl * t.lw, c4 = ch * t.lw;
                        const int r0 = t.h0, q0 = t.w0;
                        if ((unsigned)r0 < (unsigned)g.H_img && (unsigned)q0 < (unsigned)g.W) atomicAdd(gcf_b + (size_t)r0 * g.W + q0, cl - c2);
                        if ((unsigned)r0 < (unsigned)g.H_img && (unsigned)(q0 + 1) < (unsigned)g.W) atomicAdd(gcf_b + (size_t)r0 * g.W + q0 + 1, c2);
                        if ((unsigned)(r0 + 1) < (unsigned)g.H_img && (unsigned)q0 < (unsigned)g.W) atomicAdd(gcf_b + (size_t)(r0 + 1) * g.W + q0, ch - c4);
                        if ((unsigned)(r0 + 1) < (unsigned)g.H_img && (unsigned)(q0 + 1) < (unsigned)g.W) atomicAdd(gcf_b + (size_t)(r0 + 1) * g.W + q0 + 1, c4);
                    }
                } else {
                    const FastTap t = fast_tap<T, true>(tile_lo, c, h, w);  // geometry only (values unused)
                    const float gvs = ITILE ? gv * gscale : gv;       // exact: the scale is a power of two
                    const float ch = gvs * t.lh, cl = gvs - ch, c2 = cl * t.lw, c4 = ch * t.lw;
                    float* gtp = gtile_lo + ((unsigned)t.h0 - c.oy_lo) * SW + ((unsigned)t.w0 - (unsigned)c.ox);
                    if (ITILE) {
                        int* gip = reinterpret_cast<int*>(gtp);
                        int* grp = gtile_res + (gtp - gtile);
                        gi_add2(gip, grp, cl - c2);
                        gi_add2(gip + 1, grp + 1, c2);
                        gi_add2(gip + SW, grp + SW, ch - c4);
                        gi_add2(gip + SW + 1, grp + SW + 1, c4);
                    } else {
                        atomicAdd(gtp, cl - c2);
                        atomicAdd(gtp + 1, c2);
                        atomicAdd(gtp + SW, ch - c4);
                        atomicAdd(gtp + SW + 1, c4);
                    }
                }
            }
            float gaf = gt;
            if (tanh_type) {
                gaf = gt * rgam * (1.f - th[n] * th[n]) * 0.01f;
                if (affinity == AFF_TGASS) acc_gamma -= gt * th[n] * rgam * rgam;
            }
            st_stream(gc + (16 + n) * cs, gaf);
        }
    }

    if (scatter) {  // flush the accumulation tile: contributions to cells outside the image are dropped
        __syncthreads();
        const bool vec_ok = (g.W & 3) == 0 && ((reinterpret_cast<uintptr_t>(grad_conf) & 15) == 0);
        const bool poison = ITILE && s_gis.poison;
        const float ginv = poison ? __int_as_float(0x7fc00000) : s_gis.inv;
        for (int i = threadIdx.x; i < SH * (SW / 4); i += THREADS) {
            const int r = i / (SW / 4), q = (i - r * (SW / 4)) * 4;
            float4 v = *reinterpret_cast<const float4*>(gtile + r * SW + q);
            if (ITILE) {
                const int4 iv = *reinterpret_cast<const int4*>(gtile + r * SW + q);
                const int4 ir = *reinterpret_cast<const int4*>(gtile_res + r * SW + q);
                if ((iv.x | iv.y | iv.z | iv.w | ir.x | ir.y | ir.z | ir.w) == 0 && !poison) continue;
                const float k = 1.f / 65536.f;
                v = make_float4(fmaf((float)ir.x, k, (float)iv.x) * ginv, fmaf((float)ir.y, k, (float)iv.y) * ginv,
                                fmaf((float)ir.z, k, (float)iv.z) * ginv, fmaf((float)ir.w, k, (float)iv.w) * ginv);
            } else if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) {
                continue;
            }
            const int gy = c.oy + r, gx = c.ox + q;
            if ((unsigned)gy >= (unsigned)g.H_img) continue;
            float* dst = gcf_b + (size_t)gy * g.W + gx;
            if (vec_ok && gx >= 0 && gx + 3 < g.W) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z),
                             "f"(v.w)
                             : "memory");
            } else {
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((unsigned)(gx + j) < (unsigned)g.W && vv[j] != 0.f) atomicAdd(dst + j, vv[j]);
            }
        }
    }

    if (affinity == AFF_TGASS && grad_scale != nullptr) {
        acc_gamma = warp_sum(acc_gamma);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc_gamma;
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < WARPS; ++i) v += s_red[i];
            atomicAdd(&ws->sums[0], (double)v);
            __threadfence();
            const unsigned tk = atomicAdd(&ws->ticket, 1u);
            s_last = (tk == gridDim.x - 1);
            if (s_last) {
                __threadfence();
                grad_scale[0] = (float)atomicAdd(&ws->sums[0], 0.0);
                ws->sums[0] = 0.0;
                ws->ticket = 0u;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
struct AffArgs {
    const void* conv_out; const void* conf; const float* gamma;
    void* offset_out; void* aff_out;
    const void* grad_offset; const void* grad_aff; void* grad_conv; float* grad_conf; float* grad_scale; void* ws;
    Geom g; int affinity; int legacy; bool use_tma; int tile_h; CUtensorMap tmap; cudaStream_t stream;
};

template <typename T, bool CONF, bool TMA, int TH>
static void aff_fwd_launch(const AffArgs& a) {
    dim3 grid((unsigned)((size_t)a.g.tiles_x * a.g.tiles_y * a.g.B));
    nlspn_affinity_fwd_kernel<T, CONF, TMA, TH><<<grid, THREADS, 0, a.stream>>>(
        (const T*)a.conv_out, (const T*)a.conf, a.gamma, (T*)a.offset_out, (T*)a.aff_out, a.g, a.affinity, a.legacy, a.tmap);
}
static thread_local cudaError_t g_aff_attr_error = cudaSuccess;  // set by a launcher of this call, read by aff_dispatch

template <typename T, bool CONF, bool TMA, int TH, int CS, bool ITILE>
static void aff_bwd_launch_one(const AffArgs& a) {
    dim3 grid((unsigned)((size_t)a.g.tiles_x * a.g.tiles_y * a.g.B));
    const size_t dyn = ITILE ? (size_t)staged_rows(TH) * SW * sizeof(int) : 0;
    if (dyn > 0) {  // static + dynamic shared memory exceed 48 KB with the wide halo: opt in once per instantiation
        const cudaError_t e = ensure_dynamic_smem((const void*)nlspn_affinity_bwd_kernel<T, CONF, TMA, TH, CS, ITILE>, dyn);
        if (e != cudaSuccess) {  // do not launch a kernel that cannot get its shared memory: report the attribute error
            g_aff_attr_error = e;
            return;
        }
    }
    nlspn_affinity_bwd_kernel<T, CONF, TMA, TH, CS, ITILE><<<grid, THREADS, dyn, a.stream>>>(
        (const T*)a.grad_offset, (const T*)a.grad_aff, (const T*)a.conv_out, (const T*)a.conf, a.gamma, (T*)a.grad_conv,
        a.grad_conf, a.grad_scale, (ReduceWs*)a.ws, a.g, a.affinity, a.tmap);
}
template <typename T, bool CONF, bool TMA, int TH>
static void aff_bwd_launch(const AffArgs& a) {
    const bool cs128 = TMA && (size_t)a.g.H * a.g.W == 16384;  // the reference's 128x128 planes
    const bool itile = CONF && a.affinity != AFF_AS;
    if (cs128) {
        if (itile) aff_bwd_launch_one<T, CONF, TMA, TH, 16384, CONF>(a);
        else aff_bwd_launch_one<T, CONF, TMA, TH, 16384, false>(a);
    } else {
        if (itile) aff_bwd_launch_one<T, CONF, TMA, TH, 0, CONF>(a);
        else aff_bwd_launch_one<T, CONF, TMA, TH, 0, false>(a);
    }
}

template <typename T, bool FWD, bool CONF, bool TMA, int TH>
static void aff_launch(const AffArgs& a) {
    if (FWD) aff_fwd_launch<T, CONF, TMA, TH>(a);
    else aff_bwd_launch<T, CONF, TMA, TH>(a);
}

template <typename T, bool FWD, int TH>
static void aff_dispatch_th(const AffArgs& a) {
    if (a.conf == nullptr) aff_launch<T, FWD, false, false, TH>(a);
    else if (a.use_tma) aff_launch<T, FWD, true, true, TH>(a);
    else aff_launch<T, FWD, true, false, TH>(a);
}

template <typename T, bool FWD>
static cudaError_t aff_dispatch(const AffArgs& a) {
    g_aff_attr_error = cudaSuccess;
    if (a.tile_h == 8) aff_dispatch_th<T, FWD, 8>(a);
    else aff_dispatch_th<T, FWD, 2>(a);
    if (g_aff_attr_error != cudaSuccess) return g_aff_attr_error;
    return cudaGetLastError();
}

cudaError_t launch_nlspn_affinity(const void* conv_out, const void* confidence, const float* gamma, void* offset_out,
                                  void* aff_out, const void* grad_offset, const void* grad_aff, void* grad_conv_out,
                                  float* grad_confidence, float* grad_scale, void* workspace, const Geom& g, int tile_h,
                                  bool use_tma, const CUtensorMap& tmap, int affinity, int legacy, bool bf16,
                                  bool forward, cudaStream_t stream) {
    AffArgs a;
    a.conv_out = conv_out; a.conf = confidence; a.gamma = gamma; a.offset_out = offset_out; a.aff_out = aff_out;
    a.grad_offset = grad_offset; a.grad_aff = grad_aff; a.grad_conv = grad_conv_out; a.grad_conf = grad_confidence;
    a.grad_scale = grad_scale; a.ws = workspace; a.g = g; a.affinity = affinity; a.legacy = legacy;
    a.use_tma = use_tma; a.tile_h = tile_h; a.tmap = tmap; a.stream = stream;
    if (forward) return bf16 ? aff_dispatch<__nv_bfloat16, true>(a) : aff_dispatch<float, true>(a);
    return bf16 ? aff_dispatch<__nv_bfloat16, false>(a) : aff_dispatch<float, false>(a);
}

}  // namespace JSPSR_VARIANT
}  // namespace jspsr
