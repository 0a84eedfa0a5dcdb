// NLSPN affinity front-end after conv_offset_aff, one kernel each way.
// Forward replaces models/components/nlspn.py:82-175: chunk/cat/view/insert of the
// offsets, tanh/gamma scaling, eight separate 1x1 deform_conv2d confidence gathers,
// the multiply, abs-sum normalisation with its floor at 1, and the centre weight
// (about 30 launches in the reference).  Backward is its autograd.
#include "spn_kernels.cuh"

namespace jspsr {

enum { AFF_AS = 0, AFF_ASS = 1, AFF_TC = 2, AFF_TGASS = 3 };

struct Bil {
    float v1, v2, v3, v4, lh, lw;
    int h0, w0;
    bool inside;
};

// torchvision bilinear_interpolate on a [H,W] plane (1x1 deformable gather, pad 0)
template <typename T>
__device__ __forceinline__ Bil bilinear_at(const T* __restrict__ img, int H, int W, float h, float w) {
    Bil r;
    r.v1 = r.v2 = r.v3 = r.v4 = 0.f;
    r.lh = r.lw = 0.f;
    r.h0 = r.w0 = 0;
    r.inside = !((h <= -1.f) || (h >= (float)H) || (w <= -1.f) || (w >= (float)W)) && (h == h) && (w == w);
    if (h != h || w != w) { r.lh = h - h; r.lw = w - w; return r; }  // NaN stays NaN
    if (!r.inside) return r;
    r.h0 = __float2int_rd(h);
    r.w0 = __float2int_rd(w);
    r.lh = h - floorf(h);
    r.lw = w - floorf(w);
    const int h1 = r.h0 + 1, w1 = r.w0 + 1;
    if (r.h0 >= 0 && r.w0 >= 0) r.v1 = to_f32(img[(size_t)r.h0 * W + r.w0]);
    if (r.h0 >= 0 && w1 <= W - 1) r.v2 = to_f32(img[(size_t)r.h0 * W + w1]);
    if (h1 <= H - 1 && r.w0 >= 0) r.v3 = to_f32(img[(size_t)h1 * W + r.w0]);
    if (h1 <= H - 1 && w1 <= W - 1) r.v4 = to_f32(img[(size_t)h1 * W + w1]);
    return r;
}
__device__ __forceinline__ float bil_value(const Bil& r) {
    const float hh = 1.f - r.lh, hw = 1.f - r.lw;
    return hh * hw * r.v1 + hh * r.lw * r.v2 + r.lh * hw * r.v3 + r.lh * r.lw * r.v4;
}

template <int AFF>
__device__ __forceinline__ float scale_aff(float af, float gamma, float& th) {
    if (AFF == AFF_TC) {
        th = tanhf(__fdiv_rn(af, 100.f));
        return __fdiv_rn(th, gamma);
    } else if (AFF == AFF_TGASS) {
        th = tanhf(__fdiv_rn(af, 100.f));
        return __fdiv_rn(th, gamma + 1e-8f);
    }
    th = 0.f;
    return af;
}

template <typename T, int AFF, bool CONF>
__global__ void __launch_bounds__(256)
nlspn_affinity_fwd_kernel(const T* __restrict__ conv_out, const T* __restrict__ conf, const float* __restrict__ gamma_p,
                          T* __restrict__ offset_out, T* __restrict__ aff_out, int B, int H, int W, int legacy) {
    const size_t cs = (size_t)H * W, total = (size_t)B * cs;
    const float gamma = gamma_p[0];
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const size_t b = p / cs, yx = p - b * cs;
        const int y = (int)(yx / W), x = (int)(yx - (size_t)y * W);
        const T* cv = conv_out + b * 24 * cs + yx;
        T* oo = offset_out + b * 18 * cs + yx;
        T* ao = aff_out + b * 9 * cs + yx;
        const T* cf = CONF ? conf + b * cs : nullptr;
        float u[8], sabs = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int idx = n < 4 ? n : n + 1;  // tap index with the centre skipped
            // cat(o1,o2).view(B,8,2,H,W): pair n = channels (2n, 2n+1)  (nlspn.py:85)
            float oh = ld_stream(cv + (2 * n) * cs), ow = ld_stream(cv + (2 * n + 1) * cs);
            if (legacy) {  // nlspn.py:122-128 shifts the shared storage in place
                oh += (float)(idx / 3 - 1);
                ow += (float)(idx % 3 - 1);
            }
            st_stream(oo + (2 * idx) * cs, oh);
            st_stream(oo + (2 * idx + 1) * cs, ow);
            float th;
            float t = scale_aff<AFF>(ld_stream(cv + (16 + n) * cs), gamma, th);
            if (CONF) {
                const Bil r = bilinear_at<T>(cf, H, W, (float)y + oh, (float)x + ow);
                t *= bil_value(r);
            }
            u[n] = t;
            sabs += fabsf(t);
        }
        st_stream(oo + 8 * cs, 0.f);
        st_stream(oo + 9 * cs, 0.f);
        sabs += 1e-4f;
        if (AFF == AFF_ASS || AFF == AFF_TGASS) sabs = sabs < 1.f ? 1.f : sabs;
        float sum = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            if (AFF != AFF_TC) u[n] = __fdiv_rn(u[n], sabs);
            sum += u[n];
            st_stream(ao + (n < 4 ? n : n + 1) * cs, u[n]);
        }
        st_stream(ao + 4 * cs, 1.f - sum);
    }
}

template <typename T, int AFF, bool CONF>
__global__ void __launch_bounds__(256)
nlspn_affinity_bwd_kernel(const T* __restrict__ grad_offset, const T* __restrict__ grad_aff,
                          const T* __restrict__ conv_out, const T* __restrict__ conf, const float* __restrict__ gamma_p,
                          T* __restrict__ grad_conv, float* __restrict__ grad_conf, float* __restrict__ grad_scale,
                          ReduceWs* __restrict__ ws, int B, int H, int W) {
    __shared__ float s_red[8];
    __shared__ bool s_last;
    const size_t cs = (size_t)H * W, total = (size_t)B * cs;
    const float gamma = gamma_p[0];
    const float geff = AFF == AFF_TGASS ? gamma + 1e-8f : gamma;
    float acc_gamma = 0.f;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
        const size_t b = p / cs, yx = p - b * cs;
        const int y = (int)(yx / W), x = (int)(yx - (size_t)y * W);
        const T* cv = conv_out + b * 24 * cs + yx;
        const T* go = grad_offset + b * 18 * cs + yx;
        const T* ga = grad_aff + b * 9 * cs + yx;
        T* gc = grad_conv + b * 24 * cs + yx;
        const T* cf = CONF ? conf + b * cs : nullptr;
        float* gcf = (CONF && grad_conf) ? grad_conf + b * cs : nullptr;

        float t[8], th[8], cval[8], u[8], sabs = 0.f;
        Bil bil[CONF ? 8 : 1];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const int idx = n < 4 ? n : n + 1;
            // offsets feed the propagation directly; the confidence gather detaches them
            st_stream(gc + (2 * n) * cs, ld_stream(go + (2 * idx) * cs));
            st_stream(gc + (2 * n + 1) * cs, ld_stream(go + (2 * idx + 1) * cs));
            t[n] = scale_aff<AFF>(ld_stream(cv + (16 + n) * cs), gamma, th[n]);
            cval[n] = 1.f;
            if (CONF) {
                const float oh = ld_stream(cv + (2 * n) * cs), ow = ld_stream(cv + (2 * n + 1) * cs);
                bil[n] = bilinear_at<T>(cf, H, W, (float)y + oh, (float)x + ow);
                cval[n] = bil_value(bil[n]);
            }
            u[n] = t[n] * cval[n];
            sabs += fabsf(u[n]);
        }
        sabs += 1e-4f;
        bool clamped = false;
        if (AFF == AFF_ASS || AFF == AFF_TGASS) {
            clamped = sabs < 1.f;
            if (clamped) sabs = 1.f;
        }
        const float gcen = ld_stream(ga + 4 * cs);
        float gn[8], dot = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            gn[n] = ld_stream(ga + (n < 4 ? n : n + 1) * cs) - gcen;  // centre = 1 - sum
            dot += gn[n] * u[n];
        }
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            float gu = gn[n];
            if (AFF != AFF_TC) {
                gu = __fdiv_rn(gn[n], sabs);
                if (!clamped) {
                    const float sgn = u[n] > 0.f ? 1.f : (u[n] < 0.f ? -1.f : 0.f);
                    gu -= sgn * __fdiv_rn(dot, sabs * sabs);
                }
            }
            float gt = gu * cval[n];
            if (CONF && gcf) {
                const float gv = gu * t[n];  // dL/d conf_sample
                const Bil& r = bil[n];
                if (r.inside && gv != 0.f) {
                    const float hh = 1.f - r.lh, hw = 1.f - r.lw;
                    const int h1 = r.h0 + 1, w1 = r.w0 + 1;
                    if (r.h0 >= 0 && r.w0 >= 0) atomicAdd(gcf + (size_t)r.h0 * W + r.w0, gv * hh * hw);
                    if (r.h0 >= 0 && w1 <= W - 1) atomicAdd(gcf + (size_t)r.h0 * W + w1, gv * hh * r.lw);
                    if (h1 <= H - 1 && r.w0 >= 0) atomicAdd(gcf + (size_t)h1 * W + r.w0, gv * r.lh * hw);
                    if (h1 <= H - 1 && w1 <= W - 1) atomicAdd(gcf + (size_t)h1 * W + w1, gv * r.lh * r.lw);
                }
            }
            float gaf = gt;
            if (AFF == AFF_TC || AFF == AFF_TGASS) {
                gaf = __fdiv_rn(__fdiv_rn(gt, geff) * (1.f - th[n] * th[n]), 100.f);
                if (AFF == AFF_TGASS) acc_gamma -= gt * __fdiv_rn(th[n], geff * geff);
            }
            st_stream(gc + (16 + n) * cs, gaf);
        }
    }
    if (AFF == AFF_TGASS && grad_scale != nullptr) {
        acc_gamma = warp_sum(acc_gamma);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc_gamma;
        __syncthreads();
        if (threadIdx.x == 0) {
            float v = 0.f;
            for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += s_red[i];
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

static int grid_for(size_t total) { return (int)min((size_t)148 * 8, (total + 255) / 256); }

template <typename T, int AFF>
static void fwd_aff(const void* conv_out, const void* conf, const float* gamma, void* off, void* aff, int B, int H,
                    int W, int legacy, cudaStream_t st) {
    const int grid = grid_for((size_t)B * H * W);
    if (conf)
        nlspn_affinity_fwd_kernel<T, AFF, true><<<grid, 256, 0, st>>>((const T*)conv_out, (const T*)conf, gamma, (T*)off,
                                                                      (T*)aff, B, H, W, legacy);
    else
        nlspn_affinity_fwd_kernel<T, AFF, false><<<grid, 256, 0, st>>>((const T*)conv_out, nullptr, gamma, (T*)off,
                                                                       (T*)aff, B, H, W, 0);
}
template <typename T>
static void fwd_dtype(const void* conv_out, const void* conf, const float* gamma, void* off, void* aff, int B, int H,
                      int W, int affinity, int legacy, cudaStream_t st) {
    switch (affinity) {
        case AFF_AS: fwd_aff<T, AFF_AS>(conv_out, conf, gamma, off, aff, B, H, W, legacy, st); break;
        case AFF_ASS: fwd_aff<T, AFF_ASS>(conv_out, conf, gamma, off, aff, B, H, W, legacy, st); break;
        case AFF_TC: fwd_aff<T, AFF_TC>(conv_out, conf, gamma, off, aff, B, H, W, legacy, st); break;
        default: fwd_aff<T, AFF_TGASS>(conv_out, conf, gamma, off, aff, B, H, W, legacy, st); break;
    }
}

cudaError_t launch_nlspn_affinity_forward(const void* conv_out, const void* confidence, const float* gamma,
                                          void* offset_out, void* aff_out, int B, int H, int W, int affinity, int legacy,
                                          bool bf16, cudaStream_t stream) {
    if (bf16) fwd_dtype<__nv_bfloat16>(conv_out, confidence, gamma, offset_out, aff_out, B, H, W, affinity, legacy, stream);
    else fwd_dtype<float>(conv_out, confidence, gamma, offset_out, aff_out, B, H, W, affinity, legacy, stream);
    return cudaGetLastError();
}

template <typename T, int AFF>
static void bwd_aff(const void* go, const void* ga, const void* conv_out, const void* conf, const float* gamma, void* gc,
                    float* gconf, float* gscale, void* ws, int B, int H, int W, cudaStream_t st) {
    const int grid = grid_for((size_t)B * H * W);
    if (conf)
        nlspn_affinity_bwd_kernel<T, AFF, true><<<grid, 256, 0, st>>>((const T*)go, (const T*)ga, (const T*)conv_out,
                                                                      (const T*)conf, gamma, (T*)gc, gconf, gscale,
                                                                      (ReduceWs*)ws, B, H, W);
    else
        nlspn_affinity_bwd_kernel<T, AFF, false><<<grid, 256, 0, st>>>((const T*)go, (const T*)ga, (const T*)conv_out,
                                                                       nullptr, gamma, (T*)gc, nullptr, gscale,
                                                                       (ReduceWs*)ws, B, H, W);
}
template <typename T>
static void bwd_dtype(const void* go, const void* ga, const void* conv_out, const void* conf, const float* gamma,
                      void* gc, float* gconf, float* gscale, void* ws, int B, int H, int W, int affinity,
                      cudaStream_t st) {
    switch (affinity) {
        case AFF_AS: bwd_aff<T, AFF_AS>(go, ga, conv_out, conf, gamma, gc, gconf, gscale, ws, B, H, W, st); break;
        case AFF_ASS: bwd_aff<T, AFF_ASS>(go, ga, conv_out, conf, gamma, gc, gconf, gscale, ws, B, H, W, st); break;
        case AFF_TC: bwd_aff<T, AFF_TC>(go, ga, conv_out, conf, gamma, gc, gconf, gscale, ws, B, H, W, st); break;
        default: bwd_aff<T, AFF_TGASS>(go, ga, conv_out, conf, gamma, gc, gconf, gscale, ws, B, H, W, st); break;
    }
}

cudaError_t launch_nlspn_affinity_backward(const void* grad_offset, const void* grad_aff, const void* conv_out,
                                           const void* confidence, const float* gamma, void* grad_conv_out,
                                           float* grad_confidence, float* grad_scale, void* workspace, int B, int H,
                                           int W, int affinity, bool bf16, cudaStream_t stream) {
    if (bf16)
        bwd_dtype<__nv_bfloat16>(grad_offset, grad_aff, conv_out, confidence, gamma, grad_conv_out, grad_confidence,
                                 grad_scale, workspace, B, H, W, affinity, stream);
    else
        bwd_dtype<float>(grad_offset, grad_aff, conv_out, confidence, gamma, grad_conv_out, grad_confidence, grad_scale,
                         workspace, B, H, W, affinity, stream);
    return cudaGetLastError();
}

}  // namespace jspsr
