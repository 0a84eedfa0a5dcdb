/*
 * CPU oracle, C restatement (TEST INFRASTRUCTURE ONLY - never linked into the
 * product library; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs load it).
 *
 * Restates, for a 1-channel map and a 3x3 window (stride 1, pad 1, dil 1):
 *   reference call sites  models/components/spn.py:99-118 (PostProcessor.forward)
 *                         models/LRRU.py:267-298 (Post_process_deconv.forward)
 *                         models/components/nlspn.py:177-187 (_propagate_once)
 *   third-party operator  torchvision.ops.deform_conv2d (reference pin 0.16;
 *                         not vendored in the reference tree): deformable_im2col
 *                         + bilinear_interpolate forward; deformable_col2im,
 *                         deformable_col2im_coord + get_coordinate_weight backward.
 *
 * Pinned by tests/test_oracle_c.py against the fixtures under tests/golden/
 * (outputs of the reference's own modules) and against oracle/spn_oracle.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 * -ffast-math must NOT be used: the fp32 rounding order is part of the contract.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KS 3
#define KK 9
enum { NORM_NONE = 0, NORM_RESIDUAL = 1, NORM_SUM = 2 };

int spn_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define DEFINE_ORACLE(REAL, SUFFIX, FLOORF)                                                        \
    typedef struct {                                                                               \
        REAL v1, v2, v3, v4, lh, lw;                                                               \
        long h0, w0;                                                                               \
    } corners_##SUFFIX;                                                                            \
                                                                                                   \
    /* four neighbours with per-corner validity (zero outside the image) */                        \
    static corners_##SUFFIX corners_at_##SUFFIX(const REAL *img, long H, long W, REAL h, REAL w) { \
        corners_##SUFFIX c;                                                                        \
        REAL h0f = FLOORF(h), w0f = FLOORF(w);                                                     \
        c.lh = h - h0f;                                                                            \
        c.lw = w - w0f;                                                                            \
        /* clamp before the integer cast: <= -2 or >= H has no valid corner */                     \
        if (!(h0f >= -2)) h0f = -2;                                                                \
        if (h0f > (REAL)H) h0f = (REAL)H;                                                          \
        if (!(w0f >= -2)) w0f = -2;                                                                \
        if (w0f > (REAL)W) w0f = (REAL)W;                                                          \
        long h0 = (long)h0f, w0 = (long)w0f, h1 = h0 + 1, w1 = w0 + 1;                             \
        c.h0 = h0;                                                                                 \
        c.w0 = w0;                                                                                 \
        c.v1 = (h0 >= 0 && h0 <= H - 1 && w0 >= 0 && w0 <= W - 1) ? img[h0 * W + w0] : (REAL)0;    \
        c.v2 = (h0 >= 0 && h0 <= H - 1 && w1 >= 0 && w1 <= W - 1) ? img[h0 * W + w1] : (REAL)0;    \
        c.v3 = (h1 >= 0 && h1 <= H - 1 && w0 >= 0 && w0 <= W - 1) ? img[h1 * W + w0] : (REAL)0;    \
        c.v4 = (h1 >= 0 && h1 <= H - 1 && w1 >= 0 && w1 <= W - 1) ? img[h1 * W + w1] : (REAL)0;    \
        return c;                                                                                  \
    }                                                                                              \
                                                                                                   \
    /* torchvision bilinear_interpolate */                                                         \
    static REAL bilinear_##SUFFIX(const corners_##SUFFIX *c, long H, long W, REAL h, REAL w) {     \
        if (h <= -1 || (REAL)H <= h || w <= -1 || (REAL)W <= w) return (REAL)0;                    \
        REAL hh = 1 - c->lh, hw = 1 - c->lw;                                                       \
        return hh * hw * c->v1 + hh * c->lw * c->v2 + c->lh * hw * c->v3 + c->lh * c->lw * c->v4;  \
    }                                                                                              \
                                                                                                   \
    static void normalise_##SUFFIX(const REAL *a, size_t cs, int mode, REAL *m) {                  \
        REAL s = 0;                                                                                \
        for (int k = 0; k < KK; ++k) s += a[k * cs];                                               \
        if (mode == NORM_RESIDUAL) {                                                               \
            REAL mean = s / (REAL)KK;                                                              \
            for (int k = 0; k < KK; ++k) m[k] = a[k * cs] - mean;                                  \
        } else if (mode == NORM_SUM) {                                                             \
            for (int k = 0; k < KK; ++k) m[k] = a[k * cs] / s;                                     \
        } else {                                                                                   \
            for (int k = 0; k < KK; ++k) m[k] = a[k * cs];                                         \
        }                                                                                          \
    }                                                                                              \
                                                                                                   \
    void spn_oracle_forward_##SUFFIX(const REAL *init, const REAL *weight, const REAL *offset,     \
                                     const REAL *w9, const REAL *b1, REAL *out, long B, long H,    \
                                     long W, int mode, REAL scale) {                               \
        const size_t cs = (size_t)H * W;                                                           \
        _Pragma("omp parallel for collapse(2) schedule(static)")                                   \
        for (long b = 0; b < B; ++b)                                                               \
            for (long y = 0; y < H; ++y) {                                                         \
                const REAL *img = init + b * cs;                                                   \
                for (long x = 0; x < W; ++x) {                                                     \
                    size_t p = (size_t)y * W + x;                                                  \
                    REAL m[KK];                                                                    \
                    normalise_##SUFFIX(weight + b * KK * cs + p, cs, mode, m);                     \
                    REAL acc = 0;                                                                  \
                    for (int k = 0; k < KK; ++k) {                                                 \
                        REAL h = (REAL)(y - 1 + k / KS) + offset[(b * 2 * KK + 2 * k) * cs + p];   \
                        REAL w = (REAL)(x - 1 + k % KS) + offset[(b * 2 * KK + 2 * k + 1) * cs + p]; \
                        corners_##SUFFIX c = corners_at_##SUFFIX(img, H, W, h, w);                 \
                        acc += w9[k] * (m[k] * bilinear_##SUFFIX(&c, H, W, h, w));                 \
                    }                                                                              \
                    acc += b1[0];                                                                  \
                    if (mode == NORM_RESIDUAL) acc += scale * img[p];                              \
                    out[b * cs + p] = acc;                                                         \
                }                                                                                  \
            }                                                                                      \
    }                                                                                              \
                                                                                                   \
    /* grad_init may be NULL (JSPSR detaches the DEM, models/JSPSR.py:372).  grad_w9 / grad_b1     \
       are written (not accumulated); they are summed in double. */                                \
    void spn_oracle_backward_##SUFFIX(const REAL *gout, const REAL *init, const REAL *weight,      \
                                      const REAL *offset, const REAL *w9, REAL *grad_init,         \
                                      REAL *grad_weight, REAL *grad_offset, REAL *grad_w9,         \
                                      REAL *grad_b1, long B, long H, long W, int mode,             \
                                      REAL scale) {                                                \
        const size_t cs = (size_t)H * W;                                                           \
        double gw_tot[KK] = {0}, gb_tot = 0;                                                       \
        if (grad_init)                                                                             \
            for (size_t i = 0; i < (size_t)B * cs; ++i) grad_init[i] = 0;                          \
        _Pragma("omp parallel")                                                                    \
        {                                                                                          \
            double gw_loc[KK] = {0}, gb_loc = 0;                                                   \
            _Pragma("omp for collapse(2) schedule(static)")                                        \
            for (long b = 0; b < B; ++b)                                                           \
                for (long y = 0; y < H; ++y) {                                                     \
                    const REAL *img = init + b * cs;                                               \
                    for (long x = 0; x < W; ++x) {                                                 \
                        size_t p = (size_t)y * W + x;                                              \
                        REAL g = gout[b * cs + p];                                                 \
                        REAL m[KK], gm[KK];                                                        \
                        const REAL *a = weight + b * KK * cs + p;                                  \
                        normalise_##SUFFIX(a, cs, mode, m);                                        \
                        gb_loc += g;                                                               \
                        for (int k = 0; k < KK; ++k) {                                             \
                            size_t oh = (b * 2 * KK + 2 * k) * cs + p, ow = oh + cs;               \
                            REAL h = (REAL)(y - 1 + k / KS) + offset[oh];                          \
                            REAL w = (REAL)(x - 1 + k % KS) + offset[ow];                          \
                            corners_##SUFFIX c = corners_at_##SUFFIX(img, H, W, h, w);             \
                            REAL val = bilinear_##SUFFIX(&c, H, W, h, w);                          \
                            REAL gk = g * w9[k];                                                   \
                            gw_loc[k] += (double)(g * (m[k] * val));                               \
                            gm[k] = gk * val;                                                      \
                            /* get_coordinate_weight: corner validity only */                      \
                            REAL dh = c.lw * (c.v4 - c.v2) + (1 - c.lw) * (c.v3 - c.v1);           \
                            REAL dw = c.lh * (c.v4 - c.v3) + (1 - c.lh) * (c.v2 - c.v1);           \
                            grad_offset[oh] = gk * m[k] * dh;                                      \
                            grad_offset[ow] = gk * m[k] * dw;                                      \
                            if (grad_init) {                                                       \
                                REAL cc = gk * m[k];                                               \
                                long hs[4] = {c.h0, c.h0, c.h0 + 1, c.h0 + 1};                     \
                                long ws[4] = {c.w0, c.w0 + 1, c.w0, c.w0 + 1};                     \
                                REAL cw[4] = {(1 - c.lh) * (1 - c.lw), (1 - c.lh) * c.lw,          \
                                              c.lh * (1 - c.lw), c.lh * c.lw};                     \
                                for (int q = 0; q < 4; ++q)                                        \
                                    if (hs[q] >= 0 && hs[q] <= H - 1 && ws[q] >= 0 &&              \
                                        ws[q] <= W - 1) {                                          \
                                        REAL add = cc * cw[q];                                     \
                                        _Pragma("omp atomic")                                      \
                                        grad_init[b * cs + hs[q] * W + ws[q]] += add;              \
                                    }                                                              \
                            }                                                                      \
                        }                                                                          \
                        REAL *ga = grad_weight + b * KK * cs + p;                                  \
                        if (mode == NORM_RESIDUAL) {                                               \
                            REAL s = 0;                                                            \
                            for (int k = 0; k < KK; ++k) s += gm[k];                               \
                            REAL mean = s / (REAL)KK;                                              \
                            for (int k = 0; k < KK; ++k) ga[k * cs] = gm[k] - mean;                \
                            if (grad_init) {                                                       \
                                REAL add = scale * g;                                              \
                                _Pragma("omp atomic")                                              \
                                grad_init[b * cs + p] += add;                                      \
                            }                                                                      \
                        } else if (mode == NORM_SUM) {                                             \
                            REAL s = 0, dot = 0;                                                   \
                            for (int k = 0; k < KK; ++k) {                                         \
                                s += a[k * cs];                                                    \
                                dot += gm[k] * m[k];                                               \
                            }                                                                      \
                            for (int k = 0; k < KK; ++k) ga[k * cs] = (gm[k] - dot) / s;           \
                        } else {                                                                   \
                            for (int k = 0; k < KK; ++k) ga[k * cs] = gm[k];                       \
                        }                                                                          \
                    }                                                                              \
                }                                                                                  \
            _Pragma("omp critical")                                                                \
            {                                                                                      \
                for (int k = 0; k < KK; ++k) gw_tot[k] += gw_loc[k];                               \
                gb_tot += gb_loc;                                                                  \
            }                                                                                      \
        }                                                                                          \
        for (int k = 0; k < KK; ++k) grad_w9[k] = (REAL)gw_tot[k];                                 \
        grad_b1[0] = (REAL)gb_tot;                                                                 \
    }

DEFINE_ORACLE(float, f32, floorf)
DEFINE_ORACLE(double, f64, floor)
