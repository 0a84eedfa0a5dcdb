// Loss + metric epilogue of the propagation output (C ABI in include/jspsr_tiles.h).
//
//   loss_l1_l2_grad_kernel : the YAML configs' MultiLoss (L1 + L2 + 0.1 * Sobel-L1, losses/loss_schemes.py:55-72,
//                            losses/loss_functions.py:171-185) AND its gradient w.r.t. the prediction, one pass over
//                            (pred, gt): 8 B/pixel read, 4 B/pixel written, instead of ~20 elementwise / convolution
//                            kernels re-reading the 4 B/pixel output of the propagation.
//   dem_metrics_kernel     : MeterRMSE's arithmetic (evaluation/metrics.py:142-199, 361-382: border crop, clamp,
//                            de-normalise, squared error) and the absolute error, per sample, one pass.
//
// Both are HBM-bound elementwise/stencil reductions: coalesced streaming loads, the difference tile staged in shared
// memory with the replicate border the Sobel operator needs, global sums through fp64 atomics + last-CTA publish.
#include <cmath>
#include <cstdio>

#include "../../include/jspsr_tiles.h"
#include "spn_common.cuh"

int jspsr_internal_fail(int code, const char* msg);  // abi.cu: sets the thread's last-error message

namespace jspsr {

constexpr int LT_H = 16;    // rows per CTA
constexpr int LT_W = 128;   // columns per CTA
constexpr int LD_H = LT_H + 4, LD_W = LT_W + 4;   // staged difference tile (halo 2: Sobel of the halo-1 ring)
constexpr int LS_H = LT_H + 2, LS_W = LT_W + 2;   // sign tile (halo 1)

struct alignas(16) LossWs {
    double sums[3];  // sum |d|, sum d^2, sum |Sobel(pred) - Sobel(gt)|
    unsigned int ticket;
    unsigned int pad;
};
static_assert(sizeof(LossWs) <= sizeof(ReduceWs), "the propagation's reduction workspace is large enough");

__device__ __forceinline__ float sgn(float v) { return (float)((v > 0.f) - (v < 0.f)); }

// One CTA: LT_H x LT_W pixels of one plane.
template <bool WRITE_GRAD>
__global__ void __launch_bounds__(THREADS)
loss_l1_l2_grad_kernel(const float* __restrict__ pred, const float* __restrict__ gt, float* __restrict__ grad,
                       float* __restrict__ losses4, LossWs* __restrict__ ws, int H, int W, int tiles_x, int tiles_y,
                       float w_l1, float w_l2, float w_grad, float inv_n) {
    __shared__ float s_d[LD_H][LD_W];
    __shared__ signed char s_sx[LS_H][LS_W + 2];
    __shared__ signed char s_sy[LS_H][LS_W + 2];
    __shared__ float s_red[WARPS][3];
    __shared__ bool s_last;

    const int tile = blockIdx.x;
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const size_t plane = (size_t)(tile / (tiles_x * tiles_y)) * H * W;
    const int y0 = ty * LT_H, x0 = tx * LT_W;
    const float* __restrict__ p = pred + plane;
    const float* __restrict__ g = gt + plane;

    // ---- stage d = pred - gt over the tile + halo 2, replicate-clamped (kornia pads with mode "replicate") ----
    for (int i = threadIdx.x; i < LD_H * LD_W; i += THREADS) {
        const int r = i / LD_W, c = i - r * LD_W;
        const int y = min(max(y0 - 2 + r, 0), H - 1), x = min(max(x0 - 2 + c, 0), W - 1);
        const size_t o = (size_t)y * W + x;
        s_d[r][c] = ld_stream(p + o) - ld_stream(g + o);
    }
    __syncthreads();

    // ---- signs of the Sobel difference on the tile + halo 1 (zero outside the image); |.| summed inside the tile ----
    float a_grad = 0.f;
    for (int i = threadIdx.x; i < LS_H * LS_W; i += THREADS) {
        const int r = i / LS_W, c = i - r * LS_W;          // output position (y0 - 1 + r, x0 - 1 + c)
        const int y = y0 - 1 + r, x = x0 - 1 + c;
        signed char sx = 0, sy = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            // staged index of image (y + a, x + b) is [r + 1 + a][c + 1 + b]
            const float d00 = s_d[r][c], d01 = s_d[r][c + 1], d02 = s_d[r][c + 2];
            const float d10 = s_d[r + 1][c], d12 = s_d[r + 1][c + 2];
            const float d20 = s_d[r + 2][c], d21 = s_d[r + 2][c + 1], d22 = s_d[r + 2][c + 2];
            const float gx = ((d02 - d00) + 2.f * (d12 - d10) + (d22 - d20)) * 0.125f;
            const float gy = ((d20 - d00) + 2.f * (d21 - d01) + (d22 - d02)) * 0.125f;
            sx = (signed char)sgn(gx);
            sy = (signed char)sgn(gy);
            if (r >= 1 && r <= LT_H && c >= 1 && c <= LT_W) a_grad += fabsf(gx) + fabsf(gy);
        }
        s_sx[r][c] = sx;
        s_sy[r][c] = sy;
    }
    __syncthreads();

    // ---- per pixel: L1, L2 and the gradient of Total ----
    float a_l1 = 0.f, a_l2 = 0.f;
    const float c_pix_l1 = w_l1 * inv_n, c_pix_l2 = 2.f * w_l2 * inv_n, c_sob = w_grad * 0.5f * inv_n * 0.125f;
    // S(oy, ox): sign tile lookups by image position; positions outside the staged ring are outside the image
    auto SX = [&](int oy, int ox) -> float {
        return (oy < 0 || oy >= H || ox < 0 || ox >= W) ? 0.f : (float)s_sx[oy - y0 + 1][ox - x0 + 1];
    };
    auto SY = [&](int oy, int ox) -> float {
        return (oy < 0 || oy >= H || ox < 0 || ox >= W) ? 0.f : (float)s_sy[oy - y0 + 1][ox - x0 + 1];
    };
    // gradient w.r.t. the PADDED image at padded position (u, v) (image pixel (u - 1, v - 1)), times 8:
    //   sum_{i,j} kx[i][j] * sx[u - i][v - j] + ky[i][j] * sy[u - i][v - j],  kx[i][j] = r[i] * c[j], ky = kx^T,
    //   r = (1, 2, 1), c = (-1, 0, 1)
    auto GP = [&](int u, int v) -> float {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float ri = (i == 1) ? 2.f : 1.f;
            acc += ri * (SX(u - i, v - 2) - SX(u - i, v));        // j = 2 (+1) and j = 0 (-1)
            acc += ri * (SY(u - 2, v - i) - SY(u, v - i));        // ky[2][.] = +r, ky[0][.] = -r
        }
        return acc;
    };
    for (int i = threadIdx.x; i < LT_H * LT_W; i += THREADS) {
        const int r = i / LT_W, c = i - r * LT_W;
        const int y = y0 + r, x = x0 + c;
        if (y >= H || x >= W) continue;
        const float d = s_d[r + 2][c + 2];
        a_l1 += fabsf(d);
        a_l2 += d * d;
        if (WRITE_GRAD) {
            float gs;
            if (y > 0 && y < H - 1 && x > 0 && x < W - 1) {
                // interior: u = y + 1, v = x + 1; sign tile index of output (oy, ox) is [oy - y0 + 1][ox - x0 + 1]
                const int sr = r + 1, sc = c + 1;   // sign-tile index of (y, x)
                gs = 0.f;
#pragma unroll
                for (int k = -1; k <= 1; ++k) {
                    const float rk = (k == 0) ? 2.f : 1.f;
                    gs += rk * ((float)s_sx[sr - k][sc - 1] - (float)s_sx[sr - k][sc + 1]);
                    gs += rk * ((float)s_sy[sr - 1][sc - k] - (float)s_sy[sr + 1][sc - k]);
                }
            } else {
                // border pixels also receive what the replicate padding folds back onto them
                gs = 0.f;
                for (int u = (y == 0 ? 0 : y + 1); u <= (y == H - 1 ? H + 1 : y + 1); ++u)
                    for (int v = (x == 0 ? 0 : x + 1); v <= (x == W - 1 ? W + 1 : x + 1); ++v) gs += GP(u, v);
            }
            st_stream(grad + plane + (size_t)y * W + x, c_pix_l1 * sgn(d) + c_pix_l2 * d + c_sob * gs);
        }
    }

    // ---- thread -> warp -> CTA -> fp64 atomics; the last CTA publishes the four losses ----
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a_l1 = warp_sum(a_l1);
    a_l2 = warp_sum(a_l2);
    a_grad = warp_sum(a_grad);
    if (lane == 0) {
        s_red[warp][0] = a_l1;
        s_red[warp][1] = a_l2;
        s_red[warp][2] = a_grad;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
#pragma unroll
        for (int wi = 0; wi < WARPS; ++wi) v += (double)s_red[wi][threadIdx.x];
        atomicAdd(&ws->sums[threadIdx.x], v);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&ws->ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const double n_inv = (double)inv_n;
        const double l1 = atomicAdd(&ws->sums[0], 0.0) * n_inv;
        const double l2 = atomicAdd(&ws->sums[1], 0.0) * n_inv;
        const double gr = atomicAdd(&ws->sums[2], 0.0) * n_inv * 0.5;
        losses4[0] = (float)l1;
        losses4[1] = (float)l2;
        losses4[2] = (float)gr;
        losses4[3] = (float)((double)w_l1 * l1 + (double)w_l2 * l2 + (double)w_grad * gr);
        ws->sums[0] = ws->sums[1] = ws->sums[2] = 0.0;   // leave the workspace clean for the next call
        ws->ticket = 0u;
    }
}

// One CTA: a slab of rows of one sample's border-cropped window.  sums[b] = {sum d^2, sum |d|} (fp64 atomics).
__global__ void __launch_bounds__(THREADS)
dem_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, double* __restrict__ sums, int H, int W,
                   int bh, int bw, float log_range, float range, float vmin, int elev_log, int rows_per_cta) {
    __shared__ double s_red[WARPS][2];
    const int b = blockIdx.y;
    const int hc = H - 2 * bh, wc = W - 2 * bw;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, hc);
    const size_t plane = (size_t)b * H * W;
    float a_sq = 0.f, a_ab = 0.f;
    double d_sq = 0.0, d_ab = 0.0;
    for (int r = r0; r < r1; ++r) {
        const size_t row = plane + (size_t)(bh + r) * W + bw;
        for (int c = threadIdx.x; c < wc; c += THREADS) {
            const float pv = fminf(fmaxf(ld_stream(pred + row + c), 0.f), 1.f);   // MeterBase._prepare clamps pred only
            const float gv = ld_stream(gt + row + c);
            float pe, ge;
            if (elev_log) {
                pe = expf(pv * log_range) + vmin;
                ge = expf(gv * log_range) + vmin;
            } else {
                pe = __fadd_rn(__fmul_rn(pv, range), vmin);
                ge = __fadd_rn(__fmul_rn(gv, range), vmin);
            }
            const float d = pe - ge;
            a_sq += d * d;
            a_ab += fabsf(d);
        }
        // fold the fp32 row partials into fp64 so that long windows do not lose low bits
        d_sq += (double)a_sq;
        d_ab += (double)a_ab;
        a_sq = a_ab = 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d_sq += __shfl_xor_sync(0xffffffffu, d_sq, o);
        d_ab += __shfl_xor_sync(0xffffffffu, d_ab, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_red[warp][0] = d_sq;
        s_red[warp][1] = d_ab;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0.0;
#pragma unroll
        for (int wi = 0; wi < WARPS; ++wi) v += s_red[wi][threadIdx.x];
        atomicAdd(&sums[2 * b + threadIdx.x], v);
    }
}

}  // namespace jspsr

using namespace jspsr;

extern "C" int jspsr_loss_l1_l2_grad(const float* pred, const float* gt, float w_l1, float w_l2, float w_grad,
                                     float* losses4, float* grad_pred, void* workspace, int planes, int H, int W,
                                     void* stream) {
    if (planes <= 0 || H <= 0 || W <= 0)
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "loss: non-positive dimension");
    if (!pred || !gt || !losses4 || !workspace) return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "loss: null pointer");
    if (((uintptr_t)pred | (uintptr_t)gt | (uintptr_t)losses4 | (uintptr_t)grad_pred) & 3)
        return jspsr_internal_fail(JSPSR_ERR_ALIGN, "loss: a float pointer is not 4-byte aligned");
    if ((uintptr_t)workspace & 15) return jspsr_internal_fail(JSPSR_ERR_ALIGN, "loss: workspace is not 16-byte aligned");
    const int tiles_x = (W + LT_W - 1) / LT_W, tiles_y = (H + LT_H - 1) / LT_H;
    const long long ctas = (long long)planes * tiles_x * tiles_y;
    if (ctas > 0x7fffffffLL) return jspsr_internal_fail(JSPSR_ERR_UNSUPPORTED, "loss: more than 2^31 tiles");
    const float inv_n = (float)(1.0 / ((double)planes * H * W));
    if (grad_pred)
        loss_l1_l2_grad_kernel<true><<<(unsigned)ctas, THREADS, 0, (cudaStream_t)stream>>>(
            pred, gt, grad_pred, losses4, (LossWs*)workspace, H, W, tiles_x, tiles_y, w_l1, w_l2, w_grad, inv_n);
    else
        loss_l1_l2_grad_kernel<false><<<(unsigned)ctas, THREADS, 0, (cudaStream_t)stream>>>(
            pred, gt, nullptr, losses4, (LossWs*)workspace, H, W, tiles_x, tiles_y, w_l1, w_l2, w_grad, inv_n);
    const cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) {
        char msg[256];
        snprintf(msg, sizeof(msg), "loss kernel launch: %s", cudaGetErrorString(ce));
        return jspsr_internal_fail(JSPSR_ERR_CUDA, msg);
    }
    return JSPSR_OK;
}

extern "C" int jspsr_dem_metrics(const float* pred, const float* gt, double* sums, int B, int H, int W, int border_h,
                                 int border_w, float value_min, float value_max, int elev_log, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0) return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "metrics: non-positive dimension");
    if (!pred || !gt || !sums) return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "metrics: null pointer");
    if (border_h < 0 || border_w < 0 || 2 * border_h >= H || 2 * border_w >= W)
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "metrics: the border leaves no pixels");
    if (B > 65535) return jspsr_internal_fail(JSPSR_ERR_UNSUPPORTED, "metrics: more than 65535 samples per call");
    if (elev_log && !(value_max - value_min > 0.f))
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "metrics: log de-normalisation needs value_max > value_min");
    if (((uintptr_t)pred | (uintptr_t)gt) & 3 || ((uintptr_t)sums & 7))
        return jspsr_internal_fail(JSPSR_ERR_ALIGN, "metrics: misaligned pointer");
    cudaError_t ce = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)B, (cudaStream_t)stream);
    if (ce == cudaSuccess) {
        const int hc = H - 2 * border_h;
        // enough CTAs to fill the GPU twice over, at least 4 rows each
        int slabs = (2 * 148 * 4 + B - 1) / B;
        slabs = max(1, min(slabs, (hc + 3) / 4));
        const int rows_per_cta = (hc + slabs - 1) / slabs;
        slabs = (hc + rows_per_cta - 1) / rows_per_cta;
        const float range = (float)((double)value_max - (double)value_min);
        // data * log(max - min): the reference multiplies by the python float (double) rounded into the fp32 tensor op
        const float log_range = elev_log ? (float)log((double)value_max - (double)value_min) : 0.f;
        dem_metrics_kernel<<<dim3((unsigned)slabs, (unsigned)B), THREADS, 0, (cudaStream_t)stream>>>(
            pred, gt, sums, H, W, border_h, border_w, log_range, range, value_min, elev_log, rows_per_cta);
        ce = cudaGetLastError();
    }
    if (ce != cudaSuccess) {
        char msg[256];
        snprintf(msg, sizeof(msg), "metrics kernel launch: %s", cudaGetErrorString(ce));
        return jspsr_internal_fail(JSPSR_ERR_CUDA, msg);
    }
    return JSPSR_OK;
}
