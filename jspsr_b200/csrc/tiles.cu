// Tile scheduler and blended merge on the GPU (C ABI in include/jspsr_tiles.h).
//
//   tiles_crop_kernel  : [C,H,W] raster -> [n_y*n_x, C, k, k] overlapping tiles in TileCrop's row-major walk
//                        (data/data_utils.py:129-163), optionally through upscale_dem's mirrored border
//                        (utils/utils.py:1501-1522, including the one-row shift of its bottom border), in one pass.
//   tiles_merge_kernel : [S, n_y*n_x, k, k] predictions -> [S, out_h, out_w] rasters: border crop, linear-ramp
//                        weights over the overlaps and the accumulation of merge_dem (utils/utils.py:802-965) as a
//                        GATHER - one thread per output pixel, the 1..4 covering tiles visited in the reference's
//                        merge order (left to right inside a row of tiles, then top to bottom), float64 arithmetic
//                        with explicitly rounded multiplies/adds, so the result is bit-identical to the reference's
//                        numpy float64 result and needs no atomics and no zero-fill.
//
// Both are pure data movement: HBM-bound, coalesced along x, grid-stride over the destination.
#include <cstdio>

#include "../../include/jspsr_tiles.h"
#include "spn_common.cuh"

int jspsr_internal_fail(int code, const char* msg);  // abi.cu: sets the thread's last-error message

namespace jspsr {

// source row/column of padded index i (add_padding): left/top mirror n-1-i; right mirror size-1-j;
// the reference's bottom border is taken one row early (utils.py:1517) -> size-2-j
__device__ __forceinline__ int pad_source(int i, int n, int size, bool bottom) {
    if (i < n) return n - 1 - i;
    if (i < n + size) return i - n;
    const int j = i - n - size;
    return bottom ? size - 2 - j : size - 1 - j;
}

__global__ void __launch_bounds__(256)
tiles_crop_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W, int pad, int k,
                  int stride, int n_x, size_t total) {
    const size_t kk = (size_t)k * k;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % k);
        const int y = (int)((i / k) % k);
        const size_t pc = i / kk;             // tile * C + channel
        const int c = (int)(pc % C);
        const int t = (int)(pc / C);
        const int ty = t / n_x, tx = t - ty * n_x;
        int sy = stride * ty + y, sx = stride * tx + x;
        if (pad > 0) {
            sy = pad_source(sy, pad, H, true);
            sx = pad_source(sx, pad, W, false);
        }
        st_stream(dst + i, ld_stream(src + ((size_t)c * H + sy) * W + sx));
    }
}

// 1-D blend weight of local coordinate t in tile i of n along one axis (gen_weight_row / gen_weight_col):
// ramp[j] = linspace(1, 0, p + 2)[1 + j] = (j + 1) * step + 1 with step = -1 / (p + 1), product and sum rounded
// separately as numpy does
__device__ __forceinline__ double blend_weight(int t, int i, int n, int L, int p, double step) {
    if (n == 1 || p <= 0) return 1.0;
    int j = -1;
    if (i == 0) {
        if (t >= L - p) j = t - (L - p);
    } else if (i == n - 1) {
        if (t < p) j = p - 1 - t;
    } else {
        if (t >= L - p) j = t - (L - p);      // `[-p:] = weight` is assigned last and wins (utils.py:829-830)
        else if (t < p) j = p - 1 - t;
    }
    if (j < 0) return 1.0;
    return __dadd_rn(__dmul_rn((double)(j + 1), step), 1.0);
}

template <typename TO>
__global__ void __launch_bounds__(256)
tiles_merge_kernel(const float* __restrict__ tiles, TO* __restrict__ out, int n_y, int n_x, int k, int crop, int L,
                   int stride, int p, double step, int out_h, int out_w, size_t total) {
    const size_t kk = (size_t)k * k;
    const size_t plane = (size_t)out_h * out_w;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % out_w);
        const int Y = (int)((i / out_w) % out_h);
        const size_t s = i / plane;
        // tiles covering X: stride * c <= X < stride * c + L
        int c_lo, c_hi, r_lo, r_hi;
        if (n_x == 1) c_lo = c_hi = 0;
        else {
            c_hi = min(X / stride, n_x - 1);
            c_lo = max(0, (X - L + stride) / stride);      // ceil((X - L + 1) / stride)
            if (X - L + 1 <= 0) c_lo = 0;
        }
        if (n_y == 1) r_lo = r_hi = 0;
        else {
            r_hi = min(Y / stride, n_y - 1);
            r_lo = max(0, (Y - L + stride) / stride);
            if (Y - L + 1 <= 0) r_lo = 0;
        }
        const float* __restrict__ base = tiles + s * (size_t)n_y * n_x * kk;
        double acc = 0.0;
        for (int r = r_lo; r <= r_hi; ++r) {
            const int ty = Y - stride * r;
            double row = 0.0;
            for (int c = c_lo; c <= c_hi; ++c) {
                const int tx = X - stride * c;
                const float v = ld_stream(base + (size_t)(r * n_x + c) * kk + (size_t)(ty + crop) * k + (tx + crop));
                const double wv = __dmul_rn((double)v, blend_weight(tx, c, n_x, L, p, step));
                row = (c == c_lo) ? wv : __dadd_rn(row, wv);          // copy, then add (copyto_add)
            }
            const double cv = __dmul_rn(row, blend_weight(ty, r, n_y, L, p, step));
            acc = (r == r_lo) ? cv : __dadd_rn(acc, cv);
        }
        out[i] = (TO)acc;
    }
}

static int launch_blocks(size_t total) { return (int)min((size_t)148 * 16, (total + 255) / 256); }

}  // namespace jspsr

using namespace jspsr;

static int cuda_check(const char* what) {
    const cudaError_t ce = cudaGetLastError();
    if (ce == cudaSuccess) return JSPSR_OK;
    char msg[256];
    snprintf(msg, sizeof(msg), "%s: %s", what, cudaGetErrorString(ce));
    return jspsr_internal_fail(JSPSR_ERR_CUDA, msg);
}

extern "C" int jspsr_tiles_crop(const float* src, float* dst, int C, int H, int W, int pad, int k, int stride,
                                int n_y, int n_x, void* stream) {
    if (C <= 0 || H <= 0 || W <= 0 || k <= 0 || n_y <= 0 || n_x <= 0 || pad < 0 || stride < 0)
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "tiles_crop: non-positive dimension");
    if (!src || !dst) return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "tiles_crop: null pointer");
    if (((uintptr_t)src | (uintptr_t)dst) & 3) return jspsr_internal_fail(JSPSR_ERR_ALIGN, "tiles_crop: misaligned pointer");
    if (pad > W || pad > H - 1)
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "tiles_crop: the mirrored border is wider than the image");
    if ((long long)stride * (n_y - 1) + k > (long long)H + 2 * pad || (long long)stride * (n_x - 1) + k > (long long)W + 2 * pad)
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "tiles_crop: the tile walk leaves the (padded) image");
    const size_t total = (size_t)n_y * n_x * C * k * k;
    tiles_crop_kernel<<<launch_blocks(total), 256, 0, (cudaStream_t)stream>>>(src, dst, C, H, W, pad, k, stride, n_x, total);
    return cuda_check("tiles_crop launch");
}

extern "C" int jspsr_tiles_merge(const float* tiles, void* out, int S, int n_y, int n_x, int k, int crop, int stride,
                                 int out_f64, void* stream) {
    if (S <= 0 || n_y <= 0 || n_x <= 0 || k <= 0 || crop < 0 || stride < 0)
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "tiles_merge: non-positive dimension");
    if (!tiles || !out) return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "tiles_merge: null pointer");
    const int L = k - 2 * crop;
    if (L <= 0) return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "tiles_merge: the border crop leaves no pixels");
    if ((n_y > 1 || n_x > 1) && (stride <= 0 || stride > L))
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "tiles_merge: stride must be in [1, k - 2 * crop] (tiles must touch)");
    if (((uintptr_t)tiles & 3) || ((uintptr_t)out & (out_f64 ? 7 : 3)))
        return jspsr_internal_fail(JSPSR_ERR_ALIGN, "tiles_merge: misaligned pointer");
    const int p = L - stride;   // overlapped pixels (utils.py:814)
    const int out_h = stride * (n_y - 1) + L, out_w = stride * (n_x - 1) + L;
    const double step = -1.0 / (double)(p + 1);   // numpy.linspace(1, 0, p + 2): step = (0 - 1) / (p + 1)
    const size_t total = (size_t)S * out_h * out_w;
    if (out_f64)
        tiles_merge_kernel<double><<<launch_blocks(total), 256, 0, (cudaStream_t)stream>>>(
            tiles, (double*)out, n_y, n_x, k, crop, L, stride, p, step, out_h, out_w, total);
    else
        tiles_merge_kernel<float><<<launch_blocks(total), 256, 0, (cudaStream_t)stream>>>(
            tiles, (float*)out, n_y, n_x, k, crop, L, stride, p, step, out_h, out_w, total);
    return cuda_check("tiles_merge launch");
}
