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

constexpr int CROP_ROWS = 32;   // tile rows per CTA

// One CTA: CROP_ROWS rows of one (tile, channel) plane; a warp per row, lanes along x, up to four columns of a row
// and two rows in flight per lane.  All index arithmetic is 32-bit and per row / per CTA, not per element.
__global__ void __launch_bounds__(256)
tiles_crop_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int H, int W, int pad, int k,
                  int stride, int n_x) {
    const unsigned pc = blockIdx.x;             // tile * C + channel
    const int c = (int)(pc % (unsigned)C);
    const int t = (int)(pc / (unsigned)C);
    const int ty = t / n_x, tx = t - ty * n_x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y_end = min(k, (int)(blockIdx.y + 1) * CROP_ROWS);
    const float* __restrict__ plane = src + (size_t)c * H * W;
    float* __restrict__ out = dst + (size_t)pc * k * k;
    const int sx0 = stride * tx;
    for (int y = blockIdx.y * CROP_ROWS + warp; y < y_end; y += 16) {
        const int y2 = y + 8;
        const bool two = y2 < y_end;
        int sy = stride * ty + y, sy2 = stride * ty + (two ? y2 : y);
        if (pad > 0) {
            sy = pad_source(sy, pad, H, true);
            sy2 = pad_source(sy2, pad, H, true);
        }
        const float* __restrict__ r1 = plane + (size_t)sy * W;
        const float* __restrict__ r2 = plane + (size_t)sy2 * W;
        for (int x0 = lane; x0 < k; x0 += 128) {
            float v1[4], v2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int x = x0 + 32 * u;
                int sx = sx0 + (x < k ? x : 0);
                if (pad > 0) sx = pad_source(sx, pad, W, false);
                v1[u] = __ldcs(r1 + sx);
                v2[u] = __ldcs(r2 + sx);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int x = x0 + 32 * u;
                if (x < k) {
                    __stcs(out + (size_t)y * k + x, v1[u]);
                    if (two) __stcs(out + (size_t)y2 * k + x, v2[u]);
                }
            }
        }
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

// (16 rows per thread, i.e. half as many CTAs: float64 0.250 -> 0.258 ms, float32 0.238 -> 0.230 ms on the same box - 8 kept)
#ifndef JSPSR_MERGE_ROWS
#define JSPSR_MERGE_ROWS 8
#endif
constexpr int MERGE_ROWS = JSPSR_MERGE_ROWS;    // output rows per thread
constexpr int MERGE_COLS = 128;  // output columns per CTA (one per thread)

// The common case (L <= 2 * stride: at most 2 x 2 tiles cover a pixel).  The grid walks the output in BANDS of
// `stride` rows (band r = rows [r * stride, (r + 1) * stride), the last one L rows): inside band r a pixel belongs to
// tile row r and, in its first p rows, also to tile row r - 1, so no thread divides to find its tile rows and the
// row blend weights are per CTA (computed once, shared memory).  Two instantiations, one launch each: UP = false
// covers the rows owned by one tile row (most of the raster: one or two gathers per pixel), UP = true the p
// overlapped rows of bands 1 .. n_y - 1 (two or four).  A thread owns one output column (one division) and
// MERGE_ROWS rows; all gathers are issued before the float64 arithmetic, which is the generic kernel's,
// operation for operation (bit-identical results).
template <typename TO, bool UP>
__global__ void __launch_bounds__(MERGE_COLS)
tiles_merge2_kernel(const float* __restrict__ tiles, TO* __restrict__ out, int n_y, int n_x, int k, int crop, int L,
                    int stride, int p, double step, int out_h, int out_w, int chunks) {
    __shared__ double s_wy[MERGE_ROWS][2];   // blend weight of the row in tile row r - 1 / r
    const int band = blockIdx.y / chunks;
    const int r = UP ? band + 1 : band;                             // band = tile row that owns these output rows
    const int y_lo = UP ? 0 : (r > 0 ? p : 0);                      // local rows [y_lo, y_hi) of the band
    const int y_hi = UP ? p : ((r == n_y - 1) ? L : stride);
    const int y0 = y_lo + (blockIdx.y - band * chunks) * MERGE_ROWS;   // first local row (= row in tile r before the crop)
    if (y0 >= y_hi) return;
    if (threadIdx.x < MERGE_ROWS) {
        const int y = y0 + threadIdx.x;
        s_wy[threadIdx.x][0] = UP ? blend_weight(y + stride, r - 1, n_y, L, p, step) : 0.0;
        s_wy[threadIdx.x][1] = blend_weight(y, r, n_y, L, p, step);
    }
    __syncthreads();
    const int X = blockIdx.x * MERGE_COLS + threadIdx.x;
    if (X >= out_w) return;
    const size_t kk = (size_t)k * k;
    const int c_hi = (n_x == 1) ? 0 : min(X / stride, n_x - 1);
    const int tx = X - stride * c_hi;
    const bool c_two = (c_hi > 0) && (tx < p);                      // also inside tile column c_hi - 1, at tx + stride
    const double wx_hi = blend_weight(tx, c_hi, n_x, L, p, step);
    const double wx_lo = c_two ? blend_weight(tx + stride, c_hi - 1, n_x, L, p, step) : 0.0;
    // tile (r, c_hi), local pixel (y0, tx); tile column c_hi - 1 at tx + stride; tile row r - 1 at y + stride
    const float* __restrict__ t_hi = tiles + ((size_t)blockIdx.z * n_y * n_x + (size_t)r * n_x + c_hi) * kk +
                                     (size_t)(y0 + crop) * k + (tx + crop);
    const float* __restrict__ t_lo = t_hi + ((ptrdiff_t)stride - (ptrdiff_t)kk);
    const ptrdiff_t d_up = (ptrdiff_t)stride * k - (ptrdiff_t)n_x * (ptrdiff_t)kk;
    const float* __restrict__ u_hi = t_hi + d_up;
    const float* __restrict__ u_lo = t_lo + d_up;

    float v[MERGE_ROWS][UP ? 2 : 1][2];   // [row][(tile row r - 1,) tile row r][tile column c_hi - 1 / c_hi]
#pragma unroll
    for (int j = 0; j < MERGE_ROWS; ++j) {
        const int o = min(j, y_hi - 1 - y0) * k;
        v[j][UP ? 1 : 0][1] = __ldcs(t_hi + o);
        v[j][UP ? 1 : 0][0] = c_two ? __ldcs(t_lo + o) : 0.f;
        if (UP) {
            v[j][0][1] = __ldcs(u_hi + o);
            v[j][0][0] = c_two ? __ldcs(u_lo + o) : 0.f;
        }
    }
    TO* __restrict__ dst = out + ((size_t)blockIdx.z * out_h + (size_t)r * stride + y0) * out_w + X;
#pragma unroll
    for (int j = 0; j < MERGE_ROWS; ++j) {
        if (y0 + j < y_hi) {
            double row = __dmul_rn((double)v[j][UP ? 1 : 0][1], wx_hi);
            if (c_two) row = __dadd_rn(__dmul_rn((double)v[j][UP ? 1 : 0][0], wx_lo), row);   // left tile first (copyto_add)
            double acc = __dmul_rn(row, s_wy[j][1]);
            if (UP) {
                double row_up = __dmul_rn((double)v[j][0][1], wx_hi);
                if (c_two) row_up = __dadd_rn(__dmul_rn((double)v[j][0][0], wx_lo), row_up);
                acc = __dadd_rn(__dmul_rn(row_up, s_wy[j][0]), acc);                            // upper tile row first
            }
            __stcs(dst + (size_t)j * out_w, (TO)acc);
        }
    }
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
    const long long planes = (long long)n_y * n_x * C;
    const int chunks = (k + CROP_ROWS - 1) / CROP_ROWS;
    if (planes > 0x7fffffffLL || chunks > 65535) return jspsr_internal_fail(JSPSR_ERR_UNSUPPORTED, "tiles_crop: too many tiles");
    tiles_crop_kernel<<<dim3((unsigned)planes, (unsigned)chunks), 256, 0, (cudaStream_t)stream>>>(src, dst, C, H, W, pad, k,
                                                                                              stride, n_x);
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
    const unsigned gx = (unsigned)((out_w + MERGE_COLS - 1) / MERGE_COLS);
    // CTAs per band: rows owned by one tile row (up to L in the last band) / the p overlapped rows of bands >= 1
    const int chunks = (max(stride, L) + MERGE_ROWS - 1) / MERGE_ROWS, chunks_up = (p + MERGE_ROWS - 1) / MERGE_ROWS;
    const long long gy = (long long)n_y * chunks, gy_up = (long long)(n_y - 1) * chunks_up;
    if (L <= 2 * stride && gy <= 65535 && gy_up <= 65535 && S <= 65535) {   // at most two tiles overlap along an axis
#define JSPSR_LAUNCH_MERGE2(TO, UP, GY, CH)                                                               \
    tiles_merge2_kernel<TO, UP><<<dim3(gx, (unsigned)(GY), (unsigned)S), MERGE_COLS, 0, (cudaStream_t)stream>>>( \
        tiles, (TO*)out, n_y, n_x, k, crop, L, stride, p, step, out_h, out_w, CH)
        if (out_f64) JSPSR_LAUNCH_MERGE2(double, false, gy, chunks); else JSPSR_LAUNCH_MERGE2(float, false, gy, chunks);
        if (gy_up > 0) {
            if (out_f64) JSPSR_LAUNCH_MERGE2(double, true, gy_up, chunks_up);
            else JSPSR_LAUNCH_MERGE2(float, true, gy_up, chunks_up);
        }
#undef JSPSR_LAUNCH_MERGE2
    } else if (out_f64)
        tiles_merge_kernel<double><<<launch_blocks(total), 256, 0, (cudaStream_t)stream>>>(
            tiles, (double*)out, n_y, n_x, k, crop, L, stride, p, step, out_h, out_w, total);
    else
        tiles_merge_kernel<float><<<launch_blocks(total), 256, 0, (cudaStream_t)stream>>>(
            tiles, (float*)out, n_y, n_x, k, crop, L, stride, p, step, out_h, out_w, total);
    return cuda_check("tiles_merge launch");
}
