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

template <typename T, int MODE, bool TMA>
__global__ void __launch_bounds__(THREADS, FWD_MIN_BLOCKS)
spn_forward_kernel(const T* __restrict__ init, const T* __restrict__ weight, const T* __restrict__ offset,
                   const float* __restrict__ w9, const float* __restrict__ b1, T* __restrict__ out, const Geom g,
                   const float scale, int* __restrict__ status, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(128) T tile[SH * SW];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_w[10];

    TileCtx c = make_tile_ctx(g);
    stage_tile_begin<T, TMA>(tile, &bar, &tmap, init, g, c.b, c.ox, c.oy - g.init_row0);
    // w9 == nullptr: frozen unit weight / zero bias (NLSPN, nlspn.py:61-65)
    if (threadIdx.x < 9) s_w[threadIdx.x] = w9 ? w9[threadIdx.x] : 1.f;
    if (threadIdx.x == 9) s_w[9] = b1 ? b1[0] : 0.f;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t cs = (size_t)g.H * g.W;  // channel stride
    const T* wgt_b = weight + (size_t)c.b * 9 * cs;
    const T* off_b = offset + (size_t)c.b * 18 * cs;
    const T* init_b = init + (size_t)c.b * g.init_rows * g.W;
    T* out_b = out + (size_t)c.b * cs;

    // pixel `it` of this thread: row warp + WARPS*(it / 4) of the tile, column lane + 32*(it % 4)
    float a[9], oh[9], ow[9];
    auto load_inputs = [&](int it, bool& active, size_t& p) {
        const int y = c.y0 + warp + WARPS * (it / (TILE_W / 32));
        const int x = c.x0 + lane + 32 * (it % (TILE_W / 32));
        active = (y < g.H) && (x < g.W);
        p = (size_t)y * g.W + x;
        if (active) {
#pragma unroll
            for (int k = 0; k < 9; ++k) a[k] = ld_stream(wgt_b + k * cs + p);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                oh[k] = ld_stream(off_b + (2 * k) * cs + p);
                ow[k] = ld_stream(off_b + (2 * k + 1) * cs + p);
            }
        }
    };

    // the first pixel's 27 streamed loads are in flight while the tile lands
    bool active;
    size_t p;
    load_inputs(0, active, p);
    stage_tile_wait<TMA>(&bar);

#pragma unroll 1
    for (int it = 0; it < PIX_PER_THREAD; ++it) {
        if (it > 0) load_inputs(it, active, p);
        if (!active) continue;
        const int ry = warp + WARPS * (it / (TILE_W / 32));
        const int cx = lane + 32 * (it % (TILE_W / 32));
        normalise9<MODE>(a);

        const float fy = (float)(g.row0 + c.y0 + ry), fx = (float)(c.x0 + cx);
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            // torchvision: (out_y - pad + i*dil) formed as an integer, converted, + offset
            const float h = (fy + (float)(k / 3 - 1)) + oh[k];
            const float w = (fx + (float)(k % 3 - 1)) + ow[k];
            const Tap t = gather_tap<T>(tile, init_b, g, c, h, w, status);
            const float hh = 1.f - t.lh, hw = 1.f - t.lw;
            const float val = hh * hw * t.v1 + hh * t.lw * t.v2 + t.lh * hw * t.v3 + t.lh * t.lw * t.v4;
            acc += s_w[k] * (a[k] * val);
        }
        acc += s_w[9];
        if (MODE == NORM_RESIDUAL) acc += scale * to_f32(tile[(ry + HALO_T) * SW + (cx + HALO_L)]);
        st_stream(out_b + p, acc);
    }
}

template <typename T, int MODE>
static cudaError_t launch_fwd_mode(const LaunchArgs& la) {
    dim3 grid((unsigned)((size_t)la.g.tiles_x * la.g.tiles_y * la.g.B));
    if (la.use_tma)
        spn_forward_kernel<T, MODE, true><<<grid, THREADS, 0, la.stream>>>(
            (const T*)la.init, (const T*)la.weight, (const T*)la.offset, la.w9, la.b1, (T*)la.out, la.g, la.scale,
            la.status, la.tmap);
    else
        spn_forward_kernel<T, MODE, false><<<grid, THREADS, 0, la.stream>>>(
            (const T*)la.init, (const T*)la.weight, (const T*)la.offset, la.w9, la.b1, (T*)la.out, la.g, la.scale,
            la.status, la.tmap);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_fwd_dtype(const LaunchArgs& la) {
    switch (la.mode) {
        case NORM_NONE: return launch_fwd_mode<T, NORM_NONE>(la);
        case NORM_RESIDUAL: return launch_fwd_mode<T, NORM_RESIDUAL>(la);
        default: return launch_fwd_mode<T, NORM_SUM>(la);
    }
}

cudaError_t launch_spn_forward(const LaunchArgs& la) {
    return la.bf16 ? launch_fwd_dtype<__nv_bfloat16>(la) : launch_fwd_dtype<float>(la);
}

}  // namespace jspsr
