// Fused backward of one propagation application (autograd of spn.py:99-118 /
// LRRU.py:267-298 / nlspn.py:177-187; upstream: deformable_col2im,
// deformable_col2im_coord, the GEMM for grad_weight and the bias sum, plus the
// Jacobian of the affinity normalisation) in ONE pass:
//   reads  grad_out, DEM tile (TMA-staged), 9 affinities, 18 offsets
//   writes grad_affinity[9], grad_offset[18]            (dense, one owner per element)
//          grad_w[9], grad_b                            (warp shuffle -> CTA -> fp64 atomics,
//                                                        last CTA publishes and re-zeroes)
//          grad_init (optional)                         (scatter: shared-memory tile of
//                                                        atomics, flushed with global REDs)
#include "spn_kernels.cuh"

namespace jspsr {

template <typename T>
__device__ __forceinline__ void scatter_corner_global(float* __restrict__ gi_b, const Geom& g, int hi, int wi, float v) {
    if ((unsigned)hi >= (unsigned)g.H_img || (unsigned)wi >= (unsigned)g.W) return;
    const int br = hi - g.init_row0;
    if ((unsigned)br >= (unsigned)g.init_rows) return;
    atomicAdd(gi_b + (size_t)br * g.W + wi, v);
}

template <typename T, int MODE, bool TMA, bool GRAD_INIT>
__global__ void __launch_bounds__(THREADS, BWD_MIN_BLOCKS)
spn_backward_kernel(const T* __restrict__ gout, const T* __restrict__ init, const T* __restrict__ weight,
                    const T* __restrict__ offset, const float* __restrict__ w9, float* __restrict__ grad_init,
                    T* __restrict__ grad_weight, T* __restrict__ grad_offset, float* __restrict__ grad_w9,
                    float* __restrict__ grad_b1, ReduceWs* __restrict__ ws, const Geom g, const float scale,
                    const bool accumulate, const __grid_constant__ CUtensorMap tmap) {
    __shared__ __align__(128) T tile[SH * SW];
    __shared__ __align__(16) float gtile[GRAD_INIT ? SH * SW : 1];
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_w[9];
    __shared__ float s_red[WARPS][10];
    __shared__ bool s_last;

    TileCtx c = make_tile_ctx(g);
    stage_tile_begin<T, TMA>(tile, &bar, &tmap, init, g, c.b, c.ox, c.oy - g.init_row0);
    if (threadIdx.x < 9) s_w[threadIdx.x] = w9 ? w9[threadIdx.x] : 1.f;
    if (GRAD_INIT) {
        for (int i = threadIdx.x; i < SH * SW; i += THREADS) gtile[i] = 0.f;
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t cs = (size_t)g.H * g.W;
    const T* gout_b = gout + (size_t)c.b * cs;
    const T* wgt_b = weight + (size_t)c.b * 9 * cs;
    const T* off_b = offset + (size_t)c.b * 18 * cs;
    const T* init_b = init + (size_t)c.b * g.init_rows * g.W;
    T* gwgt_b = grad_weight + (size_t)c.b * 9 * cs;
    T* goff_b = grad_offset + (size_t)c.b * 18 * cs;
    float* gi_b = GRAD_INIT ? grad_init + (size_t)c.b * g.init_rows * g.W : nullptr;

    float a[9], oh[9], ow[9], go;
    auto load_inputs = [&](int it, bool& active, size_t& p) {
        const int y = c.y0 + warp + WARPS * (it / (TILE_W / 32));
        const int x = c.x0 + lane + 32 * (it % (TILE_W / 32));
        active = (y < g.H) && (x < g.W);
        p = (size_t)y * g.W + x;
        if (active) {
            go = ld_stream(gout_b + p);
#pragma unroll
            for (int k = 0; k < 9; ++k) a[k] = ld_stream(wgt_b + k * cs + p);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                oh[k] = ld_stream(off_b + (2 * k) * cs + p);
                ow[k] = ld_stream(off_b + (2 * k + 1) * cs + p);
            }
        }
    };

    float acc_w[9], acc_b = 0.f;  // this thread's share of grad_w / grad_b
#pragma unroll
    for (int k = 0; k < 9; ++k) acc_w[k] = 0.f;

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

        float s = 0.f;  // sum of raw affinities (NORM_SUM Jacobian)
        if (MODE == NORM_SUM) {
#pragma unroll
            for (int k = 0; k < 9; ++k) s += a[k];
        }
        normalise9<MODE>(a);  // a[] now holds the modulation m_k

        const float fy = (float)(g.row0 + c.y0 + ry), fx = (float)(c.x0 + cx);
        float gm[9];
        acc_b += go;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float h = (fy + (float)(k / 3 - 1)) + oh[k];
            const float w = (fx + (float)(k % 3 - 1)) + ow[k];
            const Tap t = gather_tap<T>(tile, init_b, g, c, h, w, nullptr);
            const float hh = 1.f - t.lh, hw = 1.f - t.lw;
            const float val = hh * hw * t.v1 + hh * t.lw * t.v2 + t.lh * hw * t.v3 + t.lh * t.lw * t.v4;
            // get_coordinate_weight
            const float dh = t.lw * (t.v4 - t.v2) + hw * (t.v3 - t.v1);
            const float dw = t.lh * (t.v4 - t.v3) + hh * (t.v2 - t.v1);
            const float gk = go * s_w[k];   // dL/d(column_k)
            const float gkm = gk * a[k];
            acc_w[k] += go * (a[k] * val);
            gm[k] = gk * val;
            float goh = gkm * dh, gow = gkm * dw;
            T* po = goff_b + (2 * k) * cs + p;
            if (accumulate) {
                goh += to_f32(po[0]);
                gow += to_f32(po[cs]);
            }
            st_stream(po, goh);
            st_stream(po + cs, gow);
            if (GRAD_INIT) {
                const float c1 = gkm * hh * hw, c2 = gkm * hh * t.lw, c3 = gkm * t.lh * hw, c4 = gkm * t.lh * t.lw;
                if (t.in_tile) {
                    float* gt = gtile + ((unsigned)t.h0 - (unsigned)c.oy) * SW + ((unsigned)t.w0 - (unsigned)c.ox);
                    atomicAdd(gt, c1);
                    atomicAdd(gt + 1, c2);
                    atomicAdd(gt + SW, c3);
                    atomicAdd(gt + SW + 1, c4);
                } else if (fabsf(h) < 1.0e9f && fabsf(w) < 1.0e9f) {
                    scatter_corner_global<T>(gi_b, g, t.h0, t.w0, c1);
                    scatter_corner_global<T>(gi_b, g, t.h0, t.w0 + 1, c2);
                    scatter_corner_global<T>(gi_b, g, t.h0 + 1, t.w0, c3);
                    scatter_corner_global<T>(gi_b, g, t.h0 + 1, t.w0 + 1, c4);
                }
            }
        }
        // Jacobian of the normalisation
        if (MODE == NORM_RESIDUAL) {
            float sg = gm[0];
#pragma unroll
            for (int k = 1; k < 9; ++k) sg += gm[k];
            const float mean = __fdiv_rn(sg, 9.f);
#pragma unroll
            for (int k = 0; k < 9; ++k) gm[k] -= mean;
            if (GRAD_INIT) atomicAdd(gtile + (ry + HALO_T) * SW + (cx + HALO_L), scale * go);
        } else if (MODE == NORM_SUM) {
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) dot += gm[k] * a[k];
#pragma unroll
            for (int k = 0; k < 9; ++k) gm[k] = __fdiv_rn(gm[k] - dot, s);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            T* pw = gwgt_b + k * cs + p;
            float v = gm[k];
            if (accumulate) v += to_f32(pw[0]);
            st_stream(pw, v);
        }
    }

    // ---- grad_init: flush the shared accumulation tile ----
    if (GRAD_INIT) {
        __syncthreads();
        for (int i = threadIdx.x; i < SH * SW; i += THREADS) {
            const float v = gtile[i];
            if (v != 0.f) {
                const int r = i / SW, q = i - r * SW;
                scatter_corner_global<T>(gi_b, g, c.oy + r, c.ox + q, v);
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
            if (threadIdx.x < 10) {
                const double v = atomicAdd(&ws->sums[threadIdx.x], 0.0);  // coherent read
                if (threadIdx.x < 9) grad_w9[threadIdx.x] = (float)v;
                else if (grad_b1 != nullptr) grad_b1[0] = (float)v;
                ws->sums[threadIdx.x] = 0.0;  // leave the workspace clean for the next call
            }
            if (threadIdx.x == 0) ws->ticket = 0u;
        }
    }
}

template <typename T, int MODE, bool TMA, bool GI>
static void launch_one(const LaunchArgs& la, dim3 grid) {
    spn_backward_kernel<T, MODE, TMA, GI><<<grid, THREADS, 0, la.stream>>>(
        (const T*)la.grad_out, (const T*)la.init, (const T*)la.weight, (const T*)la.offset, la.w9, la.grad_init,
        (T*)la.grad_weight, (T*)la.grad_offset, la.grad_w9, la.grad_b1, (ReduceWs*)la.workspace, la.g, la.scale,
        la.accumulate, la.tmap);
}

template <typename T, int MODE>
static cudaError_t launch_bwd_mode(const LaunchArgs& la) {
    dim3 grid((unsigned)((size_t)la.g.tiles_x * la.g.tiles_y * la.g.B));
    const bool gi = la.grad_init != nullptr;
    if (la.use_tma) {
        if (gi) launch_one<T, MODE, true, true>(la, grid);
        else launch_one<T, MODE, true, false>(la, grid);
    } else {
        if (gi) launch_one<T, MODE, false, true>(la, grid);
        else launch_one<T, MODE, false, false>(la, grid);
    }
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_bwd_dtype(const LaunchArgs& la) {
    switch (la.mode) {
        case NORM_NONE: return launch_bwd_mode<T, NORM_NONE>(la);
        case NORM_RESIDUAL: return launch_bwd_mode<T, NORM_RESIDUAL>(la);
        default: return launch_bwd_mode<T, NORM_SUM>(la);
    }
}

cudaError_t launch_spn_backward(const LaunchArgs& la) {
    return la.bf16 ? launch_bwd_dtype<__nv_bfloat16>(la) : launch_bwd_dtype<float>(la);
}

}  // namespace jspsr
