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
#include <cstdlib>

#include "../../include/jspsr_tiles.h"
#include "spn_common.cuh"

int jspsr_internal_fail(int code, const char* msg);  // abi.cu: sets the thread's last-error message

namespace jspsr {

// rows per CTA: template parameter LT_H in {32, 64}.  64 halves the share of halo rows every phase recomputes (staged
// rows 36/32 -> 68/64, sign rows marched 7 per 5 -> 12 per 10, adjoint rows 6 per 4 -> 10 per 8) at 55 KB of shared
// memory; 32 is the default (see the launch)
constexpr int LT_W = 128;   // columns per CTA (one float4 per lane)
constexpr int LD_W = LT_W + 8;      // staged difference columns: image column x0 - 4 + c (columns 2 .. 133 are used)
constexpr int LS_B = LT_W + 8;      // sign row stride in bytes: image column x0 - 4 + c (columns 3 .. 132 are used)
constexpr int loss_ld_h(int lt_h) { return lt_h + 4; }   // staged difference rows: image row y0 - 2 + r
constexpr int loss_ls_h(int lt_h) { return lt_h + 2; }   // sign rows: image row y0 - 1 + r
constexpr size_t loss_smem_bytes(int lt_h, bool grad) {
    return (size_t)loss_ld_h(lt_h) * LD_W * 4 + (grad ? 2 * (size_t)loss_ls_h(lt_h) * LS_B : 0);
}
constexpr int LOSS_SLOTS = 4;       // partial sums are spread over 4 slots of the workspace (fewer same-address atomics)

struct alignas(16) LossWs {
    double sums[LOSS_SLOTS][3];  // sum |d|, sum d^2, sum |Sobel(pred) - Sobel(gt)| * 8
    unsigned int ticket;
    unsigned int pad[3];
};
static_assert(sizeof(LossWs) <= sizeof(ReduceWs), "the propagation's reduction workspace is large enough");

// c * sign(v) (0 at v = +-0): c with its sign flipped by v's sign bit, or zero
__device__ __forceinline__ float csgn(float c, float v) {
    return (v != 0.f) ? __int_as_float(__float_as_int(c) ^ (__float_as_int(v) & 0x80000000)) : 0.f;
}
// sign + 1 in {0, 1, 2}: the sign tiles hold biased signs so that four of them add up bytewise in one 32-bit integer
__device__ __forceinline__ unsigned sgn1(float v) { return 1u + (unsigned)(v > 0.f) - (unsigned)(v < 0.f); }

// One CTA: LT_H x LT_W pixels of one plane, 8 warps; lane j owns the four columns x0 + 4j .. 4j + 3.
//
// d = pred - gt.  With kornia's replicate padding the Sobel pair is separable,
//     gx = Ay (x) Dx d,  gy = Dy (x) Ax d,  A = (1, 2, 1) with replicate border, D = (-1, 0, 1) with replicate border,
// and so is the adjoint the gradient needs: A^T s = (1, 2, 1) * s with s REPLICATED one past the border, D^T s =
// s[i - 1] - s[i + 1] with s NEGATED-and-replicated one past the border.  The sign tiles therefore carry one ring of
// such extended values around the image and every pixel, border or not, uses the same 12-tap formula (no divergence).
//
// FUSED (float4 path with the gradient): after d is staged there is no further CTA-wide step.  Each warp marches down
// its own LT_H / 8 pixel rows with the Sobel rows, the packed signs and the adjoint's row terms U, V all in registers;
// the signs of neighbouring columns come from the neighbouring lanes by shuffle, the two columns beside the tile from a
// per-warp pre-pass, and the ring one past the image border is applied in registers (x: lanes at the first / last image
// column; y: U(-1) = U(0), V(-1) = 8 - V(0) bytewise, same at the bottom).  No sign tile, no ring pass, one barrier.
template <bool WRITE_GRAD, bool VEC, int LT_H, bool FUSED>
__global__ void __launch_bounds__(THREADS, WRITE_GRAD ? 5 : 4)
loss_l1_l2_grad_kernel(const float* __restrict__ pred, const float* __restrict__ gt, float* __restrict__ grad,
                       float* __restrict__ losses4, LossWs* __restrict__ ws, int H, int W,
                       float w_l1, float w_l2, float w_grad, float inv_n) {
    constexpr int LD_H = loss_ld_h(LT_H), LS_H = loss_ls_h(LT_H);
    extern __shared__ __align__(16) unsigned char loss_smem[];
    float (*s_d)[LD_W] = reinterpret_cast<float (*)[LD_W]>(loss_smem);
    unsigned char (*s_sx)[LS_B] = reinterpret_cast<unsigned char (*)[LS_B]>(loss_smem + (size_t)LD_H * LD_W * 4);
    unsigned char (*s_sy)[LS_B] = s_sx + LS_H;      // (both only exist with WRITE_GRAD)
    __shared__ float s_red[WARPS][3];
    __shared__ bool s_last;

    // grid = (planes, tiles_x, tiles_y): no per-thread division to find the tile
    const size_t plane = (size_t)blockIdx.x * H * W;
    const int y0 = blockIdx.z * LT_H, x0 = blockIdx.y * LT_W;
    const float* __restrict__ p = pred + plane;
    const float* __restrict__ g = gt + plane;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    float a_l1 = 0.f, a_l2 = 0.f, a_grad = 0.f;

    // ---- phase 1: d over rows y0 - 2 .. y0 + LT_H + 1, columns x0 - 2 .. x0 + LT_W + 1, replicate-clamped ----
    if (VEC) {
        // W % 4 == 0 and 16-byte aligned planes: a lane's float4 is entirely inside or entirely outside the image
        constexpr int ROWS_PER_WARP = (LD_H + WARPS - 1) / WARPS;   // 5 (LT_H = 32) or 9 (64)
        constexpr int CHUNK = 5;                                    // rows in flight per lane: 10 float4 loads
        const int xc = x0 + 4 * lane;
        const bool in_x = xc < W;
        const int xl = in_x ? xc : W - 4;                   // past the image: the last float4, its .w replicated below
        // the two halo columns on either side: LD_H rows x 4 columns (threads 0 .. 4 * LD_H - 1, two passes at most)
        float hp[2] = {0.f, 0.f}, hg[2] = {0.f, 0.f};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int hi = threadIdx.x + u * THREADS;
            const int hr = hi >> 2, hq = hi & 3;                           // row, which of the 4 columns
            const int hc = (hq < 2) ? (2 + hq) : (LT_W + 2 + hq);          // staged column 2, 3, 132, 133
            if (hr < LD_H) {
                const int y = min(max(y0 - 2 + hr, 0), H - 1), x = min(max(x0 - 4 + hc, 0), W - 1);
                hp[u] = __ldg(p + (size_t)y * W + x);
                hg[u] = __ldg(g + (size_t)y * W + x);
            }
        }
#pragma unroll
        for (int i0 = 0; i0 < ROWS_PER_WARP; i0 += CHUNK) {
            float4 vp[CHUNK], vg[CHUNK];
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                const int r = min(warp + WARPS * (i0 + i), LD_H - 1);
                const int y = min(max(y0 - 2 + r, 0), H - 1);
                vp[i] = __ldcs(reinterpret_cast<const float4*>(p + (size_t)y * W + xl));
                vg[i] = __ldcs(reinterpret_cast<const float4*>(g + (size_t)y * W + xl));
            }
#pragma unroll
            for (int i = 0; i < CHUNK; ++i) {
                const int r = warp + WARPS * (i0 + i);
                if (i0 + i < ROWS_PER_WARP && r < LD_H) {
                    float4 d = make_float4(vp[i].x - vg[i].x, vp[i].y - vg[i].y, vp[i].z - vg[i].z, vp[i].w - vg[i].w);
                    if (!in_x) d.x = d.y = d.z = d.w;
                    *reinterpret_cast<float4*>(&s_d[r][4 + 4 * lane]) = d;
                    if (r >= 2 && r < LT_H + 2 && y0 - 2 + r < H && in_x) {       // own pixels: L1 and L2 from registers
                        a_l1 += (fabsf(d.x) + fabsf(d.y)) + (fabsf(d.z) + fabsf(d.w));
                        a_l2 += (d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int hi = threadIdx.x + u * THREADS;
            const int hr = hi >> 2, hq = hi & 3;
            const int hc = (hq < 2) ? (2 + hq) : (LT_W + 2 + hq);
            if (hr < LD_H) s_d[hr][hc] = hp[u] - hg[u];
        }
    } else {
        for (int i = threadIdx.x; i < LD_H * (LT_W + 4); i += THREADS) {
            const int r = i / (LT_W + 4), c = 2 + (i - r * (LT_W + 4));
            const int yy = y0 - 2 + r, xx = x0 - 4 + c;
            const int y = min(max(yy, 0), H - 1), x = min(max(xx, 0), W - 1);
            const size_t o = (size_t)y * W + x;
            const float d = __ldg(p + o) - __ldg(g + o);
            s_d[r][c] = d;
            if (r >= 2 && r < LT_H + 2 && c >= 4 && c < LT_W + 4 && yy < H && xx < W) {
                a_l1 += fabsf(d);
                a_l2 += d * d;
            }
        }
    }
    __syncthreads();

    if constexpr (FUSED) {
        static_assert(WRITE_GRAD && VEC, "the fused march is the float4 gradient path");
        constexpr int PR = LT_H / WARPS;          // pixel rows per warp
        constexpr int NS = PR + 2;                // sign rows per warp: pixel rows pr0 - 1 .. pr0 + PR
        __shared__ unsigned short s_edge[WARPS][NS][2];
        const int pr0 = warp * PR;
        if (lane < 2 * NS) {
            // signs of the columns just left / right of the tile (image columns x0 - 1, x0 + LT_W) for this warp's rows
            const int side = lane / NS, jr = lane - side * NS;
            const int r = pr0 + jr, c = side ? LT_W + 4 : 3;
            const float d00 = s_d[r][c - 1], d01 = s_d[r][c], d02 = s_d[r][c + 1];
            const float d10 = s_d[r + 1][c - 1], d12 = s_d[r + 1][c + 1];
            const float d20 = s_d[r + 2][c - 1], d21 = s_d[r + 2][c], d22 = s_d[r + 2][c + 1];
            const unsigned vx = sgn1(((d02 - d00) + (d22 - d20)) + 2.f * (d12 - d10));
            const unsigned vy = sgn1(((d20 + d22) + 2.f * d21) - ((d00 + d02) + 2.f * d01));
            s_edge[warp][jr][side] = (unsigned short)(vx | (vy << 8));
        }
        __syncwarp();
        const float c_pix_l1 = w_l1 * inv_n, c_pix_l2 = 2.f * w_l2 * inv_n, c_sob = w_grad * 0.5f * inv_n * 0.125f;
        const int xl = x0 + 4 * lane;
        const bool in_x = xl < W;
        float hd[3][4], hs[3][4];
        unsigned U[3], V[3];
#pragma unroll
        for (int i = 0; i < NS + 2; ++i) {
            {
                const int dr = pr0 + i;                      // staged d row: image row y0 - 2 + dr  (dr <= LD_H - 1)
                const float4 m = *reinterpret_cast<const float4*>(&s_d[dr][4 + 4 * lane]);
                float l = __shfl_up_sync(0xffffffffu, m.w, 1);
                float r = __shfl_down_sync(0xffffffffu, m.x, 1);
                if (lane == 0) l = s_d[dr][3];
                if (lane == 31) r = s_d[dr][LT_W + 4];
                const float e[6] = {l, m.x, m.y, m.z, m.w, r};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    hd[i % 3][q] = e[q + 2] - e[q];
                    hs[i % 3][q] = (e[q] + e[q + 2]) + 2.f * e[q + 1];
                }
            }
            if (i >= 2) {
                const int jr = i - 2;                        // sign row: image row y = y0 - 1 + pr0 + jr
                const int y = y0 - 1 + pr0 + jr;
                unsigned px = 0x01010101u, py = 0x01010101u;
                float a_row = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float gx = (hd[jr % 3][q] + hd[(jr + 2) % 3][q]) + 2.f * hd[(jr + 1) % 3][q];
                    const float gy = hs[(jr + 2) % 3][q] - hs[jr % 3][q];
                    a_row += fabsf(gx) + fabsf(gy);
                    px += (gx > 0.f) ? (1u << (8 * q)) : 0u;
                    px -= (gx < 0.f) ? (1u << (8 * q)) : 0u;
                    py += (gy > 0.f) ? (1u << (8 * q)) : 0u;
                    py -= (gy < 0.f) ? (1u << (8 * q)) : 0u;
                }
                if (jr >= 1 && jr <= PR) a_grad += (y >= 0 && y < H && in_x) ? a_row : 0.f;   // this warp's own rows
                // the neighbouring columns' signs: lanes beside, the tile's side columns, or the ring at the image border
                unsigned lx = __shfl_up_sync(0xffffffffu, px, 1), rx = __shfl_down_sync(0xffffffffu, px, 1);
                unsigned ly = __shfl_up_sync(0xffffffffu, py, 1), ry = __shfl_down_sync(0xffffffffu, py, 1);
                if (lane == 0) {
                    const unsigned e = s_edge[warp][jr][0];
                    lx = (e & 0xffu) << 24;
                    ly = (e >> 8) << 24;
                }
                if (lane == 31) {
                    const unsigned e = s_edge[warp][jr][1];
                    rx = e & 0xffu;
                    ry = e >> 8;
                }
                if (xl == 0) {                               // x = -1: d/dx sign negated, d/dy sign replicated
                    lx = (2u - (px & 0xffu)) << 24;
                    ly = (py & 0xffu) << 24;
                }
                if (xl + 4 == W) {                           // x = W
                    rx = 2u - (px >> 24);
                    ry = py >> 24;
                }
                const unsigned Lx = __funnelshift_r(lx, px, 24), Rx = __funnelshift_r(px, rx, 8);
                const unsigned Ly = __funnelshift_r(ly, py, 24), Ry = __funnelshift_r(py, ry, 8);
                unsigned Ur = Lx + (0x02020202u - Rx);       // (sx[x - 1] - sx[x + 1]) + 2 per byte
                unsigned Vr = Ly + 2u * py + Ry;             // (sy[x - 1] + 2 sy[x] + sy[x + 1]) + 4 per byte
                if (jr >= 1 && y == H) {                     // the row below the image: sx replicated, sy negated
                    Ur = U[(jr + 2) % 3];
                    Vr = 0x08080808u - V[(jr + 2) % 3];
                }
                U[jr % 3] = Ur;
                V[jr % 3] = Vr;
                if (jr >= 1 && y == 0) {                     // the row above the image, derived from row 0
                    U[(jr + 2) % 3] = Ur;
                    V[(jr + 2) % 3] = 0x08080808u - Vr;
                }
                if (jr >= 2) {
                    const int pr = pr0 + jr - 2, yp = y0 + pr;   // pixel row between sign rows jr - 2 and jr
                    // (sum + 16) per byte: U0 + 2 U1 + U2 carries 8, V0 + (8 - V2) carries 8
                    const unsigned G = U[(jr + 1) % 3] + 2u * U[(jr + 2) % 3] + U[jr % 3] + V[(jr + 1) % 3] + (0x08080808u - V[jr % 3]);
                    const float4 d = *reinterpret_cast<const float4*>(&s_d[pr + 2][4 + 4 * lane]);
                    float4 o;
                    o.x = csgn(c_pix_l1, d.x) + c_pix_l2 * d.x + c_sob * ((float)(G & 0xffu) - 16.f);
                    o.y = csgn(c_pix_l1, d.y) + c_pix_l2 * d.y + c_sob * ((float)((G >> 8) & 0xffu) - 16.f);
                    o.z = csgn(c_pix_l1, d.z) + c_pix_l2 * d.z + c_sob * ((float)((G >> 16) & 0xffu) - 16.f);
                    o.w = csgn(c_pix_l1, d.w) + c_pix_l2 * d.w + c_sob * ((float)(G >> 24) - 16.f);
                    if (yp < H && in_x) __stcs(reinterpret_cast<float4*>(grad + plane + (size_t)yp * W + xl), o);
                }
            }
        }
    } else {
    // ---- phase 2: Sobel of d, its |.| sum and (for the gradient) its biased signs ----
    // Warps 0 .. 6 march down five sign rows each with the horizontal difference / smoothing of three d rows in
    // registers (fully unrolled: the rolling window is renamed, not moved); warp 7 takes the two sign columns
    // beside the tile.  Sign cells outside the image are left as computed from the clamped d: the cells one step
    // outside are overwritten by the ring pass below and nothing reads the ones further out.
    constexpr int RPW = (LS_H + WARPS - 2) / (WARPS - 1);   // sign rows per marching warp: 5 (LT_H = 32) or 10 (64)
    static_assert(RPW * (WARPS - 1) >= LS_H, "phase-2 row split");
    if (warp < WARPS - 1) {
        const int ra = warp * RPW;                           // sign rows ra .. ra + RPW - 1: image row y0 - 1 + r
        float hd[3][4], hs[3][4];                            // d rows i - 2, i - 1, i (slot = row % 3)
#pragma unroll
        for (int i = 0; i < RPW + 2; ++i) {
            {
                const int dr = min(ra + i, LD_H - 1);
                // d at image columns x0 + 4 lane - 1 .. + 4: own float4, one value from each neighbour lane
                const float4 m = *reinterpret_cast<const float4*>(&s_d[dr][4 + 4 * lane]);
                float l = __shfl_up_sync(0xffffffffu, m.w, 1);
                float r = __shfl_down_sync(0xffffffffu, m.x, 1);
                if (lane == 0) l = s_d[dr][3];
                if (lane == 31) r = s_d[dr][LT_W + 4];
                const float e[6] = {l, m.x, m.y, m.z, m.w, r};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    hd[i % 3][q] = e[q + 2] - e[q];
                    hs[i % 3][q] = (e[q] + e[q + 2]) + 2.f * e[q + 1];
                }
            }
            if (i >= 2) {
                const int j = i - 2;                         // sign row ra + j from d rows j, j + 1, j + 2
                const int r = ra + j;
                const int y = y0 - 1 + r;
                const bool own = (y >= 0 && y < H) && r >= 1 && r <= LT_H;     // a row of this CTA's own pixels
                unsigned px = 0x01010101u, py = 0x01010101u;
                float a_row = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float gx = (hd[j % 3][q] + hd[(j + 2) % 3][q]) + 2.f * hd[(j + 1) % 3][q];
                    const float gy = hs[(j + 2) % 3][q] - hs[j % 3][q];
                    const float ag = fabsf(gx) + fabsf(gy);
                    if (VEC) a_row += ag;
                    else a_row += (x0 + 4 * lane + q < W) ? ag : 0.f;
                    if (WRITE_GRAD) {
                        px += (gx > 0.f) ? (1u << (8 * q)) : 0u;
                        px -= (gx < 0.f) ? (1u << (8 * q)) : 0u;
                        py += (gy > 0.f) ? (1u << (8 * q)) : 0u;
                        py -= (gy < 0.f) ? (1u << (8 * q)) : 0u;
                    }
                }
                a_grad += (own && (!VEC || x0 + 4 * lane < W)) ? a_row : 0.f;
                if (WRITE_GRAD && r < LS_H) {
                    *reinterpret_cast<unsigned*>(&s_sx[r][4 + 4 * lane]) = px;
                    *reinterpret_cast<unsigned*>(&s_sy[r][4 + 4 * lane]) = py;
                }
            }
        }
    } else if (WRITE_GRAD) {
        // the sign columns just left and right of the tile (image columns x0 - 1 and x0 + LT_W)
        for (int i = lane; i < 2 * LS_H; i += 32) {
            const int side = i / LS_H, r = i - side * LS_H;
            const int c = side ? LT_W + 4 : 3;                          // staged column in s_d and in the sign tiles
            const float d00 = s_d[r][c - 1], d01 = s_d[r][c], d02 = s_d[r][c + 1];
            const float d10 = s_d[r + 1][c - 1], d12 = s_d[r + 1][c + 1];
            const float d20 = s_d[r + 2][c - 1], d21 = s_d[r + 2][c], d22 = s_d[r + 2][c + 1];
            s_sx[r][c] = (unsigned char)sgn1(((d02 - d00) + (d22 - d20)) + 2.f * (d12 - d10));
            s_sy[r][c] = (unsigned char)sgn1(((d20 + d22) + 2.f * d21) - ((d00 + d02) + 2.f * d01));
        }
    }

    if (WRITE_GRAD) {
        __syncthreads();
        // ---- the ring one past the image border (only the CTAs that see it): replicate, negated along the
        //      axis the operator differentiates ----
        const int sc_r = W - x0 + 4, sr_b = H - y0 + 1;                 // staged column of x = W, staged row of y = H
        const bool left = (x0 == 0), right = (sc_r <= LT_W + 4), top = (y0 == 0), bottom = (sr_b <= LS_H - 1);
        if (left || right || top || bottom) {
            for (int i = threadIdx.x; i < 2 * LS_H + 2 * (LT_W + 2); i += THREADS) {
                int r, c;
                bool on;
                if (i < 2 * LS_H) {                                     // columns x = -1 and x = W, image rows only
                    const int side = i / LS_H;
                    r = i - side * LS_H;
                    c = side ? sc_r : 3;
                    const int y = y0 - 1 + r;
                    on = (side ? right : left) && y >= 0 && y < H;
                } else {                                                // rows y = -1 and y = H, x = -1 .. W
                    const int k = i - 2 * LS_H;
                    const int side = k / (LT_W + 2);
                    c = 3 + (k - side * (LT_W + 2));
                    r = side ? sr_b : 0;
                    const int x = x0 - 4 + c;
                    on = (side ? bottom : top) && x >= -1 && x <= W;
                }
                if (on) {
                    const int y = y0 - 1 + r, x = x0 - 4 + c;
                    const int yc = min(max(y, 0), H - 1), xc = min(max(x, 0), W - 1);
                    const unsigned vx = s_sx[yc - y0 + 1][xc - x0 + 4], vy = s_sy[yc - y0 + 1][xc - x0 + 4];
                    s_sx[r][c] = (unsigned char)((x != xc) ? 2u - vx : vx);
                    s_sy[r][c] = (unsigned char)((y != yc) ? 2u - vy : vy);
                }
            }
            // (every ring cell is derived from an IMAGE cell, corners included, so one pass needs no ordering)
        }
        __syncthreads();
    }

    if (WRITE_GRAD) {
        // ---- phase 3: dTotal/dpred; a warp marches down LT_H / WARPS pixel rows, four pixels per lane ----
        const float c_pix_l1 = w_l1 * inv_n, c_pix_l2 = 2.f * w_l2 * inv_n, c_sob = w_grad * 0.5f * inv_n * 0.125f;
        constexpr int PR = LT_H / WARPS;
        const int pr0 = warp * PR;                                    // pixel rows pr0 .. pr0 + PR - 1 (sign rows + 1)
        // U = (sx[x - 1] - sx[x + 1]) + 2 and V = (sy[x - 1] + 2 sy[x] + sy[x + 1]) + 4 per byte, for one sign row
        auto row_uv = [&](int sr, unsigned& U, unsigned& V) {
            const unsigned mx = *reinterpret_cast<const unsigned*>(&s_sx[sr][4 + 4 * lane]);
            const unsigned my = *reinterpret_cast<const unsigned*>(&s_sy[sr][4 + 4 * lane]);
            unsigned lx = __shfl_up_sync(0xffffffffu, mx, 1), rx = __shfl_down_sync(0xffffffffu, mx, 1);
            unsigned ly = __shfl_up_sync(0xffffffffu, my, 1), ry = __shfl_down_sync(0xffffffffu, my, 1);
            if (lane == 0) {
                lx = *reinterpret_cast<const unsigned*>(&s_sx[sr][0]);
                ly = *reinterpret_cast<const unsigned*>(&s_sy[sr][0]);
            }
            if (lane == 31) {
                rx = *reinterpret_cast<const unsigned*>(&s_sx[sr][LT_W + 4]);
                ry = *reinterpret_cast<const unsigned*>(&s_sy[sr][LT_W + 4]);
            }
            const unsigned Lx = __funnelshift_r(lx, mx, 24), Rx = __funnelshift_r(mx, rx, 8);
            const unsigned Ly = __funnelshift_r(ly, my, 24), Ry = __funnelshift_r(my, ry, 8);
            U = Lx + (0x02020202u - Rx);
            V = Ly + 2u * my + Ry;
        };
        unsigned U0, V0, U1, V1, U2, V2;
        row_uv(pr0, U0, V0);
        row_uv(pr0 + 1, U1, V1);
#pragma unroll
        for (int i = 0; i < PR; ++i) {
            const int pr = pr0 + i, y = y0 + pr;
            row_uv(pr + 2, U2, V2);
            // (sum + 16) per byte: U0 + 2 U1 + U2 carries 8, V0 + (8 - V2) carries 8
            const unsigned G = U0 + 2u * U1 + U2 + V0 + (0x08080808u - V2);
            const float4 d = *reinterpret_cast<const float4*>(&s_d[pr + 2][4 + 4 * lane]);
            float4 o;
            o.x = csgn(c_pix_l1, d.x) + c_pix_l2 * d.x + c_sob * ((float)(G & 0xffu) - 16.f);
            o.y = csgn(c_pix_l1, d.y) + c_pix_l2 * d.y + c_sob * ((float)((G >> 8) & 0xffu) - 16.f);
            o.z = csgn(c_pix_l1, d.z) + c_pix_l2 * d.z + c_sob * ((float)((G >> 16) & 0xffu) - 16.f);
            o.w = csgn(c_pix_l1, d.w) + c_pix_l2 * d.w + c_sob * ((float)(G >> 24) - 16.f);
            const int x = x0 + 4 * lane;
            if (y < H) {
                float* __restrict__ dst = grad + plane + (size_t)y * W + x;
                if (VEC) {
                    if (x < W) __stcs(reinterpret_cast<float4*>(dst), o);
                } else {
                    if (x < W) __stcs(dst, o.x);
                    if (x + 1 < W) __stcs(dst + 1, o.y);
                    if (x + 2 < W) __stcs(dst + 2, o.z);
                    if (x + 3 < W) __stcs(dst + 3, o.w);
                }
            }
            U0 = U1; U1 = U2;
            V0 = V1; V1 = V2;
        }
    }

    }  // !FUSED

    // ---- thread -> warp -> CTA -> fp64 atomics; the last CTA publishes the four losses ----
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
        atomicAdd(&ws->sums[(blockIdx.x + blockIdx.z) % LOSS_SLOTS][threadIdx.x], v);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&ws->ticket, 1u);
        s_last = (t == gridDim.x * gridDim.y * gridDim.z - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        double tot[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            tot[k] = 0.0;
#pragma unroll
            for (int sl = 0; sl < LOSS_SLOTS; ++sl) {
                tot[k] += atomicAdd(&ws->sums[sl][k], 0.0);
                ws->sums[sl][k] = 0.0;                         // leave the workspace clean for the next call
            }
        }
        const double n_inv = (double)inv_n;
        const double l1 = tot[0] * n_inv;
        const double l2 = tot[1] * n_inv;
        const double gr = tot[2] * n_inv * 0.5 * 0.125;
        losses4[0] = (float)l1;
        losses4[1] = (float)l2;
        losses4[2] = (float)gr;
        losses4[3] = (float)((double)w_l1 * l1 + (double)w_l2 * l2 + (double)w_grad * gr);
        ws->ticket = 0u;
    }
}

// (pred, gt) -> de-normalised elevation difference of one pixel (MeterBase._prepare clamps pred only)
template <bool ELEV_LOG>
__device__ __forceinline__ float denorm_diff(float p, float g, float log_range, float range, float vmin) {
    const float pc = fminf(fmaxf(p, 0.f), 1.f);
    float pe, ge;
    if (ELEV_LOG) {
        pe = expf(pc * log_range) + vmin;
        ge = expf(g * log_range) + vmin;
    } else {
        pe = __fadd_rn(__fmul_rn(pc, range), vmin);
        ge = __fadd_rn(__fmul_rn(g, range), vmin);
    }
    return pe - ge;
}

// One CTA: a slab of rows of one sample's border-cropped window.  sums[b] = {sum d^2, sum |d|} (fp64 atomics).
//
// VEC (W % 4 == 0, 16-byte aligned planes): a lane owns four consecutive columns of the 16-byte aligned span that
// covers the window and reads them with one 128-bit load per tensor; the (at most three) columns of the first and last
// word that belong to the border are dropped by a per-lane mask that does not depend on the row, and a warp has
// MET_ROWS rows in flight (2 x MET_ROWS 128-bit loads per lane before the first exponential).  The scalar kernel spent
// 16 of its 43 instructions per pixel on predicated addresses and per-pixel branches and sat at its issue limit
// (0.53 of the HBM rate at 4096 tiles); this form executes about half as many.
// !VEC: a warp per row, lanes along x, four columns of a row in flight per lane.
// Per-pixel arithmetic is the same expression in both, and the fp32 partial sums are folded into fp64 after at most
// a few dozen pixels, so the two differ only in the association of those short fp32 sums.
constexpr int MET_ROWS = 4;
template <bool ELEV_LOG, bool VEC>
__global__ void __launch_bounds__(THREADS)
dem_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, double* __restrict__ sums, int H, int W,
                   int bh, int bw, float log_range, float range, float vmin, int rows_per_cta) {
    __shared__ double s_red[WARPS][2];
    const int b = blockIdx.y;
    const int hc = H - 2 * bh, wc = W - 2 * bw;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, hc);
    const size_t plane = (size_t)b * H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double d_sq = 0.0, d_ab = 0.0;
    if (VEC) {
        const int ca = bw & ~3;                           // first column of the aligned span
        const int n4 = (W - bw + 3 - ca) >> 2;            // 128-bit words covering columns [bw, W - bw)
        for (int r = r0 + warp; r < r1; r += WARPS * MET_ROWS) {
            float a_sq = 0.f, a_ab = 0.f;
            for (int q = lane; q < n4; q += 32) {
                const int c = ca + 4 * q;
                float4 pv[MET_ROWS], gv[MET_ROWS];
#pragma unroll
                for (int u = 0; u < MET_ROWS; ++u) {
                    // rows past the slab re-read its last row (valid memory) and are skipped below (warp-uniform)
                    const size_t o = plane + (size_t)(bh + min(r + u * WARPS, r1 - 1)) * W + c;
                    pv[u] = __ldcs(reinterpret_cast<const float4*>(pred + o));
                    gv[u] = __ldcs(reinterpret_cast<const float4*>(gt + o));
                }
                const bool m0 = c >= bw && c < W - bw, m1 = c + 1 >= bw && c + 1 < W - bw;
                const bool m2 = c + 2 >= bw && c + 2 < W - bw, m3 = c + 3 >= bw && c + 3 < W - bw;
#pragma unroll
                for (int u = 0; u < MET_ROWS; ++u) {
                    if (r + u * WARPS < r1) {
                        float d;
                        d = denorm_diff<ELEV_LOG>(pv[u].x, gv[u].x, log_range, range, vmin);
                        d = m0 ? d : 0.f;   // a select, not a product: border pixels may hold anything
                        a_sq = fmaf(d, d, a_sq);
                        a_ab += fabsf(d);
                        d = denorm_diff<ELEV_LOG>(pv[u].y, gv[u].y, log_range, range, vmin);
                        d = m1 ? d : 0.f;
                        a_sq = fmaf(d, d, a_sq);
                        a_ab += fabsf(d);
                        d = denorm_diff<ELEV_LOG>(pv[u].z, gv[u].z, log_range, range, vmin);
                        d = m2 ? d : 0.f;
                        a_sq = fmaf(d, d, a_sq);
                        a_ab += fabsf(d);
                        d = denorm_diff<ELEV_LOG>(pv[u].w, gv[u].w, log_range, range, vmin);
                        d = m3 ? d : 0.f;
                        a_sq = fmaf(d, d, a_sq);
                        a_ab += fabsf(d);
                    }
                }
            }
            // fold the fp32 partials (4 x MET_ROWS pixels per 128 columns) into fp64 so that long windows do not lose low bits
            d_sq += (double)a_sq;
            d_ab += (double)a_ab;
        }
    } else {
        for (int r = r0 + warp; r < r1; r += WARPS) {
            const size_t row = plane + (size_t)(bh + r) * W + bw;
            float a_sq = 0.f, a_ab = 0.f;
            for (int c0 = lane; c0 < wc; c0 += 128) {
                float pv[4], gv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int c = c0 + 32 * u;
                    pv[u] = (c < wc) ? __ldcs(pred + row + c) : 0.f;
                    gv[u] = (c < wc) ? __ldcs(gt + row + c) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (c0 + 32 * u < wc) {
                        const float d = denorm_diff<ELEV_LOG>(pv[u], gv[u], log_range, range, vmin);
                        a_sq = fmaf(d, d, a_sq);
                        a_ab += fabsf(d);
                    }
                }
            }
            d_sq += (double)a_sq;
            d_ab += (double)a_ab;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d_sq += __shfl_xor_sync(0xffffffffu, d_sq, o);
        d_ab += __shfl_xor_sync(0xffffffffu, d_ab, o);
    }
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
    const int tiles_x = (W + LT_W - 1) / LT_W;
    // 32-row tiles.  The 64-row instantiation halves the halo share but measured no better on B200 (4096 tiles: 0.223 ms
    // vs 0.201 ms with the gradient, 0.119 vs 0.121 ms without: longer phases between the CTA barriers at the same four
    // resident CTAs); it stays reachable through JSPSR_LOSS_TILE_H=64 (tests hold the two to identical bits)
    int lt_h = 32;
    // float4 path: rows 16-byte aligned (W % 4 == 0 and aligned bases)
    const bool vec = (W % 4 == 0) && !(((uintptr_t)pred | (uintptr_t)gt | (uintptr_t)grad_pred) & 15);
    // With the gradient on the float4 path there are two kernels of identical per-pixel arithmetic (tests: equal bits):
    // the phased one (sign tiles in shared memory) and the fused per-warp march.  Measured on B200, 4096 tiles, same box:
    // phased/32 rows 0.188 ms, fused/32 0.197 ms, fused/64 0.183 ms, phased/64 0.223 ms - so large grids take the fused
    // march on 64-row tiles and everything else the phased kernel on 32-row tiles.  JSPSR_LOSS_FUSED / JSPSR_LOSS_TILE_H
    // override (tests).
    bool fused = vec && grad_pred && H > 32 && (long long)planes * tiles_x * ((H + 63) / 64) >= 2LL * 5 * 148;
    if (const char* ev = getenv("JSPSR_LOSS_FUSED")) fused = vec && grad_pred && atoi(ev) != 0;
    if (fused) lt_h = 64;
    if (const char* ev = getenv("JSPSR_LOSS_TILE_H")) {
        if (atoi(ev) == 32 || atoi(ev) == 64) lt_h = atoi(ev);
    }
    const int tiles_y = (H + lt_h - 1) / lt_h;
    if (tiles_x > 65535 || tiles_y > 65535) return jspsr_internal_fail(JSPSR_ERR_UNSUPPORTED, "loss: plane too large");
    if ((long long)planes * tiles_x * tiles_y > 0xffffffffLL) return jspsr_internal_fail(JSPSR_ERR_UNSUPPORTED, "loss: more than 2^32 tiles");
    const dim3 ctas((unsigned)planes, (unsigned)tiles_x, (unsigned)tiles_y);
    const float inv_n = (float)(1.0 / ((double)planes * H * W));
    cudaError_t se = cudaSuccess;
#define JSPSR_LAUNCH_LOSS(G, V, TH, FU)                                                                   \
    do {                                                                                                  \
        constexpr size_t smem = loss_smem_bytes(TH, (G) && !(FU));                                        \
        if (smem > 48 * 1024) se = ensure_dynamic_smem((const void*)loss_l1_l2_grad_kernel<G, V, TH, FU>, smem); \
        if (se == cudaSuccess)                                                                            \
            loss_l1_l2_grad_kernel<G, V, TH, FU><<<ctas, THREADS, smem, (cudaStream_t)stream>>>(         \
                pred, gt, grad_pred, losses4, (LossWs*)workspace, H, W, w_l1, w_l2, w_grad, inv_n);       \
    } while (0)
#define JSPSR_LAUNCH_LOSS_TH(G, V, FU)                                      \
    do {                                                                    \
        if (lt_h == 64) JSPSR_LAUNCH_LOSS(G, V, 64, FU); else JSPSR_LAUNCH_LOSS(G, V, 32, FU); \
    } while (0)
    if (grad_pred) {
        if (fused) JSPSR_LAUNCH_LOSS_TH(true, true, true);
        else if (vec) JSPSR_LAUNCH_LOSS_TH(true, true, false);
        else JSPSR_LAUNCH_LOSS_TH(true, false, false);
    } else {
        if (vec) JSPSR_LAUNCH_LOSS_TH(false, true, false); else JSPSR_LAUNCH_LOSS_TH(false, false, false);
    }
#undef JSPSR_LAUNCH_LOSS_TH
#undef JSPSR_LAUNCH_LOSS
    if (se != cudaSuccess) {
        char msg[256];
        snprintf(msg, sizeof(msg), "loss kernel shared-memory opt-in: %s", cudaGetErrorString(se));
        return jspsr_internal_fail(JSPSR_ERR_CUDA, msg);
    }
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
        const bool vec = (W % 4 == 0) && !(((uintptr_t)pred | (uintptr_t)gt) & 15);
        // enough CTAs for four full waves (8 resident per SM: 4096 one-CTA samples would be 3.46 waves, the last one
        // half empty), at least one row per warp each; the float4 kernel keeps MET_ROWS rows per warp in flight
        int slabs = (4 * 148 * 8 + B - 1) / B;
        const int min_rows = vec ? WARPS * MET_ROWS : WARPS;
        slabs = max(1, min(slabs, (hc + min_rows - 1) / min_rows));
        const int rows_per_cta = (hc + slabs - 1) / slabs;
        slabs = (hc + rows_per_cta - 1) / rows_per_cta;
        const float range = (float)((double)value_max - (double)value_min);
        // data * log(max - min): the reference multiplies by the python float (double) rounded into the fp32 tensor op
        const float log_range = elev_log ? (float)log((double)value_max - (double)value_min) : 0.f;
        const dim3 ctas((unsigned)slabs, (unsigned)B);
#define JSPSR_LAUNCH_METRICS(E, V)                                                                  \
    dem_metrics_kernel<E, V><<<ctas, THREADS, 0, (cudaStream_t)stream>>>(pred, gt, sums, H, W, border_h, border_w, \
                                                                       log_range, range, value_min, rows_per_cta)
        if (elev_log) {
            if (vec) JSPSR_LAUNCH_METRICS(true, true); else JSPSR_LAUNCH_METRICS(true, false);
        } else {
            if (vec) JSPSR_LAUNCH_METRICS(false, true); else JSPSR_LAUNCH_METRICS(false, false);
        }
#undef JSPSR_LAUNCH_METRICS
        ce = cudaGetLastError();
    }
    if (ce != cudaSuccess) {
        char msg[256];
        snprintf(msg, sizeof(msg), "metrics kernel launch: %s", cudaGetErrorString(ce));
        return jspsr_internal_fail(JSPSR_ERR_CUDA, msg);
    }
    return JSPSR_OK;
}
