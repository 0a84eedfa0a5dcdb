// The fixed-affinity loop of NLSPN.forward (models/components/nlspn.py:222-235) as ONE launch: all T applications of a
// sample run on chip, the affinities / offsets are read from HBM once.
//
//   feat_{t+1}(p) = sum_k aff_k(p) * bilinear(feat_t, p + k + offset_k(p))        t = 0 .. T-1, aff / offset fixed
//
// What makes the loop fusable: aff and offset do not change, so everything a tap derives from them - its position, the
// integer cell, the two bilinear fractions, the address of its 2x2 footprint in the staged tile, whether it leaves the
// tile - is ITERATION-INVARIANT.  A thread computes that once for its two pixels and keeps it in registers (4 words per
// tap); an application is then 4 shared loads + 6 flops + 2 per tap, no global loads, no conversions, no range tests.
//
// Layout: one sample (H, W <= 128) per 16-CTA cluster (non-portable size), CTA r owns rows [8r, 8r + 8); 512 threads, two
// pixels each.  The feature lives in shared memory, double buffered: every CTA holds its 8 rows plus the 6 rows above and
// 7 below (the narrow staged halo of spn_common.cuh) and 8 zero columns either side.  After computing a pixel the thread
// stores it into its own next buffer, into the neighbouring CTAs' halo rows through distributed shared memory
// (st.shared::cluster), and into list_out[t]; one cluster barrier (release / acquire) per application orders the lot.
// Taps that leave the staged tile (|row offset| > 5 or so) take the same bounds-checked global path as the plain kernel,
// reading list_out[t - 1], which the barrier has made visible.
//
// Arithmetic and summation order are those of spn_forward_kernel in NORM_NONE mode with w = 1, b = 0, so list_out is
// bit-identical to T launches (tested).  fp32, no preserve_input; everything else falls back to the T-launch loop.
#include <cooperative_groups.h>

#include "spn_kernels.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

namespace cg = cooperative_groups;

constexpr int FI_ROWS = 8;                         // rows per CTA
constexpr int FI_CTAS = 16;                        // CTAs per sample = cluster size
constexpr int FI_THREADS = 512;                    // two pixels per thread
constexpr int FI_SH = FI_ROWS + HALO_T + HALO_B;   // 21 staged rows
constexpr int FI_TILE = FI_SH * SW;                // floats per buffer
static_assert(HALO_T <= FI_ROWS && HALO_B <= FI_ROWS, "halo rows must come from the adjacent CTA only");

__device__ __forceinline__ void st_dsmem(uint32_t local_saddr, unsigned rank, float v) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(local_saddr), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(FI_THREADS, 1)
spn_iterate_fused_kernel(const float* __restrict__ feat_init, const float* __restrict__ aff, const float* __restrict__ offset,
                         float* __restrict__ list_out, const int B, const int H, const int W, const int T) {
    __shared__ __align__(16) float tile[2][FI_TILE];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    const int b = blockIdx.x / FI_CTAS;
    const int y0 = (int)rank * FI_ROWS;
    const size_t cs = (size_t)H * W;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- both buffers zero (outside the image is zero and stays zero), then generation 0 from feat_init ----
    for (int i = tid; i < 2 * FI_TILE; i += FI_THREADS) (&tile[0][0])[i] = 0.f;
    __syncthreads();
    const float* f0 = feat_init + (size_t)b * cs;
    for (int i = tid; i < FI_SH * TILE_W; i += FI_THREADS) {
        const int r = i / TILE_W, x = i - r * TILE_W;
        const int gy = y0 - HALO_T + r;
        if ((unsigned)gy < (unsigned)H && x < W) tile[0][r * SW + HALO_L + x] = f0[(size_t)gy * W + x];
    }

    // ---- iteration-invariant tap state of this thread's two pixels ----
    const int ry = warp >> 1;                         // row inside the CTA's 8
    const int y = y0 + ry;
    float a[2][9], lh[2][9], lw[2][9];
    uint32_t ad[2][9];                                // byte offset of the tap's footprint in a buffer
    unsigned slow[2] = {0u, 0u};
    bool active[2];
    int xs[2];
    Geom g;
    g.B = B; g.H = H; g.W = W; g.H_img = H; g.row0 = 0; g.init_row0 = 0; g.init_rows = H; g.tiles_x = 1; g.tiles_y = FI_CTAS;
    const int oy = y0 - HALO_T, ox = -HALO_L;         // image coordinates of staged element [0][0]
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int x = (warp & 1) * 64 + lane + 32 * j;
        xs[j] = x;
        active[j] = y < H && x < W;
        if (active[j]) {
            const size_t p = (size_t)y * W + x;
            const float* pw = aff + (size_t)b * 9 * cs + p;
            const float* po = offset + (size_t)b * 18 * cs + p;
            float oh[9], ow[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) a[j][k] = ld_stream(pw + (size_t)k * cs);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                oh[k] = ld_stream(po + (size_t)(2 * k) * cs);
                ow[k] = ld_stream(po + (size_t)(2 * k + 1) * cs);
            }
            const float fy = (float)y, fx = (float)x;
            const float hk[3] = {fy - 1.f, fy, fy + 1.f};
            const float wk[3] = {fx - 1.f, fx, fx + 1.f};
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float h = hk[k / 3] + oh[k], w = wk[k % 3] + ow[k];
                const int h0 = __float2int_rd(h), w0 = __float2int_rd(w);   // saturating; NaN -> 0
                lh[j][k] = h - floorf(h);
                lw[j][k] = w - floorf(w);
                const unsigned r = (unsigned)(h0 - oy), q = (unsigned)(w0 - ox);
                const bool ok = r < (unsigned)(FI_SH - 1) && q < (unsigned)(SW - 1);
                ad[j][k] = ok ? (r * SW + q) * 4u : 0u;
                slow[j] |= ok ? 0u : (1u << k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) { a[j][k] = 0.f; lh[j][k] = 0.f; lw[j][k] = 0.f; ad[j][k] = 0u; }
        }
    }
    cluster.sync();   // every CTA's buffers are initialised before any neighbour writes halo rows into them

    const bool up = rank > 0 && ry < HALO_B;                          // this row is a bottom-halo row of CTA rank - 1
    const bool dn = rank + 1 < FI_CTAS && ry >= FI_ROWS - HALO_T;     // ... a top-halo row of CTA rank + 1
    const float* src_g = f0;                                          // global copy of the generation being read
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        const unsigned char* cur = reinterpret_cast<const unsigned char*>(tile[t & 1]);
        float* nxt = tile[(t + 1) & 1];
        float* out_t = list_out + ((size_t)t * B + b) * cs;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (!active[j]) continue;
            float c[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float* s = reinterpret_cast<const float*>(cur + ad[j][k]);
                const float val = bilerp(s[0], s[1], s[SW], s[SW + 1], lh[j][k], lw[j][k]);
                c[k] = (slow[j] >> k) & 1u ? a[j][k] : (1.f * a[j][k]) * val;
            }
            if (slow[j]) {   // rare: taps outside the staged tile, redone through the bounds-checked global path
                const float* po = offset + (size_t)b * 18 * cs + (size_t)y * W + xs[j];
                const float fy = (float)y, fx = (float)xs[j];
#pragma unroll 1
                for (int k = 0; k < 9; ++k) {
                    if (!((slow[j] >> k) & 1u)) continue;
                    const float h = (fy + (float)(k / 3 - 1)) + po[(size_t)(2 * k) * cs];
                    const float w = (fx + (float)(k % 3 - 1)) + po[(size_t)(2 * k + 1) * cs];
                    const SlowTap st = slow_tap<float>(src_g, g, h, w, nullptr);
                    const float v = (1.f * a[j][k]) * bilerp(st.v1, st.v2, st.v3, st.v4, st.lh, st.lw);
#pragma unroll
                    for (int kk = 0; kk < 9; ++kk) c[kk] = kk == k ? v : c[kk];
                }
            }
            float acc = c[0];
#pragma unroll
            for (int k = 1; k < 9; ++k) acc += c[k];
            acc += 0.f;   // the bias slot of the plain kernel (b = 0): keeps the sign of a zero sum identical
            st_stream(out_t + (size_t)y * W + xs[j], acc);
            if (t + 1 < T) {
                const int e = (HALO_T + ry) * SW + HALO_L + xs[j];
                nxt[e] = acc;
                if (up) st_dsmem(smem_u32(nxt + (HALO_T + FI_ROWS + ry) * SW + HALO_L + xs[j]), rank - 1, acc);
                if (dn) st_dsmem(smem_u32(nxt + (ry - (FI_ROWS - HALO_T)) * SW + HALO_L + xs[j]), rank + 1, acc);
            }
        }
        src_g = out_t;
        if (t + 1 < T) cluster.sync();   // barrier.cluster arrive.release / wait.acquire: shared, distributed shared and
    }                                    // global stores of this application are visible to the next one
    cluster.sync();                      // no CTA exits while a neighbour could still address its shared memory
}

// cudaErrorNotSupported: the device cannot co-schedule a 16-CTA cluster (the caller falls back to T launches)
cudaError_t launch_spn_iterate_fused(const float* feat_init, const float* aff, const float* offset, float* list_out, int B,
                                     int H, int W, int T, cudaStream_t stream) {
    // per device: opt in to the non-portable cluster size once, and ask whether one such cluster can be resident
    static int state[64] = {0};   // 0 = unknown, 1 = supported, -1 = not
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return cudaErrorNotSupported;
    int& st = state[dev & 63];
    if (st == 0) {
        cudaError_t e = cudaFuncSetAttribute((const void*)spn_iterate_fused_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        int n = 0;
        if (e == cudaSuccess) {
            cudaLaunchConfig_t probe{};
            probe.gridDim = dim3(FI_CTAS);
            probe.blockDim = dim3(FI_THREADS);
            cudaLaunchAttribute at{};
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = FI_CTAS; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            probe.attrs = &at;
            probe.numAttrs = 1;
            e = cudaOccupancyMaxActiveClusters(&n, (const void*)spn_iterate_fused_kernel, &probe);
        }
        if (e != cudaSuccess) (void)cudaGetLastError();
        st = (e == cudaSuccess && n >= 1) ? 1 : -1;
    }
    if (st < 0) return cudaErrorNotSupported;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)B * FI_CTAS);
    cfg.blockDim = dim3(FI_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = FI_CTAS; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, spn_iterate_fused_kernel, feat_init, aff, offset, list_out, B, H, W, T);
}

}  // namespace JSPSR_VARIANT
}  // namespace jspsr
