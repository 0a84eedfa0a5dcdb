// Generator tail fused into the propagation forward:
//   weight = sigmoid(conv1x1(feature; Wc[9,C], bc[9]))            models/components/spn.py:41-44,66
//   offset = conv1x1(feature; Wo[16,C], bo[16]) + zero centre pair  models/components/spn.py:45-52,67-73
//   out    = PostProcessor(init, weight, offset)                    models/components/spn.py:99-118
// in ONE pass over HBM: C feature channels in (C*4 B/pixel instead of the 108 B/pixel of weight+offset the
// Generator would have written and the propagation re-read), 1 channel out, optionally the 27 channels of
// weight/offset for the backward.
//
// The two 1x1 convolutions are one [128 pixels x C] x [C x 25] contraction per image row segment: it runs on
// the 5th-generation tensor cores (tcgen05.mma, kind::tf32, M = 128, N = 32, K = 8; accumulator in TMEM).
// Feature rows ([C planes] x [128 pixels]) arrive by TMA in a shared-memory ring; a producer thread owns one
// pixel column: it reads its pixel's C features from the ring, splits them into tf32 hi + lo parts and writes
// them into ITS lane of tensor memory (tcgen05.st) - the A operand is read by the MMAs straight from TMEM;
// the B operand (the 25 x C weights, hi + mid parts) sits in shared memory in the canonical K-major
// no-swizzle layout; after the commit the consumer thread of the same pixel reads ITS 25 results back from
// its TMEM lane (tcgen05.ld 32x32b) - exactly the registers the 9-tap gather needs.  fp32 accuracy comes from the 3-product
// split  A_hi*B_hi + A_lo*B_hi + A_hi*B_mid  (hi parts rounded to nearest tf32, so the dropped terms are
// <= 2^-23 relative; measured with tools/umma_probe.cu).
#include <cstdlib>

#include "spn_kernels.cuh"
#include "umma_helpers.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

constexpr int GEN_STAGES = 2;    // feature rows in flight into shared memory per CTA (32 KB each at C = 64)
// Further rows pulled into L2 by TMA prefetch.  Measured and rejected (kept behind JSPSR_GEN_L2_AHEAD): 0 rows 1.52 ms,
// 2 rows 1.54 ms, 4 rows 1.73 ms with 10.5 GB read from DRAM for 8.7 GB of data - prefetched lines are evicted
// before the ring asks for them.
constexpr int GEN_L2_AHEAD = 0;
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// C: feature channels = bc * 4 of the Generator: 128 at every YAML config (models/JSPSR.py:28 hard-codes cat_only = True,
// so bc = num_feature = 32, JSPSR.py:181); 64 for cat_only = False or models/EDSR.py:104-106 with 32 features.
// TMA: the DEM box AND the feature rows arrive by TMA (needs 16-byte aligned rows); otherwise bounds-checked loads.
// TH: rows per CTA.  WRITE_WO: also store weight [B,9,H,W] and offset [B,18,H,W] (what the backward needs).
//
// Warp-specialised, 320 threads:
//   TMA warp  (warp 9)   : one lane keeps GEN_STAGES feature rows ([C][128] boxes) in flight into a shared-memory ring;
//   producers (warps 4-7): one thread per pixel column: read the pixel's C features from the ring (conflict-free),
//                          split them to tf32 hi / lo and write them into ITS TMEM lane (tcgen05.st) - the A operand
//                          never touches shared memory;
//   MMA warp  (warp 8)   : one lane issues the row's 24 MMAs (A from TMEM, B from shared memory) into accumulator
//                          (row & 1) and commits;
//   consumers (warps 0-3): read their pixel's 25 results from their TMEM lane, release the accumulator, run the
//                          Generator epilogue and the 9-tap gather, store the output.
// mbarriers: full/empty[stage] (ring), a_full[half] (128 producers have written that K half of the row's operand),
// a_free[half] (its MMAs have finished reading it) - two halves so that the producers fill one while the MMAs
// read the other -, acc_full[2] (accumulator complete), acc_empty[2] (all 128 consumers have read it).
// TMEM (256 columns per CTA, two CTAs per SM): A_hi [0,64) | A_lo [64,128) | accumulators [128,160), [160,192).
// Measured steps (2048 tiles): every thread doing all jobs in turn, 8 warps per SM: 3.4 ms (latency-bound, 35 %
// issue rate); producers + consumers with a producer lane issuing the MMAs: 2.9 ms; dedicated MMA warp: 2.6 ms
// (producers held at most 64 KB of loads in flight per SM in registers: 0.52 of the HBM roofline by Little's law);
// this version: 1.52 ms = 0.89 of the HBM roofline (0.98 when weight/offset are also written).
// FT: element type of feature and of weight_out / offset_out.  bf16 (torch.autocast: Generator.block emits bf16) is
// exactly representable in tf32, so the lo part and its MMAs vanish (16 MMAs per row), the ring stage halves, and
// weight / offset are rounded to bf16 BEFORE the gather - what the reference's autocast run propagates, and what
// makes `out` bit-identical to the propagation kernel applied to the tensors written here.
// C = 128 (the JSPSR configs): the operand needs 2 x 128 TMEM columns, so the
// CTA allocates all 512 and runs alone on its SM (ring 2 x 64 KB).
template <typename FT, int C, bool TMA, int TH, bool WRITE_WO>
__global__ void __launch_bounds__(GEN_CTA_THREADS, (C <= 64 || sizeof(FT) == 2) ? 2 : 1)
gen_spn_forward_kernel(const float* __restrict__ init, const FT* __restrict__ feature,
                       const float* __restrict__ conv_w, const float* __restrict__ conv_b,
                       const float* __restrict__ w9, const float* __restrict__ b1, float* __restrict__ out,
                       FT* __restrict__ weight_out, FT* __restrict__ offset_out, const Geom g, const int mode,
                       const float scale, const int l2_ahead, const __grid_constant__ CUtensorMap tmap,
                       const __grid_constant__ CUtensorMap tmap_feat) {
    constexpr int SH = staged_rows(TH);
    constexpr uint32_t SBO = 128;                // bytes between 8-row groups of B (rows = output channels)
    constexpr uint32_t LBO_B = GEN_N / 8 * 128;  // bytes between 16-byte K chunks of B: 512
    constexpr bool F16 = sizeof(FT) == 2;
    constexpr int STAGE_BYTES = GEN_THREADS * C * (int)sizeof(FT), B_BYTES = GEN_N * C * 4;
    constexpr int RING_BYTES = TMA ? GEN_STAGES * STAGE_BYTES : 0;
    // TMEM column map; bf16 features have no lo part, so C = 128 then fits 256 columns and two CTAs per SM again
    constexpr uint32_t COL_A_HI = 0, COL_A_LO = C, COL_ACC = F16 ? C : 2 * C;
    constexpr uint32_t TMEM_COLS = COL_ACC + 2 * GEN_N <= 256 ? 256 : 512;
    static_assert(COL_ACC + 2 * GEN_N <= TMEM_COLS, "TMEM budget");
    extern __shared__ __align__(1024) unsigned char dsm[];
    unsigned char* ring = dsm;                       // [GEN_STAGES][C][128] fp32 (TMA only)
    unsigned char* b_hi = dsm + RING_BYTES;
    unsigned char* b_mid = b_hi + B_BYTES;
    float* tile = reinterpret_cast<float*>(b_mid + B_BYTES);  // [SH][SW], 128-byte aligned (all sizes are multiples of 128)
    __shared__ __align__(8) uint64_t bar_tile, bar_full[GEN_STAGES], bar_empty[GEN_STAGES], bar_a_full[2], bar_a_free[2],
        bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ float s_w[10];
    __shared__ __align__(16) float s_bias[GEN_N];

    const int t = threadIdx.x, warp = t >> 5;
    const TileCtx c = make_tile_ctx<TH>(g);
    stage_tile_begin<float, TMA, TH, GEN_CTA_THREADS>(tile, &bar_tile, &tmap, init, g, c.b, c.ox, c.oy - g.init_row0);
    if (t < 9) s_w[t] = w9 ? w9[t] : 1.f;
    if (t == 9) s_w[9] = b1 ? b1[0] : 0.f;
    if (t < GEN_N) s_bias[t] = t < GEN_NOUT ? conv_b[t] : 0.f;
    if (t == 32) {
        for (int i = 0; i < GEN_STAGES; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], GEN_THREADS);
        }
        for (int i = 0; i < 2; ++i) {  // per K half (channels [0, C/2) and [C/2, C))
            mbar_init(&bar_a_full[i], GEN_THREADS);
            mbar_init(&bar_a_free[i], 1);
        }
        mbar_init(&bar_acc_full[0], 1);
        mbar_init(&bar_acc_full[1], 1);
        mbar_init(&bar_acc_empty[0], GEN_THREADS);
        mbar_init(&bar_acc_empty[1], GEN_THREADS);
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // B operand: row n = output channel (9 weight rows, 16 offset rows, 7 zero rows), K-major, split hi + mid
    for (int i = t; i < GEN_N * (C / 4); i += GEN_CTA_THREADS) {
        const int n = i % GEN_N, kc = i / GEN_N;
        float4 hi = make_float4(0.f, 0.f, 0.f, 0.f), mid = hi;
        if (n < GEN_NOUT) {
            const float4 v = *reinterpret_cast<const float4*>(conv_w + (size_t)n * C + kc * 4);
            hi = make_float4(tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
            mid = make_float4(tf32_rn(v.x - hi.x), tf32_rn(v.y - hi.y), tf32_rn(v.z - hi.z), tf32_rn(v.w - hi.w));
        }
        const uint32_t off = (n / 8) * SBO + kc * LBO_B + (n % 8) * 16;
        *reinterpret_cast<float4*>(b_hi + off) = hi;
        *reinterpret_cast<float4*>(b_mid + off) = mid;
    }
    fence_proxy_async();  // B operand: generic-proxy stores -> visible to the tensor core's async proxy

    const size_t cs = (size_t)g.H * g.W;
    const int px = t & (GEN_THREADS - 1);  // pixel column of this thread inside the tile (producers and consumers)
    const int x = c.x0 + px;
    const bool col_ok = x < g.W;
    // instruction descriptor: D fp32 | A, B tf32 | both K-major | N = 32 | M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((GEN_N >> 3) << 17) | ((GEN_THREADS >> 4) << 24);

    if (warp == 9) {
        // =========================== TMA warp: feature rows into the ring ===========================
        __syncthreads();  // setup complete
        if (TMA && (t & 31) == 0) {
            // optional (off by default, see GEN_L2_AHEAD): rows beyond the ring pulled into L2
            for (int r = GEN_STAGES; r < GEN_STAGES + l2_ahead && r < TH; ++r)
                tma_prefetch_3d(&tmap_feat, c.x0, c.y0 + r, c.b * C);
#pragma unroll 1
            for (int r = 0; r < TH; ++r) {
                const int s = r % GEN_STAGES;
                if (l2_ahead > 0 && r + GEN_STAGES + l2_ahead < TH)
                    tma_prefetch_3d(&tmap_feat, c.x0, c.y0 + r + GEN_STAGES + l2_ahead, c.b * C);
                if (r >= GEN_STAGES) mbar_wait(&bar_empty[s], (uint32_t)(((r / GEN_STAGES) - 1) & 1));
                mbar_arrive_expect_tx(&bar_full[s], STAGE_BYTES);
                // box {128 columns, 1 row, C planes}; columns / rows outside the image arrive as zeros
                tma_load_3d(ring + s * STAGE_BYTES, &tmap_feat, &bar_full[s], c.x0, c.y0 + r, c.b * C);
            }
        }
    } else if (warp == 8) {
        // =========================== MMA warp ===========================
        __syncthreads();  // setup complete (TMEM address, barriers, B operand)
        const uint32_t tmem = s_tmem;
#pragma unroll 1
        for (int r = 0; r < TH; ++r) {
            const uint32_t d_tmem = tmem + COL_ACC + (uint32_t)(r & 1) * GEN_N;
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // the two K halves are pipelined against the producers
                mbar_wait(&bar_a_full[h], (uint32_t)(r & 1));                                   // half h of row r is in TMEM
                if (h == 0 && r >= 2) mbar_wait(&bar_acc_empty[r & 1], (uint32_t)(((r >> 1) - 1) & 1));  // row r-2 drained
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int kq = 0; kq < C / 16; ++kq) {  // K = 8 tf32 per MMA: 8 TMEM columns of A, two 16-byte chunks of B
                    const int ks = h * (C / 16) + kq;
                    const uint64_t dbh = umma_desc_kmajor(smem_u32(b_hi) + ks * 2 * LBO_B, LBO_B, SBO);
                    const uint64_t dbm = umma_desc_kmajor(smem_u32(b_mid) + ks * 2 * LBO_B, LBO_B, SBO);
                    umma_tf32_ts(d_tmem, tmem + COL_A_HI + ks * 8, dbh, IDESC, ks > 0 ? 1u : 0u);
                    if (!F16) umma_tf32_ts(d_tmem, tmem + COL_A_LO + ks * 8, dbh, IDESC, 1u);
                    umma_tf32_ts(d_tmem, tmem + COL_A_HI + ks * 8, dbm, IDESC, 1u);
                }
                umma_commit(&bar_a_free[h]);
                if (h == 1) umma_commit(&bar_acc_full[r & 1]);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // =========================== producers ===========================
        const FT* feat_b = feature + (size_t)c.b * C * cs;
        __syncthreads();  // setup complete
        const uint32_t lane_tmem = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
        for (int r = 0; r < TH; ++r) {
            const int s = r % GEN_STAGES;
            const FT* stage = reinterpret_cast<const FT*>(ring + s * STAGE_BYTES) + px;
            const int y = c.y0 + r;
            const bool row_ok = col_ok && y < g.H;
            const FT* gp = feat_b + (size_t)y * g.W + x;
            if (TMA) mbar_wait(&bar_full[s], (uint32_t)((r / GEN_STAGES) & 1));       // the row has landed
#pragma unroll
            for (int k0 = 0; k0 < C; k0 += 16) {  // 16 channels at a time: ring / HBM -> registers -> hi, lo -> TMEM lane
                if (k0 % (C / 2) == 0 && r > 0) {  // MMAs of row r-1 no longer read this K half of A
                    mbar_wait(&bar_a_free[k0 / (C / 2)], (uint32_t)((r - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                float hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float v;
                    if (TMA) v = to_f32(stage[(k0 + j) * GEN_THREADS]);
                    else v = row_ok ? ld_stream(gp + (size_t)(k0 + j) * cs) : 0.f;
                    hi[j] = v;
                }
                if (!F16) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float v = hi[j];
                        hi[j] = tf32_rn(v);
                        lo[j] = v - hi[j];
                    }
                    tmem_st16(lane_tmem + COL_A_LO + k0, lo);
                }
                tmem_st16(lane_tmem + COL_A_HI + k0, hi);
                if ((k0 + 16) % (C / 2) == 0) {  // a K half is complete: hand it to the MMA warp
                    if (TMA && k0 + 16 == C) mbar_arrive(&bar_empty[s]);  // every read of the stage is in registers
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&bar_a_full[(k0 + 16) / (C / 2) - 1]);
                }
            }
        }
    } else {
        // =========================== consumers ===========================
        const float* init_b = init + (size_t)c.b * g.init_rows * g.W;
        float* out_b = out + (size_t)c.b * cs;
        const float* tile_lo = tile + c.r_lo * SW;
        stage_tile_wait<TMA>(&bar_tile);  // __syncthreads (pairs with the other roles') + the DEM box has landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem = s_tmem;
#pragma unroll 1
        for (int r = 0; r < TH; ++r) {
            __syncwarp();  // the previous row's out-of-tile taps may have split the warp; tcgen05.ld below is .sync.aligned
            mbar_wait(&bar_acc_full[r & 1], (uint32_t)((r >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + COL_ACC + (uint32_t)(r & 1) * GEN_N, v);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&bar_acc_empty[r & 1]);

            const int y = c.y0 + r;
            if (!col_ok || y >= g.H) continue;
            const size_t p = (size_t)y * g.W + x;
            float a[9], oh[9], ow[9];
#pragma unroll
            for (int q = 0; q < 7; ++q) {  // + bias (broadcast 16-byte reads)
                const float4 bq = reinterpret_cast<const float4*>(s_bias)[q];
                v[4 * q] += bq.x; v[4 * q + 1] += bq.y; v[4 * q + 2] += bq.z; v[4 * q + 3] += bq.w;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                // sigmoid = 1 / (1 + 2^(-z log2 e)): ex2.approx + rcp.approx (2 + 1 ulp; the argument rounding adds
                // |z| * 1e-7 relative to e^-z, i.e. < 5e-7 absolute on the weight for |z| < 20)
                a[k] = __fdividef(1.f, 1.f + __expf(-v[k]));
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (k == 4) {
                    oh[k] = ow[k] = 0.f;  // zero centre pair (spn.py:69-73)
                } else {
                    const int n = k < 4 ? k : k - 1;
                    oh[k] = v[9 + 2 * n];
                    ow[k] = v[10 + 2 * n];
                }
            }
            if (F16) {  // the Generator's outputs are bf16 tensors in this mode: propagate exactly what is written
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    a[k] = __bfloat162float(__float2bfloat16_rn(a[k]));
                    oh[k] = __bfloat162float(__float2bfloat16_rn(oh[k]));
                    ow[k] = __bfloat162float(__float2bfloat16_rn(ow[k]));
                }
            }
            if (WRITE_WO) {
                FT* pw = weight_out + (size_t)c.b * 9 * cs + p;
                FT* po = offset_out + (size_t)c.b * 18 * cs + p;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    st_stream(pw + (size_t)k * cs, a[k]);
                    st_stream(po + (size_t)(2 * k) * cs, oh[k]);
                    st_stream(po + (size_t)(2 * k + 1) * cs, ow[k]);
                }
            }
            // propagation: identical to spn_forward_kernel's pixel body
            normalise9(a, mode);
            const float fy = (float)(g.row0 + y), fx = (float)x;
            const float hk[3] = {fy - 1.f, fy, fy + 1.f};
            const float wk[3] = {fx - 1.f, fx, fx + 1.f};
            unsigned slow = 0u;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const FastTap tp = fast_tap<float>(tile_lo, c, hk[k / 3] + oh[k], wk[k % 3] + ow[k]);
                const float val = bilerp(tp.v1, tp.v2, tp.v3, tp.v4, tp.lh, tp.lw);
                a[k] = tp.ok ? (s_w[k] * a[k]) * val : a[k];
                slow |= tp.ok ? 0u : (1u << k);
            }
            if (slow) {
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (slow & (1u << k)) {
                        const SlowTap tp = slow_tap<float>(init_b, g, hk[k / 3] + oh[k], wk[k % 3] + ow[k], nullptr);
                        a[k] = (s_w[k] * a[k]) * bilerp(tp.v1, tp.v2, tp.v3, tp.v4, tp.lh, tp.lw);
                    }
                }
            }
            float acc = a[0];
#pragma unroll
            for (int k = 1; k < 9; ++k) acc += a[k];
            acc += s_w[9];
            if (mode == NORM_RESIDUAL) acc = fmaf(scale, tile[(r + HALO_T) * SW + (px + HALO_L)], acc);
            st_stream(out_b + p, acc);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(TMEM_COLS));
}

template <typename FT, int C, bool TMA, int TH, bool WO>
static cudaError_t launch_gen_one(const LaunchArgs& la, const CUtensorMap& tmap_feat, const void* feature,
                                  const float* conv_w, const float* conv_b, void* weight_out, void* offset_out) {
    size_t dyn = (TMA ? (size_t)GEN_STAGES * GEN_THREADS * C * sizeof(FT) : 0) + (size_t)2 * GEN_N * C * 4 +
                 (size_t)staged_rows(TH) * SW * 4;
    // The kernel allocates 256 or 512 TMEM columns per CTA, so at most 2 or 1 CTAs fit an SM's 512 columns.  With the TMA
    // ring the shared-memory footprint enforces that by itself; the bounds-checked fallback (no ring) is small enough for
    // more CTAs to become resident, and the extra ones would spin inside tcgen05.alloc holding registers and shared memory.
    // Pad the request so that residency never exceeds what TMEM can serve.
    constexpr size_t tmem_cols = ((sizeof(FT) == 2 ? C : 2 * C) + 2 * GEN_N <= 256) ? 256 : 512;
    constexpr size_t min_dyn = tmem_cols == 512 ? (size_t)116 * 1024 : (size_t)76 * 1024;   // > 227 KB / 2, > 227 KB / 3
    if (dyn < min_dyn) dyn = min_dyn;
    const cudaError_t attr = ensure_dynamic_smem((const void*)gen_spn_forward_kernel<FT, C, TMA, TH, WO>, dyn);
    if (attr != cudaSuccess) return attr;
    dim3 grid((unsigned)((size_t)la.g.tiles_x * la.g.tiles_y * la.g.B));
    int l2_ahead = GEN_L2_AHEAD;
    if (const char* e = getenv("JSPSR_GEN_L2_AHEAD")) l2_ahead = atoi(e);  // experiment switch
    gen_spn_forward_kernel<FT, C, TMA, TH, WO><<<grid, GEN_CTA_THREADS, dyn, la.stream>>>(
        (const float*)la.init, (const FT*)feature, conv_w, conv_b, la.w9, la.b1, (float*)la.out, (FT*)weight_out,
        (FT*)offset_out, la.g, la.mode, la.scale, l2_ahead, la.tmap, tmap_feat);
    return cudaGetLastError();
}

template <typename FT, int C, int TH>
static cudaError_t launch_gen_c(const LaunchArgs& la, const CUtensorMap& tmap_feat, const void* feature,
                                const float* conv_w, const float* conv_b, void* weight_out, void* offset_out) {
    const bool wo = weight_out != nullptr;
    if (la.use_tma) {
        return wo ? launch_gen_one<FT, C, true, TH, true>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out)
                  : launch_gen_one<FT, C, true, TH, false>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out);
    }
    return wo ? launch_gen_one<FT, C, false, TH, true>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out)
              : launch_gen_one<FT, C, false, TH, false>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out);
}

// la.tile_h must be 16 or 8 (abi.cu); la.use_tma says that BOTH tensor maps (DEM box, feature rows) are valid;
// la.bf16: feature / weight_out / offset_out are bf16 (init and out stay fp32)
cudaError_t launch_gen_spn_forward(const LaunchArgs& la, const CUtensorMap& tmap_feat, const void* feature, int C,
                                   const float* conv_w, const float* conv_b, void* weight_out, void* offset_out) {
    if (C == 128) {  // fp32: one CTA per SM, 16 rows; bf16: two CTAs per SM, 8 rows (the smaller DEM tile makes room)
        return la.bf16 ? launch_gen_c<__nv_bfloat16, 128, 8>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out)
                       : launch_gen_c<float, 128, 16>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out);
    }
    if (C != 64) return cudaErrorNotSupported;
    if (la.bf16) {
        return la.tile_h == 16 ? launch_gen_c<__nv_bfloat16, 64, 16>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out)
                               : launch_gen_c<__nv_bfloat16, 64, 8>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out);
    }
    return la.tile_h == 16 ? launch_gen_c<float, 64, 16>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out)
                           : launch_gen_c<float, 64, 8>(la, tmap_feat, feature, conv_w, conv_b, weight_out, offset_out);
}

}  // namespace JSPSR_VARIANT
}  // namespace jspsr
