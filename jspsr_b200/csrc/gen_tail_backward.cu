// Feature gradient of the Generator tail (backward of the two 1x1 convolutions of models/components/spn.py:41-52,
// 66-73 w.r.t. their input):
//   grad_feature[b, c, y, x] = sum_j gz[b, j, y, x] * W[j, c],       j = 0..24 (9 weight + 16 offset pre-activations)
// where gz is what spn_backward_kernel writes in JSPSR_BWD_GEN_PREACT mode.  One pass over HBM: 25 channels in, C out
// (100 + 4C B/pixel), the [128 pixels x 32] x [32 x C] contraction per image-row segment on tcgen05 (tf32, 3-product
// split for fp32-level accuracy; the library's SIMT batched GEMM this replaces ran at 3.8 ms for 2048 tiles).
//
// Same warp-specialised pipeline as gen_spn_forward.cu, 320 threads:
//   TMA warp  (warp 9)   : keeps GT_STAGES rows of gz ([25 planes][128 px] boxes) in flight into a shared-memory ring;
//   producers (warps 4-7): thread = pixel column: 25 values from the ring -> tf32 hi / lo -> its own TMEM lane (A operand);
//   MMA warp  (warp 8)   : 4 K-steps x 3 products per row, A from TMEM, B = W^T (hi + mid) from shared memory, N = C,
//                          into accumulator (row & 1);
//   consumers (warps 0-3): thread = the same pixel column: tcgen05.ld its C results, store them (each store
//                          instruction is one coalesced 128-byte line per warp and channel plane).
// TMEM: two A buffers (hi [0,32) | lo [32,64) and [64,96) | [96,128): the producers fill one while the MMAs read the other;
// with a single buffer the producers spent most of their time waiting for the previous row's MMAs) | accumulators
// [128, 128+C), [128+C, 128+2C)  ->  256 columns at C = 64 (two CTAs per SM), 512 at C = 128 (one CTA per SM).
#include "spn_kernels.cuh"
#include "umma_helpers.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

constexpr int GT_STAGES = 4;   // rows of gz in flight per CTA (12.5 KB each in fp32)
constexpr int GT_K = 32;       // contraction length: 25 pre-activations, zero-padded
constexpr int GT_ROWS = 16;    // image rows per CTA

// CS: compile-time channel stride H*W (0 = runtime) for the C output planes' addresses.
template <typename FT, int C, bool TMA, int CS>
__global__ void __launch_bounds__(GEN_CTA_THREADS, C <= 64 ? 2 : 1)
gen_grad_feature_kernel(const FT* __restrict__ gz, const float* __restrict__ conv_w, FT* __restrict__ grad_feature,
                        const Geom g, const __grid_constant__ CUtensorMap tmap_gz) {
    constexpr bool F16 = sizeof(FT) == 2;
    constexpr uint32_t SBO = 128;              // bytes between 8-row groups of B (rows = feature channels)
    constexpr uint32_t LBO_B = C / 8 * 128;    // bytes between 16-byte K chunks of B
    constexpr int STAGE_BYTES = GEN_NOUT * GEN_THREADS * (int)sizeof(FT), B_BYTES = C * GT_K * 4;
    constexpr int RING_BYTES = TMA ? GT_STAGES * STAGE_BYTES : 0;
    constexpr uint32_t COL_A_HI = 0, COL_A_LO = GT_K, A_BUF = 2 * GT_K, COL_ACC = 2 * A_BUF;
    constexpr uint32_t TMEM_COLS = C <= 64 ? 256 : 512;
    static_assert(2 * A_BUF + 2 * C <= TMEM_COLS && STAGE_BYTES % 128 == 0, "TMEM / ring budget");
    extern __shared__ __align__(1024) unsigned char dsm[];
    unsigned char* ring = dsm;                 // [GT_STAGES][25][128] (TMA only)
    unsigned char* b_hi = dsm + RING_BYTES;    // W^T: row n = feature channel, K = pre-activation index, K-major
    unsigned char* b_mid = b_hi + B_BYTES;
    __shared__ __align__(8) uint64_t bar_full[GT_STAGES], bar_empty[GT_STAGES], bar_a_full[2], bar_a_free[2],
        bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;

    const int t = threadIdx.x, warp = t >> 5;
    const TileCtx c = make_tile_ctx<GT_ROWS>(g);
    if (t == 32) {
        for (int i = 0; i < GT_STAGES; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], GEN_THREADS);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_a_full[i], GEN_THREADS);
            mbar_init(&bar_a_free[i], 1);
            mbar_init(&bar_acc_full[i], 1);
            mbar_init(&bar_acc_empty[i], GEN_THREADS);
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // B operand: element (n = channel, k = j) = W[j][n] for j < 25, zero for the padding; hi + mid parts
    for (int i = t; i < C * (GT_K / 4); i += GEN_CTA_THREADS) {
        const int n = i % C, kc = i / C;
        float v[4], hi[4], mid[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = kc * 4 + q;
            v[q] = j < GEN_NOUT ? conv_w[(size_t)j * C + n] : 0.f;
            hi[q] = tf32_rn(v[q]);
            mid[q] = tf32_rn(v[q] - hi[q]);
        }
        const uint32_t off = (n / 8) * SBO + kc * LBO_B + (n % 8) * 16;
        *reinterpret_cast<float4*>(b_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(b_mid + off) = make_float4(mid[0], mid[1], mid[2], mid[3]);
    }
    fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    const size_t cs = CS ? (size_t)CS : (size_t)g.H * g.W;
    const int px = t & (GEN_THREADS - 1);
    const int x = c.x0 + px;
    const bool col_ok = x < g.W;
    // instruction descriptor: D fp32 | A, B tf32 | both K-major | N = C | M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((GEN_THREADS >> 4) << 24);

    if (warp == 9) {
        // =========================== TMA warp ===========================
        if (TMA && (t & 31) == 0) {
#pragma unroll 1
            for (int r = 0; r < GT_ROWS; ++r) {
                const int s = r % GT_STAGES;
                if (r >= GT_STAGES) mbar_wait(&bar_empty[s], (uint32_t)(((r / GT_STAGES) - 1) & 1));
                mbar_arrive_expect_tx(&bar_full[s], STAGE_BYTES);
                tma_load_3d(ring + s * STAGE_BYTES, &tmap_gz, &bar_full[s], c.x0, c.y0 + r, c.b * GEN_NOUT);
            }
        }
    } else if (warp == 8) {
        // =========================== MMA warp (converged) ===========================
#pragma unroll 1
        for (int r = 0; r < GT_ROWS; ++r) {
            mbar_wait(&bar_a_full[r & 1], (uint32_t)((r >> 1) & 1));
            if (r >= 2) mbar_wait(&bar_acc_empty[r & 1], (uint32_t)(((r >> 1) - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem + COL_ACC + (uint32_t)(r & 1) * C;
            const uint32_t a_tmem = tmem + (uint32_t)(r & 1) * A_BUF;
#pragma unroll
            for (int ks = 0; ks < GT_K / 8; ++ks) {
                const uint64_t dbh = umma_desc_kmajor(smem_u32(b_hi) + ks * 2 * LBO_B, LBO_B, SBO);
                const uint64_t dbm = umma_desc_kmajor(smem_u32(b_mid) + ks * 2 * LBO_B, LBO_B, SBO);
                umma_tf32_ts(d_tmem, a_tmem + COL_A_HI + ks * 8, dbh, IDESC, ks > 0 ? 1u : 0u);
                if (!F16) umma_tf32_ts(d_tmem, a_tmem + COL_A_LO + ks * 8, dbh, IDESC, 1u);
                umma_tf32_ts(d_tmem, a_tmem + COL_A_HI + ks * 8, dbm, IDESC, 1u);
            }
            umma_commit(&bar_a_free[r & 1]);
            umma_commit(&bar_acc_full[r & 1]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // =========================== producers ===========================
        const FT* gz_b = gz + (size_t)c.b * GEN_NOUT * cs;
        const uint32_t lane_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
        for (int r = 0; r < GT_ROWS; ++r) {
            const int s = r % GT_STAGES;
            const FT* stage = reinterpret_cast<const FT*>(ring + s * STAGE_BYTES) + px;
            const int y = c.y0 + r;
            const bool row_ok = col_ok && y < g.H;
            const FT* gp = gz_b + (size_t)y * g.W + x;
            if (TMA) mbar_wait(&bar_full[s], (uint32_t)((r / GT_STAGES) & 1));
            if (r >= 2) {  // MMAs of row r-2 no longer read this A buffer
                mbar_wait(&bar_a_free[r & 1], (uint32_t)(((r >> 1) - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t a_lane = lane_tmem + (uint32_t)(r & 1) * A_BUF;
#pragma unroll
            for (int k0 = 0; k0 < GT_K; k0 += 16) {  // 16 pre-activations at a time: ring / HBM -> hi, lo -> TMEM lane
                float hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float v;
                    if (k0 + j >= GEN_NOUT) v = 0.f;
                    else if (TMA) v = to_f32(stage[(k0 + j) * GEN_THREADS]);
                    else v = row_ok ? ld_stream(gp + (size_t)(k0 + j) * cs) : 0.f;
                    hi[j] = F16 ? v : tf32_rn(v);
                    lo[j] = v - hi[j];
                }
                tmem_st16(a_lane + COL_A_HI + k0, hi);
                if (!F16) tmem_st16(a_lane + COL_A_LO + k0, lo);
            }
            if (TMA) mbar_arrive(&bar_empty[s]);  // every read of the stage is in registers / TMEM
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&bar_a_full[r & 1]);
        }
    } else {
        // =========================== consumers ===========================
        FT* gf_b = grad_feature + (size_t)c.b * C * cs;
#pragma unroll 1
        for (int r = 0; r < GT_ROWS; ++r) {
            mbar_wait(&bar_acc_full[r & 1], (uint32_t)((r >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int y = c.y0 + r;
            const bool ok = col_ok && y < g.H;
            FT* dst = gf_b + (size_t)y * g.W + x;
#pragma unroll
            for (int n0 = 0; n0 < C; n0 += 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + COL_ACC + (uint32_t)(r & 1) * C + n0, v);
                if (n0 + 32 == C) {  // the accumulator is in registers: hand it back before the stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&bar_acc_empty[r & 1]);
                }
                if (ok) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) st_stream(dst + (size_t)(n0 + j) * cs, v[j]);
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(TMEM_COLS));
}

template <typename FT, int C, bool TMA, int CS>
static cudaError_t launch_gf_cs(const LaunchArgs& la, const CUtensorMap& tmap_gz, const void* gz, const float* conv_w,
                                void* grad_feature) {
    size_t dyn = (TMA ? (size_t)GT_STAGES * GEN_NOUT * GEN_THREADS * sizeof(FT) : 0) + (size_t)2 * C * GT_K * 4;
    // residency must not exceed what the SM's 512 TMEM columns can serve (256 per CTA at C <= 64, else 512): see
    // gen_spn_forward.cu - pad the shared-memory request of the small (no TMA ring) instantiations accordingly
    constexpr size_t min_dyn = C <= 64 ? (size_t)76 * 1024 : (size_t)116 * 1024;
    if (dyn < min_dyn) dyn = min_dyn;
    const cudaError_t attr = ensure_dynamic_smem((const void*)gen_grad_feature_kernel<FT, C, TMA, CS>, dyn);
    if (attr != cudaSuccess) return attr;
    dim3 grid((unsigned)((size_t)la.g.tiles_x * la.g.tiles_y * la.g.B));
    gen_grad_feature_kernel<FT, C, TMA, CS><<<grid, GEN_CTA_THREADS, dyn, la.stream>>>((const FT*)gz, conv_w,
                                                                                     (FT*)grad_feature, la.g, tmap_gz);
    return cudaGetLastError();
}
template <typename FT, int C, bool TMA>
static cudaError_t launch_gf_one(const LaunchArgs& la, const CUtensorMap& tmap_gz, const void* gz, const float* conv_w,
                                 void* grad_feature) {
    if (TMA && (size_t)la.g.H * la.g.W == 16384) return launch_gf_cs<FT, C, TMA, TMA ? 16384 : 0>(la, tmap_gz, gz, conv_w, grad_feature);
    return launch_gf_cs<FT, C, TMA, 0>(la, tmap_gz, gz, conv_w, grad_feature);
}

// la.g: geometry with 16 rows per CTA; la.use_tma: tmap_gz is valid; la.bf16: gz / grad_feature are bf16
cudaError_t launch_gen_grad_feature(const LaunchArgs& la, const CUtensorMap& tmap_gz, const void* gz, int C,
                                    const float* conv_w, void* grad_feature) {
    if (C == 64) {
        if (la.bf16) return la.use_tma ? launch_gf_one<__nv_bfloat16, 64, true>(la, tmap_gz, gz, conv_w, grad_feature)
                                       : launch_gf_one<__nv_bfloat16, 64, false>(la, tmap_gz, gz, conv_w, grad_feature);
        return la.use_tma ? launch_gf_one<float, 64, true>(la, tmap_gz, gz, conv_w, grad_feature)
                          : launch_gf_one<float, 64, false>(la, tmap_gz, gz, conv_w, grad_feature);
    }
    if (C == 128) {
        if (la.bf16) return la.use_tma ? launch_gf_one<__nv_bfloat16, 128, true>(la, tmap_gz, gz, conv_w, grad_feature)
                                       : launch_gf_one<__nv_bfloat16, 128, false>(la, tmap_gz, gz, conv_w, grad_feature);
        return la.use_tma ? launch_gf_one<float, 128, true>(la, tmap_gz, gz, conv_w, grad_feature)
                          : launch_gf_one<float, 128, false>(la, tmap_gz, gz, conv_w, grad_feature);
    }
    return cudaErrorNotSupported;
}

}  // namespace JSPSR_VARIANT
}  // namespace jspsr
