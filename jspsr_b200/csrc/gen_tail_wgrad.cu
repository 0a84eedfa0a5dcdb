// Parameter gradients of the Generator tail (backward of the two 1x1 convolutions of models/components/spn.py:41-52
// w.r.t. their weights and biases; in the reference this is inside cuDNN's convolution backward):
//   grad_conv_w[j, c] = sum over (b, y, x) of gz[b, j, y, x] * feature[b, c, y, x]        j = 0..24, c = 0..C-1
//   grad_conv_b[j]    = sum over (b, y, x) of gz[b, j, y, x]
// with gz the pre-activation gradients spn_backward_kernel writes in JSPSR_BWD_GEN_PREACT mode.  A contraction whose K
// dimension is the PIXEL index (16.7 M at 1024 tiles) and whose output is tiny (25 x C): one pass over HBM, 100 + 4C
// bytes per pixel, everything else stays on chip.  Replaces torch.bmm (cuBLAS SIMT fp32: 3.1 ms at C = 128 / 1024 tiles,
// 0.51 of this roofline) + a separate reduction pass for the bias.
//
// Both operands are K-major in the native NCHW layout (a plane's pixels are contiguous), so the MMA is
//   D[c, j] (+)= sum_k F[c, k] * G[j, k]       M = 128 (channels = TMEM lanes), N = 32 (25 rows of gz, padded), K = 8
// on tcgen05 (kind::tf32) with the 3-product split F_hi*G_hi + F_lo*G_hi + F_hi*G_lo (hi = tf32 round-to-nearest,
// lo = the exact remainder, which the tensor core truncates to tf32: 2^-22 relative per product, fp32 level).
// Per CTA, 320 threads, two CTAs per SM, a contiguous range of 32-pixel K-blocks:
//   TMA warp  (warp 9)   : per K-block one [C planes][32 px] box of the feature (SWIZZLE_128B, so that a thread can walk
//                          its own row without bank conflicts) and one [25][32 px] box of gz into a 4-stage ring;
//   warps 4-7 (feature)  : thread = channel: its 32 pixels from the ring -> hi / lo -> its own TMEM lane (A operand,
//                          double buffered: the MMAs of block i read one buffer while block i + 1 is written);
//   warps 0-3 (gz)       : the 25 x 32 values of gz -> hi / lo -> K-major core matrices in shared memory (B operand,
//                          double buffered), plus the running bias sums; one block after the end of every accumulation
//                          run they read its accumulator (thread = channel = TMEM lane) into their fp64 totals;
//   MMA warp  (warp 8)   : 4 K-steps x 3 products per block into accumulator (run & 1).
// An accumulation run is `run_len` K-blocks (default 16 = 512 pixels): the tensor core adds into its fp32 accumulator
// with less than round-to-nearest care, so runs are kept short and combined in fp64 REGISTERS (thread = channel holds
// its 25 totals), double buffered so that the MMAs of run r + 1 proceed while run r is folded one block later (no
// pipeline drain).  A CTA touches global memory for its results once, at the end: RED.F64 into a caller-owned,
// zero-on-entry / zero-on-exit workspace; the last CTA (ticket) rounds the totals to fp32.
// TMEM: 2 x (32 hi + 32 lo) + 2 x 32 = 192 -> 256 columns.
#include "spn_kernels.cuh"
#include "umma_helpers.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

constexpr int GW_ROW_BYTES = 128;  // a K-block is one 128-byte row of every plane: 32 fp32 / 64 bf16 pixels
constexpr int GW_STAGES = 4;
constexpr int GW_N = 32;           // MMA N: 25 rows of gz, zero padded
constexpr int GW_GZ_BYTES = 4096;  // ring bytes reserved for the gz box (25 x 128 used; keeps the stages 1024-byte aligned)
template <typename FT> constexpr int gw_kb() { return GW_ROW_BYTES / (int)sizeof(FT); }   // pixels per K-block
template <typename FT> constexpr int gw_b_bytes() { return GW_N * gw_kb<FT>() * 4; }        // one B buffer (tf32)

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

// ws layout (doubles): [25 * C] weight sums | [32] bias sums | ticket (one 8-byte word)
// FT = float: hi / lo split of both operands, three products.  FT = __nv_bfloat16 (torch.autocast: Generator.block emits
// bf16 and the backward kernel writes bf16 gradients): every value is exact in tf32, one product, 64 pixels per K-block.
template <typename FT, int C, bool TMA>
__global__ void __launch_bounds__(GEN_CTA_THREADS, 2)
gen_grad_weight_kernel(const FT* __restrict__ gz, const FT* __restrict__ feature, float* __restrict__ grad_w,
                       float* __restrict__ grad_b, double* __restrict__ ws, const long long n_blocks,
                       const int blocks_per_sample, const int HW, const int run_len,
                       const __grid_constant__ CUtensorMap tmap_f, const __grid_constant__ CUtensorMap tmap_gz) {
    constexpr bool F16 = sizeof(FT) == 2;
    constexpr int GW_KB = gw_kb<FT>(), GW_B_BYTES = gw_b_bytes<FT>();
    constexpr int F_BYTES = C * GW_ROW_BYTES, STAGE_BYTES = F_BYTES + GW_GZ_BYTES;
    constexpr int RING_BYTES = TMA ? GW_STAGES * STAGE_BYTES : 0;
    constexpr uint32_t SBO = 128, LBO = GW_N / 8 * 128;         // K-major, no swizzle: 8-row groups / 16-byte K chunks of B
    // A buffer: fp32 = 32 hi + 32 lo columns, bf16 = 64 exact columns
    constexpr uint32_t A_BUF = 64, COL_LO = 32, COL_ACC = 2 * A_BUF, TMEM_COLS = 256;
    // instruction descriptor: D fp32 | A, B tf32 | both K-major | N = 32 | M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GW_N >> 3) << 17) | ((GEN_THREADS >> 4) << 24);
    extern __shared__ __align__(1024) unsigned char dsm_raw[];
    // SWIZZLE_128B boxes want a 1024-byte aligned destination: align by hand (the launch reserves the slack)
    unsigned char* dsm = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
    unsigned char* ring = dsm;                         // [GW_STAGES][ feature C x 128 B (swizzled) | gz 25 x 128 B ]
    unsigned char* b_hi = dsm + RING_BYTES;            // [2][32 rows x GW_KB] tf32, core-matrix layout
    unsigned char* b_lo = b_hi + 2 * GW_B_BYTES;       // fp32 only
    __shared__ __align__(8) uint64_t bar_full[GW_STAGES], bar_empty[GW_STAGES], bar_a_full[2], bar_b_full[2], bar_ab_free[2],
        bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t s_tmem;
    __shared__ bool s_last;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    // this CTA's contiguous range of K-blocks
    const long long kb0 = n_blocks * blockIdx.x / gridDim.x, kb1 = n_blocks * (blockIdx.x + 1) / gridDim.x;
    const int n = (int)(kb1 - kb0);

    if (t == 32) {
        for (int i = 0; i < GW_STAGES; ++i) {
            mbar_init(&bar_full[i], 1);
            mbar_init(&bar_empty[i], 2 * GEN_THREADS);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_a_full[i], GEN_THREADS);
            mbar_init(&bar_b_full[i], GEN_THREADS);
            mbar_init(&bar_ab_free[i], 1);
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
    // rows 25..31 of the B operand stay zero for the whole kernel
    for (int i = t; i < (F16 ? 2 : 4) * GW_B_BYTES / 16; i += GEN_CTA_THREADS)
        reinterpret_cast<float4*>(b_hi)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 9) {
        // =========================== TMA warp ===========================
        if (TMA && lane == 0) {
#pragma unroll 1
            for (int i = 0; i < n; ++i) {
                const int s = i % GW_STAGES;
                if (i >= GW_STAGES) mbar_wait(&bar_empty[s], (uint32_t)(((i / GW_STAGES) - 1) & 1));
                const long long kb = kb0 + i;
                const int b = (int)(kb / blocks_per_sample), pix0 = (int)(kb % blocks_per_sample) * GW_KB;
                mbar_arrive_expect_tx(&bar_full[s], F_BYTES + GEN_NOUT * GW_ROW_BYTES);
                tma_load_2d(ring + s * STAGE_BYTES, &tmap_f, &bar_full[s], pix0, b * C);
                tma_load_2d(ring + s * STAGE_BYTES + F_BYTES, &tmap_gz, &bar_full[s], pix0, b * GEN_NOUT);
            }
        }
    } else if (warp == 8) {
        // =========================== MMA warp (converged) ===========================
        int in_run = 0, run = 0;
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const int p = i & 1, q = run & 1;
            const bool first = in_run == 0, last = in_run == run_len - 1 || i == n - 1;
            mbar_wait(&bar_a_full[p], (uint32_t)((i >> 1) & 1));
            mbar_wait(&bar_b_full[p], (uint32_t)((i >> 1) & 1));
            if (first && run >= 2) mbar_wait(&bar_acc_empty[q], (uint32_t)(((run >> 1) - 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem + COL_ACC + (uint32_t)q * GW_N;
            const uint32_t a_tmem = tmem + (uint32_t)p * A_BUF;
#pragma unroll
            for (int ks = 0; ks < GW_KB / 8; ++ks) {
                const uint64_t dbh = umma_desc_kmajor(smem_u32(b_hi) + p * GW_B_BYTES + ks * 2 * LBO, LBO, SBO);
                umma_tf32_ts(d_tmem, a_tmem + ks * 8, dbh, IDESC, (first && ks == 0) ? 0u : 1u);
                if (!F16) {
                    const uint64_t dbl = umma_desc_kmajor(smem_u32(b_lo) + p * GW_B_BYTES + ks * 2 * LBO, LBO, SBO);
                    umma_tf32_ts(d_tmem, a_tmem + COL_LO + ks * 8, dbh, IDESC, 1u);
                    umma_tf32_ts(d_tmem, a_tmem + ks * 8, dbl, IDESC, 1u);
                }
            }
            umma_commit(&bar_ab_free[p]);
            if (last) {
                umma_commit(&bar_acc_full[q]);
                in_run = 0;
                ++run;
            } else {
                ++in_run;
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // =========================== feature producers: thread = channel = TMEM lane ===========================
        const int c = t - GEN_THREADS;
        const uint32_t lane_tmem = tmem + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const int s = i % GW_STAGES, p = i & 1;
            if (TMA) mbar_wait(&bar_full[s], (uint32_t)((i / GW_STAGES) & 1));
            if (i >= 2) {  // the MMAs of block i - 2 no longer read this A buffer
                mbar_wait(&bar_ab_free[p], (uint32_t)(((i >> 1) - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t a_lane = lane_tmem + (uint32_t)p * A_BUF;
            const unsigned char* row = ring + s * STAGE_BYTES + c * GW_ROW_BYTES;
            const long long kb = kb0 + i;
            const int b = (int)(kb / blocks_per_sample), pix0 = (int)(kb % blocks_per_sample) * GW_KB;
            const FT* gp = feature + ((size_t)b * C + (c < C ? c : 0)) * (size_t)HW + pix0;
#pragma unroll
            for (int k0 = 0; k0 < GW_KB; k0 += 16) {  // 16 pixels = 4 (fp32) / 2 (bf16) 16-byte chunks of the row
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = 0.f;
                if (c < C) {
                    if (TMA) {  // SWIZZLE_128B: 16-byte chunk q of row r lives at chunk q ^ (r & 7)
                        constexpr int PER = 16 / (int)sizeof(FT);  // pixels per chunk
#pragma unroll
                        for (int q = 0; q < 16 / PER; ++q) {
                            const uint4 raw = *reinterpret_cast<const uint4*>(row + (((k0 / PER + q) ^ (c & 7)) << 4));
                            if (F16) {  // a bf16 is the upper half of its fp32
                                const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    v[q * PER + 2 * e] = __uint_as_float(wds[e] << 16);
                                    v[q * PER + 2 * e + 1] = __uint_as_float(wds[e] & 0xFFFF0000u);
                                }
                            } else {
                                v[q * PER] = __uint_as_float(raw.x);
                                v[q * PER + 1] = __uint_as_float(raw.y);
                                v[q * PER + 2] = __uint_as_float(raw.z);
                                v[q * PER + 3] = __uint_as_float(raw.w);
                            }
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e)
                            if (pix0 + k0 + e < HW) v[e] = to_f32(ld_stream(gp + k0 + e));
                    }
                }
                if (F16) {
                    tmem_st16(a_lane + k0, v);
                } else {
                    float hi[16], lo[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        hi[e] = tf32_rn(v[e]);
                        lo[e] = v[e] - hi[e];
                    }
                    tmem_st16(a_lane + k0, hi);
                    tmem_st16(a_lane + COL_LO + k0, lo);
                }
            }
            if (TMA) mbar_arrive(&bar_empty[s]);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&bar_a_full[p]);
        }
    } else {
        // =========================== gz producers + accumulator folding ===========================
        const int j = t >> 2, q4 = t & 3;  // row of gz / quarter (32 bytes) of the K-block's row handled by this thread
        constexpr int QP = GW_KB / 4;      // pixels per quarter: 8 (fp32) / 16 (bf16)
        const uint32_t boff = (uint32_t)((j >> 3) * SBO + (j & 7) * 16);
        const uint32_t acc_lane = tmem + ((uint32_t)(warp * 32) << 16) + COL_ACC;
        double tot[GEN_NOUT];  // thread = channel t: its 25 weight-gradient totals
#pragma unroll
        for (int jj = 0; jj < GEN_NOUT; ++jj) tot[jj] = 0.0;
        double btot = 0.0;
        float bsum = 0.f;
        // run `r` (accumulator r & 1) is folded into the totals; called one block after its last MMA was issued
        auto fold = [&](const int r) {
            const int q = r & 1;
            __syncwarp();
            mbar_wait(&bar_acc_full[q], (uint32_t)((r >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[16];
            tmem_ld16(acc_lane + (uint32_t)q * GW_N, v);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) tot[jj] += (double)v[jj];
            tmem_ld16(acc_lane + (uint32_t)q * GW_N + 16, v);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&bar_acc_empty[q]);
#pragma unroll
            for (int jj = 16; jj < GEN_NOUT; ++jj) tot[jj] += (double)v[jj - 16];
            btot += (double)bsum;
            bsum = 0.f;
        };
        int in_run = 0, run = 0;
        bool pending = false;  // the previous block closed run `run - 1`, not folded yet
#pragma unroll 1
        for (int i = 0; i < n; ++i) {
            const int s = i % GW_STAGES, p = i & 1;
            if (TMA) mbar_wait(&bar_full[s], (uint32_t)((i / GW_STAGES) & 1));
            if (i >= 2) mbar_wait(&bar_ab_free[p], (uint32_t)(((i >> 1) - 1) & 1));
            float part = 0.f;
            if (j < GEN_NOUT) {
                float e[QP];
                if (TMA) {
                    const unsigned char* src = ring + s * STAGE_BYTES + F_BYTES + j * GW_ROW_BYTES + q4 * 32;
                    const uint4 r0 = *reinterpret_cast<const uint4*>(src), r1 = *reinterpret_cast<const uint4*>(src + 16);
                    const uint32_t wds[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        if (F16) {
                            e[2 * k] = __uint_as_float(wds[k] << 16);
                            e[2 * k + 1] = __uint_as_float(wds[k] & 0xFFFF0000u);
                        } else {
                            e[k] = __uint_as_float(wds[k]);
                        }
                    }
                } else {
                    const long long kb = kb0 + i;
                    const int b = (int)(kb / blocks_per_sample), pix = (int)(kb % blocks_per_sample) * GW_KB + q4 * QP;
                    const FT* gp = gz + ((size_t)b * GEN_NOUT + j) * (size_t)HW + pix;
#pragma unroll
                    for (int k = 0; k < QP; ++k) e[k] = pix + k < HW ? to_f32(ld_stream(gp + k)) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < QP; k += 8)
                    part += ((e[k] + e[k + 1]) + (e[k + 2] + e[k + 3])) + ((e[k + 4] + e[k + 5]) + (e[k + 6] + e[k + 7]));
                // K-major core matrices: 4 consecutive pixels of row j = one 16-byte chunk; this thread owns chunks
                // q4 * QP / 4 ... of the block
                unsigned char* dh = b_hi + p * GW_B_BYTES + boff + (q4 * (QP / 4)) * LBO;
                unsigned char* dl = b_lo + p * GW_B_BYTES + boff + (q4 * (QP / 4)) * LBO;
#pragma unroll
                for (int k = 0; k < QP; k += 4) {
                    if (F16) {
                        *reinterpret_cast<float4*>(dh + (k / 4) * LBO) = make_float4(e[k], e[k + 1], e[k + 2], e[k + 3]);
                    } else {
                        const float4 hh = make_float4(tf32_rn(e[k]), tf32_rn(e[k + 1]), tf32_rn(e[k + 2]), tf32_rn(e[k + 3]));
                        *reinterpret_cast<float4*>(dh + (k / 4) * LBO) = hh;
                        *reinterpret_cast<float4*>(dl + (k / 4) * LBO) =
                            make_float4(e[k] - hh.x, e[k + 1] - hh.y, e[k + 2] - hh.z, e[k + 3] - hh.w);
                    }
                }
            }
            if (TMA) mbar_arrive(&bar_empty[s]);
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
            mbar_arrive(&bar_b_full[p]);

            if (pending) {  // bsum still holds exactly the closed run's pixels: fold before adding this block's
                fold(run - 1);
                pending = false;
            }
            bsum += part;
            if (in_run == run_len - 1 || i == n - 1) {
                pending = true;
                in_run = 0;
                ++run;
            } else {
                ++in_run;
            }
        }
        if (pending) fold(run - 1);
        // ---- this CTA's totals -> the fp64 workspace: for each j a warp's 32 lanes hit 32 consecutive doubles ----
        if (n > 0) {
            if (t < C) {
#pragma unroll
                for (int jj = 0; jj < GEN_NOUT; ++jj) atomicAdd(ws + (size_t)jj * C + t, tot[jj]);
            }
            // bias: thread (j, q4) summed row j's quarter q4
            btot += __shfl_xor_sync(0xffffffffu, btot, 1);
            btot += __shfl_xor_sync(0xffffffffu, btot, 2);
            if (q4 == 0 && j < GEN_NOUT) atomicAdd(ws + (size_t)GEN_NOUT * C + j, btot);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __threadfence();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(TMEM_COLS));
    // ---- last CTA: fp64 totals -> fp32 results, workspace left zeroed ----
    unsigned long long* ticket = reinterpret_cast<unsigned long long*>(ws + GEN_NOUT * C + 32);
    if (t == 0) s_last = atomicAdd(ticket, 1ull) == (unsigned long long)gridDim.x - 1ull;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int i = t; i < GEN_NOUT * C + GEN_NOUT; i += GEN_CTA_THREADS) {
            const double v = atomicAdd(ws + i, 0.0);  // coherent read
            if (i < GEN_NOUT * C) {
                if (grad_w) grad_w[i] = (float)v;
            } else if (grad_b) {
                grad_b[i - GEN_NOUT * C] = (float)v;
            }
            ws[i] = 0.0;
        }
        if (t == 0) *ticket = 0ull;
    }
}

size_t gen_grad_weight_workspace_bytes() { return (size_t)(GEN_NOUT * 128 + 32 + 1) * sizeof(double); }

template <typename FT, int C, bool TMA>
static cudaError_t launch_gw(const void* gz, const void* feature, float* grad_w, float* grad_b, void* ws, int B, int HW,
                             int run_len, const CUtensorMap& tmap_f, const CUtensorMap& tmap_gz, cudaStream_t stream) {
    // two CTAs per SM by construction: 256 of the SM's 512 TMEM columns each; the request is padded for the ring-less
    // instantiation so that a third CTA can never become resident and spin in tcgen05.alloc
    constexpr int KB = gw_kb<FT>();
    size_t dyn = (TMA ? (size_t)GW_STAGES * (C * GW_ROW_BYTES + GW_GZ_BYTES) : 0) + 4 * gw_b_bytes<float>() + 1024;  // + alignment slack
    if (dyn < (size_t)76 * 1024) dyn = (size_t)76 * 1024;
    const cudaError_t attr = ensure_dynamic_smem((const void*)gen_grad_weight_kernel<FT, C, TMA>, dyn);
    if (attr != cudaSuccess) return attr;
    const int bps = (HW + KB - 1) / KB;
    const long long n_blocks = (long long)B * bps;
    int sms = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (long long)2 * sms;
    const unsigned grid = (unsigned)(n_blocks < want ? n_blocks : want);
    gen_grad_weight_kernel<FT, C, TMA><<<grid, GEN_CTA_THREADS, dyn, stream>>>((const FT*)gz, (const FT*)feature, grad_w, grad_b,
                                                                              (double*)ws, n_blocks, bps, HW, run_len, tmap_f,
                                                                              tmap_gz);
    return cudaGetLastError();
}
template <typename FT>
static cudaError_t launch_gw_c(const void* gz, const void* feature, int C, float* grad_w, float* grad_b, void* ws, int B, int HW,
                               int run_len, bool use_tma, const CUtensorMap& tmap_f, const CUtensorMap& tmap_gz,
                               cudaStream_t stream) {
    if (C == 128) {
        return use_tma ? launch_gw<FT, 128, true>(gz, feature, grad_w, grad_b, ws, B, HW, run_len, tmap_f, tmap_gz, stream)
                       : launch_gw<FT, 128, false>(gz, feature, grad_w, grad_b, ws, B, HW, run_len, tmap_f, tmap_gz, stream);
    }
    if (C == 64) {
        return use_tma ? launch_gw<FT, 64, true>(gz, feature, grad_w, grad_b, ws, B, HW, run_len, tmap_f, tmap_gz, stream)
                       : launch_gw<FT, 64, false>(gz, feature, grad_w, grad_b, ws, B, HW, run_len, tmap_f, tmap_gz, stream);
    }
    return cudaErrorInvalidValue;
}

// use_tma: tmap_f ([B*C planes][HW], box [C][128 bytes], SWIZZLE_128B) and tmap_gz ([B*25 planes][HW], box [25][128 bytes])
// are valid; bf16: gz and feature are bf16 (64 pixels per K-block)
cudaError_t launch_gen_grad_weight(const void* gz, const void* feature, int C, bool bf16, float* grad_w, float* grad_b, void* ws,
                                   int B, int HW, int run_len, bool use_tma, const CUtensorMap& tmap_f,
                                   const CUtensorMap& tmap_gz, cudaStream_t stream) {
    if (run_len < 2) run_len = 2;  // a run is folded one block after it closes: two accumulators need runs of >= 2 blocks
    if (bf16) return launch_gw_c<__nv_bfloat16>(gz, feature, C, grad_w, grad_b, ws, B, HW, run_len, use_tma, tmap_f, tmap_gz, stream);
    return launch_gw_c<float>(gz, feature, C, grad_w, grad_b, ws, B, HW, run_len, use_tma, tmap_f, tmap_gz, stream);
}

}  // namespace JSPSR_VARIANT
}  // namespace jspsr
