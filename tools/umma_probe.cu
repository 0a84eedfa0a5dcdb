// Debug probe for the tcgen05 plumbing used by the generator-tail kernel: D[128 x 32] = A[128 x K] * B[32 x K]^T
// with A/B in shared memory in the canonical K-major, no-swizzle ("interleave") layout, tf32 inputs, fp32
// accumulation in TMEM.  Prints the max error of (a) one tf32 product, (b) the 3-product split
// (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo) against an fp64 host product.   ./umma_probe [K=64]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 bytes (4 tf32), rows 16 B apart;
// SBO = bytes between 8-row groups, LBO = bytes between 16-byte K chunks
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D1, float* D3, float* D4, int K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* a_hi = reinterpret_cast<float*>(smem);
    float* a_lo = a_hi + 128 * K;
    float* b_hi = a_lo + 128 * K;
    float* b_lo = b_hi + 32 * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t LBO_A = 128 / 8 * 128, LBO_B = 32 / 8 * 128, SBO = 128;

    // stage operands: element (row m, k) at (m/8)*SBO + (k/4)*LBO + (m%8)*16 + (k%4)*4
    for (int kc = 0; kc < K / 4; ++kc) {
        float4 hi, lo;
        float* h = &hi.x;
        float* l = &lo.x;
        for (int j = 0; j < 4; ++j) {
            const float v = A[t * K + kc * 4 + j];
            h[j] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            l[j] = v - h[j];
        }
        const uint32_t off = (t / 8) * SBO + kc * LBO_A + (t % 8) * 16;
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(a_hi) + off) = hi;
        *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(a_lo) + off) = lo;
        if (t < 32) {
            for (int j = 0; j < 4; ++j) {
                const float v = B[t * K + kc * 4 + j];
                h[j] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
                l[j] = v - h[j];
            }
            const uint32_t offb = (t / 8) * SBO + kc * LBO_B + (t % 8) * 16;
            *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(b_hi) + offb) = hi;
            *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(b_lo) + offb) = lo;
        }
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy operand stores -> async proxy (UMMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    // instruction descriptor: D fp32, A/B tf32, both K-major, N = 32, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    if (t == 0) {
        for (int pass = 0; pass < 2; ++pass) {        // pass 0: hi*hi only into columns [0,32); pass 1: 3 products into [32,64)
            const uint32_t d_tmem = tm + pass * 32;
            uint32_t acc = 0;
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t dah = make_desc(smem_u32(a_hi) + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t dal = make_desc(smem_u32(a_lo) + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t dbh = make_desc(smem_u32(b_hi) + ks * 2 * LBO_B, LBO_B, SBO);
                const uint64_t dbl = make_desc(smem_u32(b_lo) + ks * 2 * LBO_B, LBO_B, SBO);
                const int nprod = pass == 0 ? 1 : 3;
                for (int pr = 0; pr < nprod; ++pr) {
                    const uint64_t da = pr == 1 ? dal : dah;
                    const uint64_t db = pr == 2 ? dbl : dbh;
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
                        "l"(da), "l"(db), "r"(idesc), "r"(acc)
                        : "memory");
                    acc = 1;
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n.reg .pred P1;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}\n" ::"r"(smem_u32(&bar)),
        "r"(0)
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int pass = 0; pass < 2; ++pass) {
        uint32_t r[32];
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + pass * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* D = pass == 0 ? D1 : D3;
        for (int j = 0; j < 32; ++j) D[t * 32 + j] = __uint_as_float(r[j]);
    }
    // ---- pass 2: A operand from TMEM (written by its owner threads with tcgen05.st), columns [64, 64+K) hi, [64+K, 64+2K) lo ----
    {
        const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
        for (int k0 = 0; k0 < K; k0 += 16) {
            uint32_t h[16], l[16];
            for (int j = 0; j < 16; ++j) {
                const float v = A[t * K + k0 + j];
                const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
                h[j] = __float_as_uint(hi);
                l[j] = __float_as_uint(v - hi);
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                         ::"r"(lane_base + 64 + k0), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]),
                         "r"(h[8]), "r"(h[9]), "r"(h[10]), "r"(h[11]), "r"(h[12]), "r"(h[13]), "r"(h[14]), "r"(h[15]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
                         ::"r"(lane_base + 64 + K + k0), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]),
                         "r"(l[8]), "r"(l[9]), "r"(l[10]), "r"(l[11]), "r"(l[12]), "r"(l[13]), "r"(l[14]), "r"(l[15]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (t == 0) {
            const uint32_t d_tmem = tm;  // reuse columns [0,32)
            uint32_t acc = 0;
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t dbh = make_desc(smem_u32(b_hi) + ks * 2 * LBO_B, LBO_B, SBO);
                const uint64_t dbl = make_desc(smem_u32(b_lo) + ks * 2 * LBO_B, LBO_B, SBO);
                for (int pr = 0; pr < 3; ++pr) {
                    const uint32_t ta = tm + 64 + (pr == 1 ? K : 0) + ks * 8;
                    const uint64_t db = pr == 2 ? dbl : dbh;
                    asm volatile(
                        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
                        "r"(ta), "l"(db), "r"(idesc), "r"(acc)
                        : "memory");
                    acc = 1;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        asm volatile(
            "{\n.reg .pred P1;\nWAIT_LOOP2:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra WAIT_DONE2;\nbra WAIT_LOOP2;\nWAIT_DONE2:\n}\n" ::"r"(smem_u32(&bar)),
            "r"(1)
            : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(lane_base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D4[t * 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256));
}

int main(int argc, char** argv) {
    const int K = argc > 1 ? atoi(argv[1]) : 64;
    std::vector<float> A(128 * K), B(32 * K), D1(128 * 32), D3(128 * 32), D4(128 * 32);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;
    float *dA, *dB, *dD1, *dD3, *dD4;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD1, D1.size() * 4); cudaMalloc(&dD3, D3.size() * 4); cudaMalloc(&dD4, D4.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = (size_t)(2 * 128 + 2 * 32) * K * 4 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<<<1, 128, smem>>>(dA, dB, dD1, dD3, dD4, K);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(D1.data(), dD1, D1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D3.data(), dD3, D3.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D4.data(), dD4, D4.size() * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e3 = 0, eh = 0, scale = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
            double ref = 0, refh = 0;
            for (int k = 0; k < K; ++k) {
                ref += (double)A[m * K + k] * B[n * K + k];
                uint32_t ua, ub;
                memcpy(&ua, &A[m * K + k], 4); memcpy(&ub, &B[n * K + k], 4);
                ua &= 0xFFFFE000u; ub &= 0xFFFFE000u;
                float fa, fb;
                memcpy(&fa, &ua, 4); memcpy(&fb, &ub, 4);
                refh += (double)fa * fb;
            }
            e1 = fmax(e1, fabs(D1[m * 32 + n] - ref));
            eh = fmax(eh, fabs(D1[m * 32 + n] - refh));
            e3 = fmax(e3, fabs(D3[m * 32 + n] - ref));
            scale = fmax(scale, fabs(ref));
        }
    printf("K=%d scale=%.4f | 1xTF32 err vs exact %.3e, vs truncated-operand product %.3e | 3xTF32 err %.3e (rel %.3e)\n", K, scale,
           e1, eh, e3, e3 / scale);
    double e4 = 0, d34 = 0;
    for (int i = 0; i < 128 * 32; ++i) d34 = fmax(d34, fabs((double)D4[i] - D3[i]));
    printf("A-from-TMEM (TS) 3xTF32: max |D4 - D3| = %.3e\n", d34);
    (void)e4;
    printf("D3[0][0..3] = %f %f %f %f\n", D3[0], D3[1], D3[2], D3[3]);
    return 0;
}
