// tcgen05 / TMEM helpers shared by the generator-tail kernels (gen_spn_forward.cu, gen_tail_backward.cu).
#pragma once
#include "spn_common.cuh"

namespace jspsr {
inline namespace JSPSR_VARIANT {

constexpr int GEN_THREADS = 128;  // one thread per pixel of a 128-pixel row segment = one TMEM lane each
constexpr int GEN_N = 32;         // MMA N: 9 weight + 16 offset rows, padded
constexpr int GEN_NOUT = 25;
constexpr int GEN_CTA_THREADS = 2 * GEN_THREADS + 64;  // consumers + producers + the MMA warp + the TMA warp

__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    // shared-memory matrix descriptor, SWIZZLE_NONE, K-major: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ float tf32_rn(float v) {  // round to nearest tf32 (10 explicit mantissa bits)
    return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The MMA warp runs these CONVERGED (all 32 lanes, one elected lane issues): inside a divergent `if (lane == 0)`
// the compiler cannot keep addresses and descriptors on the uniform datapath and wraps every MMA in an
// ELECT / R2UR / branch sequence (~100 cycles each, measured: the issue loop was the kernel's critical path).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {  // arrives on `bar` when all MMAs issued so far are complete
    asm volatile(
        "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar))
        : "memory");
}
// A operand from tensor memory (lane = row of A = pixel, one tf32 per column), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nelect.sync _|q, 0xffffffff;\n"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}


}  // namespace JSPSR_VARIANT
}  // namespace jspsr
