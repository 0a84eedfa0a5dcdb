// Shared device helpers for the propagation kernels (sm_100a only).
#pragma once
#include "spn_types.cuh"

// The kernels are compiled twice, once per staged-halo size; JSPSR_VARIANT names the inline
// namespace so both sets of symbols can live in one library (see Makefile, abi.cu).
#ifndef JSPSR_VARIANT
#define JSPSR_VARIANT narrow
#endif

namespace jspsr {
inline namespace JSPSR_VARIANT {

// ---------------------------------------------------------------------------
// Tile geometry.  One CTA owns a TILE_H x TILE_W block of output pixels of one
// sample and stages the DEM rows/cols its taps can reach in shared memory:
// HALO_* pixels around the block (offsets up to +-5 px stay on chip; anything
// further takes the bounds-checked global path).  The staged tile is a dense
// [SH][SW] box so one 3-D TMA box copy (zero-filled outside the image, which is
// exactly torchvision's zero-outside rule) can fill it.
// ---------------------------------------------------------------------------
// Rows per CTA are a template parameter TH in {16, 8, 4, 2}: 16 for large problems (least
// halo overhead, 8 pixels per thread), smaller for small batches so that the grid still
// covers all 148 SMs a few times over and a thread's serial chain of pixels stays short
// (the reference's own batch of 70 tiles is only 560 CTAs at TH = 16).
// Two halo sizes (measured, B200): the wide one keeps offsets up to +-13 rows / +-14 columns on
// chip and is 2-25 % faster on tile batches (sigma 1.5 .. 4 px offsets), the narrow one (+-5 px) is
// 5-9 % faster on whole rasters, where every staged row is a separate 32 KB-strided DRAM segment.
#ifdef JSPSR_HALO_WIDE
constexpr int HALO_T = 14, HALO_B = 15;  // rows above / below  (bottom needs the +1 bilinear row)
#else
constexpr int HALO_T = 6, HALO_B = 7;
#endif
// cols left / right.  Measured on B200 (tools/tma_probe.cu): the innermost TMA box
// coordinate must be 16-byte aligned (c0 * sizeof(T) % 16 == 0; negative is fine, an
// unaligned c0 raises "illegal instruction"), rows are free.  x0 is a multiple of 128, so
// a left halo of 8 keeps x0 - 8 aligned for fp32 and bf16; SW must be a multiple of 8
// elements for the bf16 box (inner extent multiple of 16 bytes).
#ifdef JSPSR_HALO_WIDE
constexpr int HALO_L = 16, HALO_R = 16;
#else
constexpr int HALO_L = 8, HALO_R = 8;
#endif
constexpr int SW = TILE_W + HALO_L + HALO_R;  // 144
constexpr int staged_rows(int th) { return th + HALO_T + HALO_B; }  // 29 for TH = 16
static_assert(SW % 8 == 0 && HALO_L % 8 == 0 && TILE_W % 8 == 0, "TMA inner box extent / origin must be multiples of 16 bytes");
static_assert(SW <= 256 && staged_rows(16) <= 256, "TMA box extents are limited to 256");

// ---------------------------------------------------------------------------
// element access
// ---------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// streaming (read-once) loads: keep them out of L1 so the gather fallback and the
// staged tile keep the cache
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const __nv_bfloat16* p) {
    unsigned short u;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(u) : "l"(p));
    return __uint_as_float(((unsigned)u) << 16);
}
__device__ __forceinline__ unsigned ld_stream_u16(const __nv_bfloat16* p) {  // the bf16's bits, zero-extended
    unsigned short u;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(u) : "l"(p));
    return (unsigned)u;
}
// two horizontally adjacent bf16 pixels as one 32-bit word (4-byte aligned: even x on rows of even length)
__device__ __forceinline__ uint32_t ld_stream_x2(const __nv_bfloat16* p) {
    uint32_t u;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(u) : "l"(p));
    return u;
}
// half 0 = the pixel at the lower address; a bf16 widens to fp32 by moving its 16 bits to the top
__device__ __forceinline__ float bf16x2_half(uint32_t u, int half) {
    return __uint_as_float(half == 0 ? (u << 16) : (u & 0xffff0000u));
}
// two adjacent pixels of an fp32 / bf16 plane as two floats
__device__ __forceinline__ void ld_stream_pair(const float* p, float (&v)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
}
__device__ __forceinline__ void ld_stream_pair(const __nv_bfloat16* p, float (&v)[2]) {
    const uint32_t u = ld_stream_x2(p);
    v[0] = bf16x2_half(u, 0);
    v[1] = bf16x2_half(u, 1);
}
__device__ __forceinline__ void st_stream_x2(float* p, float v0, float v1) {
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v0), "f"(v1) : "memory");
}
__device__ __forceinline__ void st_stream_x2(__nv_bfloat16* p, float v0, float v1) {
    const __nv_bfloat162 b = __floats2bfloat162_rn(v0, v1);  // .x = v0 (low half, lower address)
    asm volatile("st.global.cs.b32 [%0], %1;" ::"l"(p), "r"(*reinterpret_cast<const uint32_t*>(&b)) : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream(__nv_bfloat16* p, float v) {
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p), "h"(*reinterpret_cast<unsigned short*>(&b)) : "memory");
}

// ---------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk.tensor) wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n"  // suspend-time hint: sleep, do not spin
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 3-D tiled box copy global -> shared, completion signalled on `bar`
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---------------------------------------------------------------------------
// Stage the halo'd DEM tile.  TMA path: one elected thread issues a box copy with
// hardware zero fill.  Manual path (W*sizeof(T) not a multiple of 16, or TMA
// unavailable): bounds-checked cooperative loads.  Rows outside the init buffer are
// zero either way; whether a zero row is *legitimate* (outside the image) or a
// missing halo row is decided per tap through [r_lo, r_lo + r_span).
// ---------------------------------------------------------------------------
template <typename T, bool TMA, int TH, int NT = THREADS>
__device__ __forceinline__ void stage_tile_begin(T* tile, uint64_t* bar, const CUtensorMap* tmap, const T* init,
                                                 const Geom& g, int b, int ox, int oy_buf) {
    constexpr int SH = staged_rows(TH);
    if constexpr (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            mbar_arrive_expect_tx(bar, SH * SW * sizeof(T));
            tma_load_3d(tile, tmap, bar, ox, oy_buf, b);
        }
    } else {
        const T* src = init + (size_t)b * g.init_rows * g.W;
        for (int i = threadIdx.x; i < SH * SW; i += NT) {
            int r = i / SW, c = i - r * SW;
            int br = oy_buf + r, gx = ox + c;
            T v = from_f32<T>(0.f);
            if ((unsigned)br < (unsigned)g.init_rows && (unsigned)gx < (unsigned)g.W) v = src[(size_t)br * g.W + gx];
            tile[i] = v;
        }
    }
}
template <bool TMA>
__device__ __forceinline__ void stage_tile_wait(uint64_t* bar) {
    __syncthreads();  // manual path: tile stores visible; TMA path: barrier init visible to all waiters
    if constexpr (TMA) mbar_wait(bar, 0);
}

// Bounds-checked corner fetch from global memory (taps that leave the staged tile).
// (hi, wi) are GLOBAL coordinates.  A row inside the image but missing from the init
// buffer raises *status (row-strip calls with too small a halo).
template <typename T>
__device__ __forceinline__ float fetch_corner_global(const T* init_b, const Geom& g, int hi, int wi, int* status) {
    if ((unsigned)hi >= (unsigned)g.H_img || (unsigned)wi >= (unsigned)g.W) return 0.f;
    int br = hi - g.init_row0;
    if ((unsigned)br >= (unsigned)g.init_rows) {
        if (status) atomicOr(status, 1);
        return 0.f;
    }
    return to_f32(init_b[(size_t)br * g.W + wi]);
}

// pointer + byte stride as one opaque 64-bit add (keeps nvcc from re-deriving every channel
// address from the element index: 2 instructions per channel instead of 4)
template <typename T>
__device__ __forceinline__ T* step_ptr(T* p, size_t bytes) {
    unsigned long long r;
    asm("add.u64 %0, %1, %2;" : "=l"(r) : "l"((unsigned long long)p), "l"((unsigned long long)bytes));
    return reinterpret_cast<T*>(r);
}

// ---------------------------------------------------------------------------
// System-scope flags in peer memory (another GPU's HBM mapped through CUDA IPC, reached over NVLink).
// A producer makes its data stores visible with a system fence and then release-stores a generation stamp;
// a consumer spins on an acquire load (with back-off) until the stamp has been reached.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// A peer that never arrives (a rank died, or the ranks' call sequences diverged) must not hang the GPU: after
// PEER_WAIT_LIMIT_NS the kernel traps, which surfaces as a launch failure on the host (the fused counterpart of a
// collective's watchdog timeout).
constexpr unsigned long long PEER_WAIT_LIMIT_NS = 30ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void wait_stamp(const unsigned* flag, unsigned stamp) {
    if ((int)(ld_acquire_sys(flag) - stamp) >= 0) return;
    const unsigned long long t0 = global_ns();
    unsigned ns = 32;
    while ((int)(ld_acquire_sys(flag) - stamp) < 0) {
        __nanosleep(ns);
        if (ns < 1024) ns <<= 1;
        else if (global_ns() - t0 > PEER_WAIT_LIMIT_NS) __trap();
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace JSPSR_VARIANT
}  // namespace jspsr
