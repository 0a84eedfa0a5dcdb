// Small helper kernels around the propagation: offset range scan (strip halo
// sizing) and the `preserve_input` blend of NLSPN's loop (nlspn.py:228-229).
#include <mutex>
#include <unordered_map>

#include "spn_kernels.cuh"

namespace jspsr {

cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::unordered_map<const void*, unsigned long long> done;  // kernel -> bit mask of devices already set
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    std::lock_guard<std::mutex> lock(mu);
    unsigned long long& mask = done[kernel];
    if (mask & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) mask |= bit;
    return e;
}

template <typename T>
__global__ void __launch_bounds__(256) offset_absmax_kernel(const T* __restrict__ offset, size_t cs, int B,
                                                            float* __restrict__ out2) {
    // offset [B,18,cs]: even channels are row offsets, odd channels column offsets
    float mh = 0.f, mw = 0.f;
    const size_t total = (size_t)B * 18 * cs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const float v = fabsf(ld_stream(offset + i));
        const int ch = (int)((i / cs) % 18);
        if (ch & 1) mw = fmaxf(mw, v);
        else mh = fmaxf(mh, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mh = fmaxf(mh, __shfl_xor_sync(0xffffffffu, mh, o));
        mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // non-negative floats order like their bit patterns
        atomicMax(reinterpret_cast<int*>(out2), __float_as_int(mh));
        atomicMax(reinterpret_cast<int*>(out2) + 1, __float_as_int(mw));
    }
}

cudaError_t launch_offset_absmax(const void* offset, size_t, size_t cs, int B, bool bf16, float* out2,
                                 cudaStream_t stream) {
    const size_t total = (size_t)B * 18 * cs;
    int blocks = (int)min((size_t)148 * 8, (total + 255) / 256);
    if (bf16)
        offset_absmax_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)offset, cs, B, out2);
    else
        offset_absmax_kernel<float><<<blocks, 256, 0, stream>>>((const float*)offset, cs, B, out2);
    return cudaGetLastError();
}

// First generation of a row-strip sequence (include/jspsr_peer.h, jspsr_strip_halo_push): copy the band's own first /
// last `halo` rows into the neighbours' halo rows and raise their flags to `stamp`.  Thread 0 of every CTA first waits
// for this rank's flags >= stamp - 1: the neighbours have finished reading the buffer that is being overwritten.
template <typename T>
__global__ void __launch_bounds__(256) halo_push_kernel(const T* __restrict__ band, int Hs, int W, const StripPeerDev sp) {
    if (threadIdx.x == 0 && !(sp.debug & 4)) {
        if (sp.wait_up) wait_stamp(sp.wait_up, sp.stamp - 1u);
        if (sp.wait_dn) wait_stamp(sp.wait_dn, sp.stamp - 1u);
    }
    __syncthreads();
    const size_t n = (size_t)sp.halo * W;  // elements per direction
    T* up = static_cast<T*>(sp.up_dst);
    T* dn = static_cast<T*>(sp.dn_dst);
    const T* last = band + (size_t)(Hs - sp.halo) * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (up) up[i] = band[i];
        if (dn) dn[i] = last[i];
    }
    __syncthreads();  // every thread's peer stores are ordered before thread 0's system-scope fence (cumulativity)
    if (threadIdx.x == 0) __threadfence_system();
    if (threadIdx.x == 0 && atomicAdd(sp.tickets, 1u) == gridDim.x - 1) {
        sp.tickets[0] = 0u;
        __threadfence_system();
        if (sp.up_flag) st_release_sys(sp.up_flag, sp.stamp);
        if (sp.dn_flag) st_release_sys(sp.dn_flag, sp.stamp);
    }
}

cudaError_t launch_halo_push(const void* band, int Hs, int W, bool bf16, const StripPeerDev& sp, cudaStream_t stream) {
    const size_t n = (size_t)sp.halo * W;
    const int blocks = (int)min((size_t)148, (n + 255) / 256);
    if (bf16) halo_push_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)band, Hs, W, sp);
    else halo_push_kernel<float><<<blocks, 256, 0, stream>>>((const float*)band, Hs, W, sp);
    return cudaGetLastError();
}

template <typename T>
__global__ void __launch_bounds__(256) preserve_blend_kernel(const T* __restrict__ feat, const T* __restrict__ fix,
                                                             const float* __restrict__ mask, T* __restrict__ dst,
                                                             size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float m = mask[i];
        dst[i] = from_f32<T>((1.f - m) * to_f32(feat[i]) + m * to_f32(fix[i]));
    }
}

cudaError_t launch_preserve_blend(const void* feat, const void* feat_fix, const float* mask_fix, void* dst, size_t n,
                                  bool bf16, cudaStream_t stream) {
    int blocks = (int)min((size_t)148 * 8, (n + 255) / 256);
    if (bf16)
        preserve_blend_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(
            (const __nv_bfloat16*)feat, (const __nv_bfloat16*)feat_fix, mask_fix, (__nv_bfloat16*)dst, n);
    else
        preserve_blend_kernel<float><<<blocks, 256, 0, stream>>>((const float*)feat, (const float*)feat_fix, mask_fix,
                                                                 (float*)dst, n);
    return cudaGetLastError();
}

// The same blend with the mask derived in the kernel from a single-channel `fix` (mask = fix > 0): the six elementwise
// passes the LRRU cascade runs between its stages (LRRU.py:447-451 et seq.: sum(d_clear > 0, dim = 1) > 0, type_as,
// (1 - mask) * x + mask * d_clear) as one pass of 12 B/pixel.  Products and sum are rounded separately, as torch's
// three kernels round them, so non-finite values propagate identically ((1 - 1) * inf = NaN is kept).
template <typename T, int V>
__global__ void __launch_bounds__(256) preserve_blend_auto_kernel(const T* __restrict__ feat, const T* __restrict__ fix,
                                                                  T* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * V;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += stride) {
        T f[V], d[V], o[V];
        if (V == 4 && sizeof(T) == 4) {
            *reinterpret_cast<float4*>(f) = __ldcs(reinterpret_cast<const float4*>(feat + i));
            *reinterpret_cast<float4*>(d) = __ldcs(reinterpret_cast<const float4*>(fix + i));
        } else if (V == 4) {
            *reinterpret_cast<uint2*>(f) = __ldcs(reinterpret_cast<const uint2*>(feat + i));
            *reinterpret_cast<uint2*>(d) = __ldcs(reinterpret_cast<const uint2*>(fix + i));
        } else {
            f[0] = feat[i];
            d[0] = fix[i];
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const float dv = to_f32(d[j]);
            const float m = dv > 0.f ? 1.f : 0.f;
            o[j] = from_f32<T>(__fadd_rn(__fmul_rn(1.f - m, to_f32(f[j])), __fmul_rn(m, dv)));
        }
        if (V == 4 && sizeof(T) == 4) __stcs(reinterpret_cast<float4*>(dst + i), *reinterpret_cast<float4*>(o));
        else if (V == 4) __stcs(reinterpret_cast<uint2*>(dst + i), *reinterpret_cast<uint2*>(o));
        else dst[i] = o[0];
    }
}

cudaError_t launch_preserve_blend_auto(const void* feat, const void* fix, void* dst, size_t n, bool bf16,
                                       cudaStream_t stream) {
    const size_t align = bf16 ? 7 : 15;
    const bool vec = (n % 4 == 0) && !(((uintptr_t)feat | (uintptr_t)fix | (uintptr_t)dst) & align);
    const size_t items = vec ? n / 4 : n;
    const int blocks = (int)min((size_t)148 * 16, (items + 255) / 256);
    if (bf16) {
        if (vec) preserve_blend_auto_kernel<__nv_bfloat16, 4><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)feat, (const __nv_bfloat16*)fix, (__nv_bfloat16*)dst, n);
        else preserve_blend_auto_kernel<__nv_bfloat16, 1><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)feat, (const __nv_bfloat16*)fix, (__nv_bfloat16*)dst, n);
    } else {
        if (vec) preserve_blend_auto_kernel<float, 4><<<blocks, 256, 0, stream>>>((const float*)feat, (const float*)fix, (float*)dst, n);
        else preserve_blend_auto_kernel<float, 1><<<blocks, 256, 0, stream>>>((const float*)feat, (const float*)fix, (float*)dst, n);
    }
    return cudaGetLastError();
}

}  // namespace jspsr
