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

}  // namespace jspsr
