// Host-buffer entry point: the call made by a caller whose tensors live in host
// memory.  Batches are cut into chunks and streamed through caller-owned device
// scratch on two internal streams so that the H2D copy of chunk i+1, the kernel of
// chunk i and the D2H copy of chunk i-1 overlap (B200: separate copy engines per
// direction).  Returns when `out` is complete.
#include <cstdio>

#include "../../include/jspsr_spn.h"
#include "spn_kernels.cuh"

int jspsr_internal_fail(int code, const char* msg);  // abi.cu: sets the thread's last-error message

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static size_t slot_bytes(int chunk_B, int H, int W, size_t es) {
    const size_t px = (size_t)chunk_B * H * W;
    return align_up(px * es, 256) * 2 + align_up(px * 9 * es, 256) + align_up(px * 18 * es, 256);
}

extern "C" size_t jspsr_spn_host_scratch_bytes(int chunk_B, int H, int W, int dtype) {
    if (chunk_B <= 0 || H <= 0 || W <= 0) return 0;
    return 256 + 2 * slot_bytes(chunk_B, H, W, dtype == JSPSR_BF16 ? 2 : 4);
}

extern "C" int jspsr_spn_forward_host(const void* init, const void* weight, const void* offset, const float* w9,
                                      const float* b1, void* out, int B, int H, int W, int norm_mode, float scale,
                                      int dtype, void* dev_scratch, size_t scratch_bytes, int chunk_B) {
    if (!init || !weight || !offset || !w9 || !b1 || !out || !dev_scratch || chunk_B <= 0 || B <= 0 || H <= 0 || W <= 0)
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "forward_host: null pointer or non-positive dimension");
    if (dtype != JSPSR_F32 && dtype != JSPSR_BF16)
        return jspsr_internal_fail(JSPSR_ERR_UNSUPPORTED, "forward_host: dtype must be 0 (f32) or 1 (bf16)");
    if (scratch_bytes < jspsr_spn_host_scratch_bytes(chunk_B, H, W, dtype))
        return jspsr_internal_fail(JSPSR_ERR_BAD_ARG, "forward_host: dev_scratch is smaller than jspsr_spn_host_scratch_bytes()");
    const size_t es = dtype == JSPSR_BF16 ? 2 : 4;
    const size_t px = (size_t)H * W;
    char* base = (char*)dev_scratch;
    float* d_w9 = (float*)base;
    float* d_b1 = d_w9 + 9;
    const size_t sb = slot_bytes(chunk_B, H, W, es);
    cudaStream_t st[2];
    cudaEvent_t ready;
    int rc = JSPSR_OK;
    cudaError_t ce;
    if (cudaStreamCreateWithFlags(&st[0], cudaStreamNonBlocking) != cudaSuccess)
        return jspsr_internal_fail(JSPSR_ERR_CUDA, "forward_host: cudaStreamCreate failed");
    if (cudaStreamCreateWithFlags(&st[1], cudaStreamNonBlocking) != cudaSuccess) {
        cudaStreamDestroy(st[0]);
        return jspsr_internal_fail(JSPSR_ERR_CUDA, "forward_host: cudaStreamCreate failed");
    }
    cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
    cudaMemcpyAsync(d_w9, w9, 9 * sizeof(float), cudaMemcpyHostToDevice, st[0]);
    cudaMemcpyAsync(d_b1, b1, sizeof(float), cudaMemcpyHostToDevice, st[0]);
    cudaEventRecord(ready, st[0]);
    cudaStreamWaitEvent(st[1], ready, 0);
    int chunk = 0;
    for (int b0 = 0; b0 < B && rc == JSPSR_OK; b0 += chunk_B, ++chunk) {
        const int nb = (B - b0 < chunk_B) ? (B - b0) : chunk_B;
        cudaStream_t s = st[chunk & 1];
        char* slot = base + 256 + (size_t)(chunk & 1) * sb;
        const size_t cpx = (size_t)chunk_B * px;
        char* d_init = slot;
        char* d_out = d_init + align_up(cpx * es, 256);
        char* d_wgt = d_out + align_up(cpx * es, 256);
        char* d_off = d_wgt + align_up(cpx * 9 * es, 256);
        const size_t n = (size_t)nb * px;
        cudaMemcpyAsync(d_init, (const char*)init + (size_t)b0 * px * es, n * es, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_wgt, (const char*)weight + (size_t)b0 * px * 9 * es, n * 9 * es, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_off, (const char*)offset + (size_t)b0 * px * 18 * es, n * 18 * es, cudaMemcpyHostToDevice, s);
        rc = jspsr_spn_forward(d_init, d_wgt, d_off, d_w9, d_b1, d_out, nb, H, W, norm_mode, scale, dtype, s);
        cudaMemcpyAsync((char*)out + (size_t)b0 * px * es, d_out, n * es, cudaMemcpyDeviceToHost, s);
    }
    ce = cudaStreamSynchronize(st[0]);
    cudaError_t ce2 = cudaStreamSynchronize(st[1]);
    cudaEventDestroy(ready);
    cudaStreamDestroy(st[0]);
    cudaStreamDestroy(st[1]);
    if (rc != JSPSR_OK) return rc;
    if (ce != cudaSuccess || ce2 != cudaSuccess)
        return jspsr_internal_fail(JSPSR_ERR_CUDA, cudaGetErrorString(ce != cudaSuccess ? ce : ce2));
    return JSPSR_OK;
}
