// Debug probe: which TMA box-copy configurations work on this box.  ./tma_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NOTILE>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int bw, int bh, int c0, int c1, int c2, int rank) {
    extern __shared__ __align__(128) float tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bw * bh * 4) : "memory");
        if (rank == 3) {
            if (NOTILE)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(tile)), "l"((uint64_t)&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                             ::"r"(smem_u32(tile)), "l"((uint64_t)&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(tile)), "l"((uint64_t)&tmap), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
        }
    }
    __syncthreads();
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
                 ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    int v = argc > 1 ? atoi(argv[1]) : 0;
    int W = 128, H = 128, B = 2, rank = 3, bw = 32, bh = 8, c0 = 0, c1 = 0, c2 = 1, notile = 0;
    if (v >= 100) {  // ./tma_probe 100 c0 c1 bw bh
        c0 = atoi(argv[2]); c1 = atoi(argv[3]); bw = atoi(argv[4]); bh = atoi(argv[5]);
        if (argc > 6) c2 = atoi(argv[6]);
    }
    switch (v) {
        case 0: rank = 2; break;
        case 1: break;
        case 2: bw = 144; bh = 29; c0 = -6; c1 = -6; break;
        case 3: bw = 144; bh = 29; c0 = -6; c1 = -6; notile = 1; break;
        case 4: bw = 144; bh = 29; break;
        case 5: bw = 128; bh = 16; c0 = -6; c1 = -6; break;
        case 6: bw = 64; bh = 29; c0 = -6; c1 = -6; break;
        case 7: W = 16; H = 12; bw = 144; bh = 29; c0 = -6; c1 = -6; break;
        case 8: bw = 32; bh = 8; notile = 1; break;
    }
    std::vector<float> h((size_t)B * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003) * 0.25f;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, bw * bh * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * 4 * H};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d: encode=%d q=%d ", v, (int)r, (int)q);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    if (notile) probe<1><<<1, 128, bw * bh * 4>>>(map, o, bw, bh, c0, c1, c2, rank);
    else probe<0><<<1, 128, bw * bh * 4>>>(map, o, bw, bh, c0, c1, c2, rank);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> ho(bw * bh);
        cudaMemcpy(ho.data(), o, ho.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        int bsel = rank == 3 ? c2 : 0;
        for (int r_ = 0; r_ < bh; ++r_)
            for (int c = 0; c < bw; ++c) {
                int gy = c1 + r_, gx = c0 + c;
                float exp = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? h[((size_t)bsel * H + gy) * W + gx] : 0.f;
                if (ho[r_ * bw + c] != exp) ++bad;
            }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
