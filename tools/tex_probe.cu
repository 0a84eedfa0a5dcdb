// Probe: can the texture unit do the 9-tap deformable gather faster than the shared-memory path?
// One tex2Dgather per tap returns the tap's four corners exactly (no filtering arithmetic in the unit), the border
// address mode gives torchvision's "zero outside the image" at the left / right edge, and - for a batch of tiles stacked
// vertically in one pitch-linear 2-D texture - the rows above / below a tile are masked in registers.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tex_probe tools/tex_probe.cu && ./tools/tex_probe
// Prints Gpix/s for: gathers only (offsets from a hash), and the full forward (27 streamed fp32 / bf16 channels per pixel).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int H = 128, W = 128;

__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const __nv_bfloat16* p) {
    unsigned short v;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return __uint_as_float((unsigned)v << 16);
}
__device__ __forceinline__ void st_stream(float* p, float v) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// MODE 0: offsets from a hash, no streamed inputs.  MODE 1: the full forward (residual normalisation, w, b, + scale * centre)
template <typename T, int MODE, bool VCHECK>
__global__ void __launch_bounds__(256, 4)
tex_forward(cudaTextureObject_t tex, const T* __restrict__ weight, const T* __restrict__ offset, const float* __restrict__ w9,
            float* __restrict__ out, int B) {
    __shared__ float s_w[10];
    if (threadIdx.x < 10) s_w[threadIdx.x] = w9[threadIdx.x];
    __syncthreads();
    // CTA = 16 rows x 128 columns of one sample; warp = 32 consecutive x
    const int tiles_y = H / 16;
    const int b = blockIdx.x / tiles_y, y0 = (blockIdx.x % tiles_y) * 16;
    const size_t cs = (size_t)H * W;
    const float rowbase = (float)(b * H) + 1.0f;
    for (int it = 0; it < 8; ++it) {
        const int ry = (threadIdx.x >> 5) + 8 * (it >> 2), x = (threadIdx.x & 31) + 32 * (it & 3);
        const int y = y0 + ry;
        const size_t p = (size_t)y * W + x;
        float a[9], oh[9], ow[9];
        if (MODE == 1) {
            const T* pw = weight + (size_t)b * 9 * cs + p;
            const T* po = offset + (size_t)b * 18 * cs + p;
#pragma unroll
            for (int k = 0; k < 9; ++k) a[k] = ld_stream(pw + k * cs);
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                oh[k] = ld_stream(po + (2 * k) * cs);
                ow[k] = ld_stream(po + (2 * k + 1) * cs);
            }
            float s = a[0];
#pragma unroll
            for (int k = 1; k < 9; ++k) s += a[k];
            const float mean = __fdiv_rn(s, 9.f);
#pragma unroll
            for (int k = 0; k < 9; ++k) a[k] -= mean;
        } else {
            unsigned h = (unsigned)(p + (size_t)b * cs) * 2654435761u;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                h = h * 1664525u + 1013904223u;
                oh[k] = (float)(int)(h >> 16) * (1.0f / 65536.0f) * 6.0f - 3.0f;
                h = h * 1664525u + 1013904223u;
                ow[k] = (float)(int)(h >> 16) * (1.0f / 65536.0f) * 6.0f - 3.0f;
                a[k] = 0.1f;
            }
        }
        const float fy = (float)y, fx = (float)x;
        const float hk[3] = {fy - 1.f, fy, fy + 1.f};
        const float wk[3] = {fx - 1.f, fx, fx + 1.f};
        float acc = 0.f;
        float centre = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float hh = hk[k / 3] + oh[k], ww = wk[k % 3] + ow[k];
            const float hf = floorf(hh), wf = floorf(ww);
            const float lh = hh - hf, lw = ww - wf;
            // gather at the centre of the 2x2 footprint: (w0 + 1, h0 + 1) in unnormalised texel coordinates
            const float4 g = tex2Dgather<float4>(tex, wf + 1.0f, hf + rowbase, 0);
            // gather order: (x0,y1) (x1,y1) (x1,y0) (x0,y0)
            float top = fmaf(lw, g.z - g.w, g.w);
            float bot = fmaf(lw, g.y - g.x, g.x);
            if (VCHECK) {  // rows of the neighbouring samples in the stacked texture are not this sample's zeros
                const float t = hf - 0.5f * (float)(H - 1);
                top = fabsf(t) <= 0.5f * (float)(H - 1) ? top : 0.f;
                bot = fabsf(t + 1.0f) <= 0.5f * (float)(H - 1) ? bot : 0.f;
            }
            const float val = fmaf(lh, bot - top, top);
            acc += (s_w[k] * a[k]) * val;
            if (k == 4) centre = g.w;
        }
        acc += s_w[9];
        if (MODE == 1) {
            const float c = tex2D<float>(tex, fx + 0.5f, fy + (rowbase - 0.5f));
            acc += c;
        } else {
            acc += centre * 1e-9f;
        }
        st_stream(out + (size_t)b * cs + p, acc);
    }
}

template <typename T>
static void fill_random(std::vector<T>& v, float sigma, float clip, unsigned seed);
template <>
void fill_random<float>(std::vector<float>& v, float sigma, float clip, unsigned seed) {
    srand(seed);
    for (auto& x : v) {
        float u1 = (rand() + 1.0f) / (RAND_MAX + 2.0f), u2 = (rand() + 1.0f) / (RAND_MAX + 2.0f);
        float n = sqrtf(-2.f * logf(u1)) * cosf(6.2831853f * u2) * sigma;
        x = fminf(fmaxf(n, -clip), clip);
    }
}

template <typename T, int MODE, bool VCHECK>
static float run(cudaTextureObject_t tex, const T* w, const T* o, const float* w9, float* out, int B, int reps) {
    dim3 grid(B * (H / 16));
    for (int i = 0; i < 3; ++i) tex_forward<T, MODE, VCHECK><<<grid, 256>>>(tex, w, o, w9, out, B);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) tex_forward<T, MODE, VCHECK><<<grid, 256>>>(tex, w, o, w9, out, B);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const int B = 507;  // 507 * 128 = 64896 rows <= 65000 (maxTexture2DLinear height)
    const size_t npix = (size_t)B * H * W;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s: maxTexture2DLinear %d x %d pitch %d, textureAlignment %zu, texturePitchAlignment %zu\n", prop.name,
           prop.maxTexture2DLinear[0], prop.maxTexture2DLinear[1], prop.maxTexture2DLinear[2], prop.textureAlignment,
           prop.texturePitchAlignment);
    std::vector<float> h_dem(npix), h_w(npix * 9), h_o(npix * 18);
    srand(1);
    for (auto& x : h_dem) x = rand() / (float)RAND_MAX;
    for (auto& x : h_w) x = 1.f / (1.f + expf(-1.5f * (rand() / (float)RAND_MAX * 4.f - 2.f)));
    fill_random<float>(h_o, 1.5f, 8.f, 7);
    float *d_dem, *d_w, *d_o, *d_out, *d_w9;
    CK(cudaMalloc(&d_dem, npix * 4));
    CK(cudaMalloc(&d_w, npix * 9 * 4));
    CK(cudaMalloc(&d_o, npix * 18 * 4));
    CK(cudaMalloc(&d_out, npix * 4));
    CK(cudaMalloc(&d_w9, 40));
    CK(cudaMemcpy(d_dem, h_dem.data(), npix * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, h_w.data(), npix * 9 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_o, h_o.data(), npix * 18 * 4, cudaMemcpyHostToDevice));
    float w9[10] = {1.05f, 0.95f, 1.f, 1.02f, 0.98f, 1.f, 1.01f, 0.99f, 1.f, 0.1f};
    CK(cudaMemcpy(d_w9, w9, 40, cudaMemcpyHostToDevice));
    // bf16 copies of the streamed tensors
    std::vector<__nv_bfloat16> hb_w(npix * 9), hb_o(npix * 18);
    for (size_t i = 0; i < hb_w.size(); ++i) hb_w[i] = __float2bfloat16(h_w[i]);
    for (size_t i = 0; i < hb_o.size(); ++i) hb_o[i] = __float2bfloat16(h_o[i]);
    __nv_bfloat16 *db_w, *db_o;
    CK(cudaMalloc(&db_w, npix * 9 * 2));
    CK(cudaMalloc(&db_o, npix * 18 * 2));
    CK(cudaMemcpy(db_w, hb_w.data(), npix * 9 * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db_o, hb_o.data(), npix * 18 * 2, cudaMemcpyHostToDevice));

    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = d_dem;
    rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
    rd.res.pitch2D.width = W;
    rd.res.pitch2D.height = (size_t)B * H;
    rd.res.pitch2D.pitchInBytes = W * 4;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    cudaTextureObject_t tex = 0;
    cudaEvent_t c0, c1;
    cudaEventCreate(&c0); cudaEventCreate(&c1);
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    {   // host cost of creating + destroying a texture object
        timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (int i = 0; i < 100; ++i) {
            cudaTextureObject_t t2 = 0;
            CK(cudaCreateTextureObject(&t2, &rd, &td, nullptr));
            CK(cudaDestroyTextureObject(t2));
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        printf("cudaCreateTextureObject + Destroy: %.2f us per pair\n", ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / 100 / 1e3);
    }

    float ms;
    ms = run<float, 0, false>(tex, d_w, d_o, d_w9, d_out, B, 20);
    printf("gathers only (hash offsets), no row mask : %.4f ms  %.1f Gpix/s  %.1f G gathers/s\n", ms, npix / ms / 1e6, 9 * npix / ms / 1e6);
    ms = run<float, 0, true>(tex, d_w, d_o, d_w9, d_out, B, 20);
    printf("gathers only (hash offsets), row mask    : %.4f ms  %.1f Gpix/s\n", ms, npix / ms / 1e6);
    ms = run<float, 1, true>(tex, d_w, d_o, d_w9, d_out, B, 20);
    printf("full forward fp32 (116 B/pixel), row mask: %.4f ms  %.1f Gpix/s  %.0f GB/s\n", ms, npix / ms / 1e6, 116.0 * npix / ms / 1e6);
    ms = run<float, 1, false>(tex, d_w, d_o, d_w9, d_out, B, 20);
    printf("full forward fp32, no row mask           : %.4f ms  %.1f Gpix/s  %.0f GB/s\n", ms, npix / ms / 1e6, 116.0 * npix / ms / 1e6);
    ms = run<__nv_bfloat16, 1, true>(tex, db_w, db_o, d_w9, d_out, B, 20);
    printf("full forward bf16 weight/offset (62 B/px), row mask: %.4f ms  %.1f Gpix/s  %.0f GB/s\n", ms, npix / ms / 1e6, 62.0 * npix / ms / 1e6);
    ms = run<__nv_bfloat16, 1, false>(tex, db_w, db_o, d_w9, d_out, B, 20);
    printf("full forward bf16 weight/offset, no row mask       : %.4f ms  %.1f Gpix/s  %.0f GB/s\n", ms, npix / ms / 1e6, 62.0 * npix / ms / 1e6);

    // correctness of the gather convention on a few pixels (CPU restatement of torchvision's rule)
    std::vector<float> h_out(npix);
    run<float, 1, true>(tex, d_w, d_o, d_w9, d_out, B, 1);
    CK(cudaMemcpy(h_out.data(), d_out, npix * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int t = 0; t < 4000; ++t) {
        const int b = (t * 37) % B, y = (t * 13 + (t % 5 == 0 ? 0 : 7)) % H, x = (t * 29 + (t % 7 == 0 ? 127 : 3)) % W;
        const size_t cs = (size_t)H * W, p = (size_t)y * W + x;
        float a[9], s = 0;
        for (int k = 0; k < 9; ++k) { a[k] = h_w[(size_t)b * 9 * cs + k * cs + p]; }
        s = a[0];
        for (int k = 1; k < 9; ++k) s += a[k];
        const float mean = s / 9.f;
        float acc = 0.f;
        for (int k = 0; k < 9; ++k) {
            const float hh = (float)(y - 1 + k / 3) + h_o[(size_t)b * 18 * cs + (2 * k) * cs + p];
            const float ww = (float)(x - 1 + k % 3) + h_o[(size_t)b * 18 * cs + (2 * k + 1) * cs + p];
            const int h0 = (int)floorf(hh), w0 = (int)floorf(ww);
            const float lh = hh - floorf(hh), lw = ww - floorf(ww);
            auto at = [&](int r, int c) { return (r >= 0 && r < H && c >= 0 && c < W) ? h_dem[(size_t)b * cs + (size_t)r * W + c] : 0.f; };
            const float v1 = at(h0, w0), v2 = at(h0, w0 + 1), v3 = at(h0 + 1, w0), v4 = at(h0 + 1, w0 + 1);
            const float top = fmaf(lw, v2 - v1, v1), bot = fmaf(lw, v4 - v3, v3);
            acc += (w9[k] * (a[k] - mean)) * fmaf(lh, bot - top, top);
        }
        acc += w9[9];
        acc += h_dem[(size_t)b * cs + p];
        maxerr = fmax(maxerr, fabs((double)acc - h_out[(size_t)b * cs + p]));
    }
    printf("max |gpu - cpu| over 4000 pixels (edges included): %.3e\n", maxerr);
    return 0;
}
