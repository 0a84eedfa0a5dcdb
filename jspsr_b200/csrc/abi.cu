// C ABI of libjspsr_spn.so (see include/jspsr_spn.h): argument validation, TMA
// descriptor encoding, launches.  No torch types, no allocation, no synchronisation.
#include <initializer_list>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/jspsr_spn.h"
#include "../../include/jspsr_peer.h"
#include "spn_types.cuh"

using namespace jspsr;

namespace jspsr {
cudaError_t launch_offset_absmax(const void* offset, size_t n_pairs_block, size_t cs, int B, bool bf16, float* out2,
                                 cudaStream_t stream);
cudaError_t launch_halo_push(const void* band, int Hs, int W, bool bf16, const StripPeerDev& sp, cudaStream_t stream);
cudaError_t launch_preserve_blend(const void* feat, const void* feat_fix, const float* mask_fix, void* dst, size_t n,
                                  bool bf16, cudaStream_t stream);
cudaError_t launch_preserve_blend_auto(const void* feat, const void* fix, void* dst, size_t n, bool bf16,
                                       cudaStream_t stream);
// The tile kernels exist twice (spn_common.cuh): `wide` stages a 14/15-row, 16-column halo, `narrow` 6/7 rows and
// 8 columns.  Wide wins on tile batches, narrow on whole rasters; the launcher picks by image width.
#define JSPSR_DECLARE_VARIANT(NS)                                                                                       \
    namespace NS {                                                                                                      \
    cudaError_t launch_spn_forward(const LaunchArgs& la);                                                               \
    cudaError_t launch_spn_backward(const LaunchArgs& la);                                                              \
    int stage_box_cols();                                                                                               \
    int stage_box_rows(int th);                                                                                         \
    cudaError_t launch_nlspn_affinity(const void* conv_out, const void* confidence, const float* gamma,                 \
                                      void* offset_out, void* aff_out, const void* grad_offset, const void* grad_aff,   \
                                      void* grad_conv_out, float* grad_confidence, float* grad_scale, void* workspace,  \
                                      const Geom& g, int tile_h, bool use_tma, const CUtensorMap& tmap, int affinity,   \
                                      int legacy, bool bf16, bool forward, cudaStream_t stream);                        \
    cudaError_t launch_gen_spn_forward(const LaunchArgs& la, const CUtensorMap& tmap_feat, const void* feature, int C,  \
                                       const float* conv_w, const float* conv_b, void* weight_out, void* offset_out);   \
    cudaError_t launch_iter_carry(const float* g_a, const float* g_b, const float* aff, const float* offset,           \
                                  const float* asum_in, float* asum_out, float* carry_out, const Geom& g,               \
                                  cudaStream_t stream);                                                                 \
    cudaError_t launch_iter_grad(const float* grad_list, const float* carry, const float* feat_init,                    \
                                 const float* list_out, const float* aff, const float* offset, float* grad_aff,         \
                                 float* grad_offset, const Geom& g, int T, int tile_h, bool use_tma,                    \
                                 const CUtensorMap& tmap_init, const CUtensorMap& tmap_list, cudaStream_t stream);      \
    int iter_grad_tile_h(int T);                                                                                        \
    }
JSPSR_DECLARE_VARIANT(narrow)
JSPSR_DECLARE_VARIANT(wide)
#undef JSPSR_DECLARE_VARIANT
namespace narrow {  // gen_tail_backward.cu / gen_tail_wgrad.cu stage no DEM tile: compiled once
cudaError_t launch_gen_grad_feature(const LaunchArgs& la, const CUtensorMap& tmap_gz, const void* gz, int C,
                                    const float* conv_w, void* grad_feature);
cudaError_t launch_gen_grad_weight(const void* gz, const void* feature, int C, bool bf16, float* grad_w, float* grad_b, void* ws,
                                   int B, int HW, int run_len, bool use_tma, const CUtensorMap& tmap_f,
                                   const CUtensorMap& tmap_gz, cudaStream_t stream);
size_t gen_grad_weight_workspace_bytes();
cudaError_t launch_spn_iterate_fused(const float* feat_init, const float* aff, const float* offset, float* list_out, int B,
                                     int H, int W, int T, cudaStream_t stream);
}
}  // namespace jspsr

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
// used by host_pipeline.cu so that every entry point reports through jspsr_last_error()
int jspsr_internal_fail(int code, const char* msg) { return fail(code, "%s", msg); }

static int cuda_fail(cudaError_t e, const char* what) {
    return fail(JSPSR_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------
// TMA descriptor for the DEM buffer viewed as [B][rows][W]; box = the staged tile
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// wide halo for tile-like inputs, narrow for rasters (JSPSR_SPN_HALO=narrow|wide overrides)
static bool choose_wide(int W) {
    if (const char* e = getenv("JSPSR_SPN_HALO")) {
        if (e[0] == 'w') return true;
        if (e[0] == 'n') return false;
    }
    return W <= 1024;
}

// bf16 weight / offset streamed two pixels per thread as 32-bit words (spn_forward.cu, PAIR): every streamed pointer
// must be aligned to a pixel pair; JSPSR_SPN_PAIR=0 selects the one-pixel kernels (A/B runs, tests of both)
static bool pair_ok(int dtype, std::initializer_list<const void*> bf16_ptrs, std::initializer_list<const void*> io_ptrs) {
    if (dtype == JSPSR_F32) return false;
    if (const char* e = getenv("JSPSR_SPN_PAIR")) {
        if (e[0] == '0') return false;
    }
    const uintptr_t io_mask = dtype == JSPSR_BF16 ? 3 : 7;
    for (const void* p : bf16_ptrs) if (((uintptr_t)p & 3) != 0) return false;
    for (const void* p : io_ptrs) if (((uintptr_t)p & io_mask) != 0) return false;
    return true;
}

static bool tma_disabled_by_env() {
    // read on every call so tests can flip it; the manual tile loader is the same
    // kernel with the TMA box copy replaced by bounds-checked loads
    const char* e = getenv("JSPSR_SPN_DISABLE_TMA");
    return e && e[0] == '1';
}

// Returns true when the TMA path can be used for this buffer (and fills *map).
static bool make_init_tmap(CUtensorMap* map, const void* init, int B, int rows, int W, bool bf16, int tile_h,
                           bool wide) {
    if (tma_disabled_by_env()) return false;
    const size_t es = bf16 ? 2 : 4;
    if (((uintptr_t)init & 15) != 0) return false;
    if (((size_t)W * es) % 16 != 0) return false;  // TMA global strides are multiples of 16 bytes
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)W * es, (cuuint64_t)W * es * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)(wide ? wide::stage_box_cols() : narrow::stage_box_cols()),
                         (cuuint32_t)(wide ? wide::stage_box_rows(tile_h) : narrow::stage_box_rows(tile_h)), 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                     const_cast<void*>(init), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// TMA descriptor for the feature tensor viewed as [B*C planes][H][W]; box = one 128-pixel row segment of C planes
static bool make_feature_tmap(CUtensorMap* map, const void* feature, int B, int C, int H, int W, bool bf16) {
    if (tma_disabled_by_env()) return false;
    const size_t es = bf16 ? 2 : 4;
    if (((uintptr_t)feature & 15) != 0 || ((size_t)W * es) % 16 != 0 || C > 256) return false;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * C};
    cuuint64_t strides[2] = {(cuuint64_t)W * es, (cuuint64_t)W * es * (cuuint64_t)H};
    cuuint32_t box[3] = {(cuuint32_t)TILE_W, 1u, (cuuint32_t)C};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(feature), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// TMA descriptor for a tensor viewed as [planes][H*W] (K-major operands of the weight-gradient contraction);
// box = 128 bytes of consecutive pixels of `box_planes` planes
static bool make_plane_tmap(CUtensorMap* map, const void* base, size_t planes, size_t HW, int box_planes, bool swizzle128,
                            bool bf16) {
    if (tma_disabled_by_env()) return false;
    const size_t es = bf16 ? 2 : 4;
    if (((uintptr_t)base & 15) != 0 || (HW * es) % 16 != 0 || box_planes > 256 || planes > 0xffffffffull) return false;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)HW, (cuuint64_t)planes};
    cuuint64_t strides[1] = {(cuuint64_t)HW * es};
    cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_planes};
    cuuint32_t estr[2] = {1u, 1u};
    return enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int check_common(int B, int H, int W, int norm_mode, int dtype) {
    if (B <= 0 || H <= 0 || W <= 0) return fail(JSPSR_ERR_BAD_ARG, "non-positive dimension B=%d H=%d W=%d", B, H, W);
    if (H > (1 << 22) - 256 || W > (1 << 22) - 256)
        return fail(JSPSR_ERR_UNSUPPORTED, "H and W are limited to 2^22 - 256 (the kernels floor pixel coordinates in the fp32 mantissa)");
    if (norm_mode < 0 || norm_mode > 2) return fail(JSPSR_ERR_BAD_ARG, "norm_mode %d is not 0/1/2", norm_mode);
    if (dtype != JSPSR_F32 && dtype != JSPSR_BF16 && dtype != JSPSR_MIXED)
        return fail(JSPSR_ERR_BAD_ARG, "dtype %d is not 0 (f32) / 1 (bf16) / 2 (mixed)", dtype);
    return 0;
}
static int check_align(const void* p, size_t a, const char* name) {
    if (p && ((uintptr_t)p % a) != 0) return fail(JSPSR_ERR_ALIGN, "%s is not aligned to %zu bytes", name, a);
    return 0;
}
// Rows per CTA: the largest of max_th/../2 that still gives the grid two CTAs per SM (small batches
// are latency bound: more, shorter CTAs; large ones want the least halo overhead).  The forward is
// fastest at 16 rows, the backward at 8 (measured, see spn_kernels.cuh).  JSPSR_SPN_TILE_H overrides.
static int choose_tile_h(int B, int Hs, int W, int max_th) {
    if (const char* e = getenv("JSPSR_SPN_TILE_H")) {
        const int v = atoi(e);
        if ((v == 16 || v == 8 || v == 4 || v == 2) && v <= max_th) return v;
    }
    const size_t tx = (size_t)(W + TILE_W - 1) / TILE_W;
    for (int th = max_th; th > 2; th >>= 1) {
        const size_t tiles = tx * (size_t)((Hs + th - 1) / th) * (size_t)B;
        if (tiles >= (size_t)148 * 2) return th;
    }
    return 2;
}

static int fill_geom(LaunchArgs* la, int B, int Hs, int W, int H_img, int row0, int init_row0, int init_rows,
                     int max_th) {
    Geom* g = &la->g;
    la->tile_h = choose_tile_h(B, Hs, W, max_th);
    g->B = B; g->H = Hs; g->W = W; g->H_img = H_img; g->row0 = row0; g->init_row0 = init_row0; g->init_rows = init_rows;
    g->tiles_x = (W + TILE_W - 1) / TILE_W;
    g->tiles_y = (Hs + la->tile_h - 1) / la->tile_h;
    const size_t tiles = (size_t)g->tiles_x * g->tiles_y * B;
    if (tiles > 0x7fffffffull) return fail(JSPSR_ERR_UNSUPPORTED, "too many tiles (%zu) for one launch", tiles);
    return 0;
}

extern "C" {

int jspsr_version(void) { return JSPSR_SPN_VERSION; }
const char* jspsr_last_error(void) { return g_err; }
size_t jspsr_spn_workspace_bytes(void) { return sizeof(ReduceWs); }

// host struct (include/jspsr_peer.h) -> the kernel's view; validates the neighbour topology
static int strip_peer_dev(const jspsr_strip_peer* sp, int Hs, StripPeerDev* d) {
    if (!sp->my_flags) return fail(JSPSR_ERR_BAD_ARG, "jspsr_strip_peer: my_flags is null");
    if (sp->halo <= 0 || sp->halo > Hs)
        return fail(JSPSR_ERR_BAD_ARG, "jspsr_strip_peer: halo %d must be in [1, Hs = %d] (use fewer ranks)", sp->halo, Hs);
    if ((sp->up_dst && !sp->up_flags) || (sp->dn_dst && !sp->dn_flags))
        return fail(JSPSR_ERR_BAD_ARG, "jspsr_strip_peer: a destination without its neighbour's flag block");
    d->up_dst = sp->up_dst; d->dn_dst = sp->dn_dst;
    d->up_flag = sp->up_flags ? sp->up_flags + 1 : nullptr;   // I am the upper neighbour's LOWER neighbour
    d->dn_flag = sp->dn_flags ? sp->dn_flags + 0 : nullptr;   // and the lower neighbour's UPPER one
    d->wait_up = sp->up_flags ? sp->my_flags + 0 : nullptr;
    d->wait_dn = sp->dn_flags ? sp->my_flags + 1 : nullptr;
    d->tickets = sp->my_flags + 2;
    d->stamp = sp->stamp; d->halo = sp->halo;
    if (const char* e = getenv("JSPSR_STRIP_PEER_DEBUG")) d->debug = atoi(e);
    return 0;
}

static int spn_forward_strip_impl(const void* init, const void* weight, const void* offset, const float* w9, const float* b1,
                            void* out, int B, int Hs, int W, int H_img, int row0, int init_row0, int init_rows,
                            int norm_mode, float scale, int dtype, int* status, const jspsr_strip_peer* sp, void* stream) {
    if (int e = check_common(B, Hs, W, norm_mode, dtype)) return e;
    if (!init || !weight || !offset || !out) return fail(JSPSR_ERR_BAD_ARG, "null tensor pointer");
    if (H_img < Hs || row0 < 0 || row0 + Hs > H_img || init_row0 < 0 || init_rows <= 0 || init_row0 + init_rows > H_img ||
        H_img > (1 << 22) - 256)
        return fail(JSPSR_ERR_BAD_ARG, "inconsistent strip geometry (Hs=%d H_img=%d row0=%d init_row0=%d init_rows=%d)", Hs,
                    H_img, row0, init_row0, init_rows);
    const size_t es = dtype == JSPSR_F32 ? 4 : 2;      // weight / offset
    const size_t esi = dtype == JSPSR_BF16 ? 2 : 4;    // init / out
    if (int e = check_align(init, esi, "init")) return e;
    if (int e = check_align(weight, es, "weight")) return e;
    if (int e = check_align(offset, es, "offset")) return e;
    if (int e = check_align(out, esi, "out")) return e;
    if (int e = check_align(w9, 4, "w9")) return e;
    if (int e = check_align(b1, 4, "b1")) return e;
    LaunchArgs la;
    if (int e = fill_geom(&la, B, Hs, W, H_img, row0, init_row0, init_rows, 16)) return e;
    StripPeerDev spd;
    if (sp != nullptr) {  // fused halo exchange: rasters (B = 1), instantiated for 16 rows per CTA
        if (B != 1 || dtype == JSPSR_MIXED)
            return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_spn_forward_strip_peer: B = 1 and dtype f32 / bf16 only");
        if (int e = strip_peer_dev(sp, Hs, &spd)) return e;
        la.tile_h = 16;
        la.g.tiles_y = (Hs + 15) / 16;
        la.strip_peer = &spd;
    }
    la.init = init; la.weight = weight; la.offset = offset; la.w9 = w9; la.b1 = b1; la.out = out;
    la.mode = norm_mode; la.scale = scale; la.bf16 = dtype != JSPSR_F32; la.init_f32 = dtype == JSPSR_MIXED;
    la.status = status;
    la.stream = (cudaStream_t)stream;
    la.pair = pair_ok(dtype, {weight, offset}, {out});
    const bool wide = choose_wide(W);
    la.use_tma = make_init_tmap(&la.tmap, init, B, init_rows, W, dtype == JSPSR_BF16, la.tile_h, wide);
    cudaError_t ce = wide ? wide::launch_spn_forward(la) : narrow::launch_spn_forward(la);
    if (ce != cudaSuccess) return cuda_fail(ce, "spn_forward launch");
    return JSPSR_OK;
}

int jspsr_spn_forward_strip(const void* init, const void* weight, const void* offset, const float* w9, const float* b1,
                            void* out, int B, int Hs, int W, int H_img, int row0, int init_row0, int init_rows,
                            int norm_mode, float scale, int dtype, int* status, void* stream) {
    return spn_forward_strip_impl(init, weight, offset, w9, b1, out, B, Hs, W, H_img, row0, init_row0, init_rows, norm_mode,
                                  scale, dtype, status, nullptr, stream);
}

int jspsr_spn_forward_strip_peer(const void* init, const void* weight, const void* offset, const float* w9, const float* b1,
                                 void* out, int Hs, int W, int H_img, int row0, int init_row0, int init_rows, int norm_mode,
                                 float scale, int dtype, int* status, const jspsr_strip_peer* sp, void* stream) {
    if (!sp) return fail(JSPSR_ERR_BAD_ARG, "jspsr_spn_forward_strip_peer: sp is null (use jspsr_spn_forward_strip)");
    return spn_forward_strip_impl(init, weight, offset, w9, b1, out, 1, Hs, W, H_img, row0, init_row0, init_rows, norm_mode,
                                  scale, dtype, status, sp, stream);
}

int jspsr_strip_halo_push(const void* band, int Hs, int W, int dtype, const jspsr_strip_peer* sp, void* stream) {
    if (!band || !sp) return fail(JSPSR_ERR_BAD_ARG, "null pointer");
    if (Hs <= 0 || W <= 0) return fail(JSPSR_ERR_BAD_ARG, "non-positive dimension Hs=%d W=%d", Hs, W);
    if (dtype != JSPSR_F32 && dtype != JSPSR_BF16) return fail(JSPSR_ERR_BAD_ARG, "dtype %d is not 0 (f32) / 1 (bf16)", dtype);
    StripPeerDev spd;
    if (int e = strip_peer_dev(sp, Hs, &spd)) return e;
    if ((sp->up_flags && !sp->up_dst) || (sp->dn_flags && !sp->dn_dst))
        return fail(JSPSR_ERR_BAD_ARG, "jspsr_strip_halo_push needs a destination for every neighbour");
    if (!sp->up_flags && !sp->dn_flags) return JSPSR_OK;  // a single strip has nobody to push to
    cudaError_t ce = launch_halo_push(band, Hs, W, dtype == JSPSR_BF16, spd, (cudaStream_t)stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "halo push launch");
    return JSPSR_OK;
}

// ---- peer-mapped device memory (CUDA IPC) ----
static_assert(sizeof(cudaIpcMemHandle_t) == JSPSR_PEER_HANDLE_BYTES, "handle size");

int jspsr_peer_alloc(size_t bytes, void** dev_ptr, void* handle_out) {
    if (!dev_ptr || !handle_out || bytes == 0) return fail(JSPSR_ERR_BAD_ARG, "jspsr_peer_alloc: bad argument");
    void* p = nullptr;
    cudaError_t ce = cudaMalloc(&p, bytes);
    if (ce != cudaSuccess) return cuda_fail(ce, "jspsr_peer_alloc: cudaMalloc");
    ce = cudaMemset(p, 0, bytes);
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (ce == cudaSuccess) ce = cudaIpcGetMemHandle(&h, p);
    if (ce != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(ce, "jspsr_peer_alloc: memset / cudaIpcGetMemHandle");
    }
    memcpy(handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return JSPSR_OK;
}

int jspsr_peer_open(const void* handle, void** dev_ptr) {
    if (!handle || !dev_ptr) return fail(JSPSR_ERR_BAD_ARG, "jspsr_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t ce = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (ce != cudaSuccess) return cuda_fail(ce, "jspsr_peer_open: cudaIpcOpenMemHandle (is peer access between the GPUs available?)");
    *dev_ptr = p;
    return JSPSR_OK;
}

int jspsr_peer_close(void* dev_ptr) {
    if (!dev_ptr) return JSPSR_OK;
    cudaError_t ce = cudaIpcCloseMemHandle(dev_ptr);
    if (ce != cudaSuccess) return cuda_fail(ce, "jspsr_peer_close");
    return JSPSR_OK;
}

int jspsr_peer_free(void* dev_ptr) {
    if (!dev_ptr) return JSPSR_OK;
    cudaError_t ce = cudaFree(dev_ptr);
    if (ce != cudaSuccess) return cuda_fail(ce, "jspsr_peer_free");
    return JSPSR_OK;
}

int jspsr_spn_forward(const void* init, const void* weight, const void* offset, const float* w9, const float* b1,
                      void* out, int B, int H, int W, int norm_mode, float scale, int dtype, void* stream) {
    if (!w9 || !b1) return fail(JSPSR_ERR_BAD_ARG, "w9/b1 must be device pointers (use jspsr_spn_iterate for w=1,b=0)");
    return jspsr_spn_forward_strip(init, weight, offset, w9, b1, out, B, H, W, H, 0, 0, H, norm_mode, scale, dtype,
                                   nullptr, stream);
}

int jspsr_spn_backward(const void* grad_out, const void* init, const void* weight, const void* offset, const float* w9,
                       float* grad_init, void* grad_weight, void* grad_offset, float* grad_w9, float* grad_b1,
                       void* workspace, int B, int H, int W, int norm_mode, float scale, int dtype, unsigned flags,
                       void* stream) {
    return jspsr_spn_backward_reduce(grad_out, init, weight, offset, w9, grad_init, grad_weight, grad_offset, grad_w9, grad_b1,
                                     workspace, B, H, W, norm_mode, scale, dtype, flags, nullptr, stream);
}

int jspsr_spn_backward_reduce(const void* grad_out, const void* init, const void* weight, const void* offset,
                              const float* w9, float* grad_init, void* grad_weight, void* grad_offset, float* grad_w9,
                              float* grad_b1, void* workspace, int B, int H, int W, int norm_mode, float scale, int dtype,
                              unsigned flags, const jspsr_peer_reduce* pr, void* stream) {
    if (int e = check_common(B, H, W, norm_mode, dtype)) return e;
    PeerReduceDev prd;
    if (pr != nullptr && pr->world > 1) {
        if (pr->world > JSPSR_PEER_MAX_RANKS || pr->rank < 0 || pr->rank >= pr->world)
            return fail(JSPSR_ERR_BAD_ARG, "jspsr_peer_reduce: rank %d / world %d (at most %d ranks)", pr->rank, pr->world,
                        JSPSR_PEER_MAX_RANKS);
        if (!grad_w9) return fail(JSPSR_ERR_BAD_ARG, "jspsr_spn_backward_reduce: nothing to reduce without grad_w9");
        for (int r = 0; r < pr->world; ++r) {
            if (!pr->slots[r] || ((uintptr_t)pr->slots[r] & 15) != 0)
                return fail(JSPSR_ERR_BAD_ARG, "jspsr_peer_reduce: slots[%d] is null or unaligned", r);
            prd.slots[r] = static_cast<double*>(pr->slots[r]);
        }
        prd.rank = pr->rank; prd.world = pr->world;
        prd.mul = pr->average ? 1.f / (float)pr->world : 1.f;
    }
    const bool preact = (flags & JSPSR_BWD_GEN_PREACT) != 0;
    if (!grad_out || !init || !weight || !offset || !grad_weight || (!grad_offset && !preact))
        return fail(JSPSR_ERR_BAD_ARG, "null tensor pointer");
    if (preact && (grad_init || (flags & JSPSR_BWD_ACCUMULATE)))
        return fail(JSPSR_ERR_UNSUPPORTED, "JSPSR_BWD_GEN_PREACT excludes grad_init and JSPSR_BWD_ACCUMULATE");
    if (grad_w9 && !workspace) return fail(JSPSR_ERR_BAD_ARG, "workspace is required when grad_w9 is requested");
    if (grad_b1 && !grad_w9) return fail(JSPSR_ERR_BAD_ARG, "grad_b1 without grad_w9 is not supported");
    if ((flags & JSPSR_BWD_ACCUMULATE) && !grad_init)
        return fail(JSPSR_ERR_UNSUPPORTED, "JSPSR_BWD_ACCUMULATE is implemented for the fixed-affinity loop only "
                                           "(grad_init required)");
    const size_t es = dtype == JSPSR_F32 ? 4 : 2;      // weight / offset and their gradients
    const size_t esi = dtype == JSPSR_BF16 ? 2 : 4;    // grad_out / init
    if (int e = check_align(grad_out, esi, "grad_out")) return e;
    if (int e = check_align(init, esi, "init")) return e;
    if (int e = check_align(weight, es, "weight")) return e;
    if (int e = check_align(offset, es, "offset")) return e;
    if (int e = check_align(grad_weight, es, "grad_weight")) return e;
    if (int e = check_align(grad_offset, es, "grad_offset")) return e;
    if (int e = check_align(grad_init, 4, "grad_init")) return e;
    if (int e = check_align(workspace, 16, "workspace")) return e;
    LaunchArgs la;
    if (int e = fill_geom(&la, B, H, W, H, 0, 0, H, 8)) return e;
    la.grad_out = grad_out; la.init = init; la.weight = weight; la.offset = offset; la.w9 = w9;
    la.grad_init = grad_init; la.grad_weight = grad_weight; la.grad_offset = grad_offset;
    la.grad_w9 = grad_w9; la.grad_b1 = grad_b1; la.workspace = workspace;
    la.accumulate = (flags & JSPSR_BWD_ACCUMULATE) != 0;
    la.gen_preact = preact;
    if (prd.world > 1) la.peer_reduce = &prd;
    la.mode = norm_mode; la.scale = scale; la.bf16 = dtype != JSPSR_F32; la.init_f32 = dtype == JSPSR_MIXED;
    la.stream = (cudaStream_t)stream;
    la.pair = pair_ok(dtype, {weight, offset, grad_weight, grad_offset}, {grad_out});
    const bool wide = choose_wide(W);
    la.use_tma = make_init_tmap(&la.tmap, init, B, H, W, dtype == JSPSR_BF16, la.tile_h, wide);
    if (grad_init) {
        cudaError_t ce = cudaMemsetAsync(grad_init, 0, (size_t)B * H * W * sizeof(float), la.stream);
        if (ce != cudaSuccess) return cuda_fail(ce, "grad_init memset");
    }
    cudaError_t ce = wide ? wide::launch_spn_backward(la) : narrow::launch_spn_backward(la);
    if (ce != cudaSuccess) return cuda_fail(ce, "spn_backward launch");
    return JSPSR_OK;
}

int jspsr_gen_spn_forward(const void* init, const void* feature, const float* conv_w, const float* conv_b,
                          const float* w9, const float* b1, void* out, void* weight_out, void* offset_out, int B, int C,
                          int H, int W, int norm_mode, float scale, int dtype, void* stream) {
    if (int e = check_common(B, H, W, norm_mode, dtype)) return e;
    if (dtype != JSPSR_F32 && dtype != JSPSR_MIXED)
        return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_gen_spn_forward: dtype must be 0 (all fp32) or 2 (mixed: bf16 feature / "
                                           "weight_out / offset_out, fp32 init / out)");
    const size_t esf = dtype == JSPSR_MIXED ? 2 : 4;
    if (C != 64 && C != 128)
        return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_gen_spn_forward is instantiated for C = 128 feature channels (Generator "
                                           "bc = 32: configs/*.yml num_feature = 32 with cat_only) and C = 64 (bc = 16), got C = %d", C);
    if (!init || !feature || !conv_w || !conv_b || !w9 || !b1 || !out) return fail(JSPSR_ERR_BAD_ARG, "null tensor pointer");
    if ((weight_out == nullptr) != (offset_out == nullptr))
        return fail(JSPSR_ERR_BAD_ARG, "weight_out and offset_out must be given together");
    if (int e = check_align(init, 4, "init")) return e;
    if (int e = check_align(feature, esf, "feature")) return e;
    if (int e = check_align(conv_w, 16, "conv_w")) return e;
    if (int e = check_align(conv_b, 4, "conv_b")) return e;
    if (int e = check_align(out, 4, "out")) return e;
    if (int e = check_align(weight_out, esf, "weight_out")) return e;
    if (int e = check_align(offset_out, esf, "offset_out")) return e;
    LaunchArgs la;
    if (int e = fill_geom(&la, B, H, W, H, 0, 0, H, 16)) return e;
    if (la.tile_h < 8) {  // instantiated for 16 and 8 rows per CTA
        la.tile_h = 8;
        la.g.tiles_y = (H + 7) / 8;
    }
    if (C == 128) {  // fp32: one CTA per SM, 16 rows each; bf16 features: two CTAs per SM, 8 rows each
        la.tile_h = dtype == JSPSR_MIXED ? 8 : 16;
        la.g.tiles_y = (H + la.tile_h - 1) / la.tile_h;
    }
    la.init = init; la.w9 = w9; la.b1 = b1; la.out = out;
    la.mode = norm_mode; la.scale = scale; la.bf16 = dtype == JSPSR_MIXED; la.init_f32 = true;
    la.stream = (cudaStream_t)stream;
    // the operand buffers leave room for the narrow staged tile at two CTAs per SM (JSPSR_SPN_HALO=wide overrides)
    bool wide = false;
    if (const char* e = getenv("JSPSR_SPN_HALO")) wide = e[0] == 'w';
    CUtensorMap tmap_feat{};
    la.use_tma = make_init_tmap(&la.tmap, init, B, H, W, false, la.tile_h, wide) &&
                 make_feature_tmap(&tmap_feat, feature, B, C, H, W, la.bf16);
    cudaError_t ce = (wide ? wide::launch_gen_spn_forward : narrow::launch_gen_spn_forward)(la, tmap_feat, feature, C, conv_w,
                                                                                            conv_b, weight_out, offset_out);
    if (ce != cudaSuccess) return cuda_fail(ce, "gen_spn_forward launch");
    return JSPSR_OK;
}

int jspsr_gen_tail_grad_feature(const void* gz, const float* conv_w, void* grad_feature, int B, int C, int H, int W,
                                int dtype, void* stream) {
    if (int e = check_common(B, H, W, 0, dtype)) return e;
    if (dtype == JSPSR_MIXED) return fail(JSPSR_ERR_BAD_ARG, "jspsr_gen_tail_grad_feature: dtype is 0 (f32) or 1 (bf16)");
    if (C != 64 && C != 128)
        return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_gen_tail_grad_feature is instantiated for C = 64 and C = 128, got C = %d", C);
    if (!gz || !conv_w || !grad_feature) return fail(JSPSR_ERR_BAD_ARG, "null tensor pointer");
    const size_t es = dtype == JSPSR_BF16 ? 2 : 4;
    if (int e = check_align(gz, es, "gz")) return e;
    if (int e = check_align(grad_feature, es, "grad_feature")) return e;
    if (int e = check_align(conv_w, 4, "conv_w")) return e;
    LaunchArgs la;
    if (int e = fill_geom(&la, B, H, W, H, 0, 0, H, 16)) return e;
    la.tile_h = 16;  // the kernel always walks 16 rows per CTA
    la.g.tiles_y = (H + 15) / 16;
    la.bf16 = dtype == JSPSR_BF16;
    la.stream = (cudaStream_t)stream;
    CUtensorMap tmap_gz{};
    la.use_tma = make_feature_tmap(&tmap_gz, gz, B, 25, H, W, la.bf16);
    cudaError_t ce = narrow::launch_gen_grad_feature(la, tmap_gz, gz, C, conv_w, grad_feature);
    if (ce != cudaSuccess) return cuda_fail(ce, "gen_grad_feature launch");
    return JSPSR_OK;
}

size_t jspsr_gen_tail_workspace_bytes(void) { return narrow::gen_grad_weight_workspace_bytes(); }

int jspsr_gen_tail_grad_params(const void* gz, const void* feature, float* grad_conv_w, float* grad_conv_b, void* workspace,
                               int B, int C, int H, int W, int dtype, void* stream) {
    if (int e = check_common(B, H, W, 0, dtype)) return e;
    if (dtype == JSPSR_MIXED) return fail(JSPSR_ERR_BAD_ARG, "jspsr_gen_tail_grad_params: dtype is 0 (f32) or 1 (bf16)");
    const bool bf16 = dtype == JSPSR_BF16;
    const size_t es = bf16 ? 2 : 4;
    if (C != 64 && C != 128)
        return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_gen_tail_grad_params is instantiated for C = 64 and C = 128, got C = %d", C);
    if (!gz || !feature || !workspace || (!grad_conv_w && !grad_conv_b)) return fail(JSPSR_ERR_BAD_ARG, "null pointer");
    if (int e = check_align(gz, es, "gz")) return e;
    if (int e = check_align(feature, es, "feature")) return e;
    if (int e = check_align(grad_conv_w, 4, "grad_conv_w")) return e;
    if (int e = check_align(grad_conv_b, 4, "grad_conv_b")) return e;
    if (int e = check_align(workspace, 16, "workspace")) return e;
    const size_t HW = (size_t)H * W;
    if (HW > 0x7fffffffull) return fail(JSPSR_ERR_UNSUPPORTED, "H * W = %zu exceeds 2^31 - 1", HW);
    int run_len = 4;   // K-blocks (128 bytes of pixels per plane) per fp32 accumulation run (JSPSR_GEN_WGRAD_RUN overrides: measurements)
    if (const char* e = getenv("JSPSR_GEN_WGRAD_RUN")) {
        const int v = atoi(e);
        if (v >= 2 && v <= (1 << 20)) run_len = v;
    }
    CUtensorMap tmap_f{}, tmap_gz{};
    const bool use_tma = make_plane_tmap(&tmap_f, feature, (size_t)B * C, HW, C, true, bf16) &&
                         make_plane_tmap(&tmap_gz, gz, (size_t)B * 25, HW, 25, false, bf16);
    cudaError_t ce = narrow::launch_gen_grad_weight(gz, feature, C, bf16, grad_conv_w, grad_conv_b, workspace, B, (int)HW, run_len,
                                                    use_tma, tmap_f, tmap_gz, (cudaStream_t)stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "gen_grad_weight launch");
    return JSPSR_OK;
}

int jspsr_spn_offset_absmax(const void* offset, int B, int H, int W, int dtype, float* out2, void* stream) {
    if (int e = check_common(B, H, W, 0, dtype)) return e;
    if (dtype == JSPSR_MIXED) return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_spn_offset_absmax: the mixed dtype is implemented for jspsr_spn_forward/backward");
    if (!offset || !out2) return fail(JSPSR_ERR_BAD_ARG, "null pointer");
    cudaError_t ce = launch_offset_absmax(offset, 0, (size_t)H * W, B, dtype == JSPSR_BF16, out2, (cudaStream_t)stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "offset_absmax launch");
    return JSPSR_OK;
}

int jspsr_spn_iterate_backward(const void* grad_list, const void* feat_init, const void* list_out, const void* aff,
                               const void* offset, void* grad_feat, void* grad_aff, void* grad_offset, void* carry_scratch,
                               int B, int H, int W, int T, int dtype, void* stream) {
    if (int e = check_common(B, H, W, 0, dtype)) return e;
    if (dtype != JSPSR_F32)
        return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_spn_iterate_backward: fp32 only (run T applications of jspsr_spn_backward for bf16)");
    if (T <= 0 || T > 8)
        return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_spn_iterate_backward: T=%d is outside [1, 8] (the T staged features of a CTA live in shared memory)", T);
    if (!grad_list || !feat_init || !aff || !offset || !grad_aff || !grad_offset || (T > 1 && (!list_out || !carry_scratch)))
        return fail(JSPSR_ERR_BAD_ARG, "null tensor pointer");
    for (const void* p : {grad_list, feat_init, list_out, aff, offset, (const void*)grad_feat, (const void*)grad_aff,
                          (const void*)grad_offset, (const void*)carry_scratch})
        if (int e = check_align(p, 4, "jspsr_spn_iterate_backward tensor")) return e;
    LaunchArgs la;
    if (int e = fill_geom(&la, B, H, W, H, 0, 0, H, 8)) return e;
    la.tile_h = 8;  // both kernels are instantiated for 8 rows per CTA
    la.g.tiles_y = (H + 7) / 8;
    const size_t n = (size_t)B * H * W;
    if ((size_t)T * B > 0x7fffffffull) return fail(JSPSR_ERR_UNSUPPORTED, "T * B exceeds 2^31 - 1");
    cudaStream_t st = (cudaStream_t)stream;
    const float* gl = (const float*)grad_list;
    float* carry = (float*)carry_scratch;
    // (A) the carry chain: step t sends P^T g_t to step t - 1 (slot t - 1 of the scratch; step 0 to grad_feat)
    if (T > 1) {
        cudaError_t ce = cudaMemsetAsync(carry, 0, (size_t)(T - 1) * n * sizeof(float), st);
        if (ce != cudaSuccess) return cuda_fail(ce, "carry memset");
    }
    if (grad_feat) {
        cudaError_t ce = cudaMemsetAsync(grad_feat, 0, n * sizeof(float), st);
        if (ce != cudaSuccess) return cuda_fail(ce, "grad_feat memset");
    }
    const bool wide = choose_wide(W);
    for (int t = T - 1; t >= 0; --t) {
        float* dst = t > 0 ? carry + (size_t)(t - 1) * n : (float*)grad_feat;
        if (!dst) break;
        const float* g_b = t < T - 1 ? carry + (size_t)t * n : nullptr;
        // slot T - 1 of the scratch: sum_k |a_k| per pixel, written by the first launch of the chain, read by the others
        float* asum = T > 1 ? carry + (size_t)(T - 1) * n : nullptr;
        cudaError_t ce = (wide ? wide::launch_iter_carry : narrow::launch_iter_carry)(
            gl + (size_t)t * n, g_b, (const float*)aff, (const float*)offset, t < T - 1 ? asum : nullptr,
            t == T - 1 ? asum : nullptr, dst, la.g, st);
        if (ce != cudaSuccess) return cuda_fail(ce, "iter_carry launch");
    }
    // (B) the 27 gradients, summed over t in registers; narrow staged halo: T tiles of a CTA stay resident
    CUtensorMap tm_init{}, tm_list{};
    int gth = narrow::iter_grad_tile_h(T);
    // small batches: 8-row CTAs keep the grid at two CTAs per SM or more (measured at 2 tiles: 0.105 vs 0.112 ms)
    if (gth == 16 && !getenv("JSPSR_ITER_GRAD_TH") && (size_t)la.g.tiles_x * ((H + 15) / 16) * B < (size_t)148 * 2) gth = 8;
    Geom gg = la.g;
    gg.tiles_y = (H + gth - 1) / gth;
    bool use_tma = make_init_tmap(&tm_init, feat_init, B, H, W, false, gth, false);
    if (use_tma && T > 1) use_tma = make_init_tmap(&tm_list, list_out, (T - 1) * B, H, W, false, gth, false);
    cudaError_t ce = narrow::launch_iter_grad(gl, carry, (const float*)feat_init, (const float*)list_out, (const float*)aff,
                                              (const float*)offset, (float*)grad_aff, (float*)grad_offset, gg, T, gth,
                                              use_tma, tm_init, tm_list, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "iter_grad launch");
    return JSPSR_OK;
}

int jspsr_preserve_blend(const void* feat, const void* fix, void* dst, long long n, int dtype, void* stream) {
    if (n < 0) return fail(JSPSR_ERR_BAD_ARG, "jspsr_preserve_blend: n=%lld is negative", n);
    if (dtype != JSPSR_F32 && dtype != JSPSR_BF16)
        return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_preserve_blend: dtype %d (one dtype for the three tensors: 0 f32, 1 bf16)", dtype);
    if (n == 0) return JSPSR_OK;
    if (!feat || !fix || !dst) return fail(JSPSR_ERR_BAD_ARG, "null tensor pointer");
    const uintptr_t al = dtype == JSPSR_BF16 ? 1 : 3;
    if (((uintptr_t)feat | (uintptr_t)fix | (uintptr_t)dst) & al) return fail(JSPSR_ERR_ALIGN, "jspsr_preserve_blend: misaligned pointer");
    cudaError_t ce = launch_preserve_blend_auto(feat, fix, dst, (size_t)n, dtype == JSPSR_BF16, (cudaStream_t)stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "preserve_blend launch");
    return JSPSR_OK;
}

int jspsr_spn_iterate(const void* feat_init, const void* aff, const void* offset, const void* feat_fix,
                      const void* mask_fix, void* list_out, void* scratch, int B, int H, int W, int T, int dtype,
                      void* stream) {
    if (int e = check_common(B, H, W, 0, dtype)) return e;
    if (dtype == JSPSR_MIXED) return fail(JSPSR_ERR_UNSUPPORTED, "jspsr_spn_iterate: the mixed dtype is implemented for jspsr_spn_forward/backward");
    if (T <= 0) return fail(JSPSR_ERR_BAD_ARG, "T=%d must be positive", T);
    if (!feat_init || !aff || !offset || !list_out) return fail(JSPSR_ERR_BAD_ARG, "null tensor pointer");
    if ((feat_fix == nullptr) != (mask_fix == nullptr))
        return fail(JSPSR_ERR_BAD_ARG, "feat_fix and mask_fix must be given together");
    if (feat_fix && !scratch) return fail(JSPSR_ERR_BAD_ARG, "preserve_input needs a [B,1,H,W] scratch buffer");
    const bool bf16 = dtype == JSPSR_BF16;
    const size_t es = bf16 ? 2 : 4, px = (size_t)H * W, n = (size_t)B * px;
    // Optional single-launch form (JSPSR_SPN_ITER_FUSED=1): one 16-CTA cluster per sample keeps the iteration-invariant
    // tap state in registers and the feature in (distributed) shared memory for all T applications
    // (spn_iterate_fused.cu; bit-identical results).  OFF by default: measured on B200 it does not beat T launches - the
    // shared-memory gather, not HBM, bounds an application.
    if (const char* e = getenv("JSPSR_SPN_ITER_FUSED")) {
        if (e[0] == '1' && !bf16 && !feat_fix && T >= 2 && H <= 128 && W <= 128) {
            cudaError_t ce = narrow::launch_spn_iterate_fused((const float*)feat_init, (const float*)aff, (const float*)offset,
                                                              (float*)list_out, B, H, W, T, (cudaStream_t)stream);
            if (ce == cudaSuccess) return JSPSR_OK;
            if (ce != cudaErrorNotSupported) return cuda_fail(ce, "spn_iterate_fused launch");
        }
    }
    // Optional L2 blocking (JSPSR_SPN_ITER_CHUNK_MB > 0): run all T steps on one chunk of samples whose
    // affinities/offsets fit the L2 before moving on.  OFF by default: measured on B200 (tools/iter_bench.py,
    // 4096 tiles, T = 6) the dependent, sub-wave launches it creates are 1.6-2.8x SLOWER than T full-batch
    // launches (12.4-21.3 ms vs 7.7 ms); real T-fusion needs a persistent cluster-per-sample kernel.
    size_t budget = 0;
    if (const char* e = getenv("JSPSR_SPN_ITER_CHUNK_MB")) budget = (size_t)atol(e) << 20;
    size_t chunk = (size_t)B;
    if (budget > 0 && T > 1) {
        chunk = budget / (px * 27 * es);
        if (chunk < 1) chunk = 1;
        if (chunk > (size_t)B) chunk = (size_t)B;
    }
    for (size_t b0 = 0; b0 < (size_t)B; b0 += chunk) {
        const int nb = (int)(((size_t)B - b0 < chunk) ? ((size_t)B - b0) : chunk);
        const char* src = (const char*)feat_init + b0 * px * es;
        const char* aff_c = (const char*)aff + b0 * px * 9 * es;
        const char* off_c = (const char*)offset + b0 * px * 18 * es;
        for (int t = 0; t < T; ++t) {
            if (feat_fix) {  // nlspn.py:228-229
                char* blend = (char*)scratch + b0 * px * es;
                cudaError_t ce = launch_preserve_blend(src, (const char*)feat_fix + b0 * px * es,
                                                       (const float*)mask_fix + b0 * px, blend, (size_t)nb * px, bf16,
                                                       (cudaStream_t)stream);
                if (ce != cudaSuccess) return cuda_fail(ce, "preserve_input blend launch");
                src = blend;
            }
            char* dst = (char*)list_out + ((size_t)t * n + b0 * px) * es;
            LaunchArgs la;
            if (int e = fill_geom(&la, nb, H, W, H, 0, 0, H, 16)) return e;
            la.init = src; la.weight = aff_c; la.offset = off_c; la.w9 = nullptr; la.b1 = nullptr; la.out = dst;
            la.mode = NORM_NONE; la.scale = 0.f; la.bf16 = bf16; la.stream = (cudaStream_t)stream;
            const bool wide = choose_wide(W);
            la.use_tma = make_init_tmap(&la.tmap, src, nb, H, W, bf16, la.tile_h, wide);
            cudaError_t ce = wide ? wide::launch_spn_forward(la) : narrow::launch_spn_forward(la);
            if (ce != cudaSuccess) return cuda_fail(ce, "spn_iterate launch");
            src = dst;
        }
    }
    return JSPSR_OK;
}

int jspsr_nlspn_affinity_forward(const void* conv_out, const void* confidence, const float* aff_scale_const,
                                 void* offset_out, void* aff_out, int B, int H, int W, int affinity, int legacy,
                                 int dtype, void* stream) {
    if (int e = check_common(B, H, W, 0, dtype)) return e;
    if (!conv_out || !offset_out || !aff_out || !aff_scale_const) return fail(JSPSR_ERR_BAD_ARG, "null pointer");
    if (dtype == JSPSR_MIXED) return fail(JSPSR_ERR_UNSUPPORTED, "the mixed dtype is implemented for jspsr_spn_forward/backward");
    if (affinity < 0 || affinity > 3) return fail(JSPSR_ERR_BAD_ARG, "affinity %d is not AS/ASS/TC/TGASS", affinity);
    if (legacy && !confidence) return fail(JSPSR_ERR_BAD_ARG, "legacy only has an effect with confidence propagation");
    LaunchArgs la;  // geometry + TMA descriptor of the confidence map (the staged tensor of this kernel)
    if (int e = fill_geom(&la, B, H, W, H, 0, 0, H, 8)) return e;
    if (la.tile_h == 4) {  // the affinity kernels are instantiated for 8 and 2 rows per CTA
        la.tile_h = 2;
        la.g.tiles_y = (H + 1) / 2;
    }
    const bool bf16 = dtype == JSPSR_BF16;
    const bool wide = choose_wide(W);
    la.use_tma = confidence && make_init_tmap(&la.tmap, confidence, B, H, W, bf16, la.tile_h, wide);
    cudaError_t ce = (wide ? wide::launch_nlspn_affinity : narrow::launch_nlspn_affinity)(conv_out, confidence, aff_scale_const, offset_out, aff_out, nullptr, nullptr,
                                           nullptr, nullptr, nullptr, nullptr, la.g, la.tile_h, la.use_tma, la.tmap,
                                           affinity, legacy, bf16, true, (cudaStream_t)stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "nlspn_affinity_forward launch");
    return JSPSR_OK;
}

int jspsr_nlspn_affinity_backward(const void* grad_offset, const void* grad_aff, const void* conv_out,
                                  const void* confidence, const float* aff_scale_const, void* grad_conv_out,
                                  float* grad_confidence, float* grad_scale, void* workspace, int B, int H, int W,
                                  int affinity, int dtype, void* stream) {
    if (int e = check_common(B, H, W, 0, dtype)) return e;
    if (!grad_offset || !grad_aff || !conv_out || !grad_conv_out || !aff_scale_const)
        return fail(JSPSR_ERR_BAD_ARG, "null pointer");
    if (dtype == JSPSR_MIXED) return fail(JSPSR_ERR_UNSUPPORTED, "the mixed dtype is implemented for jspsr_spn_forward/backward");
    if (affinity < 0 || affinity > 3) return fail(JSPSR_ERR_BAD_ARG, "affinity %d is not AS/ASS/TC/TGASS", affinity);
    if (grad_scale && !workspace) return fail(JSPSR_ERR_BAD_ARG, "workspace is required when grad_scale is requested");
    if (grad_confidence && !confidence) return fail(JSPSR_ERR_BAD_ARG, "grad_confidence without confidence");
    if (grad_confidence) {
        cudaError_t ce = cudaMemsetAsync(grad_confidence, 0, (size_t)B * H * W * sizeof(float), (cudaStream_t)stream);
        if (ce != cudaSuccess) return cuda_fail(ce, "grad_confidence memset");
    }
    LaunchArgs la;
    if (int e = fill_geom(&la, B, H, W, H, 0, 0, H, 8)) return e;
    if (la.tile_h == 4) {
        la.tile_h = 2;
        la.g.tiles_y = (H + 1) / 2;
    }
    const bool bf16 = dtype == JSPSR_BF16;
    const bool wide = choose_wide(W);
    la.use_tma = confidence && make_init_tmap(&la.tmap, confidence, B, H, W, bf16, la.tile_h, wide);
    cudaError_t ce = (wide ? wide::launch_nlspn_affinity : narrow::launch_nlspn_affinity)(conv_out, confidence, aff_scale_const, nullptr, nullptr, grad_offset, grad_aff,
                                           grad_conv_out, grad_confidence, grad_scale, workspace, la.g, la.tile_h,
                                           la.use_tma, la.tmap, affinity, 0, bf16, false, (cudaStream_t)stream);
    if (ce != cudaSuccess) return cuda_fail(ce, "nlspn_affinity_backward launch");
    return JSPSR_OK;
}

}  // extern "C"
