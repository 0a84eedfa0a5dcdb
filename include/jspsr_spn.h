/*
 * jspsr_spn.h - C ABI of libjspsr_spn.so: the B200 (sm_100a) implementation of
 * JSPSR's spatial-propagation refinement step.
 *
 * The reference (xandercai/JSPSR) is pure Python and has no FFI of its own; the
 * functions below are what a binding for its propagation path has to reach.
 * Each entry point cites the reference interface it replaces (paths relative
 * to the reference tree).  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *  - plain C symbols, device pointers unless a name ends in `_host`, explicit
 *    stream (a cudaStream_t passed as void*), no torch types;
 *  - tensors are dense NCHW: init/out [B,1,H,W], weight [B,9,H,W],
 *    offset [B,18,H,W] with channel 2k = row offset and 2k+1 = column offset of
 *    tap k (k row-major over the 3x3 window) - the layout of
 *    models/components/spn.py:69-73 and torchvision.ops.deform_conv2d;
 *  - `dtype` selects the I/O element type of init/weight/offset/out and their
 *    gradients (arithmetic is always fp32); w9/b1 and their gradients are fp32;
 *  - every buffer, including outputs and the workspace, is owned by the caller;
 *    the library allocates nothing persistent and frees nothing;
 *  - all work is enqueued on `stream`, nothing synchronises, every call is
 *    CUDA-graph capturable;
 *  - return value: JSPSR_OK (0) or a negative jspsr_status; the message of the
 *    calling thread's last failure is returned by jspsr_last_error().
 *    There is no CPU fallback.
 */
#ifndef JSPSR_SPN_H_
#define JSPSR_SPN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#ifdef __GNUC__
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define JSPSR_SPN_VERSION 107 /* major*100 + minor */

typedef enum {
    JSPSR_OK = 0,
    JSPSR_ERR_BAD_ARG = -1,     /* null pointer, non-positive dimension, bad enum */
    JSPSR_ERR_UNSUPPORTED = -2, /* e.g. kernel_size != 3 */
    JSPSR_ERR_CUDA = -3,        /* launch / runtime error, see jspsr_last_error() */
    JSPSR_ERR_ALIGN = -4        /* pointer not aligned to its element type */
} jspsr_status;

/* affinity normalisation applied inside the kernel */
typedef enum {
    JSPSR_NORM_NONE = 0,     /* mask used as given: NLSPN._propagate_once, nlspn.py:177-187 */
    JSPSR_NORM_RESIDUAL = 1, /* m_k = a_k - mean_j a_j, and out += scale*init:
                                PostProcessor(residual=True) spn.py:100-101,116-117;
                                Post_process_deconv(dkn_residual=True) LRRU.py:268-269,295-296 */
    JSPSR_NORM_SUM = 2       /* m_k = a_k / sum_j a_j: spn.py:102-103, LRRU.py:270-271 */
} jspsr_norm_mode;

typedef enum {
    JSPSR_F32 = 0,
    JSPSR_BF16 = 1,
    JSPSR_MIXED = 2 /* jspsr_spn_forward / jspsr_spn_backward only: weight, offset and their gradients bf16;
                       init, out and grad_out fp32 - the tensors torch.autocast(bfloat16) hands to
                       PostProcessor.forward (bf16 Generator outputs, fp32 DEM), which torchvision's
                       operator promotes to fp32 */
} jspsr_dtype;

/* NLSPN affinity flavours, models/components/nlspn.py:35-56,92-99,162-166 */
typedef enum { JSPSR_AFF_AS = 0, JSPSR_AFF_ASS = 1, JSPSR_AFF_TC = 2, JSPSR_AFF_TGASS = 3 } jspsr_affinity;

/* flags for jspsr_spn_backward */
#define JSPSR_BWD_ACCUMULATE 1u /* grad_weight/grad_offset += (fixed-affinity T-step loop) instead of =;
                                   implemented together with grad_init only (JSPSR_ERR_UNSUPPORTED otherwise) */

#define JSPSR_BWD_GEN_PREACT 2u /* generator-tail training (jspsr_gen_spn_forward's backward): grad_weight points to a
                                   [B,25,H,W] tensor that receives the gradients w.r.t. the PRE-ACTIVATIONS of
                                   Generator.conv_weight / conv_offset (channels 0..8: dL/dweight_k * weight_k * (1 - weight_k);
                                   9..24: the 16 offset gradients without the centre pair); grad_offset is ignored
                                   (may be NULL); grad_init must be NULL (the DEM is detached, models/JSPSR.py:372) */

int jspsr_version(void);
const char *jspsr_last_error(void);

/* Bytes of caller-owned device scratch jspsr_spn_backward / jspsr_nlspn_affinity_backward
 * need for their global reductions.  The buffer must be zero-filled once when it is
 * allocated; the kernels leave it zeroed when they finish.  One buffer per stream. */
size_t jspsr_spn_workspace_bytes(void);

/*
 * Forward of one propagation application.  Replaces, in one kernel,
 *   PostProcessor.forward               models/components/spn.py:99-118
 *   Post_process_deconv.forward         models/LRRU.py:267-298
 *   NLSPN._propagate_once               models/components/nlspn.py:177-187
 * i.e. normalise -> torchvision.ops.deform_conv2d(init, offset, w, b, stride 1,
 * pad 1, dil 1, mask) -> (+ scale*init when norm_mode == RESIDUAL).
 * w9: 9 floats (the [1,1,3,3] parameter `w`), b1: 1 float (`b`), both on device.
 */
int jspsr_spn_forward(const void *init, const void *weight, const void *offset, const float *w9,
                      const float *b1, void *out, int B, int H, int W, int norm_mode, float scale,
                      int dtype, void *stream);

/*
 * Backward of jspsr_spn_forward (autograd of the above; upstream kernels
 * deformable_col2im / deformable_col2im_coord + the normalisation Jacobian).
 * grad_init may be NULL (JSPSR detaches the DEM, models/JSPSR.py:372).
 * grad_init (when given), grad_w9[9] and grad_b1[1] are OVERWRITTEN; grad_weight /
 * grad_offset are overwritten unless JSPSR_BWD_ACCUMULATE is set.  grad_init is always
 * fp32 (it is accumulated with atomics), whatever `dtype` is.
 * grad_w9 / grad_b1 may be NULL when the 3x3 weight and bias are frozen (nlspn.py:64-65).
 */
int jspsr_spn_backward(const void *grad_out, const void *init, const void *weight,
                       const void *offset, const float *w9, float *grad_init, void *grad_weight,
                       void *grad_offset, float *grad_w9, float *grad_b1, void *workspace, int B,
                       int H, int W, int norm_mode, float scale, int dtype, unsigned flags,
                       void *stream);

/*
 * Row-strip form of the forward for rasters sharded across GPUs (new; the reference
 * runs one device).  The strip owns output rows [row0, row0+Hs) of an image of
 * H_img rows; weight/offset/out hold exactly those rows.  `init` holds rows
 * [init_row0, init_row0+init_rows) of the image, i.e. the strip plus the halo rows
 * received from its neighbours.  Coordinates are formed from GLOBAL row indices so
 * every pixel sees the arithmetic of the unsharded call bit for bit.  A tap that
 * needs an image row missing from `init` sets *status (device int, may be NULL) to 1.
 */
int jspsr_spn_forward_strip(const void *init, const void *weight, const void *offset,
                            const float *w9, const float *b1, void *out, int B, int Hs, int W,
                            int H_img, int row0, int init_row0, int init_rows, int norm_mode,
                            float scale, int dtype, int *status, void *stream);

/*
 * Generator tail + propagation in one kernel (new boundary - the
 * reference runs these as separate modules).  Replaces, for the last two layers of
 * Generator.forward and the PostProcessor call that consumes them,
 *   weight = conv_weight(feature)  = sigmoid(1x1 conv, C -> 9)        models/components/spn.py:41-44,66
 *   offset = conv_offset(feature)  = 1x1 conv, C -> 16, then the zero centre pair is
 *            inserted (view/chunk/insert/cat)                          models/components/spn.py:45-52,67-73
 *   out    = PostProcessor.forward(init, weight, offset)               models/components/spn.py:99-118
 *            (called at models/JSPSR.py:375, models/EDSR.py:134)
 * feature [B,C,H,W] fp32 is the output of Generator.block (spn.py:65); conv_w is [25,C] row-major:
 * rows 0..8 = conv_weight[0].weight[9,C,1,1], rows 9..24 = conv_offset.conv[0].weight[16,C,1,1];
 * conv_b [25] the two biases in the same order (all on device, conv_w 16-byte aligned).
 * The contraction runs on the tensor cores (tcgen05, tf32 with a 3-product split: fp32-level accuracy).
 * weight_out [B,9,H,W] / offset_out [B,18,H,W]: both NULL (inference) or both given - they receive what
 * the Generator would have returned, which is what jspsr_spn_backward needs.  C = 128 (the YAML configs: models/JSPSR.py:28,181) or 64.
 * dtype JSPSR_F32: everything fp32.  dtype JSPSR_MIXED (torch.autocast): feature, weight_out, offset_out are
 * bf16 (init / out fp32); weight and offset are rounded to bf16 before the gather, i.e. `out` is exactly
 * jspsr_spn_forward(JSPSR_MIXED) of the tensors written.
 */
int jspsr_gen_spn_forward(const void *init, const void *feature, const float *conv_w,
                          const float *conv_b, const float *w9, const float *b1, void *out,
                          void *weight_out, void *offset_out, int B, int C, int H, int W,
                          int norm_mode, float scale, int dtype, void *stream);

/*
 * Feature gradient of the Generator tail: grad_feature[b,c,y,x] = sum_j gz[b,j,y,x] * conv_w[j,c] - the backward of
 * Generator.conv_weight / conv_offset (spn.py:41-52) w.r.t. their input, with gz [B,25,H,W] the pre-activation
 * gradients jspsr_spn_backward writes under JSPSR_BWD_GEN_PREACT and conv_w the [25,C] matrix of
 * jspsr_gen_spn_forward.  Tensor-core contraction (tcgen05, tf32 3-product split).  gz and grad_feature [B,C,H,W] share
 * `dtype` (JSPSR_F32 or JSPSR_BF16); C = 64 or 128.
 */
int jspsr_gen_tail_grad_feature(const void *gz, const float *conv_w, void *grad_feature, int B, int C,
                                int H, int W, int dtype, void *stream);

/*
 * Parameter gradients of the Generator tail: grad_conv_w[j,c] = sum over (b,y,x) of gz[b,j,y,x] * feature[b,c,y,x]
 * ([25,C], the layout of conv_w) and grad_conv_b[j] = sum of gz[b,j,y,x] - the backward of Generator.conv_weight /
 * conv_offset (spn.py:41-52) w.r.t. their weights and biases, which the reference gets from cuDNN's convolution
 * backward.  One pass over gz and feature, contraction over the pixel index on the tensor cores (tcgen05, tf32
 * 3-product split, short fp32 runs combined in fp64).  Either output may be NULL.  workspace: device memory of
 * jspsr_gen_tail_workspace_bytes() bytes, 16-byte aligned, ZERO on the first call; the kernel leaves it zero.
 * dtype: JSPSR_F32, or JSPSR_BF16 for bf16 gz and feature (torch.autocast; exact in tf32, one product); the outputs are
 * always fp32 (the parameters' dtype).  C = 64 or 128.
 */
size_t jspsr_gen_tail_workspace_bytes(void);
int jspsr_gen_tail_grad_params(const void *gz, const void *feature, float *grad_conv_w,
                               float *grad_conv_b, void *workspace, int B, int C, int H, int W,
                               int dtype, void *stream);

/* max |row offset| and max |column offset| over a [B,18,H,W] tensor -> out2[2] (device,
 * combined with max so several calls may fold into one pair; zero it first).  Used to
 * size strip halos. */
int jspsr_spn_offset_absmax(const void *offset, int B, int H, int W, int dtype, float *out2,
                            void *stream);

/*
 * T applications with fixed affinity and offsets (NLSPN.forward loop,
 * models/components/nlspn.py:222-235, w = 1, b = 0, mask = aff).  list_out is
 * [T,B,1,H,W]; step t reads step t-1's slice.  feat_fix/mask_fix (both or neither;
 * mask_fix fp32 0/1 [B,1,H,W]) implement `preserve_input` (nlspn.py:217-229).
 */
int jspsr_spn_iterate(const void *feat_init, const void *aff, const void *offset,
                      const void *feat_fix, const void *mask_fix, void *list_out, void *scratch,
                      int B, int H, int W, int T, int dtype, void *stream);

/*
 * Backward of jspsr_spn_iterate without feat_fix (autograd of the loop nlspn.py:222-235), fp32, 1 <= T <= 8.
 * grad_list [T,B,1,H,W]: gradient w.r.t. every step's output (zeros where unused); list_out: the forward's
 * result (steps 0 .. T-2 are read).  Writes grad_aff [B,9,H,W], grad_offset [B,18,H,W] and, unless NULL,
 * grad_feat [B,1,H,W] (gradient w.r.t. feat_init).  carry_scratch: T * B * H * W floats, caller-owned.
 * Affinities and offsets do not change over the loop, so the gradient that flows from step to step is computed by
 * T light launches (a scatter, no gather, no 27-channel gradient traffic) and the 27 gradients are summed over t in
 * registers by one kernel that holds the T staged features of its tile in shared memory - instead of T
 * applications of jspsr_spn_backward with JSPSR_BWD_ACCUMULATE, whose read-modify-write of 27 channels per step
 * is what bounds them.  Same per-step arithmetic; results agree to fp32 rounding.
 */
int jspsr_spn_iterate_backward(const void *grad_list, const void *feat_init, const void *list_out,
                               const void *aff, const void *offset, void *grad_feat, void *grad_aff,
                               void *grad_offset, void *carry_scratch, int B, int H, int W, int T,
                               int dtype, void *stream);

/*
 * Input-preservation blend of the LRRU cascade (models/LRRU.py:447-451, 460-464, 474-478, 488-492) for a
 * single-channel `fix` (d_clear): dst = (1 - m) * feat + m * fix with m = (fix > 0), element by element over n
 * elements, one pass instead of the six elementwise kernels of
 *     mask = (torch.sum(d_clear > 0.0, dim=1, keepdim=True) > 0.0).type_as(d_clear)
 *     output = (1.0 - mask) * output + mask * d_clear
 * (same roundings: both products, then the sum).  dst may alias feat.  dtype 0 f32 / 1 bf16 for all three.
 */
int jspsr_preserve_blend(const void *feat, const void *fix, void *dst, long long n, int dtype,
                         void *stream);

/*
 * NLSPN affinity front-end after conv_offset_aff (models/components/nlspn.py:82-175):
 * offset re-packing with the zero centre pair, tanh/gamma scaling, confidence gating
 * (eight 1-tap deformable gathers), abs-sum normalisation, centre weight.
 * conv_out [B,24,H,W]; confidence [B,1,H,W] or NULL (conf_prop off);
 * aff_scale_const: 1 float on device; offset_out [B,18,H,W]; aff_out [B,9,H,W].
 */
int jspsr_nlspn_affinity_forward(const void *conv_out, const void *confidence,
                                 const float *aff_scale_const, void *offset_out, void *aff_out,
                                 int B, int H, int W, int affinity, int legacy, int dtype,
                                 void *stream);

/* Backward of the above: grad_conv_out [B,24,H,W] overwritten; grad_confidence
 * (fp32 [B,1,H,W], NULL when not needed) overwritten; grad_scale (1 float, NULL when
 * the constant is frozen) overwritten. */
int jspsr_nlspn_affinity_backward(const void *grad_offset, const void *grad_aff,
                                  const void *conv_out, const void *confidence,
                                  const float *aff_scale_const, void *grad_conv_out,
                                  float *grad_confidence, float *grad_scale, void *workspace,
                                  int B, int H, int W, int affinity, int dtype, void *stream);

/*
 * Host-buffer form of the forward: the call a CPU-side caller (the reference's
 * eval loop, evaluation/evaluate_utils.py:305, or utils/utils.py:1556 upscale_dem)
 * makes with tensors that live in host memory.  Batches are streamed through
 * `dev_scratch` (device, caller-owned, >= jspsr_spn_host_scratch_bytes) in chunks,
 * H2D copy / kernel / D2H copy overlapped on internal streams; returns after the
 * result is in `out`.  Pinned host memory gives full PCIe rate; pageable works.
 */
size_t jspsr_spn_host_scratch_bytes(int chunk_B, int H, int W, int dtype);
int jspsr_spn_forward_host(const void *init, const void *weight, const void *offset,
                           const float *w9, const float *b1, void *out, int B, int H, int W,
                           int norm_mode, float scale, int dtype, void *dev_scratch,
                           size_t scratch_bytes, int chunk_B);

#ifdef __GNUC__
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* JSPSR_SPN_H_ */
