/*
 * jspsr_tiles.h - C ABI of the components either side of the propagation hot path,
 * exported by the same library as include/jspsr_spn.h (libjspsr_spn.so) under the same conventions: plain C,
 * DEVICE pointers + sizes, an explicit cudaStream_t (passed as void*), the caller owns every buffer, nothing is
 * allocated or synchronised, 0 on success / negative jspsr_status on failure with the message in jspsr_last_error(),
 * CUDA-graph capturable, no CPU fallback.
 *
 * The reference (xandercai/JSPSR) is pure Python and has no FFI; each entry point cites the Python it replaces.
 * INTEGRATION.md shows the ctypes stubs a maintainer would add.
 */
#ifndef JSPSR_TILES_H_
#define JSPSR_TILES_H_

#include "jspsr_spn.h"

#ifdef __cplusplus
extern "C" {
#endif
#ifdef __GNUC__
#pragma GCC visibility push(default)
#endif

/*
 * Tile scheduler: cut a raster into overlapping k x k tiles, optionally through a mirrored border.  Replaces
 *   TileCrop.__call__ / get_tile      data/data_utils.py:87-197   (row-major walk, origin = stride * (row, col))
 *   add_padding (upscale_dem)         utils/utils.py:1501-1522    (pad > 0; the bottom border mirrors image row
 *                                                                  H-2-i as the reference's slice does, :1517)
 *   ToTensor's HWC -> CHW             data/data_utils.py:232      (tiles are written channels-first)
 * src [C,H,W] fp32, dst [n_y*n_x, C, k, k] fp32.  The walk must stay inside the (padded) image:
 * stride*(n-1)+k <= size + 2*pad per axis.  pad <= min(W, H-1).
 */
int jspsr_tiles_crop(const float *src, float *dst, int C, int H, int W, int pad, int k, int stride,
                     int n_y, int n_x, void *stream);

/*
 * Blended merge of per-tile predictions into one raster.  Replaces merge_dem(file_list, border, method=copyto_add)
 * utils/utils.py:916-965 with gen_weight_row / gen_weight_col (:802-900) and copyto_add (:903-913):
 * every tile loses `crop` = int(k * border) pixels per side (L = k - 2*crop), tiles sit `stride` apart, the
 * p = L - stride overlapped pixels are blended with linspace(1, 0, p+2)[1:-1] ramps (first tile: ramp down at its
 * end; last: ramp up at its start; middle: both), tiles of a row are accumulated left to right, the merged rows
 * top to bottom.  Arithmetic is float64 in the reference's order: with out_f64 != 0 the result is bit-identical to
 * the reference's numpy result; out_f64 == 0 rounds it to fp32 on store.
 * tiles [S, n_y*n_x, k, k] fp32 (S samples, row-major tiles), out [S, stride*(n_y-1)+L, stride*(n_x-1)+L].
 * The reference handles n_x = n_y in {2, 3}; any n >= 1 follows the same first / middle / last rule.
 */
int jspsr_tiles_merge(const float *tiles, void *out, int S, int n_y, int n_x, int k, int crop, int stride,
                      int out_f64, void *stream);

/*
 * Training loss of the YAML configs and its gradient in one pass.  Replaces
 *   MultiLoss.forward                 losses/loss_schemes.py:55-72   (Total = sum_k weight_k * loss_k)
 *   nn.L1Loss / nn.MSELoss            losses/loss_schemes.py:8-11
 *   EdgeLoss                          losses/loss_functions.py:171-185 (L1 of kornia.filters.spatial_gradient:
 *                                     normalised 3x3 Sobel pair, replicate border)
 *   and autograd's backward of all of them w.r.t. pred.
 * pred, gt [planes, H, W] fp32 (planes = B*C).  losses4 (device) receives {L1, L2, Grad, Total}.  grad_pred
 * (nullable) receives dTotal/dpred for an upstream gradient of 1.  workspace: jspsr_spn_workspace_bytes() of
 * zero-filled device memory (left zeroed).
 */
int jspsr_loss_l1_l2_grad(const float *pred, const float *gt, float w_l1, float w_l2, float w_grad,
                          float *losses4, float *grad_pred, void *workspace, int planes, int H, int W,
                          void *stream);

/*
 * Validation metric sums per sample in one pass.  Replaces MeterBase._prepare (evaluation/metrics.py:142-199:
 * crop border_h/border_w = int(size * border) pixels, clamp pred to [0,1]) + ToDEM.descale_data
 * (data/data_utils.py:441-457: exp(x*log(max-min))+min when elev_log, else x*(max-min)+min) + MeterRMSE.update's
 * squared error (metrics.py:376-382), plus the absolute error.  pred, gt [B,1,H,W] fp32;
 * sums [B,2] float64 (zeroed by the call on `stream`) receives {sum d^2, sum |d|} over the cropped window;
 * RMSE_b = sqrt(sums[b][0] / n), MAE_b = sums[b][1] / n with n = (H-2*border_h)*(W-2*border_w).
 */
int jspsr_dem_metrics(const float *pred, const float *gt, double *sums, int B, int H, int W, int border_h,
                      int border_w, float value_min, float value_max, int elev_log, void *stream);

#ifdef __GNUC__
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* JSPSR_TILES_H_ */
