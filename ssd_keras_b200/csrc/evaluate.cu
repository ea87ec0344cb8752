// Evaluator.match_predictions on the device (reference eval_utils/average_precision_evaluator.py:538-736) and the cumulative
// true / false positive counts that compute_precision_recall starts from (:724-729).
//
// The reference walks the predictions of one class in descending confidence; each prediction is compared (element-wise iou,
// :679) with the ground-truth boxes of its image that have the same class, matched to the best one if the overlap reaches the
// threshold and that box is still free, else counted as a false positive (neutral boxes: neither).  The only sequential
// dependency is the "still free" state of a ground-truth box, which is private to one (class, image) pair -- so:
//   eval_match_kernel   one warp per (class, image) segment of the predictions (sorted by class, image, confidence desc):
//                       lanes share the image's boxes of that class, float64 IoU with the reference's arithmetic (areas with
//                       the border-pixel term, intersection without: the iou() quirk), warp arg-max with np.argmax's
//                       first-index rule, flags written at the prediction's position in (class, confidence) order.
//   eval_cumsum_kernel  one CTA per class: inclusive scans of the two flag arrays (np.cumsum, :726-727).
// Sorting (two stable key sorts) is done by the caller; precision / recall / AP are a few vector operations on the cumulative
// counts (host, same NumPy expressions as the reference).
#include "common.cuh"
#include <climits>
#include <cmath>

using namespace ssdk;

namespace {

struct EvalArgs {
  const int* seg_offsets;      // [n_seg+1] into the (class, image, confidence desc) order
  const int* pred_image;       // [n]
  const int* pred_class;       // [n]
  const float* pred_box;       // [n*4] corners
  const int* pred_rank;        // [n] position in (class, confidence desc) order
  const double* gt_rows;       // [sum G * 5] class, xmin, ymin, xmax, ymax
  const int* gt_offsets;       // [n_images+1]
  const unsigned char* gt_neutral;   // [sum G] or NULL
  unsigned char* gt_matched;   // [sum G], zero on entry
  double thr; int d;
  int* tp; int* fp;            // [n], zero on entry
};

__global__ void __launch_bounds__(256) eval_match_kernel(EvalArgs a, int n_seg) {
  const int seg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (seg >= n_seg) return;
  const int p0 = a.seg_offsets[seg], p1 = a.seg_offsets[seg + 1];
  if (p1 <= p0) return;
  const int img = a.pred_image[p0], cls = a.pred_class[p0];
  const int g0 = a.gt_offsets[img], g1 = a.gt_offsets[img + 1];
  for (int p = p0; p < p1; ++p) {                       // descending confidence inside the segment
    const double px0 = a.pred_box[p * 4], py0 = a.pred_box[p * 4 + 1], px1 = a.pred_box[p * 4 + 2], py1 = a.pred_box[p * 4 + 3];
    const double parea = __dmul_rn(__dadd_rn(__dsub_rn(px1, px0), (double)a.d), __dadd_rn(__dsub_rn(py1, py0), (double)a.d));
    // best overlap among the boxes of this class (np.argmax: first maximum; a NaN counts as the maximum)
    double best = -1.0; int best_g = INT_MAX; int best_nan = 0, any = 0;
    for (int g = g0 + lane; g < g1; g += 32) {
      const double* r = a.gt_rows + (size_t)g * 5;
      if ((int)r[0] != cls) continue;
      any = 1;
      const double iw = fmax(0.0, __dsub_rn(fmin(r[3], px1), fmax(r[1], px0)));
      const double ih = fmax(0.0, __dsub_rn(fmin(r[4], py1), fmax(r[2], py0)));
      const double inter = __dmul_rn(iw, ih);
      const double garea = __dmul_rn(__dadd_rn(__dsub_rn(r[3], r[1]), (double)a.d), __dadd_rn(__dsub_rn(r[4], r[2]), (double)a.d));
      const double v = __ddiv_rn(inter, __dsub_rn(__dadd_rn(garea, parea), inter));
      const int is_nan = (v != v) ? 1 : 0;
      if (best_nan) continue;                             // an earlier NaN of this lane already wins
      if (is_nan || v > best) { best = v; best_g = g; best_nan = is_nan; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int og = __shfl_xor_sync(0xffffffffu, best_g, o);
      const int on = __shfl_xor_sync(0xffffffffu, best_nan, o);
      any |= __shfl_xor_sync(0xffffffffu, any, o);
      bool take;
      if (on != best_nan) take = on > best_nan;           // NaN beats every number
      else if (on) take = og < best_g;                    // two NaNs: the first one
      else take = (ov > best) || (ov == best && og < best_g);
      if (take) { best = ov; best_g = og; best_nan = on; }
    }
    if (lane == 0) {
      const int r = a.pred_rank[p];
      if (!any) a.fp[r] = 1;                              // no box of this class in the image (:672-675)
      else if (best < a.thr) a.fp[r] = 1;                 // (NaN < thr is False, as in the reference)
      else if (!(a.gt_neutral && a.gt_neutral[best_g])) {
        if (!a.gt_matched[best_g]) { a.tp[r] = 1; a.gt_matched[best_g] = 1; }
        else a.fp[r] = 1;                                 // the box was matched by a more confident prediction
      }
    }
    __syncwarp();                                         // lane 0's gt_matched write is only ever read by lane 0
  }
}

__global__ void __launch_bounds__(1024) eval_cumsum_kernel(const int* __restrict__ tp, const int* __restrict__ fp, const int* __restrict__ class_offsets,
                                                           int* __restrict__ ctp, int* __restrict__ cfp) {
  __shared__ int s_t[32], s_f[32];
  __shared__ int s_base[2];
  const int c = blockIdx.x;
  const int lo = class_offsets[c], hi = class_offsets[c + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_base[0] = 0; s_base[1] = 0; }
  __syncthreads();
  for (int base = lo; base < hi; base += 1024) {
    const int i = base + threadIdx.x;
    int t = i < hi ? tp[i] : 0, f = i < hi ? fp[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int vt = __shfl_up_sync(0xffffffffu, t, o), vf = __shfl_up_sync(0xffffffffu, f, o);
      if (lane >= o) { t += vt; f += vf; }
    }
    if (lane == 31) { s_t[warp] = t; s_f[warp] = f; }
    __syncthreads();
    int bt = s_base[0], bf = s_base[1];
    for (int w = 0; w < warp; ++w) { bt += s_t[w]; bf += s_f[w]; }
    if (i < hi) { ctp[i] = bt + t; cfp[i] = bf + f; }
    __syncthreads();
    if (threadIdx.x == 1023) { s_base[0] = bt + t; s_base[1] = bf + f; }
    __syncthreads();
  }
}

}  // namespace

extern "C" int ssdk_eval_match(ssdk_ctx* ctx, int n_pred, const int* seg_offsets_dev, int n_seg, const int* pred_image_dev,
                               const int* pred_class_dev, const float* pred_box_dev, const int* pred_rank_dev, const double* gt_rows_dev,
                               const int* gt_offsets_dev, const unsigned char* gt_neutral_dev, unsigned char* gt_matched_dev,
                               double matching_iou_threshold, int border_d, int* tp_dev, int* fp_dev, void* stream_) {
  SSDK_REQUIRE(ctx && n_pred >= 0 && n_seg >= 0, "ssdk_eval_match: bad argument");
  if (n_pred == 0 || n_seg == 0) return SSDK_OK;
  SSDK_REQUIRE(seg_offsets_dev && pred_image_dev && pred_class_dev && pred_box_dev && pred_rank_dev && gt_rows_dev && gt_offsets_dev &&
               gt_matched_dev && tp_dev && fp_dev, "ssdk_eval_match: NULL argument");
  EvalArgs a{seg_offsets_dev, pred_image_dev, pred_class_dev, pred_box_dev, pred_rank_dev, gt_rows_dev, gt_offsets_dev, gt_neutral_dev,
             gt_matched_dev, matching_iou_threshold, border_d, tp_dev, fp_dev};
  eval_match_kernel<<<ceil_div(n_seg, 8), 256, 0, (cudaStream_t)stream_>>>(a, n_seg);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

extern "C" int ssdk_eval_cumsum(ssdk_ctx* ctx, const int* tp_dev, const int* fp_dev, const int* class_offsets_dev, int n_segments,
                                int* ctp_dev, int* cfp_dev, void* stream_) {
  SSDK_REQUIRE(ctx && tp_dev && fp_dev && class_offsets_dev && ctp_dev && cfp_dev && n_segments > 0, "ssdk_eval_cumsum: bad argument");
  eval_cumsum_kernel<<<n_segments, 1024, 0, (cudaStream_t)stream_>>>(tp_dev, fp_dev, class_offsets_dev, ctp_dev, cfp_dev);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}
