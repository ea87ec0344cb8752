// Batch assembly for the ground-truth encoder: the box half of the reference's augmentation chain on the device, and the
// packing of the surviving boxes into the encoder's ragged (sum G_i x 5, offsets) format.
// Reference: the label arithmetic of CropPad (data_generator/object_detection_2d_patch_sampling_ops.py:258-330, used by
// SSDExpand / SSDRandomCrop through RandomPatch), Flip (object_detection_2d_geometric_ops.py:171-195), Resize (:61-100),
// BoxFilter (object_detection_2d_image_boxes_validation_utils.py:120-200), and the degenerate-box handling + hand-off to the
// label encoder in DataGenerator.generate (object_detection_2d_data_generator.py:1095-1151).
// The random decisions of the chain (which patch, flip or not) are host-side control flow in the reference and stay with the
// caller: this file takes the decided parameters as a per-image list of box operations.
//   box_ops_kernel     one CTA per image: every box through the image's operation list in float64 (what NumPy computes on
//                      int / float64 label arrays; np.round = round-half-even = rint), a validity flag per box, ordered
//                      compaction of the survivors.
//   box_offsets_kernel one CTA: exclusive scan of the survivor counts -> row offsets, total and maximum.
//   box_pack_kernel    one CTA per image: rows to their final place.
#include "common.cuh"
#include <cmath>

using namespace ssdk;

namespace {

constexpr int kBoxThreads = 128;

__global__ void __launch_bounds__(kBoxThreads) box_ops_kernel(const void* __restrict__ gt, int gt_f64, const int* __restrict__ offs, const ssdk_box_op* __restrict__ ops,
                                                              int max_ops, float* __restrict__ tmp, int* __restrict__ counts) {
  __shared__ int s_w[kBoxThreads / 32];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const int g0 = offs[b], G = offs[b + 1] - g0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < G; base += kBoxThreads) {
    const int g = base + threadIdx.x;
    bool valid = g < G;
    double cls = 0, x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (valid) {
      if (gt_f64) { const double* r = reinterpret_cast<const double*>(gt) + (size_t)(g0 + g) * 5; cls = r[0]; x0 = r[1]; y0 = r[2]; x1 = r[3]; y1 = r[4]; }
      else { const float* r = reinterpret_cast<const float*>(gt) + (size_t)(g0 + g) * 5; cls = r[0]; x0 = r[1]; y0 = r[2]; x1 = r[3]; y1 = r[4]; }
      for (int i = 0; i < max_ops; ++i) {
        const ssdk_box_op op = ops[(size_t)b * max_ops + i];
        if (op.op == SSDK_BOXOP_END) break;
        if (op.op == SSDK_BOXOP_CROP_PAD) {
          // labels -= patch origin; BoxFilter 'center_point' against the patch; clip to the patch (CropPad.__call__)
          y0 -= op.a0; y1 -= op.a0; x0 -= op.a1; x1 -= op.a1;
          if (op.flags & 1) {
            const double cy = (y0 + y1) / 2, cx = (x0 + x1) / 2;
            valid = valid && (cy >= 0.0) && (cy <= op.a2 - 1) && (cx >= 0.0) && (cx <= op.a3 - 1);
          }
          if (op.flags & 2) {
            y0 = fmin(fmax(y0, 0.0), op.a2 - 1); y1 = fmin(fmax(y1, 0.0), op.a2 - 1);
            x0 = fmin(fmax(x0, 0.0), op.a3 - 1); x1 = fmin(fmax(x1, 0.0), op.a3 - 1);
          }
        } else if (op.op == SSDK_BOXOP_FLIP_H) {             // labels[:, [xmin, xmax]] = img_width - labels[:, [xmax, xmin]]
          const double nx0 = op.a0 - x1, nx1 = op.a0 - x0; x0 = nx0; x1 = nx1;
        } else if (op.op == SSDK_BOXOP_FLIP_V) {
          const double ny0 = op.a0 - y1, ny1 = op.a0 - y0; y0 = ny0; y1 = ny1;
        } else if (op.op == SSDK_BOXOP_RESIZE) {             // np.round(labels * (out / in), decimals=0)
          const double sy = op.a2 / op.a0, sx = op.a3 / op.a1;
          y0 = rint(y0 * sy); y1 = rint(y1 * sy); x0 = rint(x0 * sx); x1 = rint(x1 * sx);
          if (op.flags & 1) valid = valid && (x1 > x0) && (y1 > y0);
        } else if (op.op == SSDK_BOXOP_FILTER) {             // BoxFilter: degenerate and / or minimum area
          if (op.flags & 1) valid = valid && (x1 > x0) && (y1 > y0);
          if (op.flags & 2) valid = valid && ((x1 - x0) * (y1 - y0) >= op.a0);
        }
      }
    }
    // ordered compaction (box order is kept, like labels[requirements_met])
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    int before = s_base, total = 0;
    for (int w = 0; w < kBoxThreads / 32; ++w) { if (w < warp) before += s_w[w]; total += s_w[w]; }
    if (valid) {
      float* o = tmp + (size_t)(g0 + before + __popc(m & ((1u << lane) - 1))) * 5;
      o[0] = (float)cls; o[1] = (float)x0; o[2] = (float)y0; o[3] = (float)x1; o[4] = (float)y1;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = s_base;
}

__global__ void __launch_bounds__(1024) box_offsets_kernel(const int* __restrict__ counts, int B, int* __restrict__ offs_out, int* __restrict__ stats) {
  __shared__ int s_sum[1024];
  __shared__ int s_max[1024];
  const int t = threadIdx.x;
  const int per = (B + 1023) / 1024;
  const int lo = t * per, hi = min(B, lo + per);
  int sum = 0, mx = 0;
  for (int i = lo; i < hi; ++i) { sum += counts[i]; mx = max(mx, counts[i]); }
  s_sum[t] = sum; s_max[t] = mx;
  __syncthreads();
  int base = 0;
  for (int i = 0; i < t; ++i) base += s_sum[i];
  for (int i = lo; i < hi; ++i) { offs_out[i] = base; base += counts[i]; }
  if (t == 1023) {
    int m = 0;
    for (int i = 0; i < 1024; ++i) m = max(m, s_max[i]);
    offs_out[B] = base;
    if (stats) { stats[0] = base; stats[1] = m; }
  }
}

__global__ void __launch_bounds__(kBoxThreads) box_pack_kernel(const float* __restrict__ tmp, const int* __restrict__ offs_in, const int* __restrict__ offs_out,
                                                               float* __restrict__ out) {
  const int b = blockIdx.x;
  const int n = (offs_out[b + 1] - offs_out[b]) * 5;
  const float* src = tmp + (size_t)offs_in[b] * 5;
  float* dst = out + (size_t)offs_out[b] * 5;
  for (int i = threadIdx.x; i < n; i += kBoxThreads) dst[i] = src[i];
}

}  // namespace

extern "C" int ssdk_assemble_batch(ssdk_ctx* ctx, const void* gt_in_dev, int gt_in_f64, const int* offsets_in_dev, int B, int total_in,
                                   const ssdk_box_op* ops_dev, int max_ops, float* gt_out_dev, int* offsets_out_dev, int* out_stats_dev,
                                   void* stream_) {
  SSDK_REQUIRE(ctx && offsets_in_dev && gt_out_dev && offsets_out_dev && B > 0 && total_in >= 0, "ssdk_assemble_batch: bad argument");
  SSDK_REQUIRE(total_in == 0 || gt_in_dev, "ssdk_assemble_batch: gt_in_dev is NULL");
  SSDK_REQUIRE(max_ops == 0 || ops_dev, "ssdk_assemble_batch: ops_dev is NULL");
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t need = (size_t)(total_in > 0 ? total_in : 1) * 5 * sizeof(float) + (size_t)B * sizeof(int) + 256;
  int rc = ctx->ws[3].ensure(need);
  if (rc) return rc;
  float* tmp = reinterpret_cast<float*>(ctx->ws[3].ptr);
  int* counts = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ctx->ws[3].ptr) + (((size_t)(total_in > 0 ? total_in : 1) * 5 * sizeof(float) + 255) / 256 * 256));
  box_ops_kernel<<<B, kBoxThreads, 0, stream>>>(gt_in_dev, gt_in_f64, offsets_in_dev, ops_dev, max_ops, tmp, counts);
  SSDK_COUNT_LAUNCH(ctx);
  box_offsets_kernel<<<1, 1024, 0, stream>>>(counts, B, offsets_out_dev, out_stats_dev);
  SSDK_COUNT_LAUNCH(ctx);
  box_pack_kernel<<<B, kBoxThreads, 0, stream>>>(tmp, offsets_in_dev, offsets_out_dev, gt_out_dev);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}
