/* TEST / BASELINE INFRASTRUCTURE (oracle): a C restatement of the greedy non-maximum suppression that
 * keras_layers/keras_layer_DecodeDetections.py:195-199 and keras_layer_DecodeDetectionsFast.py:199-203 delegate to
 * `tf.image.non_max_suppression` (third-party: TensorFlow 1.x, version unpinned by the reference; its CPU kernel is compiled C++,
 * which is why the CPU baseline of bench.py times this file and not a Python loop).  Same arithmetic as oracle/decoder.py:tf_nms /
 * tf_nms_fast, against which tests/test_oracle_c_nms_cpu.py pins it bit for bit: float32 throughout, std::min / std::max argument
 * order (NaN behaviour), boxes canonicalised per axis, IoU of a zero-area box = 0, suppression when IoU > threshold, descending
 * score with ties to the lower index, stop at max_output_size.  Never linked into the product library.
 *
 *   gcc -O3 -ffp-contract=off -shared -fPIC -o oracle/_build/libtfnms.so oracle/tf_nms.c        (oracle/cbuild.py)            */
#include <stdlib.h>

static inline float smin(float a, float b) { return (b < a) ? b : a; }   /* std::min<float>(a, b) */
static inline float smax(float a, float b) { return (a < b) ? b : a; }   /* std::max<float>(a, b) */

typedef struct { float s; int i; } key_t_;
static int cmp_key(const void* pa, const void* pb) {
  const key_t_* a = (const key_t_*)pa; const key_t_* b = (const key_t_*)pb;
  if (a->s > b->s) return -1;
  if (a->s < b->s) return 1;
  return (a->i > b->i) - (a->i < b->i);
}

/* boxes: n x 4 float32 (x0, y0, x1, y1) in any corner order; scores: n float32 (no NaN: the caller filters by `conf > thresh`).
 * Writes the selected indices in selection order to out_idx (room for max_output_size) and returns their number, or -1.
 * The candidates still alive are kept compacted in score order (structure of arrays), so that the pass after each selection is a
 * branch-free loop the compiler vectorises -- the result does not depend on that (every candidate meets the same selected boxes,
 * and a candidate is dropped by the FIRST selected box whose IoU exceeds the threshold either way). */
int tf_nms_f32(const float* boxes, const float* scores, int n, int max_output_size, float iou_threshold, int* out_idx) {
  if (n <= 0 || max_output_size <= 0) return 0;
  key_t_* order = (key_t_*)malloc((size_t)n * sizeof(key_t_));
  float* g = (float*)malloc((size_t)n * 5 * sizeof(float));        /* x0 | x1 | y0 | y1 | area, each n long, in score order */
  int* id = (int*)malloc((size_t)n * sizeof(int));
  if (!order || !g || !id) { free(order); free(g); free(id); return -1; }
  for (int i = 0; i < n; ++i) { order[i].s = scores[i]; order[i].i = i; }
  qsort(order, (size_t)n, sizeof(key_t_), cmp_key);
  float *X0 = g, *X1 = g + n, *Y0 = g + 2 * (size_t)n, *Y1 = g + 3 * (size_t)n, *A = g + 4 * (size_t)n;
  for (int r = 0; r < n; ++r) {
    const float* b = boxes + 4 * (size_t)order[r].i;
    const float x0 = smin(b[0], b[2]), x1 = smax(b[0], b[2]), y0 = smin(b[1], b[3]), y1 = smax(b[1], b[3]);
    const float dy = y1 - y0, dx = x1 - x0;
    X0[r] = x0; X1[r] = x1; Y0[r] = y0; Y1[r] = y1; A[r] = dy * dx;
    id[r] = order[r].i;
  }
  int count = 0, m = n, head = 0;                                   /* alive candidates: positions head .. m-1 */
  while (head < m) {
    const float sx0 = X0[head], sx1 = X1[head], sy0 = Y0[head], sy1 = Y1[head], sa = A[head];
    out_idx[count++] = id[head];
    ++head;
    if (count >= max_output_size) break;
    int w = head;
    for (int t = head; t < m; ++t) {
      const float ih = smax(smin(Y1[t], sy1) - smax(Y0[t], sy0), 0.0f);
      const float iw = smax(smin(X1[t], sx1) - smax(X0[t], sx0), 0.0f);
      const float inter = ih * iw;
      const float sum = A[t] + sa;
      const float den = sum - inter;
      float v = inter / den;
      if (A[t] <= 0.0f || sa <= 0.0f) v = 0.0f;
      const int keep = !(v > iou_threshold);
      X0[w] = X0[t]; X1[w] = X1[t]; Y0[w] = Y0[t]; Y1[w] = Y1[t]; A[w] = A[t]; id[w] = id[t];
      w += keep;
    }
    m = w;
  }
  free(order); free(g); free(id);
  return count;
}
