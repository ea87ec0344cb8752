// SSDInputEncoder hot path on sm_100a: pairwise IoU (float64), greedy bipartite + multi matching,
// neutral boxes and offset encoding.  Reference: ssd_encoder_decoder/ssd_input_encoder.py:277-418,
// bounding_box_utils/bounding_box_utils.py:283-383, ssd_encoder_decoder/matching_utils.py:22-116.
//
// Kernels (all HBM/ALU-bound integer/float64 work, no tensor cores):
//   enc_rowmax_kernel    one CTA per ground-truth box: row max / first argmax over all anchors
//                        (tile bounding boxes prune anchors that cannot intersect).
//   enc_bipartite_kernel one CTA per image: the G sequential greedy rounds, with the reference's
//                        zeroed-row quirk; rows whose best anchor was taken are re-scanned.
//   enc_main_kernel      one CTA per (anchor tile, image): per-anchor best gt over the pruned
//                        candidate list, multi-match / neutral decision, offset encode, and a
//                        coalesced write of the (C+12)-float target rows staged through shared memory.
// Exactness: every decision-relevant value is computed with the reference's float64 operation
// order using non-contracting intrinsics (__dmul_rn/__dadd_rn/...).  The running arg-max compares
// IoUs by cross multiplication (inter1*union2 > inter2*union1) and divides once at the end.
#include "common.cuh"
#include <climits>
#include <cmath>

using namespace ssdk;

namespace {

constexpr int kTile = 256;      // anchors per CTA tile == threads per CTA

struct EncParams {
  const double* anchors;        // [P*4] template coords (format = coords)
  const double* tile_bbox;      // [n_tiles*4] corner bbox of each anchor tile
  int P, n_tiles, C, bg, coords, multi, d, normalize;
  double pos_thr, neg_lim, img_w, img_h;
  double var[4];
};

struct Box { double x0, y0, x1, y1, area; };

__device__ __forceinline__ Box corners_from_template(const double t[4], int coords, int d) {
  Box b;
  if (coords == SSDK_COORDS_CENTROIDS) {            // convert_coordinates 'centroids2corners' (:76-80)
    double hw = __ddiv_rn(t[2], 2.0), hh = __ddiv_rn(t[3], 2.0);
    b.x0 = __dsub_rn(t[0], hw); b.y0 = __dsub_rn(t[1], hh);
    b.x1 = __dadd_rn(t[0], hw); b.y1 = __dadd_rn(t[1], hh);
  } else if (coords == SSDK_COORDS_MINMAX) {
    b.x0 = t[0]; b.x1 = t[1]; b.y0 = t[2]; b.y1 = t[3];
  } else {
    b.x0 = t[0]; b.y0 = t[1]; b.x1 = t[2]; b.y1 = t[3];
  }
  // area uses d, the intersection never does (reference quirk, bounding_box_utils.py:345,373-374)
  b.area = __dmul_rn(__dadd_rn(__dsub_rn(b.x1, b.x0), (double)d), __dadd_rn(__dsub_rn(b.y1, b.y0), (double)d));
  return b;
}

// Ground-truth row (class,xmin,ymin,xmax,ymax) float32 pixels -> template coords in `coords` format
// (ssd_input_encoder.py:330-347).  Returns false for a degenerate box (:333).
__device__ __forceinline__ bool gt_template(const float* row, const EncParams& p, double t[4], int& cls) {
  double xmin = (double)row[1], ymin = (double)row[2], xmax = (double)row[3], ymax = (double)row[4];
  cls = (int)row[0];
  bool ok = (__dsub_rn(xmax, xmin) > 0.0) && (__dsub_rn(ymax, ymin) > 0.0);
  if (p.normalize) {
    ymin = __ddiv_rn(ymin, p.img_h); ymax = __ddiv_rn(ymax, p.img_h);
    xmin = __ddiv_rn(xmin, p.img_w); xmax = __ddiv_rn(xmax, p.img_w);
  }
  if (p.coords == SSDK_COORDS_CENTROIDS) {          // 'corners2centroids' with border_pixels (:71-75)
    t[0] = __ddiv_rn(__dadd_rn(xmin, xmax), 2.0);
    t[1] = __ddiv_rn(__dadd_rn(ymin, ymax), 2.0);
    t[2] = __dadd_rn(__dsub_rn(xmax, xmin), (double)p.d);
    t[3] = __dadd_rn(__dsub_rn(ymax, ymin), (double)p.d);
  } else if (p.coords == SSDK_COORDS_MINMAX) {
    t[0] = xmin; t[1] = xmax; t[2] = ymin; t[3] = ymax;
  } else {
    t[0] = xmin; t[1] = ymin; t[2] = xmax; t[3] = ymax;
  }
  return ok;
}

__device__ __forceinline__ double inter_area(const Box& a, const Box& b) {
  double iw = __dsub_rn(fmin(a.x1, b.x1), fmax(a.x0, b.x0));
  double ih = __dsub_rn(fmin(a.y1, b.y1), fmax(a.y0, b.y0));
  if (!(iw > 0.0) || !(ih > 0.0)) return 0.0;
  return __dmul_rn(iw, ih);
}
// union as NumPy forms it: (area_gt + area_anchor) - inter
__device__ __forceinline__ double union_area(const Box& g, const Box& a, double inter) {
  return __dsub_rn(__dadd_rn(g.area, a.area), inter);
}

__device__ __forceinline__ Box load_anchor(const EncParams& p, int a) {
  const double2* q = reinterpret_cast<const double2*>(p.anchors + (size_t)a * 4);
  double2 u = __ldg(q), v = __ldg(q + 1);
  double t[4] = {u.x, u.y, v.x, v.y};
  return corners_from_template(t, p.coords, p.d);
}

__device__ __forceinline__ bool bbox_hits(const double* bb, const Box& g) {
  return (bb[2] > g.x0) && (bb[0] < g.x1) && (bb[3] > g.y0) && (bb[1] < g.y1);
}

// ------------------------------------------------------------------------------------------
// IoU matrix (tests / microbench): out[g*P + a], bit-exact float64.
// ------------------------------------------------------------------------------------------
__global__ void iou_matrix_kernel(EncParams p, const float* __restrict__ gt, int G, double* __restrict__ out) {
  int a = blockIdx.x * blockDim.x + threadIdx.x;
  int g = blockIdx.y;
  if (a >= p.P || g >= G) return;
  double t[4]; int cls;
  gt_template(gt + (size_t)g * 5, p, t, cls);
  Box gb = corners_from_template(t, p.coords, p.d);
  Box ab = load_anchor(p, a);
  double inter = inter_area(gb, ab);
  out[(size_t)g * p.P + a] = __ddiv_rn(inter, union_area(gb, ab, inter));
}

__global__ void iou_general_kernel(const double* __restrict__ b1, int m, const double* __restrict__ b2, int n, int coords,
                                   int d, int elementwise, double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = elementwise ? (long long)(m > n ? m : n) : (long long)m * n;
  if (i >= total) return;
  int r1, r2;
  if (elementwise) { r1 = (m == 1) ? 0 : (int)i; r2 = (n == 1) ? 0 : (int)i; }
  else { r1 = (int)(i / n); r2 = (int)(i % n); }
  double t1[4] = {b1[r1 * 4], b1[r1 * 4 + 1], b1[r1 * 4 + 2], b1[r1 * 4 + 3]};
  double t2[4] = {b2[r2 * 4], b2[r2 * 4 + 1], b2[r2 * 4 + 2], b2[r2 * 4 + 3]};
  Box a = corners_from_template(t1, coords, d), b = corners_from_template(t2, coords, d);
  // np.maximum(0, ...) keeps a zero side at exactly 0, so inter is 0 (not negative) for disjoint boxes
  double iw = fmax(0.0, __dsub_rn(fmin(a.x1, b.x1), fmax(a.x0, b.x0)));
  double ih = fmax(0.0, __dsub_rn(fmin(a.y1, b.y1), fmax(a.y0, b.y0)));
  double inter = __dmul_rn(iw, ih);
  out[i] = __ddiv_rn(inter, __dsub_rn(__dadd_rn(a.area, b.area), inter));
}

// ------------------------------------------------------------------------------------------
// Block-wide (max value, min index) reduction, 256 threads.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_argmax(double& val, int& idx, double* s_val, int* s_idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, val, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) { s_val[w] = val; s_idx[w] = idx; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    double v = l < nw ? s_val[l] : -1.0;
    int i = l < nw ? s_idx[l] : INT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      double ov = __shfl_xor_sync(0xffffffffu, v, o);
      int oi = __shfl_xor_sync(0xffffffffu, i, o);
      if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
    if (l == 0) { s_val[0] = v; s_idx[0] = i; }
  }
  __syncthreads();
  val = s_val[0]; idx = s_idx[0];
  __syncthreads();
}

// Scan every anchor (tile-pruned) for ground-truth box `gb`; `removed` columns count as 0.
// Returns the row maximum and its first index over the whole block (0 / 0 for an all-zero row).
__device__ void row_scan(const EncParams& p, const Box& gb, const int* removed, int n_removed,
                         double& out_val, int& out_idx, double* s_val, int* s_idx) {
  double b_inter = 0.0, b_union = 1.0;
  int b_idx = INT_MAX;
  for (int t = 0; t < p.n_tiles; ++t) {
    if (!bbox_hits(p.tile_bbox + (size_t)t * 4, gb)) continue;      // block-uniform
    int a = t * kTile + threadIdx.x;
    if (a >= p.P) continue;
    Box ab = load_anchor(p, a);
    double inter = inter_area(gb, ab);
    if (inter > 0.0) {
      double un = union_area(gb, ab, inter);
      bool better = (p.d == 0) ? (__dmul_rn(inter, b_union) > __dmul_rn(b_inter, un))
                               : (__ddiv_rn(inter, un) > __ddiv_rn(b_inter, b_union));
      if (better) {
        bool gone = false;
        for (int r = 0; r < n_removed; ++r) gone |= (removed[r] == a);
        if (!gone) { b_inter = inter; b_union = un; b_idx = a; }
      }
    }
  }
  double val = (b_idx != INT_MAX) ? __ddiv_rn(b_inter, b_union) : 0.0;
  if (!(val > 0.0)) { val = 0.0; b_idx = INT_MAX; }
  block_argmax(val, b_idx, s_val, s_idx);
  out_val = val;
  out_idx = (val > 0.0) ? b_idx : 0;       // argmax of an all-zero row is index 0
}

// One CTA per ground-truth box.
__global__ void __launch_bounds__(kTile) enc_rowmax_kernel(EncParams p, const float* __restrict__ gt,
                                                           const int* __restrict__ gt_offsets, int B,
                                                           double* __restrict__ rowmax, int* __restrict__ rowarg,
                                                           int* __restrict__ status) {
  __shared__ double s_val[8];
  __shared__ int s_idx[8];
  int g = blockIdx.x;
  double t[4]; int cls;
  bool ok = gt_template(gt + (size_t)g * 5, p, t, cls);
  if (!ok && threadIdx.x == 0 && status) {
    int lo = 0, hi = B;                       // image of this gt row: largest b with offsets[b] <= g
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (gt_offsets[mid] <= g) lo = mid; else hi = mid; }
    atomicMax(status, lo + 1);
  }
  Box gb = corners_from_template(t, p.coords, p.d);
  double val; int idx;
  row_scan(p, gb, nullptr, 0, val, idx, s_val, s_idx);
  if (threadIdx.x == 0) { rowmax[g] = val; rowarg[g] = idx; }
}

// One CTA per image: match_bipartite_greedy (matching_utils.py:63-77).
__global__ void __launch_bounds__(kTile) enc_bipartite_kernel(EncParams p, const float* __restrict__ gt,
                                                              const int* __restrict__ gt_offsets,
                                                              const double* __restrict__ rowmax_g,
                                                              const int* __restrict__ rowarg_g,
                                                              int* __restrict__ matches) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ double s_val[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x;
  const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
  if (G <= 0) return;
  double* rv = reinterpret_cast<double*>(smem_raw);          // [G]
  int* ra = reinterpret_cast<int*>(rv + G);                   // [G]
  int* removed = ra + G;                                      // [G]
  for (int g = threadIdx.x; g < G; g += blockDim.x) { rv[g] = rowmax_g[g0 + g]; ra[g] = rowarg_g[g0 + g]; matches[g0 + g] = 0; }
  __syncthreads();
  int n_removed = 0;
  for (int round = 0; round < G; ++round) {
    double v = -1.0; int gi = INT_MAX;
    for (int g = threadIdx.x; g < G; g += blockDim.x)
      if (rv[g] > v) { v = rv[g]; gi = g; }                   // ascending g per thread: first index kept
    block_argmax(v, gi, s_val, s_idx);
    const int a_star = ra[gi];
    __syncthreads();
    if (threadIdx.x == 0) { matches[g0 + gi] = a_star; rv[gi] = 0.0; ra[gi] = 0; removed[n_removed] = a_star; }
    ++n_removed;
    __syncthreads();
    // rows whose recorded best anchor just disappeared must be re-scanned
    for (int g = 0; g < G; ++g) {
      if (rv[g] > 0.0 && ra[g] == a_star) {                   // block-uniform condition (shared memory)
        double t[4]; int cls;
        gt_template(gt + (size_t)(g0 + g) * 5, p, t, cls);
        Box gb = corners_from_template(t, p.coords, p.d);
        double nv; int ni;
        row_scan(p, gb, removed, n_removed, nv, ni, s_val, s_idx);
        if (threadIdx.x == 0) { rv[g] = nv; ra[g] = ni; }
        __syncthreads();
      }
    }
  }
}

// One CTA per (anchor tile, image).
__global__ void __launch_bounds__(kTile) enc_main_kernel(EncParams p, const float* __restrict__ gt,
                                                         const int* __restrict__ gt_offsets,
                                                         const int* __restrict__ matches,
                                                         float* __restrict__ out_y, int* __restrict__ out_match) {
  extern __shared__ unsigned char smem_raw[];
  const int tile = blockIdx.x, b = blockIdx.y;
  const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
  const int W = p.C + 12;
  // shared layout: rows[kTile*W] floats | cand boxes [G] x (x0,y0,x1,y1,area) doubles | cand idx [G] | owner[kTile]
  float* rows = reinterpret_cast<float*>(smem_raw);
  size_t off = ((size_t)kTile * W * sizeof(float) + 15) & ~(size_t)15;
  double* cb = reinterpret_cast<double*>(smem_raw + off);     // 5*G doubles
  int* cidx = reinterpret_cast<int*>(cb + 5 * (size_t)(G > 0 ? G : 1));
  int* owner = cidx + (G > 0 ? G : 1);
  __shared__ int s_ncand;

  const int a0 = tile * kTile;
  owner[threadIdx.x] = -1;
  if (threadIdx.x == 0) s_ncand = 0;
  __syncthreads();
  // bipartite owners: the highest gt index wins on duplicates (NumPy fancy assignment, :363)
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    int m = matches[g0 + g] - a0;
    if (m >= 0 && m < kTile) atomicMax(&owner[m], g);
  }
  // ordered candidate list (ascending gt index) built by warp 0
  if (threadIdx.x < 32) {
    const double* bb = p.tile_bbox + (size_t)tile * 4;
    int n = 0;
    for (int base = 0; base < G; base += 32) {
      int g = base + threadIdx.x;
      bool hit = false; Box gb{};
      if (g < G) {
        double t[4]; int cls;
        gt_template(gt + (size_t)(g0 + g) * 5, p, t, cls);
        gb = corners_from_template(t, p.coords, p.d);
        hit = bbox_hits(bb, gb);
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        int pos = n + __popc(m & ((1u << threadIdx.x) - 1));
        cb[pos * 5 + 0] = gb.x0; cb[pos * 5 + 1] = gb.y0; cb[pos * 5 + 2] = gb.x1; cb[pos * 5 + 3] = gb.y1;
        cb[pos * 5 + 4] = gb.area; cidx[pos] = g;
      }
      n += __popc(m);
    }
    if (threadIdx.x == 0) s_ncand = n;
  }
  __syncthreads();
  const int ncand = s_ncand;
  const int a = a0 + threadIdx.x;
  float* my = rows + (size_t)threadIdx.x * W;
  if (a < p.P) {
    const double2* q = reinterpret_cast<const double2*>(p.anchors + (size_t)a * 4);
    double2 u = __ldg(q), v = __ldg(q + 1);
    double at[4] = {u.x, u.y, v.x, v.y};
    Box ab = corners_from_template(at, p.coords, p.d);
    int match_g = owner[threadIdx.x];
    bool neutral = false;
    if (match_g < 0) {
      double b_inter = 0.0, b_union = 1.0; int b_g = -1;
      for (int c = 0; c < ncand; ++c) {
        Box gb; gb.x0 = cb[c * 5]; gb.y0 = cb[c * 5 + 1]; gb.x1 = cb[c * 5 + 2]; gb.y1 = cb[c * 5 + 3]; gb.area = cb[c * 5 + 4];
        double inter = inter_area(gb, ab);
        if (inter > 0.0) {
          double un = union_area(gb, ab, inter);
          bool better = (p.d == 0) ? (__dmul_rn(inter, b_union) > __dmul_rn(b_inter, un))
                                   : (__ddiv_rn(inter, un) > __ddiv_rn(b_inter, b_union));
          if (better) { b_inter = inter; b_union = un; b_g = cidx[c]; }
        }
      }
      double val = (b_g >= 0) ? __ddiv_rn(b_inter, b_union) : 0.0;
      int arg = (b_g >= 0 && val > 0.0) ? b_g : 0;           // np.argmax of an all-zero column is 0
      if (!(val > 0.0)) val = 0.0;
      if (p.multi && G > 0 && val >= p.pos_thr) { match_g = arg; val = 0.0; }     // column zeroed after matching (:381)
      if (G > 0 && val >= p.neg_lim) neutral = true;                               // :388-390
    } else {
      if (0.0 >= p.neg_lim) neutral = true;                   // matched column is all zero
    }
    // ---- fill the row ----
    for (int c = 0; c < p.C; ++c) my[c] = 0.f;
    float o4[4] = {0.f, 0.f, 0.f, 0.f};
    if (match_g >= 0) {
      double gtc[4]; int cls;
      gt_template(gt + (size_t)(g0 + match_g) * 5, p, gtc, cls);
      if (cls >= 0 && cls < p.C) my[cls] = 1.f;
      if (p.coords == SSDK_COORDS_CENTROIDS) {                // :396-400
        o4[0] = (float)__ddiv_rn(__dsub_rn(gtc[0], at[0]), __dmul_rn(at[2], p.var[0]));
        o4[1] = (float)__ddiv_rn(__dsub_rn(gtc[1], at[1]), __dmul_rn(at[3], p.var[1]));
        o4[2] = (float)__ddiv_rn(log(__ddiv_rn(gtc[2], at[2])), p.var[2]);
        o4[3] = (float)__ddiv_rn(log(__ddiv_rn(gtc[3], at[3])), p.var[3]);
      } else if (p.coords == SSDK_COORDS_CORNERS) {           // :401-405
        double w = __dsub_rn(at[2], at[0]), h = __dsub_rn(at[3], at[1]);
        o4[0] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[0], at[0]), w), p.var[0]);
        o4[1] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[1], at[1]), h), p.var[1]);
        o4[2] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[2], at[2]), w), p.var[2]);
        o4[3] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[3], at[3]), h), p.var[3]);
      } else {                                                // minmax :406-410
        double w = __dsub_rn(at[1], at[0]), h = __dsub_rn(at[3], at[2]);
        o4[0] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[0], at[0]), w), p.var[0]);
        o4[1] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[1], at[1]), w), p.var[1]);
        o4[2] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[2], at[2]), h), p.var[2]);
        o4[3] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[3], at[3]), h), p.var[3]);
      }
      if (neutral) my[p.bg] = 0.f;                            // neg_iou_limit <= 0 corner case
    } else {
      my[p.bg] = neutral ? 0.f : 1.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      my[p.C + k] = o4[k];
      my[p.C + 4 + k] = (float)at[k];
      my[p.C + 8 + k] = (float)p.var[k];
    }
    if (out_match) out_match[(size_t)b * p.P + a] = (match_g >= 0) ? match_g : (neutral ? -2 : -1);
  }
  __syncthreads();
  // coalesced copy of the staged rows
  const int n_rows = min(kTile, p.P - a0);
  const size_t n_f = (size_t)n_rows * W;
  float* dst = out_y + ((size_t)b * p.P + a0) * W;
  for (size_t i = threadIdx.x; i < n_f; i += blockDim.x) dst[i] = rows[i];
}

__global__ void anchor_tile_bbox_kernel(EncParams p, double* __restrict__ bbox) {
  __shared__ double s[4][8];
  int tile = blockIdx.x;
  int a = tile * kTile + threadIdx.x;
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  if (a < p.P) { Box ab = load_anchor(p, a); x0 = ab.x0; y0 = ab.y0; x1 = ab.x1; y1 = ab.y1; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x0 = fmin(x0, __shfl_xor_sync(0xffffffffu, x0, o)); y0 = fmin(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    x1 = fmax(x1, __shfl_xor_sync(0xffffffffu, x1, o)); y1 = fmax(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s[0][w] = x0; s[1][w] = y0; s[2][w] = x1; s[3][w] = y1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { x0 = fmin(x0, s[0][i]); y0 = fmin(y0, s[1][i]); x1 = fmax(x1, s[2][i]); y1 = fmax(y1, s[3][i]); }
    bbox[tile * 4 + 0] = x0; bbox[tile * 4 + 1] = y0; bbox[tile * 4 + 2] = x1; bbox[tile * 4 + 3] = y1;
  }
}

}  // namespace

struct ssdk_encoder {
  ssdk_ctx* ctx = nullptr;
  ssdk_encode_cfg cfg{};
  EncParams p{};
  double* d_anchors = nullptr;
  double* d_bbox = nullptr;
  Scratch rows;        // rowmax (double) + rowarg (int) + matches (int) + offsets (int)
  int* h_offsets = nullptr;   // pinned staging
  int h_offsets_cap = 0;
};

extern "C" int ssdk_encoder_create(ssdk_ctx* ctx, const ssdk_encode_cfg* cfg, const double* anchors_host, ssdk_encoder** out) {
  SSDK_REQUIRE(ctx && cfg && anchors_host && out, "ssdk_encoder_create: NULL argument");
  SSDK_REQUIRE(cfg->P > 0 && cfg->n_classes_total > 1, "ssdk_encoder_create: bad P / n_classes");
  SSDK_REQUIRE(cfg->coords >= 0 && cfg->coords <= 2, "Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.");
  SSDK_REQUIRE(cfg->background_id >= 0 && cfg->background_id < cfg->n_classes_total, "background_id out of range");
  for (int i = 0; i < 4; ++i) SSDK_REQUIRE(cfg->variances[i] > 0, "All variances must be >0");
  SSDK_CHECK_CUDA(cudaSetDevice(ctx->device));
  ssdk_encoder* e = new ssdk_encoder();
  e->ctx = ctx; e->cfg = *cfg;
  EncParams& p = e->p;
  p.P = cfg->P; p.n_tiles = ceil_div(cfg->P, kTile); p.C = cfg->n_classes_total; p.bg = cfg->background_id;
  p.coords = cfg->coords; p.multi = cfg->matching_multi; p.d = cfg->border_d; p.normalize = cfg->normalize_coords;
  p.pos_thr = cfg->pos_iou_threshold; p.neg_lim = cfg->neg_iou_limit;
  p.img_w = (double)cfg->img_width; p.img_h = (double)cfg->img_height;
  for (int i = 0; i < 4; ++i) p.var[i] = cfg->variances[i];
  SSDK_CHECK_CUDA(cudaMalloc(&e->d_anchors, (size_t)p.P * 4 * sizeof(double)));
  SSDK_CHECK_CUDA(cudaMalloc(&e->d_bbox, (size_t)p.n_tiles * 4 * sizeof(double)));
  SSDK_CHECK_CUDA(cudaMemcpy(e->d_anchors, anchors_host, (size_t)p.P * 4 * sizeof(double), cudaMemcpyHostToDevice));
  p.anchors = e->d_anchors; p.tile_bbox = e->d_bbox;
  anchor_tile_bbox_kernel<<<p.n_tiles, kTile>>>(p, e->d_bbox);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  SSDK_CHECK_CUDA(cudaDeviceSynchronize());
  *out = e;
  return SSDK_OK;
}

extern "C" int ssdk_encoder_destroy(ssdk_encoder* e) {
  if (!e) return SSDK_OK;
  cudaFree(e->d_anchors); cudaFree(e->d_bbox);
  e->rows.release();
  if (e->h_offsets) cudaFreeHost(e->h_offsets);
  delete e;
  return SSDK_OK;
}

extern "C" int ssdk_iou_matrix(ssdk_encoder* e, const float* gt_boxes_dev, int G, double* out_dev, void* stream) {
  SSDK_REQUIRE(e && gt_boxes_dev && out_dev && G > 0, "ssdk_iou_matrix: bad argument");
  dim3 grid(ceil_div(e->p.P, 256), G);
  iou_matrix_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(e->p, gt_boxes_dev, G, out_dev);
  SSDK_COUNT_LAUNCH(e->ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

extern "C" int ssdk_iou(ssdk_ctx* ctx, const double* boxes1_dev, int m, const double* boxes2_dev, int n, int coords, int border_d,
                        int elementwise, double* out_dev, void* stream) {
  SSDK_REQUIRE(ctx && boxes1_dev && boxes2_dev && out_dev && m > 0 && n > 0, "ssdk_iou: bad argument");
  SSDK_REQUIRE(coords >= 0 && coords <= 2, "Unexpected value for `coords`. Supported values are 'minmax', 'corners' and 'centroids'.");
  SSDK_REQUIRE(!elementwise || m == n || m == 1 || n == 1, "ssdk_iou: element-wise mode needs broadcast-compatible box counts");
  long long total = elementwise ? (long long)(m > n ? m : n) : (long long)m * n;
  iou_general_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, (cudaStream_t)stream>>>(boxes1_dev, m, boxes2_dev, n, coords,
                                                                                          border_d, elementwise, out_dev);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

extern "C" int ssdk_encode(ssdk_encoder* e, const float* gt_boxes_dev, const int* gt_offsets_host, int B,
                           float* out_y_dev, int* out_match_dev, int* status_dev, void* stream_) {
  SSDK_REQUIRE(e && gt_offsets_host && out_y_dev && B > 0, "ssdk_encode: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  const EncParams& p = e->p;
  const int total_g = gt_offsets_host[B];
  SSDK_REQUIRE(gt_offsets_host[0] == 0 && total_g >= 0, "ssdk_encode: gt_offsets must start at 0 and be non-decreasing");
  int max_g = 0;
  for (int b = 0; b < B; ++b) {
    int g = gt_offsets_host[b + 1] - gt_offsets_host[b];
    SSDK_REQUIRE(g >= 0, "ssdk_encode: gt_offsets must be non-decreasing");
    max_g = g > max_g ? g : max_g;
  }
  SSDK_REQUIRE(total_g == 0 || gt_boxes_dev, "ssdk_encode: gt_boxes_dev is NULL");
  // scratch: rowmax[total_g] f64 | rowarg[total_g] | matches[total_g] | offsets[B+1]
  size_t n = (size_t)(total_g > 0 ? total_g : 1);
  size_t bytes = n * 8 + n * 4 + n * 4 + (size_t)(B + 1) * 4 + 64;
  int rc = e->rows.ensure(bytes);
  if (rc) return rc;
  double* rowmax = reinterpret_cast<double*>(e->rows.ptr);
  int* rowarg = reinterpret_cast<int*>(rowmax + n);
  int* matches = rowarg + n;
  int* d_offsets = matches + n;
  if (e->h_offsets_cap < B + 1) {
    if (e->h_offsets) cudaFreeHost(e->h_offsets);
    SSDK_CHECK_CUDA(cudaMallocHost(&e->h_offsets, (size_t)(B + 1) * sizeof(int)));
    e->h_offsets_cap = B + 1;
  }
  // the pinned staging buffer is reused across calls: wait for the previous copy to be consumed
  SSDK_CHECK_CUDA(cudaStreamSynchronize(stream));
  memcpy(e->h_offsets, gt_offsets_host, (size_t)(B + 1) * sizeof(int));
  SSDK_CHECK_CUDA(cudaMemcpyAsync(d_offsets, e->h_offsets, (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
  if (total_g > 0) {
    enc_rowmax_kernel<<<total_g, kTile, 0, stream>>>(p, gt_boxes_dev, d_offsets, B, rowmax, rowarg, status_dev);
    SSDK_COUNT_LAUNCH(e->ctx);
    size_t sm_b = (size_t)max_g * (8 + 4 + 4) + 16;
    if (sm_b > 48 * 1024) SSDK_CHECK_CUDA(cudaFuncSetAttribute(enc_bipartite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
    enc_bipartite_kernel<<<B, kTile, sm_b, stream>>>(p, gt_boxes_dev, d_offsets, rowmax, rowarg, matches);
    SSDK_COUNT_LAUNCH(e->ctx);
  }
  const int W = p.C + 12;
  size_t sm_m = (((size_t)kTile * W * sizeof(float) + 15) & ~(size_t)15) + (size_t)(max_g > 0 ? max_g : 1) * (5 * 8 + 4) + kTile * 4 + 16;
  SSDK_REQUIRE(sm_m <= 227 * 1024, "ssdk_encode: n_classes (%d) / gt count (%d) need %zu bytes of shared memory", p.C, max_g, sm_m);
  if (sm_m > 48 * 1024) SSDK_CHECK_CUDA(cudaFuncSetAttribute(enc_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_m));
  dim3 grid(p.n_tiles, B);
  enc_main_kernel<<<grid, kTile, sm_m, stream>>>(p, gt_boxes_dev, d_offsets, matches, out_y_dev, out_match_dev);
  SSDK_COUNT_LAUNCH(e->ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}
