// SSDInputEncoder hot path on sm_100a: pairwise IoU (float64), greedy bipartite + multi matching,
// neutral boxes and offset encoding.  Reference: ssd_encoder_decoder/ssd_input_encoder.py:277-418,
// bounding_box_utils/bounding_box_utils.py:283-383, ssd_encoder_decoder/matching_utils.py:22-116.
//
// Kernels (all HBM/ALU-bound integer/float64 work, no tensor cores):
//   enc_fused_kernel   one CTA per (256-anchor tile, image), ONE pass over the IoU pairs: ground-truth boxes that
//                      can touch the tile form an ordered candidate list; a float32 test on outward-rounded corners
//                      rejects disjoint pairs before any float64 work; per anchor the best gt (multi-match /
//                      neutral rule) and per (gt, tile) the best anchor are reduced; the (C+12)-float target rows
//                      are staged in shared memory and written coalesced.
//   enc_greedy_kernel  one CTA per image: reduces the per-tile bests to row maxima, then one warp runs the G
//                      sequential greedy rounds of match_bipartite_greedy (zeroed-row quirk included).  When a
//                      row loses its best anchor only that anchor's tile (256 IoUs) is re-evaluated.
//   enc_fix_kernel     one warp per ground-truth box: rewrites the row of its bipartite anchor (last gt wins).
// Exactness: every decision-relevant value is computed with the reference's float64 operation order using
// non-contracting intrinsics (__dmul_rn/__dadd_rn/...); arg-max decisions compare the float64 quotients.
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
// shared helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double iou_value(const Box& g, const Box& a, double inter) {
  return __ddiv_rn(inter, union_area(g, a, inter));
}

struct RowDecision { int match_g; bool neutral; };

// One target row [one-hot class | 4 offsets | 4 anchor coords | 4 variances] (ssd_input_encoder.py:363,396-410) in compact form:
// the class vector has at most one 1 (index `one`, -1: none).
struct RowCompact { int one; float o4[4]; };
__device__ __forceinline__ RowCompact make_row(const EncParams& p, const float* gt_rows, int g0, const double at[4], RowDecision dec) {
  RowCompact r;
  r.one = -1; r.o4[0] = r.o4[1] = r.o4[2] = r.o4[3] = 0.f;
  if (dec.match_g >= 0) {
    double gtc[4]; int cls;
    gt_template(gt_rows + (size_t)(g0 + dec.match_g) * 5, p, gtc, cls);
    if (cls >= 0 && cls < p.C) r.one = cls;
    if (p.coords == SSDK_COORDS_CENTROIDS) {                // :396-400
      r.o4[0] = (float)__ddiv_rn(__dsub_rn(gtc[0], at[0]), __dmul_rn(at[2], p.var[0]));
      r.o4[1] = (float)__ddiv_rn(__dsub_rn(gtc[1], at[1]), __dmul_rn(at[3], p.var[1]));
      r.o4[2] = (float)__ddiv_rn(log(__ddiv_rn(gtc[2], at[2])), p.var[2]);
      r.o4[3] = (float)__ddiv_rn(log(__ddiv_rn(gtc[3], at[3])), p.var[3]);
    } else if (p.coords == SSDK_COORDS_CORNERS) {           // :401-405
      double w = __dsub_rn(at[2], at[0]), h = __dsub_rn(at[3], at[1]);
      r.o4[0] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[0], at[0]), w), p.var[0]);
      r.o4[1] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[1], at[1]), h), p.var[1]);
      r.o4[2] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[2], at[2]), w), p.var[2]);
      r.o4[3] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[3], at[3]), h), p.var[3]);
    } else {                                                // minmax :406-410
      double w = __dsub_rn(at[1], at[0]), h = __dsub_rn(at[3], at[2]);
      r.o4[0] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[0], at[0]), w), p.var[0]);
      r.o4[1] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[1], at[1]), w), p.var[1]);
      r.o4[2] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[2], at[2]), h), p.var[2]);
      r.o4[3] = (float)__ddiv_rn(__ddiv_rn(__dsub_rn(gtc[3], at[3]), h), p.var[3]);
    }
    if (dec.neutral && r.one == p.bg) r.one = -1;           // neg_iou_limit <= 0 corner case: the background entry is zeroed last
  } else if (!dec.neutral) {
    r.one = p.bg;
  }
  return r;
}
template <typename Store>
__device__ __forceinline__ void emit_row(const EncParams& p, const float* gt_rows, int g0, const double at[4], RowDecision dec,
                                         Store store) {
  const RowCompact r = make_row(p, gt_rows, g0, at, dec);
  for (int c = 0; c < p.C; ++c) store(c, c == r.one ? 1.f : 0.f);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    store(p.C + k, r.o4[k]);
    store(p.C + 4 + k, (float)at[k]);
    store(p.C + 8 + k, (float)p.var[k]);
  }
}

// (value desc, index asc) warp reduction
__device__ __forceinline__ void warp_argmax(double& val, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, val, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > val || (ov == val && oi < idx)) { val = ov; idx = oi; }
  }
}

// ------------------------------------------------------------------------------------------
// enc_fused_kernel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile) enc_fused_kernel(EncParams p, const float* __restrict__ gt,
                                                          const int* __restrict__ gt_offsets,
                                                          double* __restrict__ tb_val, int* __restrict__ tb_idx,
                                                          float* __restrict__ out_y, int* __restrict__ out_match,
                                                          int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tile = blockIdx.x, b = blockIdx.y;
  const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
  const int Gs = G > 0 ? G : 1;
  const int W = p.C + 12;
  // shared layout: rows[kTile*W] f32 | cand f32 bounds 4*G (float4 aligned) | cand boxes 5*G f64 | per-warp tile bests 8*G f64 |
  //                cand idx G | per-warp tile-best idx 8*G | cand slot of gt G
  // (staging compact rows and expanding them in the copy loop was measured slower on B200: 45 us vs 22 us for SSD300 B=32)
  float* rows = reinterpret_cast<float*>(smem_raw);
  size_t off = ((size_t)kTile * W * sizeof(float) + 15) & ~(size_t)15;
  float* cf = reinterpret_cast<float*>(smem_raw + off); off += (size_t)4 * Gs * sizeof(float);
  double* cb = reinterpret_cast<double*>(smem_raw + off); off += (size_t)5 * Gs * sizeof(double);
  double* wv = reinterpret_cast<double*>(smem_raw + off); off += (size_t)8 * Gs * sizeof(double);
  int* cidx = reinterpret_cast<int*>(smem_raw + off); off += (size_t)Gs * sizeof(int);
  int* wi = reinterpret_cast<int*>(smem_raw + off); off += (size_t)8 * Gs * sizeof(int);
  int* slot_of = reinterpret_cast<int*>(smem_raw + off); off += (size_t)Gs * sizeof(int);
  int* s_wfirst = reinterpret_cast<int*>(smem_raw + off);   // [8*G] first anchor index attaining the per-warp best
  __shared__ int s_ncand;

  const int a0 = tile * kTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int g = threadIdx.x; g < G; g += blockDim.x) slot_of[g] = -1;
  if (threadIdx.x == 0) s_ncand = 0;
  __syncthreads();
  // ordered candidate list (ascending gt index) built by warp 0
  if (warp == 0) {
    const double* bb = p.tile_bbox + (size_t)tile * 4;
    int n = 0;
    bool bad = false;
    for (int base = 0; base < G; base += 32) {
      int g = base + lane;
      bool hit = false; Box gb{};
      if (g < G) {
        double t[4]; int cls;
        bad |= !gt_template(gt + (size_t)(g0 + g) * 5, p, t, cls);
        gb = corners_from_template(t, p.coords, p.d);
        hit = bbox_hits(bb, gb);
      }
      unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        int pos = n + __popc(m & ((1u << lane) - 1));
        cb[pos * 5 + 0] = gb.x0; cb[pos * 5 + 1] = gb.y0; cb[pos * 5 + 2] = gb.x1; cb[pos * 5 + 3] = gb.y1; cb[pos * 5 + 4] = gb.area;
        // outward-rounded float32 corners: disjoint here => disjoint in float64
        cf[pos * 4 + 0] = __double2float_rd(gb.x0); cf[pos * 4 + 1] = __double2float_rd(gb.y0);
        cf[pos * 4 + 2] = __double2float_ru(gb.x1); cf[pos * 4 + 3] = __double2float_ru(gb.y1);
        cidx[pos] = g; slot_of[g] = pos;
      }
      n += __popc(m);
    }
    if (lane == 0) s_ncand = n;
    if (tile == 0 && status && __any_sync(0xffffffffu, bad) && lane == 0) atomicMax(status, b + 1);
  }
  __syncthreads();
  const int ncand = s_ncand;
  const int a = a0 + threadIdx.x;
  const bool live = a < p.P;
  double at[4] = {0, 0, 0, 0};
  Box ab{};
  float fx0 = 0, fy0 = 0, fx1 = 0, fy1 = 0;
  if (live) {
    const double2* q = reinterpret_cast<const double2*>(p.anchors + (size_t)a * 4);
    double2 u = __ldg(q), v = __ldg(q + 1);
    at[0] = u.x; at[1] = u.y; at[2] = v.x; at[3] = v.y;
    ab = corners_from_template(at, p.coords, p.d);
    fx0 = __double2float_rd(ab.x0); fy0 = __double2float_rd(ab.y0); fx1 = __double2float_ru(ab.x1); fy1 = __double2float_ru(ab.y1);
  }
  double best = 0.0; int best_g = -1;
  for (int c = 0; c < ncand; ++c) {
    double val = 0.0;
    const float4 gf = *reinterpret_cast<const float4*>(cf + c * 4);
    const bool maybe = live && (fminf(fx1, gf.z) > fmaxf(fx0, gf.x)) && (fminf(fy1, gf.w) > fmaxf(fy0, gf.y));
    if (maybe) {
      Box gb; gb.x0 = cb[c * 5]; gb.y0 = cb[c * 5 + 1]; gb.x1 = cb[c * 5 + 2]; gb.y1 = cb[c * 5 + 3]; gb.area = cb[c * 5 + 4];
      double inter = inter_area(gb, ab);
      if (inter > 0.0) val = iou_value(gb, ab, inter);
    }
    if (val > best) { best = val; best_g = cidx[c]; }          // strict '>' keeps the first gt on ties (np.argmax)
    // best anchor of this gt inside the warp's 32 anchors (first index on ties)
    if (__any_sync(0xffffffffu, val > 0.0)) {
      double rv = val; int ri = (val > 0.0) ? a : INT_MAX;
      warp_argmax(rv, ri);
      if (lane == 0) { wv[warp * Gs + c] = rv; s_wfirst[warp * Gs + c] = ri; }
    } else if (lane == 0) { wv[warp * Gs + c] = 0.0; s_wfirst[warp * Gs + c] = INT_MAX; }
  }
  if (live) {
    RowDecision dec{-1, false};
    double val = best;
    if (G > 0) {
      const int arg = best_g >= 0 ? best_g : 0;                // np.argmax of an all-zero column is 0
      if (p.multi && val >= p.pos_thr) { dec.match_g = arg; val = 0.0; }   // column zeroed after matching (:381)
      if (val >= p.neg_lim) dec.neutral = true;                            // :388-390
    }
    float* my = rows + (size_t)threadIdx.x * W;
    emit_row(p, gt, g0, at, dec, [&](int k, float v) { my[k] = v; });
    if (out_match) out_match[(size_t)b * p.P + a] = (dec.match_g >= 0) ? dec.match_g : (dec.neutral ? -2 : -1);
  }
  __syncthreads();
  // per (gt, tile) best anchor -> global
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double v = 0.0; int i = INT_MAX;
    const int c = slot_of[g];
    if (c >= 0) {
      for (int w = 0; w < 8; ++w) {                             // warps are in ascending anchor order: strict '>' keeps the first
        double ov = wv[w * Gs + c];
        if (ov > v) { v = ov; i = s_wfirst[w * Gs + c]; }
      }
    }
    tb_val[(size_t)(g0 + g) * p.n_tiles + tile] = v;
    tb_idx[(size_t)(g0 + g) * p.n_tiles + tile] = (v > 0.0) ? i : INT_MAX;
  }
  // coalesced copy of the staged rows
  const int n_rows = min(kTile, p.P - a0);
  const size_t n_f = (size_t)n_rows * W;
  float* dst = out_y + ((size_t)b * p.P + a0) * W;
  for (size_t i = threadIdx.x; i < n_f; i += blockDim.x) dst[i] = rows[i];
}

// ------------------------------------------------------------------------------------------
// enc_greedy_kernel: match_bipartite_greedy (matching_utils.py:63-77) for one image
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_removed(const int* removed, int n, int a) {
  bool r = false;
  for (int i = 0; i < n; ++i) r |= (removed[i] == a);
  return r;
}

// Row `g` keeps a COMPACT list of its non-zero per-tile bests (value, anchor index); the tile of an entry is idx / kTile.
// Re-evaluate tile `t` for ground-truth `gb` ignoring removed anchors (one warp) and update the row's entry for that tile.
__device__ void warp_fix_tile(const EncParams& p, const Box& gb, int t, const int* removed, int n_removed, double* tv, int* ti, int nnz) {
  const int lane = threadIdx.x & 31;
  double bv = 0.0; int bi = INT_MAX;
  for (int k = 0; k < kTile / 32; ++k) {
    const int a = t * kTile + k * 32 + lane;
    if (a < p.P) {
      Box ab = load_anchor(p, a);
      double inter = inter_area(gb, ab);
      if (inter > 0.0) {
        double v = iou_value(gb, ab, inter);
        if (v > bv && !is_removed(removed, n_removed, a)) { bv = v; bi = a; }
      }
    }
  }
  warp_argmax(bv, bi);
  for (int e = lane; e < nnz; e += 32) {
    if (__ldcg(ti + e) / kTile == t) {                       // exactly one entry per tile
      tv[e] = bv;
      if (bv > 0.0) ti[e] = bi;                              // keep the old index (tile id) when the tile is exhausted
    }
  }
  __syncwarp();
}

__device__ void warp_row_reduce(const double* tv, const int* ti, int nnz, double& val, int& idx) {
  const int lane = threadIdx.x & 31;
  double bv = 0.0; int bi = INT_MAX;
  for (int e = lane; e < nnz; e += 32) {
    double v = __ldcg(tv + e); int i = __ldcg(ti + e);
    if (v > bv || (v == bv && v > 0.0 && i < bi)) { bv = v; bi = i; }
  }
  warp_argmax(bv, bi);
  val = bv; idx = bi;
}

__global__ void __launch_bounds__(256) enc_greedy_kernel(EncParams p, const float* __restrict__ gt,
                                                         const int* __restrict__ gt_offsets, double* tb_val, int* tb_idx,
                                                         int* __restrict__ matches) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const int g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0;
  if (G <= 0) return;
  double* rv = reinterpret_cast<double*>(smem_raw);          // [G] current row maximum
  double* gbox = rv + G;                                      // [5*G] corner boxes of the ground truth
  int* ra = reinterpret_cast<int*>(gbox + 5 * (size_t)G);     // [G] its (first) anchor index
  int* removed = ra + G;                                      // [G] anchors taken so far
  int* nnz = removed + G;                                     // [G] entries in the row's compact tile list
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // phase A (one warp per row): compact the non-zero per-tile bests in place, row maximum, cached gt box
  for (int g = warp; g < G; g += (blockDim.x >> 5)) {
    double* tv = tb_val + (size_t)(g0 + g) * p.n_tiles;
    int* ti = tb_idx + (size_t)(g0 + g) * p.n_tiles;
    int k = 0;
    double bv = 0.0; int bi = INT_MAX;
    for (int base = 0; base < p.n_tiles; base += 32) {
      const int t = base + lane;
      double v = 0.0; int i = INT_MAX;
      if (t < p.n_tiles) { v = __ldcg(tv + t); i = __ldcg(ti + t); }
      const bool nz = v > 0.0;
      const unsigned m = __ballot_sync(0xffffffffu, nz);
      __syncwarp();
      if (nz) {
        const int pos = k + __popc(m & ((1u << lane) - 1));
        tv[pos] = v; ti[pos] = i;
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
      }
      k += __popc(m);
      __syncwarp();
    }
    warp_argmax(bv, bi);
    double t4[4]; int cls;
    gt_template(gt + (size_t)(g0 + g) * 5, p, t4, cls);
    Box gb = corners_from_template(t4, p.coords, p.d);
    if (lane == 0) {
      rv[g] = bv; ra[g] = (bv > 0.0) ? bi : 0; nnz[g] = k; matches[g0 + g] = 0;    // argmax of an all-zero row is 0
      gbox[g * 5] = gb.x0; gbox[g * 5 + 1] = gb.y0; gbox[g * 5 + 2] = gb.x1; gbox[g * 5 + 3] = gb.y1; gbox[g * 5 + 4] = gb.area;
    }
  }
  __threadfence_block();
  __syncthreads();
  if (warp != 0) return;
  // phase B: the G sequential rounds, one warp
  int n_removed = 0;
  for (int round = 0; round < G; ++round) {
    double v = -1.0; int gi = INT_MAX;
    for (int g = lane; g < G; g += 32)
      if (rv[g] > v) { v = rv[g]; gi = g; }                   // ascending g per lane: first index kept
    warp_argmax(v, gi);
    const int a_star = ra[gi];
    __syncwarp();
    if (lane == 0) { matches[g0 + gi] = a_star; rv[gi] = 0.0; ra[gi] = 0; removed[n_removed] = a_star; }
    ++n_removed;
    __syncwarp();
    if (!(v > 0.0)) continue;                                  // nothing left to match: no row can point at a_star
    // rows that pointed at the taken anchor: fix that tile, re-reduce, repeat while the new best is itself stale
    for (int base = 0; base < G; base += 32) {
      const int g = base + lane;
      unsigned need = __ballot_sync(0xffffffffu, g < G && rv[g] > 0.0 && ra[g] == a_star);
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const int gg = base + src;
        Box gb; gb.x0 = gbox[gg * 5]; gb.y0 = gbox[gg * 5 + 1]; gb.x1 = gbox[gg * 5 + 2]; gb.y1 = gbox[gg * 5 + 3]; gb.area = gbox[gg * 5 + 4];
        double* tv = tb_val + (size_t)(g0 + gg) * p.n_tiles;
        int* ti = tb_idx + (size_t)(g0 + gg) * p.n_tiles;
        const int nz = nnz[gg];
        int stale = a_star;
        double nv; int ni;
        while (true) {
          warp_fix_tile(p, gb, stale / kTile, removed, n_removed, tv, ti, nz);
          __threadfence_block();
          warp_row_reduce(tv, ti, nz, nv, ni);
          if (!(nv > 0.0) || !is_removed(removed, n_removed, ni)) break;   // warp-uniform
          stale = ni;                                          // a tile entry recorded before that anchor was taken
        }
        if (lane == 0) { rv[gg] = nv; ra[gg] = (nv > 0.0) ? ni : 0; }
        __syncwarp();
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// enc_fix_kernel: y_encoded[i, bipartite_matches, :-8] = labels_one_hot (:363), last writer wins
// ------------------------------------------------------------------------------------------
__global__ void enc_fix_kernel(EncParams p, const float* __restrict__ gt, const int* __restrict__ gt_offsets, int B,
                               int total_g, const int* __restrict__ matches, float* __restrict__ out_y,
                               int* __restrict__ out_match) {
  const int gg = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gg >= total_g) return;
  int lo = 0, hi = B;                                         // image of this gt row
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (gt_offsets[mid] <= gg) lo = mid; else hi = mid; }
  const int b = lo, g0 = gt_offsets[b], G = gt_offsets[b + 1] - g0, g = gg - g0;
  const int a = matches[gg];
  bool last = true;
  for (int g2 = g + 1 + lane; g2 < G; g2 += 32) last &= (matches[g0 + g2] != a);
  last = __all_sync(0xffffffffu, last);
  if (!last) return;
  const double2* q = reinterpret_cast<const double2*>(p.anchors + (size_t)a * 4);
  double2 u = __ldg(q), v = __ldg(q + 1);
  const double at[4] = {u.x, u.y, v.x, v.y};
  RowDecision dec{g, 0.0 >= p.neg_lim};                       // the matched column is all zero
  float* dst = out_y + ((size_t)b * p.P + a) * (p.C + 12);
  if (lane == 0) {
    emit_row(p, gt, g0, at, dec, [&](int k, float v2) { dst[k] = v2; });
    if (out_match) out_match[(size_t)b * p.P + a] = g;
  }
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
  Scratch rows;        // per-(gt, tile) bests + matches + offsets
  // pinned staging ring for the gt offsets: a slot is reused only after the copy issued from it has completed (no stream sync)
  static constexpr int kSlots = 8;
  int* h_offsets = nullptr;   // kSlots * h_offsets_cap ints
  int h_offsets_cap = 0;
  cudaEvent_t slot_done[kSlots] = {};
  int next_slot = 0;
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
  for (int i = 0; i < ssdk_encoder::kSlots; ++i) if (e->slot_done[i]) cudaEventDestroy(e->slot_done[i]);
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
  // scratch: tile-best values [total_g * n_tiles] f64 | tile-best indices [total_g * n_tiles] | matches[total_g] | offsets[B+1]
  const size_t n = (size_t)(total_g > 0 ? total_g : 1);
  const size_t nt = n * (size_t)p.n_tiles;
  const size_t bytes = nt * 8 + nt * 4 + n * 4 + (size_t)(B + 1) * 4 + 64;
  int rc = e->rows.ensure(bytes);
  if (rc) return rc;
  double* tb_val = reinterpret_cast<double*>(e->rows.ptr);
  int* tb_idx = reinterpret_cast<int*>(tb_val + nt);
  int* matches = tb_idx + nt;
  int* d_offsets = matches + n;
  if (e->h_offsets_cap < B + 1) {
    SSDK_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (e->h_offsets) cudaFreeHost(e->h_offsets);
    SSDK_CHECK_CUDA(cudaMallocHost(&e->h_offsets, (size_t)ssdk_encoder::kSlots * (B + 1) * sizeof(int)));
    e->h_offsets_cap = B + 1;
  }
  const int slot = e->next_slot;
  e->next_slot = (slot + 1) % ssdk_encoder::kSlots;
  if (!e->slot_done[slot]) SSDK_CHECK_CUDA(cudaEventCreateWithFlags(&e->slot_done[slot], cudaEventDisableTiming));
  else SSDK_CHECK_CUDA(cudaEventSynchronize(e->slot_done[slot]));
  int* h_off = e->h_offsets + (size_t)slot * e->h_offsets_cap;
  memcpy(h_off, gt_offsets_host, (size_t)(B + 1) * sizeof(int));
  SSDK_CHECK_CUDA(cudaMemcpyAsync(d_offsets, h_off, (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice, stream));
  SSDK_CHECK_CUDA(cudaEventRecord(e->slot_done[slot], stream));
  const int W = p.C + 12;
  const size_t gs = (size_t)(max_g > 0 ? max_g : 1);
  const size_t sm_m = (((size_t)kTile * W * sizeof(float) + 15) & ~(size_t)15) + gs * (5 * 8 + 8 * 8 + 4 * 4 + 4 + 8 * 4 + 4 + 8 * 4) + 64;
  SSDK_REQUIRE(sm_m <= 227 * 1024, "ssdk_encode: n_classes (%d) / gt count (%d) need %zu bytes of shared memory", p.C, max_g, sm_m);
  if (sm_m > 48 * 1024) SSDK_CHECK_CUDA(cudaFuncSetAttribute(enc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_m));
  dim3 grid(p.n_tiles, B);
  enc_fused_kernel<<<grid, kTile, sm_m, stream>>>(p, gt_boxes_dev, d_offsets, tb_val, tb_idx, out_y_dev, out_match_dev, status_dev);
  SSDK_COUNT_LAUNCH(e->ctx);
  if (total_g > 0) {
    const size_t sm_b = (size_t)max_g * (8 + 40 + 4 + 4 + 4) + 16;
    if (sm_b > 48 * 1024) SSDK_CHECK_CUDA(cudaFuncSetAttribute(enc_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
    enc_greedy_kernel<<<B, 256, sm_b, stream>>>(p, gt_boxes_dev, d_offsets, tb_val, tb_idx, matches);
    SSDK_COUNT_LAUNCH(e->ctx);
    enc_fix_kernel<<<ceil_div(total_g, 8), 256, 0, stream>>>(p, gt_boxes_dev, d_offsets, B, total_g, matches, out_y_dev, out_match_dev);
    SSDK_COUNT_LAUNCH(e->ctx);
  }
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}
