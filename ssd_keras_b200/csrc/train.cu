// Training step for the SSD graphs on sm_100a: backward pass + SGD.  Replaces what TensorFlow/Keras do for the reference
// in fit_generator (autodiff of models/keras_ssd300.py:263-419 and keras_loss_function/keras_ssd_loss.py:98-211, the
// l2 kernel regulariser models/keras_ssd300.py:274 and SGD(lr, momentum) ssd300_training.ipynb:169).
//
// Every convolution gradient runs on the SAME tcgen05 implicit-GEMM kernel as the forward pass (conv.cu):
//   data gradient    dX = conv(dZ, W rotated by 180 degrees with in/out channels swapped), padding dilation*(k-1)-pad;
//                    the epilogue multiplies by ReLU'(forward value) and accumulates when a tensor has several consumers.
//   weight gradient  dW[co][tap][ci] = sum_v dZT[co][v] * XT[ci][v + shift(tap)]: both operands are transposed once into
//                    [channels][pixels] matrices (K = pixels contiguous), each tap is one GEMM whose weight-side operand
//                    is read at a K offset, the pixel axis is split across CTAs (split-K) and reduced with fp32 atomics.
// Operands stay bf16 hi+lo (three MMAs per product) like the forward pass.  Small pieces (max-pool routing, L2Normalization,
// softmax/concat head, bias sums, image-facing 3-channel conv, SGD, re-packing of the bf16 planes) are plain CUDA kernels.
#include "model.cuh"
#include "wgrad.cuh"
#include <climits>

using namespace ssdk;

namespace {

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ size_t aidx(const ActBuf& a, int n, int y, int x) {
  return (((size_t)n * a.Hp() + (y + a.pad)) * a.Wp() + (x + a.pad)) * a.Cs;
}
__device__ __forceinline__ float ld2(const ActBuf& a, size_t i) {
  float v = __bfloat162float(a.hi[i]);
  if (a.lo) v += __bfloat162float(a.lo[i]);
  return v;
}
__device__ __forceinline__ void st2(const ActBuf& a, size_t i, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  a.hi[i] = h;
  if (a.lo) a.lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// dst[c][v] = src[row(v)][c] (or 0), v in [0, Kv): K-major operands for the weight-gradient GEMMs.
struct TMap {
  int identity;             // row(v) = v
  int rows_per_img, Wp;     // the X grid the GEMM iterates over
  int Ho, Wo;               // valid extent on that grid
  int src_Hp, src_Wp, src_pad;
};
__global__ void __launch_bounds__(256) transpose_kernel(const __nv_bfloat16* __restrict__ src_hi, const __nv_bfloat16* __restrict__ src_lo,
                                                        int src_ld, long long src_rows, TMap mp, long long row_off, long long Kv, int C,
                                                        __nv_bfloat16* __restrict__ dst_hi, __nv_bfloat16* __restrict__ dst_lo, long long ldT) {
  __shared__ uint16_t th[64][72], tl[64][72];
  const long long v0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {          // 64 rows x 8 chunks of 8 channels
    const int r = i >> 3, ch = (i & 7) * 8;
    const long long v = v0 + r;
    long long row = -1;
    if (v < Kv) {
      if (mp.identity) row = v + row_off;
      else {
        const long long vs = v + row_off;
        const int n = (int)(vs / mp.rows_per_img); const int rr = (int)(vs - (long long)n * mp.rows_per_img);
        const int y = rr / mp.Wp, x = rr - y * mp.Wp;
        if (y < mp.Ho && x < mp.Wo) row = ((long long)n * mp.src_Hp + (y + mp.src_pad)) * mp.src_Wp + (x + mp.src_pad);
      }
    }
    uint4 h = make_uint4(0, 0, 0, 0), l = h;
    if (row >= 0 && row < src_rows && c0 + ch < src_ld) {
      h = *reinterpret_cast<const uint4*>(src_hi + row * src_ld + c0 + ch);
      if (src_lo) l = *reinterpret_cast<const uint4*>(src_lo + row * src_ld + c0 + ch);
    }
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      th[ch + e][r] = (uint16_t)((hw[e >> 1] >> ((e & 1) * 16)) & 0xffffu);
      tl[ch + e][r] = (uint16_t)((lw[e >> 1] >> ((e & 1) * 16)) & 0xffffu);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 8; i += 256) {          // 64 channels x 8 chunks of 8 pixels
    const int c = i >> 3, vv = (i & 7) * 8;
    if (c0 + c >= C) continue;
    if (v0 + vv >= ldT) continue;
    uint4 h, l;
    h.x = th[c][vv] | ((uint32_t)th[c][vv + 1] << 16); h.y = th[c][vv + 2] | ((uint32_t)th[c][vv + 3] << 16);
    h.z = th[c][vv + 4] | ((uint32_t)th[c][vv + 5] << 16); h.w = th[c][vv + 6] | ((uint32_t)th[c][vv + 7] << 16);
    l.x = tl[c][vv] | ((uint32_t)tl[c][vv + 1] << 16); l.y = tl[c][vv + 2] | ((uint32_t)tl[c][vv + 3] << 16);
    l.z = tl[c][vv + 4] | ((uint32_t)tl[c][vv + 5] << 16); l.w = tl[c][vv + 6] | ((uint32_t)tl[c][vv + 7] << 16);
    *reinterpret_cast<uint4*>(dst_hi + (long long)(c0 + c) * ldT + v0 + vv) = h;
    if (dst_lo) *reinterpret_cast<uint4*>(dst_lo + (long long)(c0 + c) * ldT + v0 + vv) = l;
  }
}

// bias gradient: gb[c] += sum_v dZT[c][v]
__global__ void __launch_bounds__(256) rowsum_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, long long ldT,
                                                     long long Kv, float* __restrict__ gb) {
  __shared__ float s[8];
  const int c = blockIdx.y;
  const long long chunk = (Kv + gridDim.x - 1) / gridDim.x;
  const long long v0 = (long long)blockIdx.x * chunk, v1 = min(Kv, v0 + chunk);
  float acc = 0.f;
  for (long long v = v0 + threadIdx.x; v < v1; v += 256) {
    acc += __bfloat162float(hi[(long long)c * ldT + v]);
    if (lo) acc += __bfloat162float(lo[(long long)c * ldT + v]);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s[w];
    atomicAdd(gb + c, t);
  }
}

// Backward of Reshape/softmax/Concat (models/keras_ssd300.py:363-419): dY_pred rows -> gradient of the fused head conv output.
__global__ void head_bwd_kernel(const float* __restrict__ head, const float* __restrict__ dy, int B, int H, int W, int n_boxes, int C,
                                int P, int prior_off, ActBuf g) {
  const size_t wid = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)B * H * W * n_boxes;
  if (wid >= total) return;
  const int b = (int)(wid % n_boxes); const size_t pix = wid / n_boxes;
  const int hw = H * W;
  const int n = (int)(pix / hw); const int pl = (int)(pix % hw);
  const int y = pl / W, x = pl % W;
  const float* src = head + pix * (size_t)n_boxes * (C + 4) + (size_t)b * (C + 4);
  const float* d = dy + ((size_t)n * P + prior_off + (size_t)pl * n_boxes + b) * (C + 12);
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, src[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(src[c] - mx);
  sum = warp_sum(sum);
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) dot += (expf(src[c] - mx) / sum) * d[c];
  dot = warp_sum(dot);
  const size_t o = aidx(g, n, y, x) + (size_t)b * (C + 4);
  for (int c = lane; c < C; c += 32) { const float p = expf(src[c] - mx) / sum; st2(g, o + c, p * (d[c] - dot)); }
  if (lane < 4) st2(g, o + C + lane, d[C + lane]);
}

// Max-pool backward (gather form): gin(n,y,x,c) (+)= sum over windows whose FIRST maximum is (y,x) of gout; optional ReLU' mask.
// One thread per (input pixel, 8 channels): 16-byte loads of the hi / lo planes.
__device__ __forceinline__ void ld8(const ActBuf& a, size_t i, float (&v)[8]) {
  const uint4 h = *reinterpret_cast<const uint4*>(a.hi + i);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { v[2 * e] = __uint_as_float(hw[e] << 16); v[2 * e + 1] = __uint_as_float(hw[e] & 0xffff0000u); }
  if (a.lo) {
    const uint4 l = *reinterpret_cast<const uint4*>(a.lo + i);
    const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[2 * e] += __uint_as_float(lw[e] << 16); v[2 * e + 1] += __uint_as_float(lw[e] & 0xffff0000u); }
  }
}
__device__ __forceinline__ void st8(const ActBuf& a, size_t i, const float (&v)[8]) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
    hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * e] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v[2 * e + 1] - __bfloat162float(h1));
    lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(a.hi + i) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  if (a.lo) *reinterpret_cast<uint4*>(a.lo + i) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
}
__global__ void __launch_bounds__(256) pool_bwd_kernel(ActBuf in, ActBuf gout, ActBuf gin, int Hout, int Wout, int KH, int KW, int stride,
                                                       int pad_t, int pad_l, int relu_mask, int accumulate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int CG = in.C >> 3;
  const size_t total = (size_t)in.B * in.H * in.W * CG;
  if (i >= total) return;
  const int c = (int)(i % CG) * 8; const size_t pix = i / CG;
  const int x = (int)(pix % in.W); const int y = (int)((pix / in.W) % in.H); const int n = (int)(pix / ((size_t)in.W * in.H));
  float v[8], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  ld8(in, aidx(in, n, y, x) + c, v);
  const int yo_lo = max(0, (y + pad_t - KH + stride) / stride), yo_hi = min(Hout - 1, (y + pad_t) / stride);
  const int xo_lo = max(0, (x + pad_l - KW + stride) / stride), xo_hi = min(Wout - 1, (x + pad_l) / stride);
  for (int yo = yo_lo; yo <= yo_hi; ++yo) {
    const int y0 = yo * stride - pad_t;
    if (y < y0 || y >= y0 + KH) continue;
    for (int xo = xo_lo; xo <= xo_hi; ++xo) {
      const int x0 = xo * stride - pad_l;
      if (x < x0 || x >= x0 + KW) continue;
      // per channel: is (y, x) the first maximum of this window (row-major scan, strict '>')?
      unsigned first = 0xffu;
      for (int ky = 0; ky < KH && first; ++ky) {
        const int yy = y0 + ky;
        if (yy < 0 || yy >= in.H) continue;
        for (int kx = 0; kx < KW; ++kx) {
          const int xx = x0 + kx;
          if (xx < 0 || xx >= in.W || (yy == y && xx == x)) continue;
          float u[8];
          ld8(in, aidx(in, n, yy, xx) + c, u);
          const bool before = (yy < y) || (yy == y && xx < x);
#pragma unroll
          for (int e = 0; e < 8; ++e) if (u[e] > v[e] || (u[e] == v[e] && before)) first &= ~(1u << e);
        }
      }
      if (first) {
        float g[8];
        ld8(gout, aidx(gout, n, yo, xo) + c, g);
#pragma unroll
        for (int e = 0; e < 8; ++e) if (first & (1u << e)) acc[e] += g[e];
      }
    }
  }
  const size_t o = aidx(gin, n, y, x) + c;
  if (relu_mask) {
#pragma unroll
    for (int e = 0; e < 8; ++e) if (!(v[e] > 0.f)) acc[e] = 0.f;
  }
  if (accumulate) {
    float old[8];
    ld8(gin, o, old);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += old[e];
  }
  st8(gin, o, acc);
}

// L2Normalization backward (y_c = gamma_c * x_c * s, s = rsqrt(max(sum x^2, 1e-12))): one warp per pixel.
__global__ void l2norm_bwd_kernel(ActBuf x, ActBuf gy, ActBuf gx, const float* __restrict__ gamma, float* __restrict__ ggamma,
                                  int relu_mask, int accumulate) {
  const size_t pix = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)x.B * x.H * x.W;
  if (pix >= total) return;
  const int xx = (int)(pix % x.W); const int yy = (int)((pix / x.W) % x.H); const int n = (int)(pix / ((size_t)x.W * x.H));
  const size_t sx = aidx(x, n, yy, xx), sg = aidx(gy, n, yy, xx), so = aidx(gx, n, yy, xx);
  float ss = 0.f, dot = 0.f;
  for (int c = lane; c < x.C; c += 32) {
    const float v = ld2(x, sx + c);
    ss += v * v;
    dot += gamma[c] * ld2(gy, sg + c) * v;
  }
  ss = warp_sum(ss); dot = warp_sum(dot);
  const bool clamped = !(ss > 1e-12f);
  const float s = rsqrtf(fmaxf(ss, 1e-12f));
  for (int c = lane; c < x.C; c += 32) {
    const float v = ld2(x, sx + c), d = ld2(gy, sg + c);
    float g = s * gamma[c] * d;
    if (!clamped) g -= v * s * s * s * dot;
    atomicAdd(ggamma + c, d * v * s);
    if (relu_mask && !(v > 0.f)) g = 0.f;
    if (accumulate) g += ld2(gx, so + c);
    st2(gx, so + c, g);
  }
}

// Data gradient of a strided convolution: dX = col2im(dCol), dCol = dZ * W^T computed by the GEMM kernel (one tap).
__global__ void col2im_kernel(const float* __restrict__ dcol, int Ho, int Wo, int KH, int KW, int stride, int dil, int pad_t, int pad_l,
                              int ld, ActBuf fwd, ActBuf gin, int relu_mask, int accumulate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)gin.B * gin.H * gin.W * gin.C;
  if (i >= total) return;
  const int c = (int)(i % gin.C); const size_t pix = i / gin.C;
  const int x = (int)(pix % gin.W); const int y = (int)((pix / gin.W) % gin.H); const int n = (int)(pix / ((size_t)gin.W * gin.H));
  float acc = 0.f;
  for (int kh = 0; kh < KH; ++kh) {
    const int yy = y + pad_t - kh * dil;
    if (yy < 0 || yy % stride) continue;
    const int yo = yy / stride;
    if (yo >= Ho) continue;
    for (int kw = 0; kw < KW; ++kw) {
      const int xx = x + pad_l - kw * dil;
      if (xx < 0 || xx % stride) continue;
      const int xo = xx / stride;
      if (xo >= Wo) continue;
      acc += dcol[(((size_t)n * Ho + yo) * Wo + xo) * ld + (size_t)(kh * KW + kw) * gin.C + c];
    }
  }
  const size_t o = aidx(gin, n, y, x) + c;
  if (relu_mask && !(ld2(fwd, aidx(fwd, n, y, x) + c) > 0.f)) acc = 0.f;
  if (accumulate) acc += ld2(gin, o);
  st2(gin, o, acc);
}

// Weight gradient of the image-facing conv (Cin <= 4): gw[co][tap][ci] += sum_pix dZ[pix][co] * X[pix + tap][ci].
// A block owns `rows_per_block` output rows; thread -> (k = tap*cin + ci, 8 output channels), 16-byte loads of dZ.
__global__ void __launch_bounds__(256) wgrad_direct_kernel(ActBuf in, ActBuf g, float* __restrict__ gw, int KH, int KW, int dil,
                                                           int pad_t, int pad_l, int rows_per_block) {
  extern __shared__ float s_acc[];               // [K][Cout]
  const int K = KH * KW * in.C, Cout = g.C;
  for (int i = threadIdx.x; i < K * Cout; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int total_rows = g.B * g.H;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(total_rows, r0 + rows_per_block);
  for (int idx = threadIdx.x; idx < K * (Cout / 8); idx += 256) {
    const int k = idx / (Cout / 8), cg = (idx % (Cout / 8)) * 8;
    const int c = k % in.C, tap = k / in.C, kw = tap % KW, kh = tap / KW;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = r0; r < r1; ++r) {
      const int n = r / g.H, yo = r - n * g.H;
      const int y = yo + kh * dil - pad_t;
      if (y < 0 || y >= in.H) continue;
      const int xo_lo = max(0, pad_l - kw * dil), xo_hi = min(g.W, in.W + pad_l - kw * dil);
      size_t xi = aidx(in, n, y, xo_lo + kw * dil - pad_l) + c;
      size_t go = aidx(g, n, yo, xo_lo) + cg;
      for (int xo = xo_lo; xo < xo_hi; ++xo, xi += in.Cs, go += g.Cs) {
        const float xv = ld2(in, xi);
        float gv[8];
        ld8(g, go, gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += xv * gv[e];
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) s_acc[k * Cout + cg + e] += acc[e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * Cout; i += 256) {
    const int k = i / Cout, co = i % Cout;
    atomicAdd(gw + (size_t)co * K + k, s_acc[i]);     // OHWI: [co][tap][ci], k = tap*cin + ci
  }
}

// Fast path of the above for the 3x3x3 image-facing kernel (conv1_1): thread -> (2 output channels, pixel lane), all 27
// accumulators per channel in registers; the 3-channel input is read through 8-byte broadcast loads, dZ through coalesced
// 4-byte loads.
__global__ void __launch_bounds__(256) wgrad_direct3x3_kernel(ActBuf in, ActBuf g, float* __restrict__ gw, int pad_t, int pad_l,
                                                              int rows_per_block) {
  extern __shared__ float s_acc[];               // [27][Cout]
  const int Cout = g.C, CG = Cout >> 1, PL = 256 / CG;
  for (int i = threadIdx.x; i < 27 * Cout; i += 256) s_acc[i] = 0.f;
  __syncthreads();
  const int cq = threadIdx.x % CG, pl = threadIdx.x / CG;
  if (pl < PL) {
    float acc[27][2];
#pragma unroll
    for (int k = 0; k < 27; ++k) { acc[k][0] = 0.f; acc[k][1] = 0.f; }
    const int total_rows = g.B * g.H;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(total_rows, r0 + rows_per_block);
    for (int r = r0; r < r1; ++r) {
      const int n = r / g.H, yo = r - n * g.H;
      for (int xo = pl; xo < g.W; xo += PL) {
        const size_t go = aidx(g, n, yo, xo) + 2 * cq;
        const uint32_t gh = *reinterpret_cast<const uint32_t*>(g.hi + go);
        float g0 = __uint_as_float(gh << 16), g1 = __uint_as_float(gh & 0xffff0000u);
        if (g.lo) {
          const uint32_t gl = *reinterpret_cast<const uint32_t*>(g.lo + go);
          g0 += __uint_as_float(gl << 16); g1 += __uint_as_float(gl & 0xffff0000u);
        }
        if (g0 == 0.f && g1 == 0.f) continue;
        // window origin; the buffer's zero border (>= 2 - pad) makes every tap a constant offset from it
        const size_t x00 = (((size_t)n * in.Hp() + (yo - pad_t + in.pad)) * in.Wp() + (xo - pad_l + in.pad)) * in.Cs;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const size_t xi = x00 + (size_t)(kh * in.Wp() + kw) * in.Cs;
            const uint2 xh = *reinterpret_cast<const uint2*>(in.hi + xi);
            float x0 = __uint_as_float(xh.x << 16), x1 = __uint_as_float(xh.x & 0xffff0000u), x2 = __uint_as_float(xh.y << 16);
            if (in.lo) {
              const uint2 xl = *reinterpret_cast<const uint2*>(in.lo + xi);
              x0 += __uint_as_float(xl.x << 16); x1 += __uint_as_float(xl.x & 0xffff0000u); x2 += __uint_as_float(xl.y << 16);
            }
            const int k = (kh * 3 + kw) * 3;
            acc[k][0] += x0 * g0; acc[k][1] += x0 * g1;
            acc[k + 1][0] += x1 * g0; acc[k + 1][1] += x1 * g1;
            acc[k + 2][0] += x2 * g0; acc[k + 2][1] += x2 * g1;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      atomicAdd(&s_acc[k * Cout + 2 * cq], acc[k][0]);
      atomicAdd(&s_acc[k * Cout + 2 * cq + 1], acc[k][1]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * Cout; i += 256) {
    const int k = i / Cout, co = i % Cout;
    atomicAdd(gw + (size_t)co * 27 + k, s_acc[i]);
  }
}

// SGD with momentum (Keras: v = m*v - lr*g; w += v); kernels get the l2 regulariser's gradient 2*l2*w.
// `w` is HWIO [taps][cin][cout]; the gradient is OHWI [cout][taps][cin].
__global__ void sgd_kernel_w(float* __restrict__ w, float* __restrict__ v, const float* __restrict__ g, int taps, int cin, int cout,
                             float lr, float mom, float l2, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)taps * cin * cout;
  if (i >= total) return;
  const int co = (int)(i % cout); const size_t r = i / cout; const int ci = (int)(r % cin); const int t = (int)(r / cin);
  const float grad = g[((size_t)co * taps + t) * cin + ci] * scale + 2.f * l2 * w[i];
  const float nv = mom * v[i] - lr * grad;
  v[i] = nv;
  w[i] += nv;
}
__global__ void sgd_kernel_flat(float* __restrict__ w, float* __restrict__ v, const float* __restrict__ g, size_t n, float lr, float mom,
                                float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float nv = mom * v[i] - lr * g[i] * scale;
  v[i] = nv;
  w[i] += nv;
}

// master HWIO -> packed K-major bf16 hi/lo planes.  mode 0: forward virtual path [cout][tap][kblocks*64];
// mode 1: forward im2col path [cout][k = tap*cin + c]; mode 2: data-gradient kernel [cin][taps-1-tap][kb2*64 over cout];
// mode 3: [tap*cin + c][cout] (col-gradient GEMM of strided convolutions).
__global__ void repack_kernel(const float* __restrict__ w, int taps, int cin, int cout, int mode, int kblocks, size_t krow, int rows,
                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)rows * krow;
  if (i >= total) return;
  const int row = (int)(i / krow); const size_t k = i % krow;
  float val = 0.f;
  if (mode == 0) {
    const int t = (int)(k / ((size_t)kblocks * 64)), c = (int)(k % ((size_t)kblocks * 64));
    if (t < taps && c < cin) val = w[((size_t)t * cin + c) * cout + row];
  } else if (mode == 1) {
    if (k < (size_t)taps * cin) val = w[k * cout + row];
  } else if (mode == 2) {
    const int t2 = (int)(k / ((size_t)kblocks * 64)), co = (int)(k % ((size_t)kblocks * 64));
    if (t2 < taps && co < cout) val = w[((size_t)(taps - 1 - t2) * cin + row) * cout + co];
  } else {                      // mode 3: W^T for the col-gradient GEMM of strided convs: [k = tap*cin + c][cout]
    if (k < (size_t)cout) val = w[(size_t)row * cout + k];
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(val);
  hi[i] = h;
  if (lo) lo[i] = __float2bfloat16_rn(val - __bfloat162float(h));
}

__global__ void hwio_to_ohwi_kernel(const float* __restrict__ w, int taps, int cin, int cout, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)taps * cin * cout;
  if (i >= total) return;
  const int ci = (int)(i % cin); const size_t r = i / cin; const int t = (int)(r % taps); const int co = (int)(r / taps);
  out[i] = w[((size_t)t * cin + ci) * cout + co];
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
struct TLayer {
  int li = -1, op = -1;
  // parameters (conv / head): master kernel HWIO in L.w_f32, bias in L.bias (shared with the forward plan)
  int cin = 0, cout = 0, taps = 0;
  float *vw = nullptr, *vb = nullptr;       // momentum
  long long off_w = -1, off_b = -1, off_g = -1, off_bng = -1, off_bnb = -1;
  float *v2w = nullptr, *v2b = nullptr, *v2gamma = nullptr;     // Adam second moments (vw / vb / vgamma hold the first moments)
  float *m_bng = nullptr, *v_bng = nullptr, *m_bnb = nullptr, *v_bnb = nullptr;   // BatchNormalization gamma / beta optimiser state
  ActBuf g;                                 // gradient of this layer's output (pre-activation); same geometry as the output
  bool has_g = false;
  // data gradient
  bool has_dgrad = false;
  ConvLaunch dgrad{};
  __nv_bfloat16 *w2_hi = nullptr, *w2_lo = nullptr; size_t w2_krow = 0; int w2_kblocks = 0;
  int* dgrad_tiles = nullptr;
  bool dgrad_strided = false;               // stride != 1: col-gradient GEMM (dZ * W^T) + col2im
  float* dcol = nullptr; int dcol_ld = 0;
  // weight gradient
  bool wg_native = false;                   // stride-1 layers: wgrad.cu reads dZ / X in place (no transposed copies)
  WgradLaunch wg;
  std::vector<ConvLaunch> wgrad;            // fallback: one GEMM per tap (or one for the im2col path) on transposed operands
  int Wq = 0;                               // row pitch of the transposed operands' pixel grid
  std::vector<int> wgrad_res;               // tap shift mod 8: TMA needs 16-byte aligned K offsets, so XT is built once per residue
  long long Kv = 0, ldT = 0;
  TMap dy_map{}, x_map{};
  float* vgamma = nullptr;
};

}  // namespace

struct ssdk_trainer {
  ssdk_model* m = nullptr;
  std::vector<TLayer> tl;
  float* grad = nullptr; bool own_grad = false;
  long long n_params = 0;
  __nv_bfloat16 *xT_hi = nullptr, *xT_lo = nullptr, *dyT_hi = nullptr, *dyT_lo = nullptr;   // scratch, sized for the largest layer
  long long xT_elems = 0, dyT_elems = 0;
  float* dypred = nullptr;                  // [B*P*(C+12)]
  std::vector<char> written;                // backward pass state: which activations already hold a partial gradient
  std::vector<void*> allocs;
};

namespace {

template <typename T>
int t_alloc(ssdk_trainer* t, T** out, size_t count, bool zero) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(count * sizeof(T), 16);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return SSDK_ERR_NOMEM; }
  if (zero) cudaMemset(p, 0, bytes);
  t->allocs.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return SSDK_OK;
}

int launch_repack(ssdk_ctx* ctx, const float* w, int taps, int cin, int cout, int mode, int kblocks, size_t krow, int rows,
                  __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t s) {
  const size_t total = (size_t)rows * krow;
  repack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(w, taps, cin, cout, mode, kblocks, krow, rows, hi, lo);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

bool is_conv(int op) { return op == SSDK_OP_CONV || op == SSDK_OP_HEAD; }

}  // namespace

extern "C" int ssdk_trainer_create(ssdk_model* m, float* flat_grad_dev, ssdk_trainer** out) {
  SSDK_REQUIRE(m && out, "ssdk_trainer_create: NULL argument");
  SSDK_REQUIRE(m->training, "ssdk_trainer_create: the model plan was not created with training=1");
  SSDK_CHECK_CUDA(cudaSetDevice(m->ctx->device));
  ssdk_trainer* t = new ssdk_trainer();
  t->m = m;
  const int n = (int)m->layers.size();
  t->tl.resize(n);
  int rc = SSDK_OK;
  auto fail = [&](int code) { ssdk_trainer_destroy(t); return code; };
  // parameter spans
  long long off = 0;
  for (int i = 0; i < n; ++i) {
    LayerPlan& L = m->layers[i];
    TLayer& T = t->tl[i];
    T.li = i; T.op = L.d.op;
    if (is_conv(L.d.op)) {
      if ((L.d.act == SSDK_ACT_ELU || L.bn_scale) && !L.bn_train) {
        set_error("training: layer %d has an ELU / folded BatchNormalization without the raw BatchNormalization parameters (bn_gamma, ...)", i);
        return fail(SSDK_ERR_UNSUPPORTED);
      }
      T.cin = m->layers[L.d.input].C; T.cout = L.C; T.taps = L.d.kh * L.d.kw;
      T.off_w = off; off += (long long)T.cout * T.taps * T.cin;
      T.off_b = off; off += T.cout;
      if (L.bn_train) { T.off_bng = off; off += T.cout; T.off_bnb = off; off += T.cout; }
    } else if (L.d.op == SSDK_OP_L2NORM) {
      T.off_g = off; off += L.C;
    }
  }
  t->n_params = off;
  if (flat_grad_dev) t->grad = flat_grad_dev;
  else { rc = t_alloc(t, &t->grad, (size_t)off, true); if (rc) return fail(rc); t->own_grad = true; }
  rc = t_alloc(t, &t->dypred, (size_t)m->B * m->P * (m->Ctot + 12), true); if (rc) return fail(rc);
  // gradient buffers: same geometry as the forward outputs
  for (int i = 0; i < n; ++i) {
    LayerPlan& L = m->layers[i];
    TLayer& T = t->tl[i];
    if (L.d.op == SSDK_OP_INPUT) continue;
    ActBuf& g = T.g;
    if (L.d.op == SSDK_OP_HEAD) { g.B = m->B; g.H = L.H; g.W = L.W; g.C = L.C; g.Cs = (L.C + 7) / 8 * 8; g.pad = 1; }
    else { g = L.out; g.hi = nullptr; g.lo = nullptr; }
    const size_t ne = g.elems() + 64 * 8;
    rc = t_alloc(t, &g.hi, ne, true); if (rc) return fail(rc);
    if (m->split) { rc = t_alloc(t, &g.lo, ne, true); if (rc) return fail(rc); }
    T.has_g = true;
  }
  // per conv: momentum, data-gradient plan, weight-gradient plans
  std::vector<char> written(n, 0);            // does the producer's gradient buffer already hold a contribution?
  long long max_xT = 0, max_dyT = 0;
  for (int i = n - 1; i >= 0; --i) {
    LayerPlan& L = m->layers[i];
    TLayer& T = t->tl[i];
    const ssdk_layer_desc& d = L.d;
    if (d.op == SSDK_OP_INPUT) continue;
    const int pi = d.input;
    LayerPlan& PL = m->layers[pi];
    TLayer& PT = t->tl[pi];
    const bool prod_needs_grad = PL.d.op != SSDK_OP_INPUT;
    if (d.op == SSDK_OP_L2NORM) { rc = t_alloc(t, &T.vgamma, (size_t)L.C, true); if (rc) return fail(rc); }
    if (!is_conv(d.op)) { if (prod_needs_grad) written[pi] = 1; continue; }
    rc = t_alloc(t, &T.vw, (size_t)T.cout * T.taps * T.cin, true); if (rc) return fail(rc);
    rc = t_alloc(t, &T.vb, (size_t)T.cout, true); if (rc) return fail(rc);
    if (L.bn_train) {
      rc = t_alloc(t, &T.m_bng, (size_t)T.cout, true); if (rc) return fail(rc);
      rc = t_alloc(t, &T.v_bng, (size_t)T.cout, true); if (rc) return fail(rc);
      rc = t_alloc(t, &T.m_bnb, (size_t)T.cout, true); if (rc) return fail(rc);
      rc = t_alloc(t, &T.v_bnb, (size_t)T.cout, true); if (rc) return fail(rc);
    }
    // ---- data gradient
    if (prod_needs_grad) {
      if (T.cin % 8 != 0) { set_error("training: input channels must be a multiple of 8 (layer %d)", i); return fail(SSDK_ERR_UNSUPPORTED); }
      if (d.stride != 1) {
        // dCol[M][taps*cin] = dZ[M][cout] * W^T, then col2im
        T.has_dgrad = true; T.dgrad_strided = true;
        const int kcol = T.taps * T.cin;
        T.w2_kblocks = (T.g.Cs + 63) / 64;
        T.w2_krow = (size_t)T.w2_kblocks * 64;
        rc = t_alloc(t, &T.w2_hi, (size_t)kcol * T.w2_krow, true); if (rc) return fail(rc);
        if (m->split) { rc = t_alloc(t, &T.w2_lo, (size_t)kcol * T.w2_krow, true); if (rc) return fail(rc); }
        T.dcol_ld = kcol;
        rc = t_alloc(t, &T.dcol, (size_t)m->B * L.H * L.W * kcol, true); if (rc) return fail(rc);
        ConvGeom gg;
        gg.in = &T.g; gg.kh = 1; gg.kw = 1; gg.dilation = 1; gg.pad_t = 0; gg.pad_l = 0;
        gg.Ho = L.H; gg.Wo = L.W; gg.B = m->B; gg.cout = kcol;
        rc = plan_conv_gemm(m, T.dgrad, gg, T.w2_hi, T.w2_lo, T.w2_krow, T.w2_kblocks, (T.g.Cs - (T.w2_kblocks - 1) * 64 + 15) / 16, &T.dgrad_tiles);
        if (rc) return fail(rc);
        ConvArgs& a = T.dgrad.args;
        a.epi = EPI_F32; a.bias = nullptr; a.act = SSDK_ACT_NONE; a.out_f32 = T.dcol;
        written[pi] = 1;
        goto wgrad_plan;
      }
      T.has_dgrad = true;
      T.w2_kblocks = (T.g.Cs + 63) / 64;
      T.w2_krow = (size_t)T.taps * T.w2_kblocks * 64;
      rc = t_alloc(t, &T.w2_hi, (size_t)T.cin * T.w2_krow, true); if (rc) return fail(rc);
      if (m->split) { rc = t_alloc(t, &T.w2_lo, (size_t)T.cin * T.w2_krow, true); if (rc) return fail(rc); }
      ConvGeom gg;
      gg.in = &T.g; gg.kh = d.kh; gg.kw = d.kw; gg.dilation = d.dilation;
      gg.pad_t = d.dilation * (d.kh - 1) - d.pad_t; gg.pad_l = d.dilation * (d.kw - 1) - d.pad_l;
      gg.Ho = PL.H; gg.Wo = PL.W; gg.B = m->B; gg.cout = T.cin;
      rc = plan_conv_gemm(m, T.dgrad, gg, T.w2_hi, T.w2_lo, T.w2_krow, T.w2_kblocks, (T.g.Cs - (T.w2_kblocks - 1) * 64 + 15) / 16, &T.dgrad_tiles);
      if (rc) return fail(rc);
      ConvArgs& a = T.dgrad.args;
      a.epi = EPI_SPLIT; a.bias = nullptr; a.act = SSDK_ACT_NONE;
      a.out_hi = PT.g.hi; a.out_lo = PT.g.lo; a.out_Hp = PT.g.Hp(); a.out_Wp = PT.g.Wp(); a.out_pad = PT.g.pad; a.out_Cs = PT.g.Cs;
      a.mask_hi = (PL.d.op == SSDK_OP_CONV && PL.d.act == SSDK_ACT_RELU) ? PL.out.hi : nullptr;
      a.accumulate = written[pi] ? 1 : 0;
      written[pi] = 1;
    }
    // ---- weight gradient
    wgrad_plan:
    if (L.direct) continue;                     // handled by wgrad_direct_kernel
    const ActBuf& X = PL.out;
    long long Kv;
    if (!L.im2col && !getenv("SSDK_WGRAD_TRANSPOSED") && wgrad_supported(X, T.g, d.kh, d.kw, d.stride, d.dilation)) {
      T.wg_native = true;
      rc = plan_wgrad(m->ctx, T.wg, X, T.g, L.H, L.W, d.kh, d.kw, d.dilation, d.pad_t, d.pad_l, m->split, nullptr);
      if (rc) return fail(rc);
      continue;
    }
    if (L.im2col) {
      Kv = (long long)m->B * L.H * L.W;
      T.dy_map = TMap{0, L.H * L.W, L.W, L.H, L.W, T.g.Hp(), T.g.Wp(), T.g.pad};
      T.x_map = TMap{1, 0, 0, 0, 0, 0, 0, 0};
      max_xT = std::max(max_xT, (long long)L.Kpad * ((Kv + 7) / 8 * 8 + 64));
    } else {
      // the GEMM's pixel axis runs over X's padded grid with the row pitch rounded up to 8 elements: tap shifts are then
      // kh*Wq + kw (+ const) and only the kw part can break TMA's 16-byte alignment -> one XT copy per distinct kw residue
      const int Wq = (X.Wp() + 7) / 8 * 8;
      T.Wq = Wq;
      Kv = (long long)m->B * X.Hp() * Wq;
      T.dy_map = TMap{0, X.Hp() * Wq, Wq, L.H, L.W, T.g.Hp(), T.g.Wp(), T.g.pad};
      T.x_map = TMap{0, X.Hp() * Wq, Wq, X.Hp(), X.Wp(), X.Hp(), X.Wp(), 0};
      max_xT = std::max(max_xT, (long long)X.Cs * ((Kv + 7) / 8 * 8 + 64));
    }
    T.Kv = Kv; T.ldT = (Kv + 7) / 8 * 8 + 64;
    max_dyT = std::max(max_dyT, (long long)T.g.Cs * T.ldT);
  }
  t->xT_elems = max_xT; t->dyT_elems = max_dyT;
  rc = t_alloc(t, &t->xT_hi, (size_t)max_xT + 4096, true); if (rc) return fail(rc);
  rc = t_alloc(t, &t->dyT_hi, (size_t)max_dyT + 4096, true); if (rc) return fail(rc);
  if (m->split) {
    rc = t_alloc(t, &t->xT_lo, (size_t)max_xT + 4096, true); if (rc) return fail(rc);
    rc = t_alloc(t, &t->dyT_lo, (size_t)max_dyT + 4096, true); if (rc) return fail(rc);
  }
  // weight-gradient GEMM plans (need the scratch addresses)
  for (int i = 0; i < n; ++i) {
    LayerPlan& L = m->layers[i];
    TLayer& T = t->tl[i];
    if (!is_conv(L.d.op) || L.direct) continue;
    if (T.wg_native) { T.wg.args.dw = t->grad + T.off_w; continue; }
    const ssdk_layer_desc& d = L.d;
    const ActBuf& X = m->layers[d.input].out;
    const int n_gemm = L.im2col ? 1 : T.taps;
    const int ncols = L.im2col ? L.Kpad : T.cin;            // N of the GEMM
    const int kblocks = (int)((T.Kv + 63) / 64);
    const int last_ksteps = (int)((T.Kv - (long long)(kblocks - 1) * 64 + 15) / 16);
    T.wgrad.resize(n_gemm); T.wgrad_res.assign(n_gemm, 0);
    for (int tp = 0; tp < n_gemm; ++tp) {
      ConvGeom wg;
      wg.a_hi = t->dyT_hi; wg.a_lo = t->dyT_lo; wg.a_inner = (uint64_t)T.ldT; wg.a_ld = (uint64_t)T.ldT; wg.a_rows = (uint64_t)T.cout;
      wg.Ho = T.cout; wg.Wo = 1; wg.B = 1; wg.cout = ncols;
      ConvLaunch& cl = T.wgrad[tp];
      // the transposed operands are [channels][ldT] with zeros in [Kv, ldT)
      rc = plan_conv_gemm(m, cl, wg, t->xT_hi, t->xT_lo, (size_t)T.ldT, kblocks, last_ksteps, nullptr);
      if (rc) return fail(rc);
      ConvArgs& a = cl.args;
      a.epi = EPI_ATOMIC; a.bias = nullptr; a.act = SSDK_ACT_NONE;
      a.out_f32 = t->grad + T.off_w;
      if (L.im2col) { a.out_ld = T.taps * T.cin; a.out_col_off = 0; a.b_k_offset = 0; a.cout = T.taps * T.cin; a.n_tiles_n = (a.cout + a.BN - 1) / a.BN; }
      else {
        const int kh = tp / d.kw, kw = tp % d.kw;
        a.out_ld = T.taps * T.cin; a.out_col_off = tp * T.cin;
        const int shift = (kh * d.dilation - d.pad_t + X.pad) * T.Wq + (kw * d.dilation - d.pad_l + X.pad);
        T.wgrad_res[tp] = shift & 7;                 // XT_r[c][v] = X[v + r][c], read at the aligned offset shift - r
        a.b_k_offset = shift - (shift & 7);
      }
      const int units = a.n_tiles_m * a.n_tiles_n;
      int ks = std::max(1, (4 * m->ctx->sm_count + units - 1) / units);
      ks = std::min(ks, kblocks);
      a.kb_per = (kblocks + ks - 1) / ks;
      a.k_split = (kblocks + a.kb_per - 1) / a.kb_per;
      cl.grid = std::max(1, std::min(units * a.k_split, m->ctx->sm_count));
    }
  }
  // rotated kernels for the data gradients
  for (int i = 0; i < n; ++i) {
    TLayer& T = t->tl[i];
    if (!T.has_dgrad) continue;
    if (T.dgrad_strided) rc = launch_repack(m->ctx, m->layers[i].w_f32, T.taps, T.cin, T.cout, 3, T.w2_kblocks, T.w2_krow, T.taps * T.cin, T.w2_hi, T.w2_lo, 0);
    else rc = launch_repack(m->ctx, m->layers[i].w_f32, T.taps, T.cin, T.cout, 2, T.w2_kblocks, T.w2_krow, T.cin, T.w2_hi, T.w2_lo, 0);
    if (rc) return fail(rc);
  }
  SSDK_CHECK_CUDA(cudaDeviceSynchronize());
  *out = t;
  return SSDK_OK;
}

extern "C" int ssdk_trainer_destroy(ssdk_trainer* t) {
  if (!t) return SSDK_OK;
  for (void* p : t->allocs) cudaFree(p);
  delete t;
  return SSDK_OK;
}

extern "C" int ssdk_trainer_num_params(const ssdk_trainer* t, long long* out_n) {
  SSDK_REQUIRE(t && out_n, "ssdk_trainer_num_params: NULL argument");
  *out_n = t->n_params;
  return SSDK_OK;
}

extern "C" int ssdk_trainer_param_span(const ssdk_trainer* t, int layer, int which, long long* out_offset, long long* out_count) {
  SSDK_REQUIRE(t && out_offset && out_count && layer >= 0 && layer < (int)t->tl.size(), "ssdk_trainer_param_span: bad argument");
  const TLayer& T = t->tl[layer];
  *out_offset = -1; *out_count = 0;
  if (which == 0 && T.off_w >= 0) { *out_offset = T.off_w; *out_count = (long long)T.cout * T.taps * T.cin; }
  else if (which == 1 && T.off_b >= 0) { *out_offset = T.off_b; *out_count = T.cout; }
  else if (which == 2 && T.off_g >= 0) { *out_offset = T.off_g; *out_count = t->m->layers[layer].C; }
  else if (which == 3 && T.off_bng >= 0) { *out_offset = T.off_bng; *out_count = T.cout; }
  else if (which == 4 && T.off_bnb >= 0) { *out_offset = T.off_bnb; *out_count = T.cout; }
  return SSDK_OK;
}

extern "C" float* ssdk_trainer_grad_buffer(ssdk_trainer* t) { return t ? t->grad : nullptr; }

namespace {

int do_transpose(ssdk_trainer* t, const __nv_bfloat16* hi, const __nv_bfloat16* lo, int src_ld, long long src_rows, const TMap& mp,
                 long long Kv, int C, __nv_bfloat16* dhi, __nv_bfloat16* dlo, long long ldT, cudaStream_t s, long long row_off = 0) {
  dim3 grid((unsigned)((ldT + 63) / 64), (unsigned)((C + 63) / 64));
  transpose_kernel<<<grid, 256, 0, s>>>(hi, lo, src_ld, src_rows, mp, row_off, Kv, C, dhi, dlo, ldT);
  SSDK_COUNT_LAUNCH(t->m->ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

}  // namespace

namespace {
int backward_layers(ssdk_trainer* t, const float* dypred, int hi, int lo, cudaStream_t s);
}

// Loss (optional) and d loss / d y_pred from one launch of the fused loss kernel; clears the gradient buffer.
extern "C" int ssdk_train_backward_begin(ssdk_trainer* t, const float* y_true, const float* y_pred, int neg_pos_ratio, int n_neg_min,
                                         float alpha, float* out_loss, void* stream_) {
  SSDK_REQUIRE(t && y_true && y_pred, "ssdk_train_backward_begin: NULL argument");
  ssdk_model* m = t->m;
  ssdk_ctx* ctx = m->ctx;
  int rc;
  SSDK_CHECK_CUDA(cudaMemsetAsync(t->grad, 0, (size_t)t->n_params * sizeof(float), (cudaStream_t)stream_));
  if (out_loss) rc = ssdk_ssd_loss_fwd_bwd(ctx, y_true, y_pred, m->B, m->P, m->Ctot, neg_pos_ratio, n_neg_min, alpha, nullptr, out_loss, nullptr, t->dypred, stream_);
  else rc = ssdk_ssd_loss_bwd(ctx, y_true, y_pred, m->B, m->P, m->Ctot, neg_pos_ratio, n_neg_min, alpha, nullptr, t->dypred, stream_);
  return rc;
}

// The backward pass of layers hi, hi-1, ..., lo (graph indices; a full pass is n_layers-1 .. 0 in one or several calls, top
// down).  When it returns (stream order), the parameter gradients of exactly these layers are final -- their span of the
// flat buffer can be all-reduced while the lower layers are still being differentiated.  dypred_dev NULL: the gradient
// ssdk_train_backward_begin left in the trainer; otherwise d loss / d y_pred provided by the caller (the buffer is cleared
// when the pass starts at the top layer).
extern "C" int ssdk_train_backward_layers(ssdk_trainer* t, const float* dypred_dev, int hi, int lo, void* stream_) {
  SSDK_REQUIRE(t, "ssdk_train_backward_layers: NULL trainer");
  const int n = (int)t->m->layers.size();
  SSDK_REQUIRE(hi < n && lo >= 0 && lo <= hi, "ssdk_train_backward_layers: bad layer range [%d, %d] of %d layers", lo, hi, n);
  if (dypred_dev && hi == n - 1) SSDK_CHECK_CUDA(cudaMemsetAsync(t->grad, 0, (size_t)t->n_params * sizeof(float), (cudaStream_t)stream_));
  return backward_layers(t, dypred_dev ? dypred_dev : t->dypred, hi, lo, (cudaStream_t)stream_);
}

extern "C" int ssdk_train_backward(ssdk_trainer* t, const float* y_true, const float* y_pred, int neg_pos_ratio, int n_neg_min,
                                   float alpha, float* out_loss, void* stream_) {
  int rc = ssdk_train_backward_begin(t, y_true, y_pred, neg_pos_ratio, n_neg_min, alpha, out_loss, stream_);
  if (rc) return rc;
  return backward_layers(t, t->dypred, (int)t->m->layers.size() - 1, 0, (cudaStream_t)stream_);
}

extern "C" int ssdk_train_backward_dy(ssdk_trainer* t, const float* dypred_dev, void* stream_) {
  SSDK_REQUIRE(t && dypred_dev, "ssdk_train_backward_dy: NULL argument");
  return ssdk_train_backward_layers(t, dypred_dev, (int)t->m->layers.size() - 1, 0, stream_);
}

namespace {

// parameter gradients of layers hi .. lo from d loss / d y_pred (B,P,C+12), walking the layers in reverse
int backward_layers(ssdk_trainer* t, const float* dypred, int hi, int lo, cudaStream_t s) {
  ssdk_model* m = t->m;
  ssdk_ctx* ctx = m->ctx;
  int rc;
  const int n = (int)m->layers.size();
  if (hi == n - 1) t->written.assign(n, 0);
  SSDK_REQUIRE((int)t->written.size() == n, "ssdk_train_backward_layers: a pass must start at the top layer");
  std::vector<char>& written = t->written;
  for (int i = hi; i >= lo; --i) {
    LayerPlan& L = m->layers[i];
    TLayer& T = t->tl[i];
    const ssdk_layer_desc& d = L.d;
    if (d.op == SSDK_OP_INPUT) continue;
    const int pi = d.input;
    LayerPlan& PL = m->layers[pi];
    TLayer& PT = t->tl[pi];
    const bool prod_needs_grad = PL.d.op != SSDK_OP_INPUT;
    const int relu_mask = (PL.d.op == SSDK_OP_CONV && PL.d.act == SSDK_ACT_RELU) ? 1 : 0;
    if (d.op == SSDK_OP_HEAD) {
      const size_t total = (size_t)m->B * L.H * L.W * d.n_boxes;
      head_bwd_kernel<<<(unsigned)((total + 7) / 8), 256, 0, s>>>(L.head_f32, dypred, m->B, L.H, L.W, d.n_boxes, m->Ctot, m->P, L.prior_off, T.g);
      SSDK_COUNT_LAUNCH(ctx);
    }
    if (d.op == SSDK_OP_MAXPOOL) {
      if (prod_needs_grad) {
        if (PL.out.C % 8) { set_error("training: max-pool channels must be a multiple of 8"); return SSDK_ERR_UNSUPPORTED; }
        const size_t total = (size_t)PL.out.B * PL.out.H * PL.out.W * (PL.out.C / 8);
        pool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(PL.out, T.g, PT.g, L.H, L.W, d.kh, d.kw, d.stride, d.pad_t, d.pad_l,
                                                                         relu_mask, written[pi] ? 1 : 0);
        SSDK_COUNT_LAUNCH(ctx);
        written[pi] = 1;
      }
      continue;
    }
    if (d.op == SSDK_OP_L2NORM) {
      if (prod_needs_grad) {
        const size_t total = (size_t)PL.out.B * PL.out.H * PL.out.W;
        l2norm_bwd_kernel<<<(unsigned)((total + 7) / 8), 256, 0, s>>>(PL.out, T.g, PT.g, L.gamma, t->grad + T.off_g, relu_mask, written[pi] ? 1 : 0);
        SSDK_COUNT_LAUNCH(ctx);
        written[pi] = 1;
      }
      continue;
    }
    // ---- conv + BatchNormalization + activation: the gradient planes hold d loss / d activation; turn them into the gradient of
    //      the raw conv output (and produce dgamma / dbeta) before the conv's own gradients are formed
    if (L.bn_train) { rc = launch_bn_backward(ctx, L, d.act, T.g, t->grad + T.off_bng, t->grad + T.off_bnb, s); if (rc) return rc; }
    // ---- convolution / head: weight + bias gradients
    if (L.direct) {
      const int K = T.taps * T.cin;
      const size_t smem = (size_t)K * T.cout * sizeof(float);
      const int total_rows = T.g.B * T.g.H;
      const int rpb = std::max(1, (total_rows + 8 * ctx->sm_count - 1) / (8 * ctx->sm_count));
      const bool fast3 = d.kh == 3 && d.kw == 3 && d.dilation == 1 && T.cin == 3 && PL.out.Cs >= 4 && T.cout % 2 == 0 && T.cout <= 512 &&
                         PL.out.pad >= d.pad_t && PL.out.pad >= d.pad_l && PL.out.pad >= 2 - d.pad_t && PL.out.pad >= 2 - d.pad_l;
      if (fast3) wgrad_direct3x3_kernel<<<(unsigned)((total_rows + rpb - 1) / rpb), 256, smem, s>>>(PL.out, T.g, t->grad + T.off_w, d.pad_t, d.pad_l, rpb);
      else wgrad_direct_kernel<<<(unsigned)((total_rows + rpb - 1) / rpb), 256, smem, s>>>(PL.out, T.g, t->grad + T.off_w, d.kh, d.kw, d.dilation, d.pad_t, d.pad_l, rpb);
      SSDK_COUNT_LAUNCH(ctx);
      rc = launch_bias_grad(ctx, T.g, t->grad + T.off_b, s); if (rc) return rc;
    } else if (T.wg_native) {
      rc = launch_wgrad(ctx, T.wg, s); if (rc) return rc;
      rc = launch_bias_grad(ctx, T.g, t->grad + T.off_b, s); if (rc) return rc;
    } else {
      // transposed operands
      rc = do_transpose(t, T.g.hi, T.g.lo, T.g.Cs, (long long)T.g.rows(), T.dy_map, T.Kv, T.cout, t->dyT_hi, t->dyT_lo, T.ldT, s); if (rc) return rc;
      if (L.im2col) {
        rc = do_transpose(t, L.col_hi, L.col_lo, L.Kpad, T.Kv, T.x_map, T.Kv, L.Kpad, t->xT_hi, t->xT_lo, T.ldT, s); if (rc) return rc;
        rc = launch_conv(ctx, T.wgrad[0], s); if (rc) return rc;
      } else {
        for (int r = 0; r < 8; ++r) {
          bool any = false;
          for (size_t tp = 0; tp < T.wgrad.size(); ++tp) any |= (T.wgrad_res[tp] == r);
          if (!any) continue;
          rc = do_transpose(t, PL.out.hi, PL.out.lo, PL.out.Cs, (long long)PL.out.rows(), T.x_map, T.Kv, T.cin, t->xT_hi, t->xT_lo, T.ldT, s, r);
          if (rc) return rc;
          for (size_t tp = 0; tp < T.wgrad.size(); ++tp)
            if (T.wgrad_res[tp] == r) { rc = launch_conv(ctx, T.wgrad[tp], s); if (rc) return rc; }
        }
      }
      dim3 gr(64, T.cout);
      rowsum_kernel<<<gr, 256, 0, s>>>(t->dyT_hi, t->dyT_lo, T.ldT, T.Kv, t->grad + T.off_b);
      SSDK_COUNT_LAUNCH(ctx);
    }
    // ---- data gradient
    if (T.has_dgrad && T.dgrad_strided) {
      rc = launch_conv(ctx, T.dgrad, s); if (rc) return rc;
      const size_t total = (size_t)PT.g.B * PT.g.H * PT.g.W * PT.g.C;
      col2im_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(T.dcol, L.H, L.W, d.kh, d.kw, d.stride, d.dilation, d.pad_t, d.pad_l, T.dcol_ld,
                                                                     PL.out, PT.g, relu_mask, written[pi] ? 1 : 0);
      SSDK_COUNT_LAUNCH(ctx);
      written[pi] = 1;
    } else if (T.has_dgrad) {
      T.dgrad.args.accumulate = written[pi] ? 1 : 0;
      rc = launch_conv(ctx, T.dgrad, s); if (rc) return rc;
      written[pi] = 1;
    }
  }
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

}  // namespace

extern "C" int ssdk_train_apply(ssdk_trainer* t, float lr, float momentum, float l2_reg, float grad_scale, void* stream_) {
  SSDK_REQUIRE(t, "ssdk_train_apply: NULL trainer");
  ssdk_model* m = t->m;
  ssdk_ctx* ctx = m->ctx;
  cudaStream_t s = (cudaStream_t)stream_;
  int rc;
  for (size_t i = 0; i < t->tl.size(); ++i) {
    TLayer& T = t->tl[i];
    LayerPlan& L = m->layers[i];
    if (T.off_g >= 0) {
      sgd_kernel_flat<<<(unsigned)((L.C + 255) / 256), 256, 0, s>>>(L.gamma, T.vgamma, t->grad + T.off_g, (size_t)L.C, lr, momentum, grad_scale);
      SSDK_COUNT_LAUNCH(ctx);
    }
    if (T.off_w < 0) continue;
    const size_t nw = (size_t)T.taps * T.cin * T.cout;
    sgd_kernel_w<<<(unsigned)((nw + 255) / 256), 256, 0, s>>>(L.w_f32, T.vw, t->grad + T.off_w, T.taps, T.cin, T.cout, lr, momentum, l2_reg, grad_scale);
    SSDK_COUNT_LAUNCH(ctx);
    sgd_kernel_flat<<<(unsigned)((T.cout + 255) / 256), 256, 0, s>>>(L.bias, T.vb, t->grad + T.off_b, (size_t)T.cout, lr, momentum, grad_scale);
    SSDK_COUNT_LAUNCH(ctx);
    if (T.off_bng >= 0) {
      sgd_kernel_flat<<<(unsigned)((T.cout + 255) / 256), 256, 0, s>>>(L.bn_gamma, T.m_bng, t->grad + T.off_bng, (size_t)T.cout, lr, momentum, grad_scale);
      sgd_kernel_flat<<<(unsigned)((T.cout + 255) / 256), 256, 0, s>>>(L.bn_beta, T.m_bnb, t->grad + T.off_bnb, (size_t)T.cout, lr, momentum, grad_scale);
      SSDK_COUNT_LAUNCH(ctx); SSDK_COUNT_LAUNCH(ctx);
    }
    if (!L.direct) {
      rc = launch_repack(ctx, L.w_f32, T.taps, T.cin, T.cout, L.im2col ? 1 : 0, L.kblocks, L.w_krow, T.cout, L.w_hi, L.w_lo, s); if (rc) return rc;
    }
    if (T.has_dgrad) {
      if (T.dgrad_strided) rc = launch_repack(ctx, L.w_f32, T.taps, T.cin, T.cout, 3, T.w2_kblocks, T.w2_krow, T.taps * T.cin, T.w2_hi, T.w2_lo, s);
      else rc = launch_repack(ctx, L.w_f32, T.taps, T.cin, T.cout, 2, T.w2_kblocks, T.w2_krow, T.cin, T.w2_hi, T.w2_lo, s);
      if (rc) return rc;
    }
  }
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

namespace {

__global__ void adam_kernel_w(float* __restrict__ w, float* __restrict__ m1, float* __restrict__ m2, const float* __restrict__ g, int taps, int cin,
                              int cout, float lr_t, float b1, float b2, float eps, float l2, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)taps * cin * cout;
  if (i >= total) return;
  const int co = (int)(i % cout); const size_t r = i / cout; const int ci = (int)(r % cin); const int t = (int)(r / cin);
  const float grad = g[((size_t)co * taps + t) * cin + ci] * scale + 2.f * l2 * w[i];
  const float a = b1 * m1[i] + (1.f - b1) * grad;
  const float b = b2 * m2[i] + (1.f - b2) * grad * grad;
  m1[i] = a; m2[i] = b;
  w[i] -= lr_t * a / (sqrtf(b) + eps);
}
__global__ void adam_kernel_flat(float* __restrict__ w, float* __restrict__ m1, float* __restrict__ m2, const float* __restrict__ g, size_t n,
                                 float lr_t, float b1, float b2, float eps, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float grad = g[i] * scale;
  const float a = b1 * m1[i] + (1.f - b1) * grad;
  const float b = b2 * m2[i] + (1.f - b2) * grad * grad;
  m1[i] = a; m2[i] = b;
  w[i] -= lr_t * a / (sqrtf(b) + eps);
}

int repack_after_update(ssdk_trainer* t, int i, cudaStream_t s) {
  ssdk_model* m = t->m;
  TLayer& T = t->tl[i];
  LayerPlan& L = m->layers[i];
  int rc;
  if (!L.direct) {
    rc = launch_repack(m->ctx, L.w_f32, T.taps, T.cin, T.cout, L.im2col ? 1 : 0, L.kblocks, L.w_krow, T.cout, L.w_hi, L.w_lo, s); if (rc) return rc;
  }
  if (T.has_dgrad) {
    if (T.dgrad_strided) rc = launch_repack(m->ctx, L.w_f32, T.taps, T.cin, T.cout, 3, T.w2_kblocks, T.w2_krow, T.taps * T.cin, T.w2_hi, T.w2_lo, s);
    else rc = launch_repack(m->ctx, L.w_f32, T.taps, T.cin, T.cout, 2, T.w2_kblocks, T.w2_krow, T.cin, T.w2_hi, T.w2_lo, s);
    if (rc) return rc;
  }
  return SSDK_OK;
}

}  // namespace

extern "C" int ssdk_train_apply_adam(ssdk_trainer* t, float lr, float beta1, float beta2, float eps, float l2_reg, float grad_scale, int step,
                                     void* stream_) {
  SSDK_REQUIRE(t && step >= 1, "ssdk_train_apply_adam: bad argument");
  ssdk_model* m = t->m;
  ssdk_ctx* ctx = m->ctx;
  cudaStream_t s = (cudaStream_t)stream_;
  // Keras: lr_t = lr * sqrt(1 - beta_2^t) / (1 - beta_1^t)
  const float lr_t = lr * (float)(std::sqrt(1.0 - std::pow((double)beta2, (double)step)) / (1.0 - std::pow((double)beta1, (double)step)));
  int rc;
  for (size_t i = 0; i < t->tl.size(); ++i) {
    TLayer& T = t->tl[i];
    LayerPlan& L = m->layers[i];
    if (T.off_g >= 0) {
      if (!T.v2gamma) { rc = t_alloc(t, &T.v2gamma, (size_t)L.C, true); if (rc) return rc; }
      adam_kernel_flat<<<(unsigned)((L.C + 255) / 256), 256, 0, s>>>(L.gamma, T.vgamma, T.v2gamma, t->grad + T.off_g, (size_t)L.C, lr_t, beta1, beta2, eps, grad_scale);
      SSDK_COUNT_LAUNCH(ctx);
    }
    if (T.off_w < 0) continue;
    const size_t nw = (size_t)T.taps * T.cin * T.cout;
    if (!T.v2w) { rc = t_alloc(t, &T.v2w, nw, true); if (rc) return rc; rc = t_alloc(t, &T.v2b, (size_t)T.cout, true); if (rc) return rc; }
    adam_kernel_w<<<(unsigned)((nw + 255) / 256), 256, 0, s>>>(L.w_f32, T.vw, T.v2w, t->grad + T.off_w, T.taps, T.cin, T.cout, lr_t, beta1, beta2, eps, l2_reg, grad_scale);
    SSDK_COUNT_LAUNCH(ctx);
    adam_kernel_flat<<<(unsigned)((T.cout + 255) / 256), 256, 0, s>>>(L.bias, T.vb, T.v2b, t->grad + T.off_b, (size_t)T.cout, lr_t, beta1, beta2, eps, grad_scale);
    SSDK_COUNT_LAUNCH(ctx);
    if (T.off_bng >= 0) {
      adam_kernel_flat<<<(unsigned)((T.cout + 255) / 256), 256, 0, s>>>(L.bn_gamma, T.m_bng, T.v_bng, t->grad + T.off_bng, (size_t)T.cout, lr_t, beta1, beta2, eps, grad_scale);
      adam_kernel_flat<<<(unsigned)((T.cout + 255) / 256), 256, 0, s>>>(L.bn_beta, T.m_bnb, T.v_bnb, t->grad + T.off_bnb, (size_t)T.cout, lr_t, beta1, beta2, eps, grad_scale);
      SSDK_COUNT_LAUNCH(ctx); SSDK_COUNT_LAUNCH(ctx);
    }
    rc = repack_after_update(t, (int)i, s); if (rc) return rc;
  }
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

extern "C" int ssdk_trainer_read_bn_stats(ssdk_trainer* t, int layer, float* mean_dev, float* var_dev, void* stream_) {
  SSDK_REQUIRE(t && mean_dev && var_dev && layer >= 0 && layer < (int)t->tl.size(), "ssdk_trainer_read_bn_stats: bad argument");
  LayerPlan& L = t->m->layers[layer];
  SSDK_REQUIRE(L.bn_train, "ssdk_trainer_read_bn_stats: layer %d has no BatchNormalization", layer);
  SSDK_CHECK_CUDA(cudaMemcpyAsync(mean_dev, L.bn_mmean, (size_t)L.C * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
  SSDK_CHECK_CUDA(cudaMemcpyAsync(var_dev, L.bn_mvar, (size_t)L.C * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
  return SSDK_OK;
}

extern "C" int ssdk_trainer_read_params(ssdk_trainer* t, float* out_dev, void* stream_) {
  SSDK_REQUIRE(t && out_dev, "ssdk_trainer_read_params: NULL argument");
  ssdk_model* m = t->m;
  cudaStream_t s = (cudaStream_t)stream_;
  for (size_t i = 0; i < t->tl.size(); ++i) {
    TLayer& T = t->tl[i];
    LayerPlan& L = m->layers[i];
    if (T.off_g >= 0) SSDK_CHECK_CUDA(cudaMemcpyAsync(out_dev + T.off_g, L.gamma, (size_t)L.C * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (T.off_w < 0) continue;
    const size_t nw = (size_t)T.taps * T.cin * T.cout;
    hwio_to_ohwi_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, s>>>(L.w_f32, T.taps, T.cin, T.cout, out_dev + T.off_w);
    SSDK_COUNT_LAUNCH(m->ctx);
    SSDK_CHECK_CUDA(cudaMemcpyAsync(out_dev + T.off_b, L.bias, (size_t)T.cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (T.off_bng >= 0) {
      SSDK_CHECK_CUDA(cudaMemcpyAsync(out_dev + T.off_bng, L.bn_gamma, (size_t)T.cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
      SSDK_CHECK_CUDA(cudaMemcpyAsync(out_dev + T.off_bnb, L.bn_beta, (size_t)T.cout * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
  }
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}
