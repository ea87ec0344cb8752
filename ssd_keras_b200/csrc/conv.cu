// tcgen05 implicit-GEMM convolution for sm_100a plus the small data-movement kernels around it.
// Reference ops: Conv2D / MaxPooling2D / Lambda preprocessing / L2Normalization / Reshape+softmax+Concat in
// models/keras_ssd300.py:263-419, keras_layers/keras_layer_L2Normalization.py:61-63.
//
// conv_tcgen05_kernel -- persistent, warp-specialised (DESIGN.md section "conv"):
//   GEMM view   D[M = output pixels, N = Cout] = sum over (tap, cin-block) A_tap[M, 64] * W_tap[N, 64]^T.
//   "im2col in the TMA descriptor": the activation tensor is a zero-bordered NHWC buffer viewed as a 2-D
//   matrix [B*Hp*Wp, C]; the A tile of filter tap (kh, kw) for output rows [m0, m0+128) is simply rows
//   [m0 + shift(kh,kw), ...) of that matrix, so each tap is ONE 2-D TMA box load with a row offset, and
//   padding / dilation come for free (the border is zero, rows past the end are TMA zero-filled).
//   Outputs are computed for every position of the padded grid ("virtual rows"); the epilogue stores the
//   valid ones, and m-tiles without any valid row are not scheduled.
//   warp 0: TMA producer (one elected lane)      warp 1: tcgen05.mma issuer (one lane)
//   warp 2: TMEM allocator                        warps 4-7: epilogue (tcgen05.ld -> bias/BN/act -> store)
//   smem ring of `stages` {A_hi, A_lo, W_hi, W_lo} 128B-swizzled K-major tiles, full/empty mbarriers;
//   two TMEM accumulators (2*BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//   Precision: operands are bf16 "hi + lo" pairs; three MMAs per k-step (hi*hi, hi*lo, lo*hi) accumulate
//   in fp32 TMEM, which reproduces an fp32 convolution to ~1e-5 relative (split=0 issues hi*hi only).
#include "conv.cuh"
#include "tc.cuh"
#include <cudaTypedefs.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

namespace ssdk {

// ------------------------------------------------------------------------------------------------
// TMA descriptor creation (driver entry point resolved at run time)
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int tma_init() {
  if (g_encode) return SSDK_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SSDK_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || !fn) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return SSDK_ERR_CUDA; }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return SSDK_OK;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows) {
  int rc = tma_init();
  if (rc) return rc;
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu stride=%llu box=%ux%u base=%p", (int)r,
              (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_inner, box_rows, base);
    return SSDK_ERR_CUDA;
  }
  return SSDK_OK;
}

int make_tmap_4d(CUtensorMap* out, const void* base, const uint64_t dims_[4], const uint32_t box_[4]) {
  int rc = tma_init();
  if (rc) return rc;
  cuuint64_t dims[4] = {dims_[0], dims_[1], dims_[2], dims_[3]};
  cuuint64_t strides[3] = {dims_[0] * 2, dims_[0] * dims_[1] * 2, dims_[0] * dims_[1] * dims_[2] * 2};
  cuuint32_t box[4] = {box_[0], box_[1], box_[2], box_[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(4d) failed (%d): dims=%llu,%llu,%llu,%llu box=%u,%u,%u,%u base=%p", (int)r,
              (unsigned long long)dims_[0], (unsigned long long)dims_[1], (unsigned long long)dims_[2], (unsigned long long)dims_[3],
              box_[0], box_[1], box_[2], box_[3], base);
    return SSDK_ERR_CUDA;
  }
  return SSDK_OK;
}

static constexpr int kBM = 128;          // UMMA M (rows per tile)
static constexpr int kBK = 64;           // channels per k-block = one 128-byte swizzle atom of bf16
static constexpr int kATile = kBM * kBK * 2;   // 16 KB

static int mt_of(const ConvArgs& a) { return a.mt > 1 ? a.mt : 1; }
static size_t slot_a_bytes(const ConvArgs& a) { return (size_t)(a.slab_rows + (mt_of(a) - 1) * kBM) * kBK * 2 * (a.split ? 2 : 1); }
static size_t slot_b_bytes(const ConvArgs& a) { return (size_t)a.BN * kBK * 2 * (a.split ? 2 : 1); }
static size_t epi_param_bytes(const ConvArgs& a) { return ((size_t)a.cout * 3 * sizeof(float) + 127) / 128 * 128; }   // bias | bn scale | bn shift
size_t conv_smem_bytes(const ConvArgs& a) { return 1024 + slot_a_bytes(a) * a.stages_a + slot_b_bytes(a) * a.stages_b + 512 + epi_param_bytes(a); }
// Two rings: A slabs (one per (kh, channel block), shared by the KW taps of that row) and weight tiles (one per tap).
void conv_pick_stages(ConvArgs& a) {
  const size_t budget = 218 * 1024 - 1536 - epi_param_bytes(a);
  if (a.resident_b) {                       // weights resident: stages_b counts the tiles, the rest goes to A slabs
    a.stages_b = a.KH * a.KW * a.kblocks;
    a.stages_a = 2;
    while (a.stages_a < 4 && slot_a_bytes(a) * (a.stages_a + 1) + slot_b_bytes(a) * a.stages_b <= budget) ++a.stages_a;
    return;
  }
  a.stages_a = 2; a.stages_b = 2;
  int max_a = 4, max_b = 6;
  if (const char* e = getenv("SSDK_SA_MAX")) max_a = atoi(e);
  if (const char* e = getenv("SSDK_SB_MAX")) max_b = atoi(e);
  bool grew = true;
  while (grew) {
    grew = false;
    if (a.stages_b < max_b && slot_a_bytes(a) * a.stages_a + slot_b_bytes(a) * (a.stages_b + 1) <= budget) { ++a.stages_b; grew = true; }
    if (a.stages_a < max_a && slot_a_bytes(a) * (a.stages_a + 1) + slot_b_bytes(a) * a.stages_b <= budget) { ++a.stages_a; grew = true; }
  }
}

// ------------------------------------------------------------------------------------------------
// UMMA descriptors (PTX wrappers: tc.cuh)
// ------------------------------------------------------------------------------------------------
// K-major, 128B-swizzled operand tile: start address, SBO = 1024 B (8 rows x 128 B), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, int bo_mode = 0) {
  uint64_t d = 0;
  const uint32_t phase = (addr >> 7) & 0x7u;          // start row inside the 8-row (1024 B) swizzle atom
  if (bo_mode == 1) d |= (uint64_t)phase << 49;       // [49,52) base offset
  else if (bo_mode == 2) d |= (uint64_t)((8u - phase) & 7u) << 49;
  d |= (uint64_t)((addr >> 4) & 0x3FFFu);            // [0,14)  start address >> 4
  d |= (uint64_t)0 << 16;                            // [16,30) leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;     // [32,46) stride byte offset
  d |= (uint64_t)1 << 46;                            // [46,48) version = 1
  d |= (uint64_t)2 << 61;                            // [61,64) layout = SWIZZLE_128B
  return d;
}
// constant upper part of that descriptor: SBO = 1024 B at [32,46), version 1 at [46,48), SWIZZLE_128B at [61,64)
static constexpr uint64_t kDescHi = ((uint64_t)(1024u >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
__device__ __forceinline__ uint32_t make_idesc(int n) {
  // c_format F32 (bit 4), a/b format BF16 (bits 7, 10), K-major A and B, N>>3 at bit 17, M>>4 at bit 24
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == SSDK_ACT_RELU) return fmaxf(x, 0.f);
  if (act == SSDK_ACT_ELU) return x > 0.f ? x : expm1f(x);
  return x;
}

// Accumulator read: 32 columns of this thread's TMEM lane; with a separate cross-term accumulator (`xoff` columns further)
// the two partial sums are added here in fp32 round-to-nearest.
__device__ __forceinline__ void ld_acc32(uint32_t taddr, uint32_t xoff, uint32_t (&v)[32]) {
  if (!xoff) { tmem_ld32(taddr, v); return; }
  {
    uint32_t w[32];
    tmem_ld32_nowait(taddr, v);            // both loads in flight, one wait (SSDK: the epilogue of 256-wide tiles with a cross-term
    tmem_ld32_nowait(taddr + xoff, w);     // accumulator does not overlap the next tile's MMAs, so its latency is on the critical path)
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
  }
}

// Predictor-head epilogue for a compile-time row width CP4 = n_classes + 4 (25: Pascal VOC, the benchmark configuration): one thread =
// one pixel = n_boxes prior rows.  With CP4 known the accumulator columns of a box are compile-time register indices, so each box
// costs one sweep: <= 2 TMEM loads, C bias adds, C exponentials (kept in registers), one reciprocal, C+12 stores.
template <int CP4>
__device__ __forceinline__ void epi_head_fixed(const ConvArgs& args, uint32_t t_row, uint32_t xoff, bool valid, int n, int pix,
                                               const float* s_bias) {
  constexpr int C = CP4 - 4, RW = C + 12;
#pragma unroll
  for (int bx = 0; bx < 8; ++bx) {
    if (bx >= args.head_nb) break;                               // uniform
    constexpr int kMaxCol = 256;
    const int c_lo = bx * CP4;                                   // compile-time after unrolling
    const int k_lo = c_lo >> 5, k_hi = (c_lo + CP4 - 1) >> 5;
    if (c_lo + CP4 > kMaxCol) break;
    uint32_t v0[32], v1[32];
    ld_acc32(t_row + (uint32_t)(k_lo * 32), xoff, v0);
    if (k_hi != k_lo) ld_acc32(t_row + (uint32_t)(k_hi * 32), xoff, v1);
    if (!valid) continue;
    float e[CP4];
#pragma unroll
    for (int r = 0; r < CP4; ++r) {
      const int col = c_lo + r;
      e[r] = __uint_as_float(((col >> 5) == k_lo) ? v0[col & 31] : v1[col & 31]) + s_bias[col];
    }
    float mx = e[0];
#pragma unroll
    for (int r = 1; r < C; ++r) mx = fmaxf(mx, e[r]);
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < C; ++r) { e[r] = expf(e[r] - mx); sum += e[r]; }
    const float inv = 1.0f / sum;
    const int prior = args.head_prior_off + pix * args.head_nb + bx;
    float* dst = args.out_f32 + ((size_t)n * args.head_P + prior) * RW;
    const float4 an = __ldg(reinterpret_cast<const float4*>(args.head_anchors) + prior);
#pragma unroll
    for (int r = 0; r < C; ++r) dst[r] = e[r] * inv;
#pragma unroll
    for (int r = C; r < CP4; ++r) dst[r] = e[r];
    dst[C + 4] = an.x; dst[C + 5] = an.y; dst[C + 6] = an.z; dst[C + 7] = an.w;
    dst[C + 8] = args.head_var[0]; dst[C + 9] = args.head_var[1]; dst[C + 10] = args.head_var[2]; dst[C + 11] = args.head_var[3];
  }
}

// Epilogue of the activation-producing launches: bias / folded BatchNorm / activation -> bf16 hi+lo planes, 8 channels per
// 16-byte store.  BWD adds what the data-gradient launches need: ReLU'(forward value) mask and accumulation into the output.
template <bool BWD, bool PIPE = false>
__device__ __forceinline__ void epi_split(const ConvArgs& args, uint32_t t_row, uint32_t xoff, int ncols, int n0, size_t o, bool valid,
                                          const float* s_bias, const float* s_scale, const float* s_shift) {
  if (!BWD && !args.bn_scale && args.act == SSDK_ACT_RELU && args.out_lo) {
    // the common forward case (bias + ReLU, hi/lo planes) without per-element branches: packed conversions (two values per
    // cvt.rn.bf16x2.f32), biases fetched four at a time, and 32-byte stores (one full sector per lane and instruction: a thread's
    // row is 2*Cout bytes of its own, so 16-byte stores leave every sector half written per instruction).  Bit-identical to the
    // generic path below.
    const bool wide = (args.out_Cs % 16 == 0) && ((n0 & 15) == 0);
    // one chunk of 32 accumulator columns: bias + ReLU, hi/lo split, stores
    auto process = [&](const uint32_t (&vr)[32], int c0) {
#pragma unroll
      for (int g2 = 0; g2 < 2; ++g2) {
        uint32_t ph[8], pl[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int g = g2 * 2 + h;
          if (c0 + g * 8 < ncols) {
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + n0 + c0 + g * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + n0 + c0 + g * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float f0 = fmaxf(__uint_as_float(vr[g * 8 + j * 2]) + bb[j * 2], 0.f);
              const float f1 = fmaxf(__uint_as_float(vr[g * 8 + j * 2 + 1]) + bb[j * 2 + 1], 0.f);
              const __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
              const uint32_t hp = *reinterpret_cast<const uint32_t*>(&hh);
              const __nv_bfloat162 ll = __floats2bfloat162_rn(f0 - __uint_as_float(hp << 16), f1 - __uint_as_float(hp & 0xffff0000u));
              ph[h * 4 + j] = hp; pl[h * 4 + j] = *reinterpret_cast<const uint32_t*>(&ll);
            }
          }
        }
        const int cA = c0 + g2 * 16;
        if (wide && cA + 16 <= ncols) {
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(args.out_hi + o + cA), "r"(ph[0]), "r"(ph[1]), "r"(ph[2]),
                       "r"(ph[3]), "r"(ph[4]), "r"(ph[5]), "r"(ph[6]), "r"(ph[7]) : "memory");
          asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(args.out_lo + o + cA), "r"(pl[0]), "r"(pl[1]), "r"(pl[2]),
                       "r"(pl[3]), "r"(pl[4]), "r"(pl[5]), "r"(pl[6]), "r"(pl[7]) : "memory");
        } else {
          if (cA < ncols) {
            *reinterpret_cast<uint4*>(args.out_hi + o + cA) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            *reinterpret_cast<uint4*>(args.out_lo + o + cA) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
          }
          if (cA + 8 < ncols) {
            *reinterpret_cast<uint4*>(args.out_hi + o + cA + 8) = make_uint4(ph[4], ph[5], ph[6], ph[7]);
            *reinterpret_cast<uint4*>(args.out_lo + o + cA + 8) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
          }
        }
      }
    
    };
    if constexpr (PIPE) {
      // software-pipelined accumulator reads (conv_tcgen05_kernel): the TMEM loads of chunk c+1 are issued before chunk c is converted
      // and stored, so their latency hides behind that work (tcgen05.wait::ld waits for every load this thread has issued).  It
      // matters for the 256-wide tiles with a cross-term accumulator: they have ONE accumulator set, so their epilogue does not run
      // under the next tile's MMAs.  args.epi_pipe (SSDK_EPI_PIPE): 2 = as described (default), 1 = both loads of a chunk in flight
      // but no prefetch, 0 = one load at a time.
      const int pipe = args.epi_pipe;
      uint32_t va[32], vb[32];
      if (pipe == 2) {
        tmem_ld32_nowait(t_row, va);
        if (xoff) tmem_ld32_nowait(t_row + xoff, vb);
      }
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t vr[32];
        if (pipe != 2) {
          tmem_ld32_nowait(t_row + (uint32_t)c0, va);
          if (pipe == 0) tmem_ld_wait();
          if (xoff) tmem_ld32_nowait(t_row + (uint32_t)c0 + xoff, vb);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) vr[j] = xoff ? __float_as_uint(__uint_as_float(va[j]) + __uint_as_float(vb[j])) : va[j];
        if (pipe == 2 && c0 + 32 < ncols) {
          tmem_ld32_nowait(t_row + (uint32_t)(c0 + 32), va);
          if (xoff) tmem_ld32_nowait(t_row + (uint32_t)(c0 + 32) + xoff, vb);
        }
        if (valid) process(vr, c0);
      }
    } else {
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t vr[32];
        ld_acc32(t_row + (uint32_t)c0, xoff, vr);
        if (valid) process(vr, c0);
      }
    }
    return;
  }
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t vr[32];
    ld_acc32(t_row + (uint32_t)c0, xoff, vr);
    if (!valid) continue;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (c0 + g * 8 < ncols) {
        uint32_t ph[4], pl[4];
        uint4 mk = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u), oh = make_uint4(0, 0, 0, 0), ol = oh;
        if (BWD) {
          if (args.mask_hi) mk = *reinterpret_cast<const uint4*>(args.mask_hi + o + c0 + g * 8);
          if (args.accumulate) {
            oh = *reinterpret_cast<const uint4*>(args.out_hi + o + c0 + g * 8);
            if (args.out_lo) ol = *reinterpret_cast<const uint4*>(args.out_lo + o + c0 + g * 8);
          }
        }
        const uint32_t mkw[4] = {mk.x, mk.y, mk.z, mk.w}, ohw[4] = {oh.x, oh.y, oh.z, oh.w}, olw[4] = {ol.x, ol.y, ol.z, ol.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float f[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = n0 + c0 + g * 8 + j * 2 + e;
            float xv = __uint_as_float(vr[g * 8 + j * 2 + e]) + s_bias[col];
            if (args.bn_scale) xv = xv * s_scale[col] + s_shift[col];
            xv = apply_act(xv, args.act);
            if (BWD) {
              if (!(__uint_as_float(((mkw[j] >> (e * 16)) & 0xffffu) << 16) > 0.f)) xv = 0.f;       // ReLU'(forward value)
              xv += __uint_as_float(((ohw[j] >> (e * 16)) & 0xffffu) << 16) + __uint_as_float(((olw[j] >> (e * 16)) & 0xffffu) << 16);
            }
            f[e] = xv;
          }
          __nv_bfloat16 h0 = __float2bfloat16_rn(f[0]), h1 = __float2bfloat16_rn(f[1]);
          __nv_bfloat16 l0 = __float2bfloat16_rn(f[0] - __bfloat162float(h0));
          __nv_bfloat16 l1 = __float2bfloat16_rn(f[1] - __bfloat162float(h1));
          ph[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          pl[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
        *reinterpret_cast<uint4*>(args.out_hi + o + c0 + g * 8) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        if (args.out_lo) *reinterpret_cast<uint4*>(args.out_lo + o + c0 + g * 8) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The convolution kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
conv_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                    const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                    const __grid_constant__ ConvArgs args) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SA = args.stages_a, SB = args.stages_b, BN = args.BN, split = args.split;
  const int MT = args.mt > 1 ? args.mt : 1;                            // m-tiles per work unit
  const uint32_t a_plane = (uint32_t)(args.slab_rows + (MT - 1) * kBM) * kBK * 2;   // one A slab plane (hi or lo)
  const uint32_t a_box = (uint32_t)args.slab_rows * kBK * 2;           // bytes one TMA box delivers
  const uint32_t b_tile = (uint32_t)BN * kBK * 2;
  const uint32_t slot_a = a_plane * (split ? 2 : 1), slot_b = b_tile * (split ? 2 : 1);
  const uint32_t ring_b = smem_base + slot_a * SA;
  const uint32_t bar_base = ring_b + slot_b * SB;
  // barrier slots (8 B each): fullA[SA] | emptyA[SA] | fullB[SB] | emptyB[SB] | tmem_full[2] | tmem_empty[2] | tmem ptr
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (SA + s); };
  auto fullB = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto emptyB = [&](int s) { return bar_base + 8u * (2 * SA + SB + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SA + 2 * SB + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));
  // epilogue parameters live in shared memory (L1 is almost entirely carved out for the rings, global loads would miss)
  float* s_bias = reinterpret_cast<float*>(smem_dyn + (bar_base + 512u - smem_u32(smem_dyn)));
  float* s_scale = s_bias + args.cout;
  float* s_shift = s_scale + args.cout;
  for (int i = threadIdx.x; i < args.cout; i += blockDim.x) {
    s_bias[i] = args.bias ? args.bias[i] : 0.f;
    if (args.bn_scale) { s_scale[i] = args.bn_scale[i]; s_shift[i] = args.bn_shift[i]; }
  }

  int tmem_cols = 32;
  const int XS = args.acc_split ? 2 : 1;                               // accumulators per output tile (main | cross terms)
  const int NB = args.acc_bufs == 1 ? 1 : 2;                           // accumulator sets
  while (tmem_cols < NB * XS * MT * BN) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b_hi)) : "memory");
    if (split) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a_lo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b_lo)) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int KS = args.k_split > 1 ? args.k_split : 1;
  const int total_tiles = args.n_tiles_m * args.n_tiles_n * KS;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      if (args.resident_b && (int)blockIdx.x < total_tiles) {      // every weight tile once, each on its own barrier
        for (int i = 0; i < SB; ++i) {
          const uint32_t db = ring_b + slot_b * i;
          mbar_expect_tx(fullB(i), slot_b);
          tma_load_2d(db, &tm_b_hi, i * kBK + args.b_k_offset, 0, fullB(i));
          if (split) tma_load_2d(db + b_tile, &tm_b_lo, i * kBK + args.b_k_offset, 0, fullB(i));
        }
      }
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int tt = t / KS, ks = t - tt * KS;
        const int m0 = args.tile_list[tt / args.n_tiles_n] * kBM;
        const int n0 = (tt % args.n_tiles_n) * BN;
        const int kb0 = KS > 1 ? ks * args.kb_per : 0;
        const int kb1 = KS > 1 ? min(args.kblocks, kb0 + args.kb_per) : args.kblocks;
        for (int kh = 0; kh < args.KH; ++kh) {
          const int row0 = m0 + args.row_shift[kh];
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(emptyA(sa), pa ^ 1u);
            const uint32_t da = smem_base + slot_a * sa;
            mbar_expect_tx(fullA(sa), a_box * (uint32_t)MT * (split ? 2u : 1u));
            for (int mt = 0; mt < MT; ++mt) {                // the second box overlaps the first by slab_rows-128 identical rows
              tma_load_2d(da + mt * kATile, &tm_a_hi, kb * kBK, row0 + mt * kBM, fullA(sa));
              if (split) tma_load_2d(da + a_plane + mt * kATile, &tm_a_lo, kb * kBK, row0 + mt * kBM, fullA(sa));
            }
            if (++sa == SA) { sa = 0; pa ^= 1u; }
            for (int kw = 0; kw < args.KW && !args.resident_b; ++kw) {
              mbar_wait(emptyB(sb), pb ^ 1u);
              const uint32_t db = ring_b + slot_b * sb;
              mbar_expect_tx(fullB(sb), slot_b);
              const int kcol = ((kh * args.KW + kw) * args.kblocks + kb) * kBK + args.b_k_offset;
              tma_load_2d(db, &tm_b_hi, kcol, n0, fullB(sb));
              if (split) tma_load_2d(db + b_tile, &tm_b_lo, kcol, n0, fullB(sb));
              if (++sb == SB) { sb = 0; pb ^= 1u; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int tt = t / KS, ks = t - tt * KS;
      const int n0 = (tt % args.n_tiles_n) * BN;
      const int kb0 = KS > 1 ? ks * args.kb_per : 0;
      const int kb1 = KS > 1 ? min(args.kblocks, kb0 + args.kb_per) : args.kblocks;
      const int acc = NB == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = (uint32_t)(NB == 2 ? (it >> 1) : it) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      int n_eff = args.cout - n0;
      n_eff = n_eff > BN ? BN : ((n_eff + 15) & ~15);
      const uint32_t idesc = make_idesc(n_eff);
      const uint32_t idesc2 = make_idesc(BN + n_eff);       // fuse_b: [B_hi (BN rows, the tail zero-filled by TMA) ; B_lo (n_eff rows)]
      const bool fuse_b = args.fuse_b != 0;
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * XS * MT * BN);
      const uint32_t x_cols = (uint32_t)((XS - 1) * BN);     // column offset of the cross-term accumulator
      uint32_t accumulate = 0;
      for (int kh = 0; kh < args.KH; ++kh) {
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(fullA(sa), pa);
          const uint32_t a_hi = smem_base + slot_a * sa, a_lo = a_hi + a_plane;
          const int ksteps = (kb == args.kblocks - 1) ? args.last_ksteps : 4;
          for (int kw = 0; kw < args.KW; ++kw) {
            int sbi = sb;
            if (args.resident_b) { sbi = (kh * args.KW + kw) * args.kblocks + kb; if (it == 0) mbar_wait(fullB(sbi), 0); }
            else mbar_wait(fullB(sb), pb);
            tc_fence_after();
            if (elect_one()) {
              // The issuing thread is a single in-order instruction stream: for N <= 128 an MMA retires in 32-64 clocks, so the
              // descriptor arithmetic between two issues must be a couple of integer adds.  Only the 14-bit start-address field
              // (address >> 4) changes: +2 per k-step (32 B), +8*rows for a row-shifted tap, +1024 for the second m-tile.
              const uint32_t bh = (ring_b + slot_b * sbi) >> 4, bl = bh + (b_tile >> 4);
              const uint32_t ah = (a_hi + (uint32_t)(kw * args.kw_rows) * 128u) >> 4, al = ah + (a_plane >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < ksteps) {
                  const uint64_t db = kDescHi | (uint64_t)(bh + 2 * k), dbl = kDescHi | (uint64_t)(bl + 2 * k);
#pragma unroll
                  for (int mt = 0; mt < 2; ++mt) {
                    if (mt < MT) {
                      const uint64_t da = kDescHi | (uint64_t)(ah + mt * (kATile >> 4) + 2 * k);
                      const uint32_t d_main = d_tmem + (uint32_t)(mt * XS * BN);
                      if (fuse_b) {
                        tc_mma(d_main, da, db, idesc2, accumulate);          // [main | cross] (+)= A_hi * [B_hi ; B_lo]
                        tc_mma(d_main + x_cols, kDescHi | (uint64_t)(al + mt * (kATile >> 4) + 2 * k), db, idesc, 1);
                      } else {
                        tc_mma(d_main, da, db, idesc, accumulate);
                        if (split) {
                          tc_mma(d_main + x_cols, da, dbl, idesc, XS == 2 ? accumulate : 1u);
                          tc_mma(d_main + x_cols, kDescHi | (uint64_t)(al + mt * (kATile >> 4) + 2 * k), db, idesc, 1);
                        }
                      }
                    }
                  }
                  accumulate = 1;
                }
              }
              if (!args.resident_b) tc_commit(emptyB(sb));   // weight slot is free once these MMAs retire
              if (kw == args.KW - 1) tc_commit(emptyA(sa));  // ... and the slab after its last tap
            }
            __syncwarp();
            if (!args.resident_b && ++sb == SB) { sb = 0; pb ^= 1u; }
          }
          if (++sa == SA) { sa = 0; pa ^= 1u; }
        }
      }
      if (elect_one()) tc_commit(tfull_bar(acc));
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp - 4;                                // TMEM lane quarter of this warp (== warp % 4)
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int tt = t / KS;
      const int m0 = args.tile_list[tt / args.n_tiles_n] * kBM;
      const int n0 = (tt % args.n_tiles_n) * BN;
      const int acc = NB == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = (uint32_t)(NB == 2 ? (it >> 1) : it) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      for (int mt = 0; mt < MT; ++mt) {
      const int v = m0 + mt * kBM + q * 32 + lane;         // virtual row of this thread
      bool valid = v < args.M_total;
      int n = 0, y = 0, x = 0;
      if (valid) {
        n = v / args.rows_per_img;
        const int r = v - n * args.rows_per_img;
        y = r / args.in_Wp;
        x = r - y * args.in_Wp;
        valid = (y < args.Ho) && (x < args.Wo);
      }
      int ncols = args.cout - n0;
      ncols = ncols > BN ? BN : ncols;
      const uint32_t t_row = tmem_base + (uint32_t)((acc * MT + mt) * XS * BN) + ((uint32_t)(q * 32) << 16);
      const uint32_t xoff = (uint32_t)((XS - 1) * BN);
      if (args.epi == EPI_SPLIT) {
        const size_t o = (((size_t)n * args.out_Hp + (y + args.out_pad)) * args.out_Wp + (x + args.out_pad)) * args.out_Cs + n0;
        if (args.mask_hi || args.accumulate) epi_split<true>(args, t_row, xoff, ncols, n0, o, valid, s_bias, s_scale, s_shift);
        else epi_split<false, true>(args, t_row, xoff, ncols, n0, o, valid, s_bias, s_scale, s_shift);
      } else if (args.epi == EPI_ATOMIC) {
        float* dstp = args.out_f32 + (size_t)v * args.out_ld + args.out_col_off + n0;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t vr[32];
          ld_acc32(t_row + (uint32_t)c0, xoff, vr);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < ncols) atomicAdd(dstp + c0 + j, __uint_as_float(vr[j]));
          }
        }
      } else if (args.epi == EPI_HEAD) {
        // one thread = one pixel = n_boxes prior rows.  Per box three sweeps over its C+4 accumulator columns (TMEM reads are cheap
        // and this warp group runs under the MMAs of the next tile): maximum, sum of exponentials, normalised store.
        const int C = args.head_C, CP4 = C + 4, RW = C + 12;
        const int pix = y * args.Wo + x;
        if (CP4 == 25 && args.head_nb <= 8) { epi_head_fixed<25>(args, t_row, xoff, valid, n, pix, s_bias); continue; }
        for (int bx = 0; bx < args.head_nb; ++bx) {
          const int c_lo = bx * CP4, c_hi = c_lo + CP4;
          const int k_lo = c_lo >> 5, k_hi = (c_hi - 1) >> 5;
          float mx = -INFINITY;
          for (int k = k_lo; k <= k_hi; ++k) {
            uint32_t vr[32];
            ld_acc32(t_row + (uint32_t)(k * 32), xoff, vr);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = k * 32 + j;
              if (col >= c_lo && col < c_lo + C) mx = fmaxf(mx, __uint_as_float(vr[j]) + s_bias[col]);
            }
          }
          float sum = 0.f;
          for (int k = k_lo; k <= k_hi; ++k) {
            uint32_t vr[32];
            ld_acc32(t_row + (uint32_t)(k * 32), xoff, vr);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = k * 32 + j;
              if (col >= c_lo && col < c_lo + C) sum += expf(__uint_as_float(vr[j]) + s_bias[col] - mx);
            }
          }
          const int prior = args.head_prior_off + pix * args.head_nb + bx;
          float* dst = args.out_f32 + ((size_t)n * args.head_P + prior) * RW;
          for (int k = k_lo; k <= k_hi; ++k) {
            uint32_t vr[32];
            ld_acc32(t_row + (uint32_t)(k * 32), xoff, vr);
            if (valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = k * 32 + j;
                if (col >= c_lo && col < c_hi) {
                  const float v = __uint_as_float(vr[j]) + s_bias[col];
                  const int r = col - c_lo;
                  dst[r] = r < C ? expf(v - mx) / sum : v;
                }
              }
            }
          }
          if (valid) {
            const float4 an = __ldg(reinterpret_cast<const float4*>(args.head_anchors) + prior);
            dst[C + 4] = an.x; dst[C + 5] = an.y; dst[C + 6] = an.z; dst[C + 7] = an.w;
            dst[C + 8] = args.head_var[0]; dst[C + 9] = args.head_var[1]; dst[C + 10] = args.head_var[2]; dst[C + 11] = args.head_var[3];
          }
        }
      } else {
        const size_t o = (((size_t)n * args.Ho + y) * args.Wo + x) * (size_t)args.cout + n0;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t vr[32];
          ld_acc32(t_row + (uint32_t)c0, xoff, vr);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (c0 + j < ncols) {
                const int col = n0 + c0 + j;
                float xv = __uint_as_float(vr[j]) + s_bias[col];
                args.out_f32[o + c0 + j] = apply_act(xv, args.act);
              }
            }
          }
        }
      }
      }   // mt
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

int launch_conv(ssdk_ctx* ctx, const ConvLaunch& L, cudaStream_t stream, int grid_cap) {
  static bool attr_set = false;
  if (!attr_set) {
    SSDK_CHECK_CUDA(cudaFuncSetAttribute(conv_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  // the kernel is persistent with a static stride over its work units: any grid size computes the same result
  const int grid = grid_cap > 0 ? std::max(1, std::min(L.grid, grid_cap)) : L.grid;
  conv_tcgen05_kernel<<<grid, 256, L.smem, stream>>>(L.a_hi, L.a_lo, L.b_hi, L.b_lo, L.args);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

// ------------------------------------------------------------------------------------------------
// helpers: split store
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_store(const ActBuf& o, size_t idx, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  o.hi[idx] = h;
  if (o.lo) o.lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ float split_load(const ActBuf& a, size_t idx) {
  float v = __bfloat162float(a.hi[idx]);
  if (a.lo) v += __bfloat162float(a.lo[idx]);
  return v;
}
__device__ __forceinline__ size_t act_index(const ActBuf& a, int n, int y, int x) {
  return (((size_t)n * a.Hp() + (y + a.pad)) * a.Wp() + (x + a.pad)) * a.Cs;
}

// (x - mean) / std, channel swap (models/keras_ssd300.py:247-272) -> split bf16 planes
__global__ void preprocess_kernel(const float* __restrict__ img, int B, int H, int W, int Cimg, float3 mean, float3 inv_std,
                                  int has_std, int3 swap, ActBuf out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * H * W;
  if (i >= total) return;
  const int x = (int)(i % W); const int y = (int)((i / W) % H); const int n = (int)(i / ((size_t)W * H));
  const float* p = img + i * Cimg;
  float v[3];
  const float m[3] = {mean.x, mean.y, mean.z};
  const float s[3] = {inv_std.x, inv_std.y, inv_std.z};
  for (int c = 0; c < Cimg && c < 3; ++c) {
    float t = p[c] - m[c];
    if (has_std) t = t / s[c];          // inv_std holds the divisor itself when has_std
    v[c] = t;
  }
  const int sw[3] = {swap.x, swap.y, swap.z};
  const size_t o = act_index(out, n, y, x);
  if (out.Cs == 8 && Cimg == 3) {               // one 16-byte word per plane and pixel (channels 3..7 are zero)
    const float w0 = v[sw[0]], w1 = v[sw[1]], w2 = v[sw[2]];
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(w0, w1);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(w2);
    const uint32_t hp = *reinterpret_cast<const uint32_t*>(&h01);
    *reinterpret_cast<uint4*>(out.hi + o) = make_uint4(hp, (uint32_t)__bfloat16_as_ushort(h2), 0u, 0u);
    if (out.lo) {
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(w0 - __uint_as_float(hp << 16), w1 - __uint_as_float(hp & 0xffff0000u));
      const __nv_bfloat16 l2 = __float2bfloat16_rn(w2 - __bfloat162float(h2));
      *reinterpret_cast<uint4*>(out.lo + o) = make_uint4(*reinterpret_cast<const uint32_t*>(&l01), (uint32_t)__bfloat16_as_ushort(l2), 0u, 0u);
    }
    return;
  }
  for (int c = 0; c < Cimg && c < 3; ++c) split_store(out, o + c, v[sw[c]]);
}

int launch_preprocess(ssdk_ctx* ctx, const float* images, int B, int H, int W, int Cimg, const float* mean, const float* stddev,
                      const int* swap, const ActBuf& out, cudaStream_t stream) {
  SSDK_REQUIRE(Cimg == 3, "only 3-channel images are supported (got %d)", Cimg);
  float3 m = mean ? make_float3(mean[0], mean[1], mean[2]) : make_float3(0, 0, 0);
  float3 s = stddev ? make_float3(stddev[0], stddev[1], stddev[2]) : make_float3(1, 1, 1);
  int3 sw = swap ? make_int3(swap[0], swap[1], swap[2]) : make_int3(0, 1, 2);
  const size_t total = (size_t)B * H * W;
  preprocess_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(images, B, H, W, Cimg, m, s, stddev ? 1 : 0, sw, out);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

// Explicit im2col for the few layers the descriptor trick does not cover (3-channel input, strided convs):
// out[row = (n, yo, xo)][k = (kh*KW + kw)*Cin + c], zero padded to Kpad columns.
__global__ void im2col_kernel(ActBuf in, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int Ho, int Wo,
                              int KH, int KW, int stride, int dil, int pad_t, int pad_l, int Kpad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)in.B * Ho * Wo * Kpad;
  if (i >= total) return;
  const int k = (int)(i % Kpad);
  const size_t row = i / Kpad;
  const int xo = (int)(row % Wo); const int yo = (int)((row / Wo) % Ho); const int n = (int)(row / ((size_t)Wo * Ho));
  __nv_bfloat16 h = __float2bfloat16_rn(0.f), l = h;
  if (k < KH * KW * in.C) {
    const int c = k % in.C; const int tap = k / in.C; const int kw = tap % KW; const int kh = tap / KW;
    const int y = yo * stride + kh * dil - pad_t, x = xo * stride + kw * dil - pad_l;
    if (y >= 0 && y < in.H && x >= 0 && x < in.W) {
      const size_t s = act_index(in, n, y, x) + c;
      h = in.hi[s];
      if (in.lo) l = in.lo[s];
    }
  }
  out_hi[i] = h;
  if (out_lo) out_lo[i] = l;
}

// The same with eight channels (one 16-byte word of a plane) per thread: whenever the input has a multiple of 8 channels, eight
// consecutive K indices are eight consecutive channels of one tap.
__global__ void im2col8_kernel(ActBuf in, __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, int Ho, int Wo,
                               int KH, int KW, int stride, int dil, int pad_t, int pad_l, int Kpad) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int K8 = Kpad / 8;
  const size_t total = (size_t)in.B * Ho * Wo * K8;
  if (i >= total) return;
  const int k = (int)(i % K8) * 8;
  const size_t row = i / K8;
  const int xo = (int)(row % Wo); const int yo = (int)((row / Wo) % Ho); const int n = (int)(row / ((size_t)Wo * Ho));
  uint4 h = make_uint4(0, 0, 0, 0), l = h;
  if (k < KH * KW * in.C) {
    const int c = k % in.C; const int tap = k / in.C; const int kw = tap % KW; const int kh = tap / KW;
    const int y = yo * stride + kh * dil - pad_t, x = xo * stride + kw * dil - pad_l;
    if (y >= 0 && y < in.H && x >= 0 && x < in.W) {
      const size_t s = act_index(in, n, y, x) + c;
      h = *reinterpret_cast<const uint4*>(in.hi + s);
      if (in.lo) l = *reinterpret_cast<const uint4*>(in.lo + s);
    }
  }
  *reinterpret_cast<uint4*>(out_hi + row * Kpad + k) = h;
  if (out_lo) *reinterpret_cast<uint4*>(out_lo + row * Kpad + k) = l;
}

int launch_im2col(ssdk_ctx* ctx, const ActBuf& in, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int Ho, int Wo, int kh, int kw,
                  int stride, int dil, int pad_t, int pad_l, int Kpad, cudaStream_t stream) {
  if (in.C % 8 == 0 && in.Cs == in.C && Kpad % 8 == 0 && (reinterpret_cast<uintptr_t>(out_hi) & 15) == 0 &&
      (!out_lo || (reinterpret_cast<uintptr_t>(out_lo) & 15) == 0)) {
    const size_t total8 = (size_t)in.B * Ho * Wo * (Kpad / 8);
    im2col8_kernel<<<(unsigned)((total8 + 255) / 256), 256, 0, stream>>>(in, out_hi, out_lo, Ho, Wo, kh, kw, stride, dil, pad_t, pad_l, Kpad);
    SSDK_COUNT_LAUNCH(ctx);
    SSDK_CHECK_CUDA(cudaGetLastError());
    return SSDK_OK;
  }
  const size_t total = (size_t)in.B * Ho * Wo * Kpad;
  im2col_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out_hi, out_lo, Ho, Wo, kh, kw, stride, dil, pad_t, pad_l, Kpad);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

// Direct fp32 convolution for the image-facing layer (Cin < 8, e.g. conv1_1 3x3x3 or SSD7's conv1 5x5x3): K = kh*kw*cin is
// far too thin for a tensor-core tile, so each thread computes FOUR horizontally adjacent output pixels x 16 output channels
// with plain FMAs (one shared-memory weight fetch feeds 4 pixels); bias / BN / activation / hi-lo split are fused.
// HWIO weights are used as they are ([K][Cout]).
constexpr int kDirectPx = 4;
template <int CIN, int KHW>     // CIN > 0 / KHW > 0: compile-time input channels / square kernel size (full unrolling); 0: run-time
__global__ void __launch_bounds__(256) conv_direct_kernel(ActBuf in, ActBuf out, const float* __restrict__ w, const float* __restrict__ bias,
                                                          const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, int act,
                                                          int KH_, int KW_, int dil, int pad_t, int pad_l) {
  extern __shared__ float s_w[];                 // [K][Cout]
  const int KH = KHW > 0 ? KHW : KH_, KW = KHW > 0 ? KHW : KW_, Cin = CIN > 0 ? CIN : in.C;
  const int K = KH * KW * Cin, Cout = out.C;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int groups = Cout / 16;
  const int wq = (out.W + kDirectPx - 1) / kDirectPx;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)out.B * out.H * wq * groups;
  if (gid >= total) return;
  const int g = (int)(gid % groups);
  const size_t q = gid / groups;
  const int xo0 = (int)(q % wq) * kDirectPx; const int yo = (int)((q / wq) % out.H); const int n = (int)(q / ((size_t)wq * out.H));
  float acc[kDirectPx][16];
#pragma unroll
  for (int p = 0; p < kDirectPx; ++p)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[p][e] = 0.f;
#pragma unroll
  for (int kh = 0; kh < KH; ++kh) {
    const int y = yo + kh * dil - pad_t;
    if (y < 0 || y >= in.H) continue;
#pragma unroll
    for (int kw = 0; kw < KW; ++kw) {
      float v[kDirectPx][4];
#pragma unroll
      for (int p = 0; p < kDirectPx; ++p) {
        const int x = xo0 + p + kw * dil - pad_l;
        uint2 h2 = make_uint2(0, 0), l2 = make_uint2(0, 0);
        if (x >= 0 && x < in.W) {
          const size_t s = act_index(in, n, y, x);
          h2 = *reinterpret_cast<const uint2*>(in.hi + s);                    // channels 0..3 (Cin <= 4 used)
          if (in.lo) l2 = *reinterpret_cast<const uint2*>(in.lo + s);
        }
        v[p][0] = __uint_as_float(h2.x << 16) + __uint_as_float(l2.x << 16);
        v[p][1] = __uint_as_float(h2.x & 0xffff0000u) + __uint_as_float(l2.x & 0xffff0000u);
        v[p][2] = __uint_as_float(h2.y << 16) + __uint_as_float(l2.y << 16);
        v[p][3] = __uint_as_float(h2.y & 0xffff0000u) + __uint_as_float(l2.y & 0xffff0000u);
      }
#pragma unroll
      for (int c = 0; c < Cin; ++c) {
        const float4* wr = reinterpret_cast<const float4*>(s_w + (size_t)((kh * KW + kw) * Cin + c) * Cout + g * 16);
        const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
        const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
        for (int p = 0; p < kDirectPx; ++p) {
          const float vv = c == 0 ? v[p][0] : (c == 1 ? v[p][1] : (c == 2 ? v[p][2] : v[p][3]));
#pragma unroll
          for (int e = 0; e < 16; ++e) acc[p][e] = fmaf(vv, wv[e], acc[p][e]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < kDirectPx; ++p) {
    const int xo = xo0 + p;
    if (xo >= out.W) break;
    uint32_t ph[8], pl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float f[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = g * 16 + j * 2 + e;
        float xv = acc[p][j * 2 + e] + __ldg(bias + col);
        if (bn_scale) xv = xv * __ldg(bn_scale + col) + __ldg(bn_shift + col);
        f[e] = apply_act(xv, act);
      }
      __nv_bfloat16 h0 = __float2bfloat16_rn(f[0]), h1 = __float2bfloat16_rn(f[1]);
      __nv_bfloat16 l0 = __float2bfloat16_rn(f[0] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(f[1] - __bfloat162float(h1));
      ph[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      pl[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t o = act_index(out, n, yo, xo) + (size_t)g * 16;
    *reinterpret_cast<uint4*>(out.hi + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(out.hi + o + 8) = make_uint4(ph[4], ph[5], ph[6], ph[7]);
    if (out.lo) {
      *reinterpret_cast<uint4*>(out.lo + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
      *reinterpret_cast<uint4*>(out.lo + o + 8) = make_uint4(pl[4], pl[5], pl[6], pl[7]);
    }
  }
}

int launch_conv_direct(ssdk_ctx* ctx, const ActBuf& in, const ActBuf& out, const float* w, const float* bias, const float* bn_scale,
                       const float* bn_shift, int act, int kh, int kw, int dil, int pad_t, int pad_l, cudaStream_t stream) {
  const size_t smem = (size_t)kh * kw * in.C * out.C * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    SSDK_CHECK_CUDA(cudaFuncSetAttribute(conv_direct_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  const size_t total = (size_t)out.B * out.H * ((out.W + kDirectPx - 1) / kDirectPx) * (out.C / 16);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  // (a fully unrolled <3, 3> instantiation was measured slower on B200: 771 vs 683 us for conv1_1 at batch 32)
  conv_direct_kernel<0, 0><<<blocks, 256, smem, stream>>>(in, out, w, bias, bn_scale, bn_shift, act, kh, kw, dil, pad_t, pad_l);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

// ------------------------------------------------------------------------------------------------
// conv_first_kernel: the image-facing layer (Cin <= 4: conv1_1 3x3x3 of models/keras_ssd300.py:263, conv1 5x5x3 of
// models/keras_ssd7.py:277) on the tensor cores.  K = taps * 4 (channels padded to 4) is far too thin for the TMA ring of
// conv_tcgen05_kernel -- an explicit im2col would write and re-read ten times the layer's input -- so the A tile is GATHERED:
//   warps  0- 7  epilogue, two groups of four (TMEM lane quarter = warp % 4), group g drains accumulator g
//   warps  8-15  gather, two groups of 128 threads; thread = one output pixel: reads the taps' 8-byte (4-channel) hi / lo words
//                from the zero-bordered input planes (L1/L2 hits: every word is used by taps of neighbouring pixels) and writes
//                them K-major into the 128B-swizzled A stage; group g fills the stages of tiles g, g+2, ...
//   warp   16    TMEM allocation, tcgen05.mma issue (<= 8 k-steps x 3 products per tile), commits
// The weights (K-major, swizzled image prepared on the host) stay in shared memory for the whole launch.  The layer is bound by
// writing its output (B*H*W*Cout*4 bytes of hi+lo planes); the epilogue is the one of the other activation-producing launches.
// ------------------------------------------------------------------------------------------------
struct FirstArgs {
  ActBuf in;
  const __nv_bfloat16* w_hi; const __nv_bfloat16* w_lo;   // [kblocks][BN][64] swizzled images
  int KH, KW, dil, pad_t, pad_l;
  int tap_off[32];            // element offset of tap t's 4-channel word relative to the output pixel's own word in the input planes
                              // (the planes' zero border covers every tap: checked on the host)
  int kblocks, ksteps;        // 64-wide blocks / 16-wide MMA steps that cover taps * 4
  int BN, stages, split;
  long long M;                // B * Ho * Wo output pixels
  int n_tiles, Ho, Wo;
  ConvArgs epi;               // bias / bn / act / output planes as the shared epilogue expects them
};
constexpr int kFirstThreads = 17 * 32;

__global__ void __launch_bounds__(kFirstThreads, 1) conv_first_kernel(const __grid_constant__ FirstArgs fa) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* smem_al = smem_dyn + (smem_base - smem_u32(smem_dyn));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SA = fa.stages, BN = fa.BN, split = fa.split, KB = fa.kblocks;
  const uint32_t a_plane = (uint32_t)KB * kATile;                      // one stage plane (hi or lo)
  const uint32_t a_stage = a_plane * (split ? 2 : 1);
  const uint32_t b_block = (uint32_t)BN * 128u;                        // one k-block of the weight tile
  const uint32_t b_plane = b_block * KB;
  const uint32_t ring_b = smem_base + a_stage * SA;
  const uint32_t bar_base = ring_b + b_plane * (split ? 2 : 1);
  auto fullA = [&](int i) { return bar_base + 8u * i; };
  auto emptyA = [&](int i) { return bar_base + 8u * (SA + i); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * SA + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SA + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_al + (tmem_slot - smem_base));
  float* s_bias = reinterpret_cast<float*>(smem_al + (bar_base + 256u - smem_base));
  float* s_scale = s_bias + fa.epi.cout;
  float* s_shift = s_scale + fa.epi.cout;
  for (int i = threadIdx.x; i < fa.epi.cout; i += blockDim.x) {
    s_bias[i] = fa.epi.bias ? fa.epi.bias[i] : 0.f;
    if (fa.epi.bn_scale) { s_scale[i] = fa.epi.bn_scale[i]; s_shift[i] = fa.epi.bn_shift[i]; }
  }
  {   // resident weights: straight copy of the swizzled images
    const uint4* src_h = reinterpret_cast<const uint4*>(fa.w_hi);
    const uint4* src_l = reinterpret_cast<const uint4*>(fa.w_lo);
    uint4* dst = reinterpret_cast<uint4*>(smem_al + (ring_b - smem_base));
    const int n16 = (int)(b_plane >> 4);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) { dst[i] = src_h[i]; if (split) dst[n16 + i] = src_l[i]; }
  }
  int tmem_cols = 32;
  while (tmem_cols < 2 * BN) tmem_cols <<= 1;
  if (warp == 16) {
    if (lane == 0) {
      for (int i = 0; i < SA; ++i) { mbar_init(fullA(i), 128); mbar_init(emptyA(i), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // the weight copy above is read by the MMA (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int taps = fa.KH * fa.KW;

  if (warp == 16) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(BN);
    int it = 0;
    for (int t = blockIdx.x; t < fa.n_tiles; t += gridDim.x, ++it) {
      const int st = it % SA;
      const int acc = it & 1;
      mbar_wait(fullA(st), (uint32_t)(it / SA) & 1u);
      mbar_wait(tempty_bar(acc), ((uint32_t)(it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        const uint32_t a0 = smem_base + a_stage * st;
        for (int ks = 0; ks < fa.ksteps; ++ks) {
          const int kb = ks >> 2, k = ks & 3;
          const uint32_t ah = (a0 + (uint32_t)kb * kATile) >> 4, bh = (ring_b + (uint32_t)kb * b_block) >> 4;
          const uint64_t da = kDescHi | (uint64_t)(ah + 2 * k), db = kDescHi | (uint64_t)(bh + 2 * k);
          tc_mma(d, da, db, idesc, ks ? 1u : 0u);
          if (split) {
            tc_mma(d, da, kDescHi | (uint64_t)(bh + (b_plane >> 4) + 2 * k), idesc, 1u);
            tc_mma(d, kDescHi | (uint64_t)(ah + (a_plane >> 4) + 2 * k), db, idesc, 1u);
          }
        }
        tc_commit(emptyA(st));
        tc_commit(tfull_bar(acc));
      }
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ===================== gather =====================
    const int grp = (warp - 8) >> 2;
    const int r = ((warp - 8) & 3) * 32 + lane;                        // row of the tile
    const int slots = fa.ksteps * 4;                                   // 8-byte (4-channel) K slots; slots >= taps hold zeros
    int it = 0;
    for (int t = blockIdx.x; t < fa.n_tiles; t += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int st = it % SA;
      const long long v = (long long)t * kBM + r;
      const bool valid = v < fa.M;
      int n = 0, y = 0, x = 0;
      if (valid) {
        n = (int)(v / ((long long)fa.Ho * fa.Wo));
        const int rem = (int)(v - (long long)n * fa.Ho * fa.Wo);
        y = rem / fa.Wo; x = rem - y * fa.Wo;
      }
      const size_t src0 = valid ? act_index(fa.in, n, y, x) : 0;
      // all loads of a round are issued before the first store (twelve taps = the whole 3x3 kernel in one round trip)
      constexpr int kRound = 12;
      uint2 h[kRound], l[kRound];
      for (int s0 = 0; s0 < slots; s0 += kRound) {
#pragma unroll
        for (int j = 0; j < kRound; ++j) {
          const int tap = s0 + j;
          h[j] = make_uint2(0, 0); l[j] = make_uint2(0, 0);
          if (valid && tap < taps) {
            const long long src = (long long)src0 + fa.tap_off[tap];
            h[j] = __ldg(reinterpret_cast<const uint2*>(fa.in.hi + src));
            if (split) l[j] = __ldg(reinterpret_cast<const uint2*>(fa.in.lo + src));
          }
        }
        if (s0 == 0) mbar_wait(emptyA(st), ((uint32_t)(it / SA) & 1u) ^ 1u);   // (the loads above do not touch the stage)
        unsigned char* stage = smem_al + (a_stage * st);
#pragma unroll
        for (int j = 0; j < kRound; ++j) {
          const int tap = s0 + j;
          if (tap < slots) {
            const uint32_t off = (uint32_t)(tap >> 4) * kATile + (uint32_t)r * 128u + ((uint32_t)(((tap >> 1) & 7) ^ (r & 7)) << 4) + (uint32_t)(tap & 1) * 8u;
            *reinterpret_cast<uint2*>(stage + off) = h[j];
            if (split) *reinterpret_cast<uint2*>(stage + a_plane + off) = l[j];
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA's async proxy
      mbar_arrive(fullA(st));
    }
  } else {
    // ===================== epilogue =====================
    const int grp = warp >> 2, q = warp & 3;
    int it = 0;
    for (int t = blockIdx.x; t < fa.n_tiles; t += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int acc = it & 1;
      mbar_wait(tfull_bar(acc), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const long long v = (long long)t * kBM + q * 32 + lane;
      const bool valid = v < fa.M;
      size_t o = 0;
      if (valid) {
        const int n = (int)(v / ((long long)fa.Ho * fa.Wo));
        const int rem = (int)(v - (long long)n * fa.Ho * fa.Wo);
        const int y = rem / fa.Wo, x = rem - y * fa.Wo;
        o = (((size_t)n * fa.epi.out_Hp + (y + fa.epi.out_pad)) * fa.epi.out_Wp + (x + fa.epi.out_pad)) * fa.epi.out_Cs;
      }
      const uint32_t t_row = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(q * 32) << 16);
      epi_split<false>(fa.epi, t_row, 0u, fa.epi.cout, 0, o, valid, s_bias, s_scale, s_shift);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// K-major, 128B-swizzled shared-memory image of the image-facing layer's weights: row = output channel (padded to BN), element
// k = tap * 4 + c.  HWIO kernel in.
void first_weight_image(const float* hwio, int taps, int cin, int cout, int BN, int kblocks, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo) {
  hi.assign((size_t)kblocks * BN * 64, 0); lo.assign((size_t)kblocks * BN * 64, 0);
  auto f2bf = [](float f) { uint32_t u; memcpy(&u, &f, 4); if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);
                            u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); };
  auto bf2f = [](uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; };
  for (int o = 0; o < cout; ++o)
    for (int t = 0; t < taps; ++t)
      for (int c = 0; c < cin; ++c) {
        const int k = t * 4 + c, kb = k >> 6, kk = k & 63;
        const size_t byte = (size_t)kb * BN * 128 + (size_t)o * 128 + ((size_t)((kk >> 3) ^ (o & 7)) << 4) + (size_t)(kk & 7) * 2;
        const float w = hwio[((size_t)t * cin + c) * cout + o];
        const uint16_t h = f2bf(w);
        hi[byte / 2] = h; lo[byte / 2] = f2bf(w - bf2f(h));
      }
}

bool first_border_ok(const ActBuf& in, int kh, int kw, int dil, int pad_t, int pad_l) {
  return in.pad >= pad_t && in.pad >= pad_l && in.pad >= (kh - 1) * dil - pad_t && in.pad >= (kw - 1) * dil - pad_l;
}
int first_tc_supported(int taps, int cin, int cout) { return cin <= 4 && taps * 4 <= 128 && cout % 8 == 0 && cout <= 128; }

int launch_conv_first(ssdk_ctx* ctx, const ActBuf& in, const ActBuf& out, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo,
                      const float* bias, const float* bn_scale, const float* bn_shift, int act, int kh, int kw, int dil, int pad_t,
                      int pad_l, cudaStream_t stream) {
  FirstArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.in = in; fa.w_hi = w_hi; fa.w_lo = w_lo;
  fa.KH = kh; fa.KW = kw; fa.dil = dil; fa.pad_t = pad_t; fa.pad_l = pad_l;
  SSDK_REQUIRE(first_border_ok(in, kh, kw, dil, pad_t, pad_l), "image-facing convolution: the input planes' zero border is too small");
  for (int t = 0; t < kh * kw; ++t) {
    const int dy = (t / kw) * dil - pad_t, dx = (t % kw) * dil - pad_l;
    fa.tap_off[t] = (dy * in.Wp() + dx) * in.Cs;
  }
  const int K = kh * kw * 4;
  fa.ksteps = (K + 15) / 16; fa.kblocks = (K + 63) / 64;
  fa.BN = (out.C + 15) / 16 * 16;
  fa.split = (in.lo && w_lo) ? 1 : 0;
  fa.M = (long long)out.B * out.H * out.W; fa.Ho = out.H; fa.Wo = out.W;
  fa.n_tiles = (int)((fa.M + kBM - 1) / kBM);
  const size_t a_stage = (size_t)fa.kblocks * kATile * (fa.split ? 2 : 1);
  const size_t b_bytes = (size_t)fa.kblocks * fa.BN * 128 * (fa.split ? 2 : 1);
  const size_t fixed = 1024 + b_bytes + 256 + ((size_t)out.C * 3 * sizeof(float) + 127) / 128 * 128 + 128;
  fa.stages = (int)std::min<size_t>(4, (200 * 1024 - fixed) / a_stage);
  SSDK_REQUIRE(fa.stages >= 2, "image-facing convolution: the gathered A tile does not fit in shared memory twice");
  fa.epi.cout = out.C; fa.epi.bias = bias; fa.epi.bn_scale = bn_scale; fa.epi.bn_shift = bn_shift; fa.epi.act = act;
  fa.epi.out_hi = out.hi; fa.epi.out_lo = out.lo; fa.epi.out_Hp = out.Hp(); fa.epi.out_Wp = out.Wp(); fa.epi.out_pad = out.pad; fa.epi.out_Cs = out.Cs;
  const size_t smem = fixed + a_stage * fa.stages;
  static bool attr_set = false;
  if (!attr_set) {
    SSDK_CHECK_CUDA(cudaFuncSetAttribute(conv_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = std::min(fa.n_tiles, ctx->sm_count);
  conv_first_kernel<<<grid, kFirstThreads, smem, stream>>>(fa);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

// Max pooling; out-of-range window positions are ignored (TF 'same' pooling pads with -inf).
// One thread per (pixel, group of 8 channels); the max is taken on the reconstructed value hi + lo.
__global__ void maxpool_kernel(ActBuf in, ActBuf out, int KH, int KW, int stride, int pad_t, int pad_l) {
  const int groups = in.Cs / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)out.B * out.H * out.W * groups;
  if (i >= total) return;
  const int g = (int)(i % groups);
  const size_t pix = i / groups;
  const int xo = (int)(pix % out.W); const int yo = (int)((pix / out.W) % out.H); const int n = (int)(pix / ((size_t)out.W * out.H));
  float best[8];
  uint32_t bh[8], bl[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; bh[e] = 0; bl[e] = 0; }
  for (int ky = 0; ky < KH; ++ky) {
    const int y = yo * stride + ky - pad_t;
    if (y < 0 || y >= in.H) continue;
    for (int kx = 0; kx < KW; ++kx) {
      const int x = xo * stride + kx - pad_l;
      if (x < 0 || x >= in.W) continue;
      const size_t s = act_index(in, n, y, x) + (size_t)g * 8;
      const uint4 h4 = *reinterpret_cast<const uint4*>(in.hi + s);
      uint4 l4 = make_uint4(0, 0, 0, 0);
      if (in.lo) l4 = *reinterpret_cast<const uint4*>(in.lo + s);
      const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t hb = (hw[e >> 1] >> ((e & 1) * 16)) & 0xffffu, lb = (lw[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
        const float v = __uint_as_float(hb << 16) + __uint_as_float(lb << 16);
        if (v > best[e]) { best[e] = v; bh[e] = hb; bl[e] = lb; }
      }
    }
  }
  const size_t o = act_index(out, n, yo, xo) + (size_t)g * 8;
  uint4 oh, ol;
  oh.x = bh[0] | (bh[1] << 16); oh.y = bh[2] | (bh[3] << 16); oh.z = bh[4] | (bh[5] << 16); oh.w = bh[6] | (bh[7] << 16);
  ol.x = bl[0] | (bl[1] << 16); ol.y = bl[2] | (bl[3] << 16); ol.z = bl[4] | (bl[5] << 16); ol.w = bl[6] | (bl[7] << 16);
  *reinterpret_cast<uint4*>(out.hi + o) = oh;
  if (out.lo) *reinterpret_cast<uint4*>(out.lo + o) = ol;
}

// 2x2 / stride 2 without top-left padding (every pool of the VGG trunk but pool5): the four taps are loaded up front (eight
// independent 16-byte loads per thread); a tap that falls off the bottom / right edge ('same' pooling of an odd extent) is
// clamped onto its in-range neighbour, which cannot change a first-maximum scan.
__global__ void __launch_bounds__(256) maxpool2x2_kernel(ActBuf in, ActBuf out) {
  const int groups = in.Cs / 8;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)out.B * out.H * out.W * groups;
  if (i >= total) return;
  const int g = (int)(i % groups);
  const size_t pix = i / groups;
  const int xo = (int)(pix % out.W); const int yo = (int)((pix / out.W) % out.H); const int n = (int)(pix / ((size_t)out.W * out.H));
  const int y0 = 2 * yo, x0 = 2 * xo, y1 = min(y0 + 1, in.H - 1), x1 = min(x0 + 1, in.W - 1);
  const size_t s00 = act_index(in, n, y0, x0) + (size_t)g * 8, s01 = act_index(in, n, y0, x1) + (size_t)g * 8;
  const size_t s10 = act_index(in, n, y1, x0) + (size_t)g * 8, s11 = act_index(in, n, y1, x1) + (size_t)g * 8;
  uint4 h[4], l[4];
  h[0] = __ldcs(reinterpret_cast<const uint4*>(in.hi + s00)); h[1] = __ldcs(reinterpret_cast<const uint4*>(in.hi + s01));
  h[2] = __ldcs(reinterpret_cast<const uint4*>(in.hi + s10)); h[3] = __ldcs(reinterpret_cast<const uint4*>(in.hi + s11));
  if (in.lo) {
    l[0] = __ldcs(reinterpret_cast<const uint4*>(in.lo + s00)); l[1] = __ldcs(reinterpret_cast<const uint4*>(in.lo + s01));
    l[2] = __ldcs(reinterpret_cast<const uint4*>(in.lo + s10)); l[3] = __ldcs(reinterpret_cast<const uint4*>(in.lo + s11));
  } else {
    l[0] = l[1] = l[2] = l[3] = make_uint4(0, 0, 0, 0);
  }
  float best[8];
  uint32_t bh[8], bl[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; bh[e] = 0; bl[e] = 0; }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const uint32_t hw[4] = {h[t].x, h[t].y, h[t].z, h[t].w}, lw[4] = {l[t].x, l[t].y, l[t].z, l[t].w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t hb = (hw[e >> 1] >> ((e & 1) * 16)) & 0xffffu, lb = (lw[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
      const float v = __uint_as_float(hb << 16) + __uint_as_float(lb << 16);
      if (v > best[e]) { best[e] = v; bh[e] = hb; bl[e] = lb; }
    }
  }
  const size_t o = act_index(out, n, yo, xo) + (size_t)g * 8;
  uint4 oh, ol;
  oh.x = bh[0] | (bh[1] << 16); oh.y = bh[2] | (bh[3] << 16); oh.z = bh[4] | (bh[5] << 16); oh.w = bh[6] | (bh[7] << 16);
  ol.x = bl[0] | (bl[1] << 16); ol.y = bl[2] | (bl[3] << 16); ol.z = bl[4] | (bl[5] << 16); ol.w = bl[6] | (bl[7] << 16);
  *reinterpret_cast<uint4*>(out.hi + o) = oh;
  if (out.lo) *reinterpret_cast<uint4*>(out.lo + o) = ol;
}

int launch_maxpool(ssdk_ctx* ctx, const ActBuf& in, const ActBuf& out, int kh, int kw, int stride, int pad_t, int pad_l,
                   cudaStream_t stream) {
  const size_t total = (size_t)out.B * out.H * out.W * (in.Cs / 8);
  if (kh == 2 && kw == 2 && stride == 2 && pad_t == 0 && pad_l == 0 && 2 * (out.H - 1) < in.H && 2 * (out.W - 1) < in.W) {
    maxpool2x2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out);
    SSDK_COUNT_LAUNCH(ctx);
    SSDK_CHECK_CUDA(cudaGetLastError());
    return SSDK_OK;
  }
  maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out, kh, kw, stride, pad_t, pad_l);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

// L2Normalization (keras_layer_L2Normalization.py:61-63): x * rsqrt(max(sum_c x^2, 1e-12)) * gamma_c, one warp per pixel.
__global__ void l2norm_kernel(ActBuf in, ActBuf out, const float* __restrict__ gamma) {
  const size_t pix = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)in.B * in.H * in.W;
  if (pix >= total) return;
  const int x = (int)(pix % in.W); const int y = (int)((pix / in.W) % in.H); const int n = (int)(pix / ((size_t)in.W * in.H));
  const size_t s = act_index(in, n, y, x), o = act_index(out, n, y, x);
  float ss = 0.f;
  for (int c = lane; c < in.C; c += 32) { float v = split_load(in, s + c); ss += v * v; }
  ss = warp_sum(ss);
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
  for (int c = lane; c < in.C; c += 32) split_store(out, o + c, split_load(in, s + c) * inv * __ldg(gamma + c));
}

// The same with eight channels (16 bytes of each plane) per lane and step: up to 512 channels stay in registers between the two
// passes.  The sum of squares is formed in a different order than above (per-lane partial sums of 8, then the warp tree).
__global__ void __launch_bounds__(256) l2norm8_kernel(ActBuf in, ActBuf out, const float* __restrict__ gamma) {
  const size_t pix = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)in.B * in.H * in.W;
  if (pix >= total) return;
  const int x = (int)(pix % in.W); const int y = (int)((pix / in.W) % in.H); const int n = (int)(pix / ((size_t)in.W * in.H));
  const size_t s = act_index(in, n, y, x), o = act_index(out, n, y, x);
  float v[2][8];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int c = (k * 32 + lane) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[k][e] = 0.f;
    if (c < in.C) {
      const uint4 h4 = *reinterpret_cast<const uint4*>(in.hi + s + c);
      uint4 l4 = make_uint4(0, 0, 0, 0);
      if (in.lo) l4 = *reinterpret_cast<const uint4*>(in.lo + s + c);
      const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t hb = (hw[e >> 1] >> ((e & 1) * 16)) & 0xffffu, lb = (lw[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
        v[k][e] = __uint_as_float(hb << 16) + __uint_as_float(lb << 16);
        ss += v[k][e] * v[k][e];
      }
    }
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int c = (k * 32 + lane) * 8;
    if (c < in.C) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float f0 = v[k][j * 2] * inv * gm[j * 2], f1 = v[k][j * 2 + 1] * inv * gm[j * 2 + 1];
        const __nv_bfloat162 hh = __floats2bfloat162_rn(f0, f1);
        const uint32_t hp = *reinterpret_cast<const uint32_t*>(&hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(f0 - __uint_as_float(hp << 16), f1 - __uint_as_float(hp & 0xffff0000u));
        ph[j] = hp; pl[j] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      *reinterpret_cast<uint4*>(out.hi + o + c) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
      if (out.lo) *reinterpret_cast<uint4*>(out.lo + o + c) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
  }
}

int launch_l2norm(ssdk_ctx* ctx, const ActBuf& in, const ActBuf& out, const float* gamma, cudaStream_t stream) {
  const size_t total = (size_t)in.B * in.H * in.W;
  if (in.C % 8 == 0 && in.C <= 512 && in.Cs == in.C && out.Cs == out.C && (reinterpret_cast<uintptr_t>(gamma) & 15) == 0) {
    l2norm8_kernel<<<(unsigned)((total + 7) / 8), 256, 0, stream>>>(in, out, gamma);
    SSDK_COUNT_LAUNCH(ctx);
    SSDK_CHECK_CUDA(cudaGetLastError());
    return SSDK_OK;
  }
  l2norm_kernel<<<(unsigned)((total + 7) / 8), 256, 0, stream>>>(in, out, gamma);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

// Reshape / softmax / Concat (models/keras_ssd300.py:363-419): one warp per prior.
// head row layout per pixel: n_boxes x [C class logits | 4 box offsets].
__global__ void head_finalize_kernel(const float* __restrict__ head, int B, int HW, int n_boxes, int C, int P, int prior_off,
                                     const float* __restrict__ anchors, float4 var, float* __restrict__ y_pred) {
  const size_t wid = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const size_t total = (size_t)B * HW * n_boxes;
  if (wid >= total) return;
  const int b = (int)(wid % n_boxes); const size_t pix = wid / n_boxes;
  const int n = (int)(pix / HW); const int p_local = (int)(pix % HW) * n_boxes + b;
  const float* src = head + pix * (size_t)n_boxes * (C + 4) + (size_t)b * (C + 4);
  float* dst = y_pred + ((size_t)n * P + prior_off + p_local) * (C + 12);
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, src[c]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(src[c] - mx);
  sum = warp_sum(sum);
  for (int c = lane; c < C; c += 32) dst[c] = expf(src[c] - mx) / sum;
  if (lane < 4) {
    dst[C + lane] = src[C + lane];
    dst[C + 4 + lane] = anchors[(size_t)(prior_off + p_local) * 4 + lane];
    const float v[4] = {var.x, var.y, var.z, var.w};
    dst[C + 8 + lane] = v[lane];
  }
}

int launch_head_finalize(ssdk_ctx* ctx, const float* head, int B, int HW, int n_boxes, int C, int P, int prior_off,
                         const float* anchors, const float* variances, float* y_pred, cudaStream_t stream) {
  const size_t total = (size_t)B * HW * n_boxes;
  head_finalize_kernel<<<(unsigned)((total + 7) / 8), 256, 0, stream>>>(head, B, HW, n_boxes, C, P, prior_off, anchors,
                                                                        make_float4(variances[0], variances[1], variances[2], variances[3]), y_pred);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

__global__ void l2norm_f32_kernel(const float* __restrict__ x, long long rows, int C, const float* __restrict__ gamma,
                                  float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* p = x + r * C;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) { float v = p[c]; ss += v * v; }
  ss = warp_sum(ss);
  const float inv = rsqrtf(fmaxf(ss, 1e-12f));
  for (int c = lane; c < C; c += 32) out[r * C + c] = p[c] * inv * __ldg(gamma + c);
}

__global__ void unpack_kernel(ActBuf in, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)in.B * in.H * in.W * in.C;
  if (i >= total) return;
  const int c = (int)(i % in.C); const size_t pix = i / in.C;
  const int x = (int)(pix % in.W); const int y = (int)((pix / in.W) % in.H); const int n = (int)(pix / ((size_t)in.W * in.H));
  out[i] = split_load(in, act_index(in, n, y, x) + c);
}

// float32 NHWC tensor -> split bf16 planes of an activation buffer (SSDK_OP_TENSOR); channels beyond C stay zero
__global__ void pack_kernel(const float* __restrict__ in, ActBuf out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)out.B * out.H * out.W * out.C;
  if (i >= total) return;
  const int c = (int)(i % out.C); const size_t pix = i / out.C;
  const int x = (int)(pix % out.W); const int y = (int)((pix / out.W) % out.H); const int n = (int)(pix / ((size_t)out.W * out.H));
  split_store(out, act_index(out, n, y, x) + c, in[i]);
}

int launch_pack(ssdk_ctx* ctx, const float* in, const ActBuf& out, cudaStream_t stream) {
  const size_t total = (size_t)out.B * out.H * out.W * out.C;
  pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

int launch_unpack(ssdk_ctx* ctx, const ActBuf& in, float* out, cudaStream_t stream) {
  const size_t total = (size_t)in.B * in.H * in.W * in.C;
  unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, out);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}

}  // namespace ssdk

extern "C" int ssdk_l2_normalize(ssdk_ctx* ctx, const float* x_dev, long long rows, int C, const float* gamma_dev, float* out_dev,
                                 void* stream) {
  using namespace ssdk;
  SSDK_REQUIRE(ctx && x_dev && gamma_dev && out_dev && rows > 0 && C > 0, "ssdk_l2_normalize: bad argument");
  l2norm_f32_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x_dev, rows, C, gamma_dev, out_dev);
  SSDK_COUNT_LAUNCH(ctx);
  SSDK_CHECK_CUDA(cudaGetLastError());
  return SSDK_OK;
}
