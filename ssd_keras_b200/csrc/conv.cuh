// Internal interface between model.cu (plan building) and conv.cu (kernels).
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <vector>

namespace ssdk {

// An activation tensor: NHWC, channel-padded to a multiple of 8, spatially padded with a zero border
// of `pad` pixels, stored as TWO bf16 planes: value = hi + lo (lo is absent in single-pass bf16 mode).
// hi + lo carries 16 significant bits, which keeps the tcgen05 path within ~1e-5 of an fp32 conv.
struct ActBuf {
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
  int B = 0, H = 0, W = 0, C = 0;   // logical shape
  int Cs = 0;                        // stored channels (multiple of 8)
  int pad = 0;
  // shared = 1 (inference plans): the row pitch is W + pad, not W + 2*pad -- the right border of a row IS the left border of the
  // next one (the same zeros).  Every index stays ((n*Hp + y + pad)*Wp + x + pad)*Cs; a tap that runs off the right edge lands
  // in the next row's left border.  The implicit GEMM computes every position of the stored grid, so this cuts its padding rows:
  // fc6 (19 wide, border 6) goes from 61 % to 76 % valid rows, the 19-wide conv5_x from 90 % to 95 %.
  int shared = 0;
  __host__ __device__ int Hp() const { return H + 2 * pad; }
  __host__ __device__ int Wp() const { return W + (shared ? pad : 2 * pad); }
  size_t rows() const { return (size_t)B * Hp() * Wp(); }
  size_t elems() const { return rows() * Cs; }
};

enum { EPI_SPLIT = 0, EPI_F32 = 1, EPI_ATOMIC = 2, EPI_HEAD = 3 };

constexpr int kMaxTaps = 25;

struct ConvArgs {
  // GEMM view
  int M_total;          // virtual rows (B * rows_per_img)
  int rows_per_img;     // rows of the virtual grid per image
  int in_Wp;            // virtual row pitch
  int Ho, Wo, B;        // valid extent: virtual (y, x) is a real output iff y < Ho && x < Wo
  int KH, KW, kblocks;  // K loop = KH * kblocks A-slabs, each feeding KW weight tiles (taps (kh, 0..KW-1))
  int last_ksteps;      // UMMA k-steps (of 16) in the last k-block of every tap (1..4)
  int row_shift[8];     // row offset of tap (kh, kw=0)
  int kw_rows;          // rows between consecutive kw taps (= dilation); tap kw reads slab rows [kw*kw_rows, +128)
  int slab_rows;        // rows per A slab = round_up8(128 + (KW-1)*kw_rows) <= 256
  int stages_a, stages_b;
  int acc_split;        // 1: the two cross terms (hi*lo, lo*hi) accumulate in their own TMEM columns and are added in the epilogue
  int fuse_b;           // 1 (needs acc_split): A_hi * [B_hi ; B_lo] as ONE MMA of N = BN + n into [main | cross], then A_lo * B_hi into cross
  int acc_bufs;         // accumulator sets in TMEM (2: the epilogue of tile i overlaps the MMAs of tile i+1; 1 when columns run out)
  int resident_b;       // 1: all KH*KW*kblocks weight tiles of the (single) n-tile stay in shared memory for the whole launch
  int mt;               // m-tiles per work unit (1 or 2): with 2, every weight tile feeds two 128-row MMAs (halves the weight traffic)
  int bo_mode;          // how the UMMA descriptor's base-offset field is filled for row-shifted A tiles (debug knob)
  int epi_pipe;         // forward epilogue: 2 = TMEM loads of the next 32 columns issued before the current ones are converted / stored, 1 / 0 = experiment knobs
  int n_tiles_m;        // number of entries in tile_list
  int n_tiles_n;
  const int* tile_list; // m-tile indices that contain at least one valid row
  int BN;               // accumulator tile width (64/128/256); TMA box rows of the weight tile
  int cout;
  int split;            // 1: bf16x3 (hi*hi + hi*lo + lo*hi), 0: single bf16 pass
  // epilogue
  int epi;
  const float* bias; const float* bn_scale; const float* bn_shift;
  int act;
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int out_Hp, out_Wp, out_pad, out_Cs;
  float* out_f32;       // EPI_F32: compact [B*Ho*Wo][cout]; EPI_ATOMIC: atomicAdd into out_f32[row*out_ld + out_col_off + col]
  // --- extensions used by the backward pass ---
  int k_split;          // >1: the K loop (KH == KW == 1 only) is cut into k_split ranges, one work unit each (EPI_ATOMIC)
  int kb_per;           // k-blocks per range
  int b_k_offset;       // added to the K coordinate of the weight-side operand (row-shifted taps of a transposed tensor)
  int out_ld, out_col_off;
  const __nv_bfloat16* mask_hi;   // EPI_SPLIT: zero the result where this plane (same geometry as the output) is <= 0 (ReLU')
  int accumulate;       // EPI_SPLIT: add to the value already stored in the output planes
  // --- EPI_HEAD: Reshape / softmax / Concat of models/keras_ssd300.py:363-419 in the predictor conv's epilogue.  The tile's columns
  //     are n_boxes x [C class logits | 4 offsets]; every (pixel, box) becomes one row of y_pred (out_f32):
  //     [softmax(C) | 4 offsets | 4 anchor coordinates | 4 variances] at prior head_prior_off + pixel * n_boxes + box.
  int head_nb, head_C, head_P, head_prior_off;
  const float* head_anchors;      // [P*4]
  float head_var[4];
};

struct ConvLaunch {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  ConvArgs args;
  int grid;
  size_t smem;
  double flops_algo, flops_issued;
};

int tma_init();   // resolves cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency)
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_rows);
// 4-D map over a bf16 NHWC tensor, dims / box innermost first (channels, x, y, image); 128B swizzle, zero fill
int make_tmap_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint32_t box[4]);
size_t conv_smem_bytes(const ConvArgs& a);
void conv_pick_stages(ConvArgs& a);
int launch_conv(ssdk_ctx* ctx, const ConvLaunch& L, cudaStream_t stream, int grid_cap = 0);   // grid_cap > 0: at most that many CTAs

// elementwise / data movement kernels
int launch_preprocess(ssdk_ctx* ctx, const float* images, int B, int H, int W, int Cimg, const float* mean, const float* stddev,
                      const int* swap, const ActBuf& out, cudaStream_t stream);
int launch_im2col(ssdk_ctx* ctx, const ActBuf& in, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int Ho, int Wo, int kh, int kw,
                  int stride, int dil, int pad_t, int pad_l, int Kpad, cudaStream_t stream);
int launch_conv_direct(ssdk_ctx* ctx, const ActBuf& in, const ActBuf& out, const float* w, const float* bias, const float* bn_scale,
                       const float* bn_shift, int act, int kh, int kw, int dil, int pad_t, int pad_l, cudaStream_t stream);
// image-facing layer on the tensor cores (gathered A tile, weights resident in shared memory as a swizzled image)
int first_tc_supported(int taps, int cin, int cout);
bool first_border_ok(const ActBuf& in, int kh, int kw, int dil, int pad_t, int pad_l);
void first_weight_image(const float* hwio, int taps, int cin, int cout, int BN, int kblocks, std::vector<uint16_t>& hi, std::vector<uint16_t>& lo);
int launch_conv_first(ssdk_ctx* ctx, const ActBuf& in, const ActBuf& out, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo,
                      const float* bias, const float* bn_scale, const float* bn_shift, int act, int kh, int kw, int dil, int pad_t,
                      int pad_l, cudaStream_t stream);
int launch_maxpool(ssdk_ctx* ctx, const ActBuf& in, const ActBuf& out, int kh, int kw, int stride, int pad_t, int pad_l,
                   cudaStream_t stream);
int launch_l2norm(ssdk_ctx* ctx, const ActBuf& in, const ActBuf& out, const float* gamma, cudaStream_t stream);
int launch_head_finalize(ssdk_ctx* ctx, const float* head, int B, int HW, int n_boxes, int C, int P, int prior_off,
                         const float* anchors, const float* variances, float* y_pred, cudaStream_t stream);
int launch_unpack(ssdk_ctx* ctx, const ActBuf& in, float* out, cudaStream_t stream);
int launch_pack(ssdk_ctx* ctx, const float* in, const ActBuf& out, cudaStream_t stream);      // float32 NHWC -> hi/lo planes (interior only)

}  // namespace ssdk
