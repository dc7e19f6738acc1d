// tc_gen.h - general-shape tcgen05 / TMEM / TMA convolution engine (CUDA build only; not emulated).
//
// Every Conv2D / Conv2DTranspose of the model (src/abstract_cvae.py:30-33, 81-89) and each of their two backward
// products is an implicit GEMM over *plane tensors*: bf16 [B][PL][H][W][8], one 16-byte unit per (pixel, 8-channel
// chunk), PL chunk planes per image.  Three element maps are used:
//   PLAIN  plane = chunk                         (decoder activations, encoder gradients)
//   S2D    plane = (row parity*2 + col parity)*KC + chunk at (y/2, x/2)   (encoder activations, decoder gradients)
//   X3     the 3-channel input image, 2x2 space-to-depth, 12 values packed into 2 planes
//   X27    the 3-channel input image as the 27-value 3x3 stride-2 patch of every output pixel (4 planes): the first layer's
//          weight gradient is then ONE tap (2 MMAs per K step instead of 5)
// A stride-2 Conv2D over an S2D tensor and a stride-2 Conv2DTranspose over a PLAIN tensor are both stride-1 products
// whose taps are descriptor start offsets into one TMA-loaded halo tile (tc_common.cuh), so one forward kernel
// (tc_gconv_kernel, MMA list built on the host) covers forward and data-gradient of every layer, and one
// pixel-K kernel (tc_gwgrad_kernel) covers every weight / bias gradient.  SPLIT operands (bf16 hi + lo planes,
// xh*wh + xl*wh + xh*wl) give fp32-grade products where the loss needs them.
#pragma once
#include "common.cuh"

namespace kc {

enum GenLayout { GEN_PLAIN = 0, GEN_S2D = 1, GEN_X3 = 2, GEN_X27 = 3 };
enum GenKind {
  GEN_CONV_S2 = 0,    // K side: S2D / X3 input, taps (di,dj) in {0,1}^2   (Conv2D s2 forward, Conv2DTranspose s2 dgrad)
  GEN_CONVT_S2 = 1,   // K side: PLAIN input, N side: (col parity, channel), one group per row parity (ConvT s2 fwd, Conv2D s2 dgrad)
  GEN_CONV_S1 = 2,    // K side: PLAIN input, 9 taps (output layer forward and dgrad)
  GEN_DENSE = 3,      // a Dense layer product seen from its LONG dimension: GEMM rows = the n of a [R][n] matrix viewed as
                      // [n/32][32] "pixels", K side = that matrix as planes [R/8][n][8] (gen_pack_rows_T), columns = the short
                      // dimension (batch rows of z / latent columns) supplied through the weight image; one tap
};
enum GenPre { GEN_PRE_NONE = 0, GEN_PRE_BIAS_RELU = 1, GEN_PRE_BIAS_SIGMOID = 2, GEN_PRE_BIAS = 3 };

// shape of one product; fixed per layer (independent of the batch)
struct GenConvSpec {
  int kind;        // GenKind
  int in_layout;   // GenLayout of the K-side tensor
  int Ck;          // real K-side channels (per parity for S2D)
  int Cn;          // real N-side channels (per parity for GEN_CONVT_S2)
  int KCk;         // 8-channel chunks per parity on the K side (planes = KCk, 4*KCk or 2 for X3), hi planes only
  int w_mode;      // fp32 weight element (tap, k, n): 0 -> (tap*Ck + k)*Cn + n ; 1 -> (tap*Cn + n)*Ck + k
  int flip;        // GEN_CONV_S1: 1 = in[y+1-kh] (Conv2DTranspose s1 forward), 0 = in[y-1+kh]
  int split;       // K-side tensor carries lo planes behind the hi planes; weight image has a lo half
  int w_stride;    // GEN_DENSE: row stride of the fp32 "weight" source (0: Cn for w_mode 0, Ck for w_mode 1)
  int ones_col1;   // GEN_DENSE: 1 + index of a column whose B operand is all ones (row sums = a bias gradient), 0 = none
  int ones_src;    // GEN_DENSE: index of a 1.0f in the fp32 "weight" source (read for the ones column)
  int w_col0;      // GEN_DENSE: first source column (wide weight gradients run as several column blocks of <= 256)
  int Hg, Wg;      // GEMM grid: output pixels per image (CONV_S2 / CONV_S1) or input pixels (CONVT_S2)
};

struct GenConvPlan;   // opaque: device-resident MMA list, weight gather table, tile geometry
// upload = false: host-side plan only (no device copies): what kcvae_gen_plan_dump hands to the CPU-side plan simulator
GenConvPlan* gen_conv_plan_create(const GenConvSpec& s, const char** why_not, bool upload = true);
// flat int32 description of a plan (tests/engine_sim.py documents the layout); returns the length needed
int64_t gen_conv_plan_dump(const GenConvPlan* p, int32_t* out, int64_t capacity);
void gen_conv_plan_free(GenConvPlan* p);
size_t gen_conv_weight_image_bytes(const GenConvPlan* p);
int gen_conv_Cop(const GenConvPlan* p);          // padded N channels (per parity)
// host copy of the plan's gather table: one entry per bf16 element of the image; >= 0: index into the layer's fp32 weight
// tensor (| GEN_LO_FLAG: the lo part of that weight), -1: structural zero.  A model concatenates the tables of all its
// plans (adding each variable's offset in the flat parameter vector) and builds every image with ONE gen_gather_weights.
constexpr int32_t GEN_LO_FLAG = 0x40000000;
const int32_t* gen_conv_table(const GenConvPlan* p, size_t* n);
void gen_gather_weights(const float* w, const int32_t* table_dev, int64_t n, void* img, cudaStream_t st);
// fp32 weights -> bf16 B-operand image of this plan (hi half, then lo half when split)
void gen_conv_prep_weights(const GenConvPlan* p, const float* w, void* img, cudaStream_t st);

struct GenPlanes {     // a plane tensor
  void* base;
  int layout;          // GenLayout
  int KC;              // chunks per parity
  int split;           // lo planes present
  int H, W;            // plane dims (pixels)
  int planes() const { return (layout == GEN_S2D ? 4 * KC : (layout == GEN_X3 ? 2 : (layout == GEN_X27 ? 4 : KC))) * (split ? 2 : 1); }
  size_t units(int B) const { return (size_t)B * planes() * H * W; }
};

struct GenEpilogue {
  int pre;                 // GenPre
  const float* bias;       // [Cn]
  const GenPlanes* mask;   // multiply by (mask > 0): activation stored at the OUTPUT pixel (PLAIN or S2D), hi planes; or nullptr
  const GenPlanes* out;    // bf16 plane output (PLAIN or S2D; split -> hi + lo) or nullptr
  float* out_f32;          // fp32 NHWC [B,Ho,Wo,Cn] or nullptr (GEN_DENSE: [Cn][dense_ld], element (column, row))
  int dense_n, dense_ld;   // GEN_DENSE: valid GEMM rows n and leading dimension of out_f32; bias is per ROW; `out` planes are
  int dense_cc;            //   written as [column][chunk(c)][pixel p][8] with row n = p * dense_cc + c (the Dense output as an image)
  int dense_tr;            // GEN_DENSE: out_f32 element (column, row) at row * dense_ld + column instead (a [rows][columns] matrix);
                           //   mask_f32 is read at the same index as out_f32
  const float* mask_f32;   // fp32 NHWC mask [B,Ho,Wo,Cn] (alternative to `mask`) or nullptr
};
// in: K-side plane tensor; returns 0 on success (1: no driver entry point, 2: tensor map encode failed)
int gen_conv_run(const GenConvPlan* p, const GenPlanes& in, const void* wimg, const GenEpilogue& e, int B, int* error_flag,
                 const char* name, cudaStream_t st);

// ---- weight / bias gradient: dW[tap] = sum_pixels S[pixel + shift(tap)] (x) U[pixel] --------------------------------
struct GenWgradSpec {
  int kind;            // GenKind of the FORWARD layer
  int flip;            // GEN_CONV_S1 forward flip
  int s_layout, s_KC;  // shifted operand (the layer input): S2D / X3 for CONV_S2, PLAIN otherwise; hi planes
  int u_layout, u_KC;  // unshifted operand (the output gradient): PLAIN for CONV_S2 / CONV_S1, S2D for CONVT_S2
  int Cs, Cu;          // real channels of the input / output-gradient tensors (per parity for S2D)
  int w_mode;          // dW element (tap, s-channel, u-channel): 0 -> (tap*Cs + cs)*Cu + cu ; 1 -> (tap*Cu + cu)*Cs + cs
  int Hg, Wg;          // pixel grid both operands are indexed on (plane dims of U)
  int split_dense;     // GEN_DENSE: both tensors carry lo planes right behind their hi planes (s_KC / u_KC count hi + lo); every
                       // output sums the hi*hi, hi*lo and lo*hi accumulators (fp32-grade Dense forward over a long K)
};
struct GenWgradPlan;
GenWgradPlan* gen_wgrad_plan_create(const GenWgradSpec& s, const char** why_not, bool upload = true);
int64_t gen_wgrad_plan_dump(const GenWgradPlan* p, int32_t* out, int64_t capacity);
void gen_wgrad_plan_free(GenWgradPlan* p);
size_t gen_wgrad_partial_floats(const GenWgradPlan* p);
// dW (9*Cs*Cu floats) and db (Cu floats, may be nullptr) from S (shifted input planes) and U (gradient planes)
int gen_wgrad_run(const GenWgradPlan* p, const GenPlanes& S, const GenPlanes& U, float* dW, float* db, float* partial, int B,
                  int* error_flag, const char* name, cudaStream_t st);

// ---- packers -----------------------------------------------------------------------------------------------------------
// x fp32 NHWC [B,H,W,3] (H, W even) -> X3 planes [B][2 (x2 when split)][H/2][W/2][8]
void gen_pack_x3(const float* x, int B, int H, int W, int split, void* out, cudaStream_t st, void* x27_out = nullptr);
// (x27_out: also the X27 patch planes [B][4][H/2][W/2][8], hi only - the shifted operand of the first layer's weight gradient)
// fp32 NHWC [B,H,W,C] -> PLAIN / S2D planes (channels padded with zeros; split -> lo planes too)
void gen_pack_nhwc(const float* in, int B, int H, int W, int C, const GenPlanes& out, cudaStream_t st);
// fp32 [R][N] row-major -> planes [ceil(R/8) (x2 when split)][N][8]: unit (chunk, n) holds rows 8*chunk .. 8*chunk+7 of column n
// (rows >= R are zero).  The K-side tensor of GEN_DENSE products (n viewed as [N/32][32] pixels; N % 32 == 0).
void gen_pack_rows_T(const float* in, int R, int N, int split, void* out, cudaStream_t st, int Np = 0, int ones_n = -1);
// (Np > N: plane pitch, columns N..Np-1 are zero except column ones_n, which holds 1.0 for every row < R: a "ones pixel"
// through which a pixel-K product adds a bias)
// fp32 [N][C] row-major -> planes [ceil(C/8) (x2 when split)][Np][8]: unit (chunk, n) = in[n][8*chunk .. 8*chunk+7]; column
// ones_n (>= N) holds bias[c]; other columns >= N are zero
void gen_pack_cols(const float* in, const float* bias, int N, int C, int split, void* out, int Np, int ones_n, cudaStream_t st);
// planes -> fp32 NHWC [B,H,W,C] (hi + lo when split); tests / debug
void gen_unpack_nhwc(const GenPlanes& in, int B, int H, int W, int C, float* out, cudaStream_t st);

}  // namespace kc
