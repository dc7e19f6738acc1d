// tc_conv.h - launchers of the tcgen05 / TMA kernels (CUDA build only; not emulated).
#pragma once
#include "common.cuh"

namespace kc {

bool tc_out_conv_supported(int Cin, int Cout);
size_t tc_out_weight_image_elems(int Cin);   // bf16 elements
// fp32 NHWC [B,HW,C] -> chunk-planar bf16 [B][C/8][HW][8], the layout every tensor-core consumer of the last
// decoder activation reads (one TMA box row = 32 pixels x 16 B contiguous)
void cast_f32_to_bf16_planar(const float* in, void* out_bf16, int B, int64_t HW, int C, cudaStream_t st);
// W [3,3,Cout,Cin] fp32 (Keras Conv2DTranspose layout) -> UMMA B-operand image (bf16)
void tc_prep_out_weights(const float* w, int Cout, int Cin, void* img_bf16, cudaStream_t st);
// x_hat = [sigmoid](bias + conv3x3_s1_flipped(act)) ; act bf16 NHWC [B,H,W,Cin]; returns 0 on success
int tc_out_conv(const void* act_bf16, const void* wimg_bf16, const float* bias, float* xhat, int B, int H, int W,
                int Cin, int Cout, int apply_sigmoid, int* error_flag, cudaStream_t st);

// output-layer data gradient: dl8 = d(loss)/d(logit) as bf16 NHWC padded to 8 channels
bool tc_out_dgrad_supported(int Cin, int Cout);
size_t tc_dgrad_weight_image_elems();
void tc_prep_dgrad_weights(const float* w, int Cout, int Cin, void* img_bf16, cudaStream_t st);
// g_out fp32 NHWC and/or g_s2d bf16 space-to-depth, chunk-planar [B][4 parities][Cin/8][H/2][W/2][8] (either may be nullptr)
// chan_sum [Cin] (optional) = sum over all pixels of g = bias gradient of the producing layer;
// chan_partial >= 148*8*32 floats of scratch
// relu_bits (optional, [B,H,W] uint32 written by tc_tail_fused): replaces the reads of mask_bf16
int tc_out_dgrad(const void* dl8_bf16, const void* wimg_bf16, const void* mask_bf16, const uint32_t* relu_bits, float* g_out,
                 void* g_s2d_bf16,
                 float* chan_sum, float* chan_partial, int B, int H, int W, int Cin, int* error_flag, cudaStream_t st);
// backward of the last Conv2DTranspose s2 (Cin <= 8 -> 32) from the space-to-depth gradient
bool tc_convT_bwd_supported(int Cin, int Cout, int h, int w);
size_t tc_convT_dgrad_weight_image_elems();
size_t tc_convT_wgrad_partial_floats(int Cin);
void tc_prep_convT_dgrad_weights(const float* w, int Cout, int Cin, void* img_bf16, cudaStream_t st);
int tc_convT_dgrad(const void* g_s2d, const void* wimg, const float* mask, float* g_prev, int B, int h, int w, int Cin,
                   int* error_flag, cudaStream_t st, void* g_planes = nullptr, int g_KC = 0);
int tc_convT_wgrad(const void* g_s2d, const void* a_prev8, float* dW, float* partial, int B, int h, int w, int Cin,
                   int* error_flag, cudaStream_t st);

// Conv2DTranspose s2 (Cin <= 8 -> 32) + bias + ReLU by sub-pixel phases; bf16 NHWC output
bool tc_convT_fwd_supported(int Cin, int Cout);
size_t tc_convT_weight_image_elems();
void pack_c8_bf16(const float* in, int64_t npix, int C, void* out_bf16x8, cudaStream_t st);
void tc_prep_convT_weights(const float* w, int Cout, int Cin, void* img_bf16, cudaStream_t st);
int tc_convT_fwd(const void* in8_bf16, const void* wimg_bf16, const float* bias, void* out_bf16, int B, int h, int w,
                 int* error_flag, cudaStream_t st);

// fused decoder tail: Conv2DTranspose s2 (Cprev <= 8 -> 32) -> Conv2DTranspose s1 (32 -> Cout) with the
// 32-channel activation kept in shared memory; optional sigmoid, x_hat, error map and per-frame score
// Conv2DTranspose s2 forward 32 -> few channels (Cout <= 8) from a chunk-planar bf16 input [B][4][h][w][8]:
// out8 = bf16 [B,2h,2w,8] (the input layout of the few -> 32 layers), out_f32 = fp32 [B,2h,2w,Cout]; either may be nullptr
bool tc_convT_few_fwd_supported(int Cin, int Cout);
size_t tc_convT_few_weight_image_elems();
void tc_prep_convT_few_weights(const float* w, int Cout, int Cin, void* img_bf16, cudaStream_t st);
// split != 0: the input holds bf16 hi and lo planes ([B][hi 4 | lo 4][h][w][8]) and the products are xh*wh + xl*wh + xh*wl
// (fp32-grade accuracy; the training forward)
int tc_convT_few_fwd(const void* in_planar_bf16, const void* wimg, const float* bias, void* out8_bf16, float* out_f32, int B, int h,
                     int w, int Cout, int split, int* error_flag, cudaStream_t st);
bool tc_tail_fused_supported(int Cprev, int Clast, int Cout, int H, int W);
// weight image the fused tail reads for its output convolution (layout depends on Cout); img = tc_out_weight_image_elems()
void tc_prep_tail_weights(const float* w, int Cout, int Cin, void* img_bf16, cudaStream_t st);
// the four weight images a training step needs (convT fwd, fused tail, out dgrad, convT dgrad) in one launch
void tc_prep_all_weights(const float* w_convT, const float* w_out, int Cprev, int Clast, int Cout, void* img_convT,
                         void* img_tail, void* img_dgrad, void* img_convT_dgrad, cudaStream_t st);
size_t tc_tail_score_partial_floats(int B, int H, int W);
// a_last_planar (optional): chunk-planar bf16 copy of the intermediate activation for the backward pass
int tc_tail_fused(const void* in8_bf16, const void* wimgA, const void* wimgB, const float* biasA, const float* biasB,
                  const float* x, float* xhat, void* a_last_planar, uint32_t* relu_bits, float* err, float* score,
                  float* err_minmax, float* score_partial, int B, int H, int W, int Cout, int apply_sigmoid, int* error_flag,
                  cudaStream_t st);

// output-layer weight gradient (MN-major tcgen05, K = pixels); partial >= tc_out_wgrad_partial_floats()
bool tc_out_wgrad_supported(int Cin, int Cout);
size_t tc_out_wgrad_partial_floats(int Cin, int Cout);
int tc_out_wgrad(const void* dl8_bf16, const void* act_bf16, float* dW, float* partial, int B, int H, int W, int Cin,
                 int Cout, int* error_flag, cudaStream_t st);

}  // namespace kc
