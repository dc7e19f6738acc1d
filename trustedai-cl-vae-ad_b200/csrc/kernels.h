// kernels.h - host-callable launchers of the kcvae CUDA kernels.
// All tensors fp32, images / activations NHWC.  Every launcher enqueues on `st` and
// returns immediately.
#pragma once
#include "common.cuh"

namespace kc {

// ---------------------------------------------------------------- convolutions (conv.cu)
enum ConvMode { CONV_S2 = 0, CONV_S1 = 1, CONVT_S2 = 2 };
enum Epilogue { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_BIAS_SIGMOID = 2, EPI_MASK = 3 };

struct ConvArgs {
  const float* in;    // [B,Hi,Wi,Ci]
  const float* w;     // 9 taps; element (tap,ci,co) at tap*Ci*Co + ci*w_sci + co*w_sco
  const float* bias;  // [Co] or nullptr
  const float* mask;  // EPI_MASK: out *= (mask > 0), same shape as out
  float* out;         // [B,Ho,Wo,Co]
  int B, Hi, Wi, Ci, Ho, Wo, Co;
  int w_sci, w_sco;
  int flip;           // CONV_S1: 1 = in[oy+1-kh] (Conv2DTranspose s1), 0 = in[oy-1+kh]
  int pad_t, pad_l;   // CONV_S2: in[2oy+kh-pad_t]; CONVT_S2: out[2i+kh-pad_t] += in[i]
};
void conv_forward(int mode, int epi, const ConvArgs& a, cudaStream_t st);

// dW[tap,a,b] = sum_{n,i,j} P[n,i,j,a] * Q[n, sy*i+dy*kh+oy, sx*j+dx*kw+ox, b]
// written at out[tap*Ca*Cb + a*o_sa + b*o_sb]
struct WgradArgs {
  const float* P;  // [B,Hp,Wp,Ca]
  const float* Q;  // [B,Hq,Wq,Cb]
  float* out;      // 9*Ca*Cb
  float* partial;  // workspace, >= wgrad_partial_floats()
  float* pcolsum;  // optional [Ca]: sum over all pixels of P (bias gradient when P is a gradient), same pass
  int B, Hp, Wp, Ca, Hq, Wq, Cb;
  int s, d, oy, ox;  // qy = s*i + d*kh + oy ; qx = s*j + d*kw + ox
  int o_sa, o_sb;
};
size_t wgrad_partial_floats(int B, int Hp, int Ca, int Cb);
void conv_wgrad(const WgradArgs& a, cudaStream_t st);

// out[c] = sum_r in[r*C + c]  (bias gradients); partial >= colsum_partial_floats()
size_t colsum_partial_floats(int64_t rows, int C);
void colsum(const float* in, int64_t rows, int C, float* out, float* partial, cudaStream_t st);

// ------------------------------------------------------------------- dense (dense.cu)
struct GemmArgs {
  const float* A; int64_t a_sm, a_sk;   // A[m,k] at m*a_sm + k*a_sk
  const float* Bm; int64_t b_sk, b_sn;  // B[k,n] at k*b_sk + n*b_sn
  float* C;                             // [M,N] row-major
  const float* bias;                    // [N] or nullptr
  const float* mask;                    // [M,N] or nullptr: C *= (mask > 0)
  int relu;
  int M, N, K;
  float* partial;                       // split-K workspace >= gemm_partial_floats()
};
size_t gemm_partial_floats(int M, int N, int K);
void gemm(const GemmArgs& a, cudaStream_t st);

// Wide Dense (K small, N huge, everything row-major and 16-byte aligned): streaming kernels for
// the decoder Dense.  dense_wide_ok() says whether the shapes / pointers qualify.
bool dense_wide_ok(const void* A, const void* W, const void* C, const void* bias, int M, int N, int K);
void dense_wide_forward(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int relu,
                        cudaStream_t st, void* Cp = nullptr, int Cc = 0, int split = 0);
size_t dense_wide_partial_floats(int M, int N, int K);
// dW[K,N] = A^T G, db[N] = colsum(G) (db may be nullptr), dA[M,K] = G W^T (dA may be nullptr)
void dense_wide_backward(const float* A, const float* G, const float* W, float* dW, float* db, float* dA, float* partial,
                         int M, int N, int K, cudaStream_t st, cudaStream_t st_w);

// --------------------------------------------------------------------- loss (loss.cu)
// slots of the fp64 `sums` vector shared by the loss kernels (all-reduced under DP)
enum SumSlot {
  S_SE = 0,      // sum (x - xhat)^2
  S_XHX = 1,     // sum xhat * x
  S_XH = 2,      // sum xhat
  S_EX = 3,      // sum exp(x)
  S_ABSZ = 4,    // sum |z|
  S_KL = 5,      // sum |1 + lv^2 - m^2 - exp(lv^2)|
  S_Z1 = 6,      // Global: S_Z1..S_Z1+3 = sum z, z^2, z^3, z^4 ; Single: 4 per column from here
};
constexpr int kMaxLatent = 1024;
constexpr int kSumsLen = S_Z1 + 4 * kMaxLatent;

// one pass over (x, xhat): image sums, min/max, optional d(loss)/d(logit), optional
// per-position batch moments for x_std_loss.
struct ImageStatsArgs {
  const float* x; const float* xhat;
  int B; int64_t P;           // P = H*W*C positions
  double* sums;               // += into S_SE..S_EX  (device, zeroed by caller)
  float* minmax;              // [2]: min xhat, max xhat
  double* std_acc;            // [1]: sum_p (std_b x - std_b xhat)^2 (local batch) or nullptr
  double* pos_sums;           // [4*P] per-position sum x, x^2, xhat, xhat^2 (DP) or nullptr
  float* dlogit;              // [B,P] or nullptr: grad_scale*(xhat-x)*xhat*(1-xhat)
  uint16_t* dl8;              // optional bf16 copy of dlogit as NHWC padded to 8 channels (tensor-core dgrad)
  int C;                      // image channels (for dl8 indexing)
  float* dbias;               // optional [C]: sum over batch and pixels of d(loss)/d(logit) (output-layer bias gradient)
  float grad_scale;
  int want_ce;                // accumulate S_XHX, S_XH, S_EX
  double* partial;            // workspace >= image_stats_partial_doubles()
};
size_t image_stats_partial_doubles();
void image_stats(const ImageStatsArgs& a, cudaStream_t st);
// x_std_loss numerator from all-reduced per-position sums (DP, FULL tier)
void image_std_from_pos_sums(const double* pos_sums, int64_t P, int B_global, double* std_acc,
                             double* partial, cudaStream_t st);

// z = mean + 0.5*logvar + eps from the encoder head output [B,2L]; eps may be nullptr
// (zeros) or, when philox_seed != 0 and eps == nullptr and training, on-device N(0,1).
void reparameterize(const float* head, int B, int L, const float* eps, int gen_eps,
                    uint64_t seed, uint64_t counter, float* z, float* mean, float* logvar,
                    float* eps_out, cudaStream_t st);
void reparam_from_parts(const float* mean, const float* logvar, int B, int L, const float* eps,
                        int gen_eps, uint64_t seed, uint64_t counter, float* z, cudaStream_t st);
// latent power sums -> sums[S_ABSZ..]; model_type 0 global, 1 single
void latent_sums(const float* z, const float* mean, const float* logvar, int B, int L,
                 int model_type, double* sums, cudaStream_t st);

struct LossWeights { float kurtosis_target, w_mse, w_kurtosis, w_skew, w_z_l1_reg; };
// metrics (float[16], reference dict order) from the (global) sums; B_global = total batch.
// std_acc/minmax may be nullptr (LOSS_ONLY tier -> those entries are NaN).
void finalize_metrics(const double* sums, const float* minmax, const double* std_acc,
                      int B_global, int L, int64_t P, int model_type, LossWeights lw,
                      int have_ce, float* metrics, cudaStream_t st);
// dhead[B,2L] = [g, 0.5 g], g = g_z (decoder path, may be nullptr) + moment/L1 terms
void latent_backward(const float* z, const float* g_z, const double* sums, int B_local,
                     int B_global, int L, int model_type, LossWeights lw, float* dhead,
                     cudaStream_t st);

// ------------------------------------------------------------------ scoring (loss.cu)
// err[b,h,w] = sum_c (x - xhat)^2 ; score[b] = sum_hw err.  err may be nullptr.
size_t score_partial_floats(int B, int64_t HW);
// err_minmax [B,2] = per-frame (min, max) of err, may be nullptr.
void score(const float* x, const float* xhat, int B, int64_t HW, int C, float* err,
           float* score_out, float* err_minmax, float* partial, cudaStream_t st);
void normalize_scores(const float* err, const float* score_in, int B, int64_t HW, float meu,
                      float sigma, float emin, float emax, float thr, float* norm, float* z,
                      uint8_t* flags, cudaStream_t st);

// -------------------------------------------------------------------- misc (loss.cu)
void add_noise(const float* x, const float* noise, int64_t n, float stddev, uint64_t seed,
               uint64_t counter, float* out, cudaStream_t st);
void adam_update(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1,
                 float b2, float eps, cudaStream_t st);
void glorot_fill(float* p, int64_t n, float limit, uint64_t seed, uint32_t stream_id,
                 cudaStream_t st);

// ---- frontend.cu: uint8 front end, streaming score, scorer outputs ----
struct ResizePlan;
ResizePlan* resize_plan_create(int in_h, int in_w, int out_h, int out_w, int C);
void resize_plan_free(ResizePlan* p);
bool resize_plan_matches(const ResizePlan* p, int in_h, int in_w, int out_h, int out_w, int C);
int preprocess_u8(const uint8_t* in, int B, ResizePlan* plan, int64_t same_size_elems, float* out, cudaStream_t st);
void render_outputs(const float* norm_err, const float* rec, int64_t npix, int C, uint8_t* err_u8, uint8_t* heatmap,
                    uint8_t* overlay, uint8_t* rec_u8, cudaStream_t st);

}  // namespace kc
