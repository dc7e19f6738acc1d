/*
 * kcvae.h - C ABI of libkcvae.so: the B200 (sm_100a) engine behind the Python-facing
 * KurtosisGlobalCVAE / KurtosisSingleCVAE classes of gtemplin/TrustedAI-CL-VAE-AD.
 *
 * The reference has no FFI of its own (it is pure Python over TensorFlow); the seam this
 * library sits behind is the object returned by src/load_model.py:70-72
 * (load_model_from_config).  Each entry point below names the reference method it
 * replaces; the Python mirror in trustedai-cl-vae-ad_b200/ binds them with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.  All tensors are contiguous
 *    float32, images NHWC ([B,H,W,C]) exactly like the tf.data pipeline yields them
 *    (src/data_loader.py:86-90).
 *  - pointers named d_* are DEVICE pointers (16-byte aligned), h_* are HOST pointers
 *    (pinned memory makes the copies asynchronous).  Caller owns every I/O buffer; the
 *    library owns weights, gradients, Adam state, activations and the RNG counter.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls
 *    are asynchronous on that stream unless stated otherwise.
 *  - every function returns 0 on success, a negative kcvae_status otherwise;
 *    kcvae_last_error(h) returns a human-readable message for the last failure.
 *  - a handle is bound to one GPU and is not thread-safe (the reference is
 *    single-threaded Python: camera_streamer_qt.py:1283-1285).
 */
#ifndef KCVAE_H_
#define KCVAE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KCVAE_ABI_VERSION 1
#define KCVAE_MAX_LAYERS 8
#define KCVAE_NUM_METRICS 16 /* metrics vectors are float[16]; 12 (Global) / 10 (Single) used */

typedef enum {
  KCVAE_OK = 0,
  KCVAE_ERR_INVALID = -1,   /* bad argument / config / shape */
  KCVAE_ERR_CUDA = -2,      /* CUDA runtime error */
  KCVAE_ERR_NCCL = -3,      /* NCCL error */
  KCVAE_ERR_COLLAPSE = -4,  /* decoder width/height collapse (src/abstract_cvae.py:65-68) */
  KCVAE_ERR_UNSUPPORTED = -5
} kcvae_status;

typedef enum { KCVAE_GLOBAL = 0, KCVAE_SINGLE = 1 } kcvae_model_type; /* src/load_model.py:9-31 */

/* metric tiers (SURVEY 7 hard part 5): LOSS_ONLY skips the reported-only terms that need
 * extra passes / collectives (x_std_loss, cross_entropy); FULL reproduces the whole dict. */
typedef enum { KCVAE_METRICS_FULL = 0, KCVAE_METRICS_LOSS_ONLY = 1 } kcvae_metric_tier;

/* arithmetic of the decoder convolutions: FP32 = CUDA-core fp32 everywhere;
 * BF16_TC = bf16 operands, fp32 accumulate on tcgen05 tensor cores where a kernel exists */
typedef enum { KCVAE_PREC_FP32 = 0, KCVAE_PREC_BF16_TC = 1 } kcvae_precision;

/* config.yml schema (README.md:52-85; src/abstract_cvae.py:14-16,30,42,62;
 * src/kurtosis_global_cvae.py:15-21) flattened to plain C */
typedef struct {
  int32_t image_h, image_w, image_c;          /* data.image_size                        */
  int32_t n_layers;
  int32_t layers[KCVAE_MAX_LAYERS];           /* model.layers                           */
  int32_t encoder_dense_filters;              /* 0 = key absent (src/abstract_cvae.py:43) */
  int32_t decoder_dense_filters;
  int32_t latent_dimensions;
  int32_t model_type;                         /* kcvae_model_type                       */
  float kurtosis_target;                      /* loss.kurtosis                          */
  float w_mse, w_kurtosis, w_skew, w_kl_divergence, w_z_l1_reg, w_x_std;
  float beta;                                 /* training.beta: image-noise stddev      */
  float learning_rate;                        /* training.learning_rate                 */
  int32_t max_batch;                          /* workspace is sized for this; grows on demand */
  int32_t precision;                          /* kcvae_precision                        */
} kcvae_config;

typedef struct kcvae_model* kcvae_handle;

/* ---- lifetime ---------------------------------------------------------------------- */
int kcvae_abi_version(void);
/* AbstractCVAE.__init__ + _build_encoder/_build_decoder (src/abstract_cvae.py:9-92).
 * Weights are zero until kcvae_set_weights / kcvae_init_glorot. */
int kcvae_create(const kcvae_config* cfg, int device, kcvae_handle* out);
int kcvae_destroy(kcvae_handle h);
const char* kcvae_last_error(kcvae_handle h); /* h may be NULL: error of the last create */

/* ---- variables: Keras trainable_weights order and layouts ---------- */
int kcvae_num_variables(kcvae_handle h);
int64_t kcvae_param_count(kcvae_handle h);
/* rank, dims[4] and offset (in floats) of variable idx inside the flat parameter vector */
int kcvae_variable_info(kcvae_handle h, int idx, int32_t* rank, int64_t dims[4], int64_t* offset);
int kcvae_set_weights(kcvae_handle h, const float* h_flat, int64_t n);  /* synchronous */
int kcvae_get_weights(kcvae_handle h, float* h_flat, int64_t n);        /* synchronous */
/* dL/dw of the last train_step / loss_and_grads (tape.gradient, src/abstract_cvae.py:160) */
int kcvae_get_grads(kcvae_handle h, float* h_flat, int64_t n);          /* synchronous */
/* device pointers of the flat fp32 parameter / gradient / Adam vectors (library-owned) */
float* kcvae_weights_device(kcvae_handle h);
float* kcvae_grads_device(kcvae_handle h);
/* Keras defaults (glorot_uniform kernels, zero biases) from an on-device Philox stream */
int kcvae_init_glorot(kcvae_handle h, uint64_t seed, void* stream);

/* ---- optimizer: tf.keras.optimizers.Adam (train.py:99-101) --------------------------- */
int kcvae_adam_reset(kcvae_handle h);                                   /* m = v = 0, t = 0 */
int kcvae_set_adam_state(kcvae_handle h, const float* h_m, const float* h_v, int64_t n, int64_t t);
int kcvae_get_adam_state(kcvae_handle h, float* h_m, float* h_v, int64_t n, int64_t* t);
/* learning rate is mutable at run time (camera_streamer_qt.py:1329); beta is model.beta
 * (train.py:46-47 BetaAnnealingCallback) */
int kcvae_set_learning_rate(kcvae_handle h, float lr);
int kcvae_set_beta(kcvae_handle h, float beta);
/* opt-in README behaviour (src/abstract_cvae.py:117-118 applied to train_step, SURVEY Note A): when on, kcvae_train_step adds
 * on-device Philox N(0, beta^2) noise to the encoder input unless the caller passes its own d_img_noise */
int kcvae_set_train_image_noise(kcvae_handle h, int on);
int kcvae_set_loss_weights(kcvae_handle h, float kurtosis_target, float w_mse, float w_kurtosis,
                           float w_skew, float w_z_l1_reg);
int kcvae_seed(kcvae_handle h, uint64_t seed); /* Philox key for on-device eps / image noise */

/* ---- data parallel (new work; the reference has no collectives) ---------------------- */
/* 128-byte ncclUniqueId created on rank 0 and shipped to the other ranks by the caller */
int kcvae_comm_unique_id(void* out_id128);
int kcvae_comm_init(kcvae_handle h, const void* id128, int rank, int world_size);
int kcvae_comm_world(kcvae_handle h);
int kcvae_broadcast_weights(kcvae_handle h, int root, void* stream);

/* ---- forward --------------------------------------------------------------------------- */
/* encode(x, training) (src/abstract_cvae.py:115-122).  img_noise: NULL and training!=0 =>
 * on-device N(0, beta^2); non-NULL => added to x as is (parity mode). */
int kcvae_encode(kcvae_handle h, const float* d_x, int batch, int training,
                 const float* d_img_noise, float* d_mean, float* d_logvar, void* stream);
/* reparameterize(mean, logvar, training) (:124-129): z = mean + 0.5*logvar + eps;
 * d_eps NULL: eps = 0 if !training else on-device N(0,1) */
int kcvae_reparameterize(kcvae_handle h, const float* d_mean, const float* d_logvar, int batch,
                         int training, const float* d_eps, float* d_z, void* stream);
/* decode(z, apply_sigmoid) (:131-137) */
int kcvae_decode(kcvae_handle h, const float* d_z, int batch, int apply_sigmoid, float* d_out,
                 void* stream);
/* call_detailed(x, training) / call (:139-149); any of d_z/d_mean/d_logvar may be NULL */
int kcvae_forward(kcvae_handle h, const float* d_x, int batch, int training, const float* d_eps,
                  float* d_xhat, float* d_z, float* d_mean, float* d_logvar, void* stream);

/* ---- loss / training step ------------------------------------------------------------ */
/* compute_loss(x, training, return_inf) (src/kurtosis_global_cvae.py:32-110,
 * src/kurtosis_single_cvae.py:25-77).  d_metrics: float[KCVAE_NUM_METRICS] in the
 * reference's dict order.  d_xhat may be NULL. */
int kcvae_loss(kcvae_handle h, const float* d_x, int batch, int training, const float* d_eps,
               float* d_metrics, float* d_xhat, int tier, void* stream);
/* train_step / train_step_and_run (src/abstract_cvae.py:154-178): forward, loss, backward,
 * (gradient all-reduce when a communicator is attached), Adam.  d_img_noise is the opt-in
 * image noise the README describes (unreachable in the reference, SURVEY Note A): NULL = off. */
int kcvae_train_step(kcvae_handle h, const float* d_x, int batch, const float* d_eps,
                     const float* d_img_noise, float* d_metrics, float* d_xhat, int tier,
                     void* stream);
/* same as kcvae_train_step without the optimizer update: leaves dL/dw in the gradient
 * vector (tape.gradient, :160) - used by parity tests and gradient inspection */
int kcvae_loss_and_grads(kcvae_handle h, const float* d_x, int batch, const float* d_eps,
                         float* d_metrics, float* d_xhat, int tier, void* stream);

/* ---- anomaly scoring (do_anomaly_detection.py:57-117) -------------------------------- */
/* x_rec = call(x, False); err = sum_c (x - x_rec)^2 -> d_err [B,H,W] (may be NULL);
 * d_score[b] = sum_hw err; d_err_minmax [B,2] = per-frame (min, max) of err (may be NULL; the
 * global min/max of get_data_scale :70-71 are their min/max); d_xhat may be NULL */
int kcvae_score(kcvae_handle h, const float* d_x, int batch, float* d_err, float* d_score,
                float* d_err_minmax, float* d_xhat, void* stream);
/* evaluate_anomalies' elementwise tail (:89-91): z = (score-meu)/sigma,
 * norm = (err-min)/(max-min), flags = z > threshold */
int kcvae_normalize_scores(kcvae_handle h, const float* d_err, const float* d_score, int batch,
                           float meu, float sigma, float emin, float emax, float threshold,
                           float* d_norm, float* d_z, uint8_t* d_flags, void* stream);

/* ---- host-buffer entry points (H2D / D2H inside the call; end-to-end path) ------------ */
/* optional pipelining: start the H2D copy of the NEXT call's frames on an internal copy stream
 * while the current step computes; the next *_host call given the same pointer only waits for it */
int kcvae_prefetch_host(kcvae_handle h, const float* h_x, int batch);
/* synchronous: returns after metrics are in h_metrics */
int kcvae_train_step_host(kcvae_handle h, const float* h_x, int batch, const float* h_eps,
                          float* h_metrics, float* h_xhat, int tier, void* stream);
int kcvae_score_host(kcvae_handle h, const float* h_x, int batch, float* h_err, float* h_score,
                     void* stream);

/* ---- front end: uint8 frames -> model input --------------------------------- */
/* src/data_loader.py:10-20 (_normalize_img, _resize_img) / camera_streamer_qt.py:1296:
 * x = tf.image.resize(uint8 / 255., image_size[:2], antialias=True) for d_frames [B,in_h,in_w,C] uint8 NHWC;
 * in_h == H and in_w == W: the cast alone.  d_x [B,H,W,C] fp32. */
int kcvae_preprocess_u8(kcvae_handle h, const uint8_t* d_frames, int batch, int in_h, int in_w,
                        float* d_x, void* stream);
/* kcvae_score_host / kcvae_train_step_host fed with uint8 HOST frames (a quarter of the H2D bytes) */
/* optional pipelining, as kcvae_prefetch_host: start the H2D copy of the NEXT call's uint8 frames now */
int kcvae_prefetch_host_u8(kcvae_handle h, const uint8_t* h_frames, int batch, int in_h, int in_w);
int kcvae_score_host_u8(kcvae_handle h, const uint8_t* h_frames, int batch, int in_h, int in_w,
                        float* h_err, float* h_score, void* stream);
int kcvae_train_step_host_u8(kcvae_handle h, const uint8_t* h_frames, int batch, int in_h, int in_w,
                             const float* h_eps, float* h_metrics, float* h_xhat, int tier, void* stream);

/* ---- streaming anomaly score of the camera tool ----------------------------- */
/* camera_streamer_qt.py:1364-1400: per-pixel EMA mean / second moment of the error map, z scores, count of
 * pixels whose z-of-z exceeds 3, EMA of that count -> anomaly_score; EMA-normalised uint8 error image. */
typedef struct kcvae_stream* kcvae_stream_handle;
int kcvae_stream_create(int H, int W, int device, kcvae_stream_handle* out);
int kcvae_stream_destroy(kcvae_stream_handle s);
int kcvae_stream_reset(kcvae_stream_handle s);
const char* kcvae_stream_last_error(kcvae_stream_handle s);
/* one frame.  d_err [H,W] = sum_c (x - x_rec)^2 (kcvae_score's error map); ma = stream_error_ma (:213, 0.99);
 * d_err_u8 [H,W] (may be NULL) = round(255 (err - ema_min) / (ema_max - ema_min)), saturated;
 * h_out[8] = anomaly_count, anomaly_score, frame min, frame max, ema_min, ema_max, z mean, z std.
 * Synchronous (the reference reads the count on the host every frame). */
int kcvae_stream_update(kcvae_stream_handle s, const float* d_err, double ma, uint8_t* d_err_u8,
                        float* h_out, void* stream);

/* ---- scorer outputs ------------------------------------------------------------ */
/* do_anomaly_detection.py:166-170, output_reconstructions.py:68-83, camera_streamer_qt.py:1417-1418:
 * err_u8 = round(255 norm_err); heatmap = cv2.applyColorMap(err_u8, COLORMAP_JET) (OpenCV channel order);
 * rec_u8 = round(255 rec); overlay = cv2.addWeighted(heatmap, .5, rec_u8, .5, 0).  d_norm_err [B,H,W],
 * d_rec [B,H,W,C]; every output may be NULL. */
int kcvae_render_outputs(const float* d_norm_err, const float* d_rec, int batch, int H, int W, int C,
                         uint8_t* d_err_u8, uint8_t* d_heatmap, uint8_t* d_overlay, uint8_t* d_rec_u8,
                         void* stream);

/* ---- introspection --------------------------------------------------------------------- */
/* number of kernels this library launched on behalf of handle h since creation */
int64_t kcvae_launch_count(kcvae_handle h);
/* 1 = tcgen05 tensor-core kernels active for this handle, 0 = fp32 CUDA-core path only;
 * negative = a tensor-core pipeline reported an error.  Synchronises the device. */
int kcvae_tc_status(kcvae_handle h);
/* per-launch device timing with CUDA events on the launching stream (bench.py roofline):
 * enable, run steps, then report "<layer tag>/<kernel> <calls> <total ms>" lines */
int kcvae_profile_enable(int on);
int64_t kcvae_profile_report(char* buf, int64_t capacity);
/* copies an internal activation for layer-level parity tests: which = 0..(encoder convs),
 * then decoder stages; returns element count or negative status */
int64_t kcvae_debug_activation(kcvae_handle h, int which, float* h_out, int64_t capacity);

/* layer-level hooks of the general tensor-core convolution engine (csrc/tc_gen.cu), fp32 NHWC device tensors in and
 * out: one forward-type product (kind 0: Conv2D k3 s2 / Conv2DTranspose s2 data gradient over a space-to-depth input,
 * 1: Conv2DTranspose k3 s2 / Conv2D s2 data gradient, 2: 3x3 stride-1 - src/abstract_cvae.py:30-33, 81-89) or one weight +
 * bias gradient.  w_mode 0: weight element (tap,k,n) at (tap*Ck + k)*Cn + n (HWIO), 1: (tap*Cn + n)*Ck + k ([kh,kw,out,in]).
 * pre: 0 none, 1 bias+ReLU, 2 bias+sigmoid, 3 bias; out_mode 0 fp32 NHWC, 1 / 2 bf16 hi+lo planes (plain / space-to-depth)
 * unpacked again; mask_mode 0 fp32 mask, 1 / 2 mask read from bf16 planes.  Used by the parity tests only. */
int kcvae_gen_conv_test(int kind, int w_mode, int flip, int split, int pre, int in_x3, int out_mode, int mask_mode,
                        const float* d_in, const float* d_w, const float* d_bias, const float* d_mask, float* d_out,
                        int B, int Hi, int Wi, int Ck, int Cn, void* stream);
int kcvae_gen_wgrad_test(int kind, int w_mode, int flip, int s_x3, const float* d_s, const float* d_u, float* d_dW, float* d_db,
                         int B, int Hs, int Ws, int Cs, int Cu, void* stream);
/* Dense-layer products of the same engine (src/abstract_cvae.py:41-45, 75-77) on fp32 device matrices, A = [R][N] row-major:
 * mode 0 forward out[c][n] = act(bias[n] + sum_r A[r][n] Bm[c][r]); mode 1 weight + bias gradient out[c][n] = sum_r A[r][n] Bm[r][c],
 * out[C][n] = sum_r A[r][n]; mode 2 data gradient out[u][s] = sum_n A[u][n] Bm[s][n].  Used by the parity tests only. */
int kcvae_gen_dense_test(int mode, int split, int relu, const float* d_a, const float* d_b, const float* d_bias, float* d_out,
                         int R, int N, int C, void* stream);
/* the host-side plan (MMA list, K slabs, weight gather table / accumulator roles, scatter table) of one product as a flat
 * int32 array; needs no GPU.  The CPU test-suite interprets it with numpy against the oracle.  Returns the length needed. */
int64_t kcvae_gen_plan_dump(int which, const int32_t* spec, int nspec, int32_t* out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* KCVAE_H_ */
