// frontend.cu - the stages either side of the KurtosisCVAE hot path, all small
// HBM-bound byte / float kernels:
//   * uint8 camera / dataset frames -> fp32 NHWC in [0,1], optionally through tf.image.resize(antialias=True)
//     (src/data_loader.py:10-20, camera_streamer_qt.py:1296)
//   * the streaming anomaly score of the camera tool: per-pixel EMA moments of the error map, z-of-z threshold
//     count, EMA-normalised error image (camera_streamer_qt.py:1364-1400)
//   * scorer outputs: uint8 error image, JET heat map, 0.5/0.5 overlay, uint8 reconstruction
//     (do_anomaly_detection.py:166-170, output_reconstructions.py:68-83, camera_streamer_qt.py:1417-1418)
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kcvae.h"
#include "kernels.h"

namespace kc {

// ================================================================= uint8 -> float (+ antialiased resize)
// TensorFlow's ScaleAndTranslate with the triangle kernel (what tf.image.resize(..., antialias=True) runs for the
// default bilinear method): span and weight tables per output index, computed in fp32 exactly as the op does.
// TF is not vendored with the reference, so this restates its published algorithm.
static void compute_spans(int in_size, int out_size, std::vector<int>& start, std::vector<int>& count,
                          std::vector<float>& weights, int& span) {
  const float scale = (float)out_size / (float)in_size;
  const float inv_scale = 1.0f / scale;
  const float kernel_scale = inv_scale > 1.0f ? inv_scale : 1.0f;       // antialias: widen the kernel when shrinking
  const float radius = 1.0f;                                            // triangle kernel
  span = 2 * (int)std::ceil(radius * kernel_scale) + 1;
  if (span > in_size) span = in_size;
  const float one_over_kernel_scale = 1.0f / kernel_scale;
  start.assign(out_size, 0);
  count.assign(out_size, 0);
  weights.assign((size_t)out_size * span, 0.0f);
  for (int x = 0; x < out_size; ++x) {
    const float col_f = (float)x + 0.5f;
    const float sample_f = col_f * inv_scale;
    if (sample_f < 0.0f || sample_f > (float)in_size) continue;
    int64_t s = (int64_t)std::ceil(sample_f - radius * kernel_scale - 0.5f);
    int64_t e = (int64_t)std::floor(sample_f + radius * kernel_scale - 0.5f);
    s = s < 0 ? 0 : (s > in_size - 1 ? in_size - 1 : s);
    e = (e < 0 ? 0 : (e > in_size - 1 ? in_size - 1 : e)) + 1;
    int n = (int)(e - s);
    if (n > span) n = span;
    float total = 0.0f;
    float* w = &weights[(size_t)x * span];
    for (int j = 0; j < n; ++j) {
      const float kernel_pos = (float)(s + j) + 0.5f - sample_f;
      const float a = std::fabs(kernel_pos * one_over_kernel_scale);
      w[j] = a < 1.0f ? 1.0f - a : 0.0f;
      total += w[j];
    }
    if (std::fabs(total) >= 1000.0f * 1.17549435e-38f) {
      const float inv = 1.0f / total;
      for (int j = 0; j < n; ++j) w[j] *= inv;
    }
    start[x] = (int)s;
    count[x] = n;
  }
}

struct ResizePlan {
  int in_h = 0, in_w = 0, out_h = 0, out_w = 0, C = 0;
  int span_r = 0, span_c = 0;
  int *row_start = nullptr, *row_count = nullptr, *col_start = nullptr, *col_count = nullptr;
  float *row_w = nullptr, *col_w = nullptr;
  float* tmp = nullptr;          // [B, out_h, in_w, C] after the row pass
  size_t tmp_floats = 0;
};

void resize_plan_free(ResizePlan* p) {
  if (!p) return;
  cudaFree(p->row_start); cudaFree(p->row_count); cudaFree(p->col_start); cudaFree(p->col_count);
  cudaFree(p->row_w); cudaFree(p->col_w); cudaFree(p->tmp);
  delete p;
}

// nullptr on allocation failure
ResizePlan* resize_plan_create(int in_h, int in_w, int out_h, int out_w, int C) {
  ResizePlan* p = new ResizePlan();
  p->in_h = in_h; p->in_w = in_w; p->out_h = out_h; p->out_w = out_w; p->C = C;
  std::vector<int> rs, rc, cs, cc;
  std::vector<float> rw, cw;
  compute_spans(in_h, out_h, rs, rc, rw, p->span_r);
  compute_spans(in_w, out_w, cs, cc, cw, p->span_c);
  auto up = [](auto** d, const auto& v) {
    if (cudaMalloc(reinterpret_cast<void**>(d), v.size() * sizeof(v[0])) != cudaSuccess) return false;
    return cudaMemcpy(*d, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice) == cudaSuccess;
  };
  if (!up(&p->row_start, rs) || !up(&p->row_count, rc) || !up(&p->col_start, cs) || !up(&p->col_count, cc) ||
      !up(&p->row_w, rw) || !up(&p->col_w, cw)) {
    resize_plan_free(p);
    return nullptr;
  }
  return p;
}

bool resize_plan_matches(const ResizePlan* p, int in_h, int in_w, int out_h, int out_w, int C) {
  return p && p->in_h == in_h && p->in_w == in_w && p->out_h == out_h && p->out_w == out_w && p->C == C;
}

// x = float(u8) / 255 through a 256-entry table (bit-identical to the division, src/data_loader.py:12)
__global__ void __launch_bounds__(256) u8_to_unit_kernel(const uint8_t* in, int64_t n, float* out) {
  __shared__ float lut[256];
  lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
  __syncthreads();
  const int64_t n4 = n >> 2;
  const uint32_t* in4 = reinterpret_cast<const uint32_t*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = __ldg(in4 + i);
    out4[i] = make_float4(lut[v & 255u], lut[(v >> 8) & 255u], lut[(v >> 16) & 255u], lut[v >> 24]);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = lut[in[i]];
}

// row pass: tmp[b, y, xc] = sum_i w[y][i] * unit(in[b, start[y] + i, xc]),  xc over in_w * C (contiguous)
// products and sums are separate roundings (no fused multiply-add), in span order
__global__ void __launch_bounds__(256) resize_rows_kernel(const uint8_t* in, int B, int in_h, int64_t rowlen, int out_h,
                                                          const int* start, const int* count, const float* w, int span,
                                                          float* tmp) {
  __shared__ float lut[256];
  lut[threadIdx.x] = (float)threadIdx.x / 255.0f;
  __syncthreads();
  const int64_t total = (int64_t)B * out_h * rowlen;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t xc = i % rowlen;
    const int y = (int)((i / rowlen) % out_h);
    const int64_t b = i / (rowlen * out_h);
    const int s = start[y], n = count[y];
    const float* wy = w + (int64_t)y * span;
    const uint8_t* src = in + (b * in_h + s) * rowlen + xc;
    float acc = 0.0f;
    for (int j = 0; j < n; ++j) acc = __fadd_rn(acc, __fmul_rn(wy[j], lut[src[(int64_t)j * rowlen]]));
    tmp[i] = acc;
  }
}

// column pass: out[b, y, x, c] = sum_i w[x][i] * tmp[b, y, start[x] + i, c]
__global__ void __launch_bounds__(256) resize_cols_kernel(const float* tmp, int64_t rows, int in_w, int C, int out_w,
                                                          const int* start, const int* count, const float* w, int span,
                                                          float* out) {
  const int64_t total = rows * out_w * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int x = (int)((i / C) % out_w);
    const int64_t r = i / ((int64_t)C * out_w);
    const int s = start[x], n = count[x];
    const float* wx = w + (int64_t)x * span;
    const float* src = tmp + (r * in_w + s) * C + c;
    float acc = 0.0f;
    for (int j = 0; j < n; ++j) acc = __fadd_rn(acc, __fmul_rn(wx[j], src[(int64_t)j * C]));
    out[i] = acc;
  }
}

// frames [B, in_h, in_w, C] uint8 -> out [B, out_h, out_w, C] fp32.  plan == nullptr: same size, cast only.
// returns 0, or 1 when the scratch could not be allocated
int preprocess_u8(const uint8_t* in, int B, ResizePlan* plan, int64_t same_size_elems, float* out, cudaStream_t st) {
  if (!plan) {
    ProfScope prof_("u8_to_unit", st);
    ++g_launches;
    KC_LAUNCH(u8_to_unit_kernel, grid_for(same_size_elems / 4 + 1, 256), 256, 0, st, in, same_size_elems, out);
    return 0;
  }
  const int64_t rowlen = (int64_t)plan->in_w * plan->C;
  const size_t need = (size_t)B * plan->out_h * rowlen;
  if (need > plan->tmp_floats) {
    cudaFree(plan->tmp);
    plan->tmp = nullptr; plan->tmp_floats = 0;
    if (cudaMalloc(reinterpret_cast<void**>(&plan->tmp), need * sizeof(float)) != cudaSuccess) return 1;
    plan->tmp_floats = need;
  }
  ProfScope prof_("resize_antialias", st);
  g_launches += 2;
  KC_LAUNCH(resize_rows_kernel, grid_for((int64_t)need, 256), 256, 0, st, in, B, plan->in_h, rowlen, plan->out_h,
            plan->row_start, plan->row_count, plan->row_w, plan->span_r, plan->tmp);
  KC_LAUNCH(resize_cols_kernel, grid_for((int64_t)B * plan->out_h * plan->out_w * plan->C, 256), 256, 0, st, plan->tmp,
            (int64_t)B * plan->out_h, plan->in_w, plan->C, plan->out_w, plan->col_start, plan->col_count, plan->col_w,
            plan->span_c, out);
  return 0;
}

// ================================================================= streaming anomaly score
constexpr int kStreamBlocks = 148;     // one block per SM for the per-pixel pass
constexpr int kStreamThreads = 256;

// pass 1 (camera_streamer_qt.py:1381-1389): per pixel  S1 = w0*S1 + w1*e,  S2 = w0*S2 + w1*e^2  (the first frame seeds
// S1 = e, S2 = e^2 and then takes the same update), var = |S2 - S1^2|, z = (e - S1) / sqrt(var + 1e-10);
// per block: sum of z (fp64), min / max of e
__global__ void __launch_bounds__(kStreamThreads) stream_update_kernel(const float* err, int n, float w0, float w1, int first,
                                                                      float* s1, float* s2, float* z, double* psum,
                                                                      float* pmin, float* pmax) {
  __shared__ double dscratch[32];
  __shared__ float fs[64];
  double t = 0.0;
  float mn = 3.4e38f, mx = -3.4e38f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float e = err[i];
    const float e2 = __fmul_rn(e, e);
    float a = first ? e : s1[i];
    float b = first ? e2 : s2[i];
    a = __fadd_rn(__fmul_rn(w0, a), __fmul_rn(w1, e));
    b = __fadd_rn(__fmul_rn(w0, b), __fmul_rn(w1, e2));
    s1[i] = a; s2[i] = b;
    const float var = fabsf(__fadd_rn(b, -__fmul_rn(a, a)));
    const float zi = __fadd_rn(e, -a) / sqrtf(__fadd_rn(var, 1e-10f));
    z[i] = zi;
    t += (double)zi;
    mn = fminf(mn, e); mx = fmaxf(mx, e);
  }
  const double r = block_sum(t, dscratch);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  mn = warp_min(mn); mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) { fs[wid] = mn; fs[32 + wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w) { mn = fminf(mn, fs[w]); mx = fmaxf(mx, fs[32 + w]); }
    psum[blockIdx.x] = r; pmin[blockIdx.x] = mn; pmax[blockIdx.x] = mx;
  }
}

// pass 2, one block: z mean / population std (two-pass, fp64), count of (z - mean)/std > 3 (:1391-1395), EMA of the
// frame min / max (:1372-1373) and the normalised uint8 error image (:1375-1376).
// state: [0] ema_min, [1] ema_max (fp32, device resident).  out: [0] count, [1] frame min, [2] frame max,
// [3] ema_min, [4] ema_max, [5] z mean, [6] z std
__global__ void __launch_bounds__(1024) stream_finish_kernel(const float* err, const float* z, int n, int nblocks,
                                                             const double* psum, const float* pmin, const float* pmax,
                                                             float w0, float w1, float* state, uint8_t* err_u8,
                                                             float* out) {
  __shared__ double dscratch[32];
  __shared__ double sh_mean, sh_std;
  __shared__ float sh_min, sh_max;
  __shared__ int iscratch[32];
  if (threadIdx.x == 0) {
    double s = 0.0;
    float mn = 3.4e38f, mx = -3.4e38f;
    for (int b = 0; b < nblocks; ++b) { s += psum[b]; mn = fminf(mn, pmin[b]); mx = fmaxf(mx, pmax[b]); }
    sh_mean = s / (double)n;
    const float emin = __fadd_rn(__fmul_rn(w0, state[0]), __fmul_rn(w1, mn));
    const float emax = __fadd_rn(__fmul_rn(w0, state[1]), __fmul_rn(w1, mx));
    state[0] = emin; state[1] = emax;
    sh_min = emin; sh_max = emax;
    out[1] = mn; out[2] = mx; out[3] = emin; out[4] = emax;
  }
  __syncthreads();
  const double mean = sh_mean;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const double d = (double)z[i] - mean; q += d * d; }
  const double qs = block_sum(q, dscratch);
  if (threadIdx.x == 0) sh_std = sqrt(qs / (double)n);
  __syncthreads();
  const float fmean = (float)mean, fstd = (float)sh_std;
  const float emin = sh_min, range = __fadd_rn(sh_max, -sh_min);
  int cnt = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float zz = __fadd_rn(z[i], -fmean) / fstd;
    cnt += zz > 3.0f ? 1 : 0;
    if (err_u8) {
      const float v = rintf(255.0f * (__fadd_rn(err[i], -emin) / range));     // np.round: half to even
      err_u8[i] = (uint8_t)(v >= 0.0f ? (v <= 255.0f ? v : 255.0f) : 0.0f);    // NaN and negatives -> 0, saturating
    }
  }
  cnt = warp_sum(cnt);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) iscratch[wid] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += iscratch[w];
    out[0] = (float)total; out[5] = fmean; out[6] = fstd;
  }
}

// ================================================================= scorer outputs
__constant__ uint8_t c_jet_lut[256][3] = {
#include "jet_lut.inc"
};

__device__ __forceinline__ uint8_t unit_to_u8(float v) {     // np.round(255. * v).astype(np.uint8), saturating
  const float r = rintf(255.0f * v);
  return (uint8_t)(r >= 0.0f ? (r <= 255.0f ? r : 255.0f) : 0.0f);
}

// norm_err [B,H,W] in [0,1], rec [B,H,W,3] in [0,1] -> err_u8 [B,H,W], heatmap [B,H,W,3] (cv2 channel order),
// overlay = cv2.addWeighted(heatmap, .5, rec_u8, .5, 0), rec_u8; any output may be nullptr
__global__ void __launch_bounds__(256) render_outputs_kernel(const float* norm_err, const float* rec, int64_t npix, int C,
                                                             uint8_t* err_u8, uint8_t* heatmap, uint8_t* overlay,
                                                             uint8_t* rec_u8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t e = norm_err ? unit_to_u8(norm_err[i]) : (uint8_t)0;
    if (err_u8) err_u8[i] = e;
    for (int c = 0; c < C; ++c) {
      const uint8_t h = c < 3 ? c_jet_lut[e][c] : (uint8_t)0;
      const uint8_t r = rec ? unit_to_u8(rec[i * C + c]) : (uint8_t)0;
      if (heatmap) heatmap[i * C + c] = h;
      if (rec_u8) rec_u8[i * C + c] = r;
      if (overlay) overlay[i * C + c] = (uint8_t)rintf(0.5f * (float)h + 0.5f * (float)r);   // cvRound: half to even
    }
  }
}

void render_outputs(const float* norm_err, const float* rec, int64_t npix, int C, uint8_t* err_u8, uint8_t* heatmap,
                    uint8_t* overlay, uint8_t* rec_u8, cudaStream_t st) {
  ProfScope prof_("render_outputs", st);
  ++g_launches;
  KC_LAUNCH(render_outputs_kernel, grid_for(npix, 256), 256, 0, st, norm_err, rec, npix, C, err_u8, heatmap, overlay, rec_u8);
}

}  // namespace kc

// ================================================================= C ABI: streaming score + scorer outputs
using namespace kc;

struct kcvae_stream {
  int H = 0, W = 0, device = 0;
  float *s1 = nullptr, *s2 = nullptr, *z = nullptr, *pmin = nullptr, *pmax = nullptr, *state = nullptr, *out_dev = nullptr;
  double* psum = nullptr;
  bool primed = false;
  // scalar EMA state, same mixed precision as the reference's Python / TF arithmetic (camera_streamer_qt.py:1397-1400):
  // the running mean is a Python float (double), the running second moment a float32 tensor
  double anomaly_sum = 0.0;
  float anomaly_sum_2 = 0.0f;
  std::string err;
};

static std::string g_stream_create_error;

extern "C" {

const char* kcvae_stream_last_error(kcvae_stream_handle s) { return s ? s->err.c_str() : g_stream_create_error.c_str(); }

int kcvae_stream_create(int H, int W, int device, kcvae_stream_handle* out) {
  if (!out || H <= 0 || W <= 0) { g_stream_create_error = "stream_create: invalid arguments"; return KCVAE_ERR_INVALID; }
  kcvae_stream* s = new kcvae_stream();
  s->H = H; s->W = W; s->device = device;
  const size_t n = (size_t)H * W;
  bool ok = cudaSetDevice(device) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&s->s1), n * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&s->s2), n * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&s->z), n * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&s->psum), kStreamBlocks * sizeof(double)) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&s->pmin), kStreamBlocks * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&s->pmax), kStreamBlocks * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&s->state), 2 * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(reinterpret_cast<void**>(&s->out_dev), 8 * sizeof(float)) == cudaSuccess;
  if (!ok) { g_stream_create_error = "stream_create: device allocation failed"; kcvae_stream_destroy(s); return KCVAE_ERR_CUDA; }
  *out = s;
  return kcvae_stream_reset(s);
}

int kcvae_stream_destroy(kcvae_stream_handle s) {
  if (!s) return KCVAE_OK;
  cudaSetDevice(s->device);
  cudaFree(s->s1); cudaFree(s->s2); cudaFree(s->z); cudaFree(s->psum); cudaFree(s->pmin); cudaFree(s->pmax);
  cudaFree(s->state); cudaFree(s->out_dev);
  delete s;
  return KCVAE_OK;
}

// back to the tool's start-up state (camera_streamer_qt.py:211-221): EMA min = max = 0, no pixel moments yet
int kcvae_stream_reset(kcvae_stream_handle s) {
  if (!s) return KCVAE_ERR_INVALID;
  if (cudaSetDevice(s->device) != cudaSuccess || cudaMemset(s->state, 0, 2 * sizeof(float)) != cudaSuccess) {
    s->err = "stream_reset: cudaMemset failed";
    return KCVAE_ERR_CUDA;
  }
  s->primed = false;
  s->anomaly_sum = 0.0;
  s->anomaly_sum_2 = 0.0f;
  return KCVAE_OK;
}

int kcvae_stream_update(kcvae_stream_handle s, const float* d_err, double ma, uint8_t* d_err_u8, float* h_out, void* stream) {
  if (!s || !d_err || !h_out) { if (s) s->err = "stream_update: null pointer"; return KCVAE_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaSetDevice(s->device) != cudaSuccess) { s->err = "stream_update: cudaSetDevice failed"; return KCVAE_ERR_CUDA; }
  const int n = s->H * s->W;
  const float w0 = (float)ma, w1 = (float)(1.0 - ma);    // Python doubles cast to float32 when they meet the tensors
  {
    ProfScope prof_("stream_update", st);
    g_launches += 2;
    KC_LAUNCH(stream_update_kernel, kStreamBlocks, kStreamThreads, 0, st, d_err, n, w0, w1, s->primed ? 0 : 1, s->s1, s->s2,
              s->z, s->psum, s->pmin, s->pmax);
    KC_LAUNCH(stream_finish_kernel, 1, 1024, 0, st, d_err, s->z, n, kStreamBlocks, s->psum, s->pmin, s->pmax, w0, w1,
              s->state, d_err_u8, s->out_dev);
  }
  s->primed = true;
  float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyAsync(o, s->out_dev, 7 * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    s->err = std::string("stream_update: ") + cudaGetErrorString(cudaGetLastError());
    return KCVAE_ERR_CUDA;
  }
  // :1397-1400  (count is a Python float; sum stays double, sum_2 becomes a float32 tensor)
  const double count = (double)o[0];
  s->anomaly_sum = ma * s->anomaly_sum + (1.0 - ma) * count;
  const float c2 = (float)count * (float)count;
  s->anomaly_sum_2 = w0 * s->anomaly_sum_2 + w1 * c2;
  const float sum_f = (float)s->anomaly_sum;
  const float var = s->anomaly_sum_2 - sum_f * sum_f;
  const float score = (float)(count - s->anomaly_sum) / std::sqrt(var);
  h_out[0] = o[0];                 // anomaly_count
  h_out[1] = score;                // anomaly_score (NaN / inf exactly where the reference produces them)
  h_out[2] = o[1]; h_out[3] = o[2];  // frame min / max of the error map
  h_out[4] = o[3]; h_out[5] = o[4];  // EMA min / max
  h_out[6] = o[5]; h_out[7] = o[6];  // mean / std of the per-pixel z scores
  return KCVAE_OK;
}

int kcvae_render_outputs(const float* d_norm_err, const float* d_rec, int batch, int H, int W, int C, uint8_t* d_err_u8,
                         uint8_t* d_heatmap, uint8_t* d_overlay, uint8_t* d_rec_u8, void* stream) {
  if (batch <= 0 || H <= 0 || W <= 0 || C <= 0 || C > 4) return KCVAE_ERR_INVALID;
  if ((d_heatmap || d_err_u8) && !d_norm_err) return KCVAE_ERR_INVALID;
  if ((d_overlay || d_rec_u8) && !d_rec) return KCVAE_ERR_INVALID;
  render_outputs(d_norm_err, d_rec, (int64_t)batch * H * W, C, d_err_u8, d_heatmap, d_overlay, d_rec_u8, (cudaStream_t)stream);
  return cudaPeekAtLastError() == cudaSuccess ? KCVAE_OK : KCVAE_ERR_CUDA;
}

}  // extern "C"
