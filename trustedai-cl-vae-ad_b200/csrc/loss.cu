// loss.cu - fused bandwidth kernels: image statistics + d(loss)/d(logit), latent moments
// and their gradient, metric assembly, anomaly score, noise, Adam.
//
// Reference semantics: src/kurtosis_global_cvae.py:40-110, src/kurtosis_single_cvae.py:25-77,
// do_anomaly_detection.py:57-117, Keras optimizer_v2 Adam (train.py:99-101).
// Reductions are two-level and deterministic (no float atomics).  Moments use fp64 power
// sums (hard part 3).
#include "kernels.h"

namespace kc {

// ======================================================================= image statistics
constexpr int kStatBlocks = kNumSMs * 8;
constexpr int kStatSlots = 16;  // se, xhx, xh, ex, std, min, max, -, then 8 per-channel sums of dlogit
size_t image_stats_partial_doubles() { return (size_t)kStatBlocks * kStatSlots + 1; }   // + the ticket of the pixel kernel (zero between launches)

__global__ void __launch_bounds__(256) image_stats_kernel(ImageStatsArgs a) {
  __shared__ double scratch[32];
  double t_se = 0, t_xhx = 0, t_xh = 0, t_ex = 0, t_std = 0;
  double t_g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  float mn = 3.4e38f, mx = -3.4e38f;
  const bool want_std = a.std_acc != nullptr || a.pos_sums != nullptr;
  const float* __restrict__ X = a.x;
  const float* __restrict__ XH = a.xhat;
  float* __restrict__ DL = a.dlogit;
  uint16_t* __restrict__ DL8 = a.dl8;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.P;
       p += (int64_t)gridDim.x * blockDim.x) {
    float se = 0, xhx = 0, xh = 0, ex = 0, gsum = 0;
    // per-position batch moments about the fixed pivot 0.5 (frames and reconstructions live in [0,1]):
    // fp32 sums of d and d^2 lose nothing to cancellation, and pivoted sums of different ranks still add
    float sx = 0, sxx = 0, sh = 0, shh = 0;
    const int64_t pix8 = (p / a.C) * 8 + (p % a.C);   // slot inside one frame of the 8-channel bf16 layout
    const int64_t frame8 = (a.P / a.C) * 8;
    // four frames per trip: the eight loads are issued before any dependent store
    for (int b0 = 0; b0 < a.B; b0 += 4) {
      float xv4[4], hv4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool in = b0 + u < a.B;
        const int64_t i = (int64_t)(in ? b0 + u : b0) * a.P + p;
        xv4[u] = __ldg(X + i); hv4[u] = __ldg(XH + i);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (b0 + u >= a.B) break;
        const int b = b0 + u;
        const int64_t i = (int64_t)b * a.P + p;
        const float xv = xv4[u], hv = hv4[u];
        const float d = xv - hv;
        se = fmaf(d, d, se);
        mn = fminf(mn, hv);
        mx = fmaxf(mx, hv);
        if (a.want_ce) {
          xhx = fmaf(hv, xv, xhx);
          xh += hv;
          ex += expf(xv);
        }
        if (want_std) {
          const float dx = xv - 0.5f, dh = hv - 0.5f;
          sx += dx; sxx = fmaf(dx, dx, sxx);
          sh += dh; shh = fmaf(dh, dh, shh);
        }
        if (DL || DL8) {
          const float gl = a.grad_scale * (hv - xv) * hv * (1.0f - hv);
          if (DL) DL[i] = gl;
          gsum += gl;
          if (DL8) {  // bf16 round-to-nearest-even, channel-padded layout [b][pixel][8]
            const uint32_t u32 = __float_as_uint(gl);
            const uint32_t rb = u32 + 0x7FFFu + ((u32 >> 16) & 1u);
            DL8[(int64_t)b * frame8 + pix8] = (uint16_t)(rb >> 16);
          }
        }
      }
    }
    t_se += se; t_xhx += xhx; t_xh += xh; t_ex += ex;
    if (a.dbias) {
      const int ch = (int)(p % a.C);
#pragma unroll
      for (int c = 0; c < 8; ++c) t_g[c] += (c == ch) ? (double)gsum : 0.0;
    }
    if (a.pos_sums) {
      a.pos_sums[p] = sx; a.pos_sums[a.P + p] = sxx;
      a.pos_sums[2 * a.P + p] = sh; a.pos_sums[3 * a.P + p] = shh;
    } else if (a.std_acc) {
      const double inv = 1.0 / a.B;
      const double vx = fmax((double)sxx * inv - ((double)sx * inv) * ((double)sx * inv), 0.0);
      const double vh = fmax((double)shh * inv - ((double)sh * inv) * ((double)sh * inv), 0.0);
      const double dd = sqrt(vx) - sqrt(vh);
      t_std += dd * dd;
    }
  }
  double* out = a.partial + (int64_t)blockIdx.x * kStatSlots;
  double r;
  r = block_sum(t_se, scratch);  if (threadIdx.x == 0) out[0] = r;
  r = block_sum(t_xhx, scratch); if (threadIdx.x == 0) out[1] = r;
  r = block_sum(t_xh, scratch);  if (threadIdx.x == 0) out[2] = r;
  r = block_sum(t_ex, scratch);  if (threadIdx.x == 0) out[3] = r;
  r = block_sum(t_std, scratch); if (threadIdx.x == 0) out[4] = r;
  if (a.dbias) {
    for (int c = 0; c < a.C && c < 8; ++c) {
      r = block_sum(t_g[c], scratch);
      if (threadIdx.x == 0) out[8 + c] = r;
    }
  }
  // min / max through the same scratch
  __shared__ float fs[64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  mn = warp_min(mn); mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) { fs[wid] = mn; fs[32 + wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w) { mn = fminf(mn, fs[w]); mx = fmaxf(mx, fs[32 + w]); }
    out[5] = mn; out[6] = mx;
  }
}

__device__ __forceinline__ void image_stats_finish_body(const double* partial, int blocks, int want_ce,
                                                        double* sums, float* minmax, double* std_acc, float* dbias, int C) {
  __shared__ double scratch[32];
  double v[5] = {0, 0, 0, 0, 0};
  float mn = 3.4e38f, mx = -3.4e38f;
  for (int b = threadIdx.x; b < blocks; b += blockDim.x) {
    const double* p = partial + (int64_t)b * kStatSlots;
#pragma unroll
    for (int k = 0; k < 5; ++k) v[k] += p[k];
    mn = fminf(mn, (float)p[5]);
    mx = fmaxf(mx, (float)p[6]);
  }
  for (int k = 0; k < 5; ++k) {
    const double r = block_sum(v[k], scratch);
    if (threadIdx.x == 0) {
      if (k == 0) sums[S_SE] = r;
      else if (k < 4) { if (want_ce) sums[S_SE + k] = r; else sums[S_SE + k] = 0.0; }
      else if (std_acc) std_acc[0] = r;
    }
  }
  if (dbias) {
    for (int c = 0; c < C && c < 8; ++c) {
      double g = 0;
      for (int b = threadIdx.x; b < blocks; b += blockDim.x) g += partial[(int64_t)b * kStatSlots + 8 + c];
      const double r = block_sum(g, scratch);
      if (threadIdx.x == 0) dbias[c] = (float)r;
    }
  }
  __shared__ float fs[64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  mn = warp_min(mn); mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) { fs[wid] = mn; fs[32 + wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0 && minmax) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w) { mn = fminf(mn, fs[w]); mx = fmaxf(mx, fs[32 + w]); }
    minmax[0] = mn; minmax[1] = mx; minmax[2] = -mx;   // [2]: lets data-parallel ranks fold min and max in ONE MIN all-reduce
  }
}

__global__ void image_stats_finish_kernel(const double* partial, int blocks, int want_ce,
                                          double* sums, float* minmax, double* std_acc, float* dbias, int C) {
  image_stats_finish_body(partial, blocks, want_ce, sums, minmax, std_acc, dbias, C);
}

// Tensor-core training path: thread = one PIXEL (all C channels) down the batch.  d(loss)/d(logit) leaves as one
// 16-byte bf16 unit per pixel and frame (coalesced full sectors; the element-per-thread kernel above writes the
// same units as scattered 2-byte stores), and the block that finishes last folds the per-block partials, so the
// launch of a separate single-block kernel is gone.  Same partial layout and the same fixed summation orders.
template <int C>
__global__ void __launch_bounds__(256) image_stats_pix_kernel(ImageStatsArgs a, int64_t HW) {
  __shared__ double scratch[32];
  __shared__ bool is_last;
  double t_se = 0, t_xhx = 0, t_xh = 0, t_ex = 0, t_std = 0;
  double t_g[C];
#pragma unroll
  for (int c = 0; c < C; ++c) t_g[c] = 0;
  float mn = 3.4e38f, mx = -3.4e38f;
  const bool want_std = a.std_acc != nullptr || a.pos_sums != nullptr;
  const float* __restrict__ X = a.x;
  const float* __restrict__ XH = a.xhat;
  uint4* __restrict__ DL8 = reinterpret_cast<uint4*>(a.dl8);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < HW; q += (int64_t)gridDim.x * blockDim.x) {
    float se = 0, xhx = 0, xh = 0, ex = 0;
    float gsum[C], sx[C], sxx[C], sh[C], shh[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { gsum[c] = 0; sx[c] = 0; sxx[c] = 0; sh[c] = 0; shh[c] = 0; }
    for (int b0 = 0; b0 < a.B; b0 += 4) {
      float xv4[4][C], hv4[4][C];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = ((int64_t)(b0 + u < a.B ? b0 + u : b0) * HW + q) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) { xv4[u][c] = __ldg(X + i + c); hv4[u][c] = __ldg(XH + i + c); }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (b0 + u >= a.B) break;
        uint32_t h16[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float xv = xv4[u][c], hv = hv4[u][c];
          const float d = xv - hv;
          se = fmaf(d, d, se);
          mn = fminf(mn, hv);
          mx = fmaxf(mx, hv);
          if (a.want_ce) {
            xhx = fmaf(hv, xv, xhx);
            xh += hv;
            ex += expf(xv);
          }
          if (want_std) {
            const float dx = xv - 0.5f, dh = hv - 0.5f;
            sx[c] += dx; sxx[c] = fmaf(dx, dx, sxx[c]);
            sh[c] += dh; shh[c] = fmaf(dh, dh, shh[c]);
          }
          const float gl = a.grad_scale * (hv - xv) * hv * (1.0f - hv);
          gsum[c] += gl;
          const uint32_t u32 = __float_as_uint(gl);                 // bf16 round-to-nearest-even
          h16[c] = (u32 + 0x7FFFu + ((u32 >> 16) & 1u)) >> 16;
        }
        DL8[(int64_t)(b0 + u) * HW + q] = make_uint4(h16[0] | (h16[1] << 16), h16[2] | (h16[3] << 16), h16[4] | (h16[5] << 16),
                                                     h16[6] | (h16[7] << 16));
      }
    }
    t_se += se; t_xhx += xhx; t_xh += xh; t_ex += ex;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      t_g[c] += (double)gsum[c];
      const int64_t p = q * C + c;
      if (a.pos_sums) {
        a.pos_sums[p] = sx[c]; a.pos_sums[a.P + p] = sxx[c];
        a.pos_sums[2 * a.P + p] = sh[c]; a.pos_sums[3 * a.P + p] = shh[c];
      } else if (a.std_acc) {
        const double inv = 1.0 / a.B;
        const double vx = fmax((double)sxx[c] * inv - ((double)sx[c] * inv) * ((double)sx[c] * inv), 0.0);
        const double vh = fmax((double)shh[c] * inv - ((double)sh[c] * inv) * ((double)sh[c] * inv), 0.0);
        const double dd = sqrt(vx) - sqrt(vh);
        t_std += dd * dd;
      }
    }
  }
  double* out = a.partial + (int64_t)blockIdx.x * kStatSlots;
  double r;
  r = block_sum(t_se, scratch);  if (threadIdx.x == 0) out[0] = r;
  r = block_sum(t_xhx, scratch); if (threadIdx.x == 0) out[1] = r;
  r = block_sum(t_xh, scratch);  if (threadIdx.x == 0) out[2] = r;
  r = block_sum(t_ex, scratch);  if (threadIdx.x == 0) out[3] = r;
  r = block_sum(t_std, scratch); if (threadIdx.x == 0) out[4] = r;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    r = block_sum(t_g[c], scratch);
    if (threadIdx.x == 0) out[8 + c] = r;
  }
  __shared__ float fs[64];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  mn = warp_min(mn); mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) { fs[wid] = mn; fs[32 + wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w) { mn = fminf(mn, fs[w]); mx = fmaxf(mx, fs[32 + w]); }
    out[5] = mn; out[6] = mx;
    // ticket: the block that arrives last sees every other block's partials (release / acquire through the fences)
    __threadfence();
    unsigned* ticket = reinterpret_cast<unsigned*>(a.partial + (int64_t)kStatBlocks * kStatSlots);
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (is_last) *ticket = 0u;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    image_stats_finish_body(a.partial, (int)gridDim.x, a.want_ce, a.sums, a.minmax, a.pos_sums ? nullptr : a.std_acc, a.dbias, C);
  }
}

void image_stats(const ImageStatsArgs& a, cudaStream_t st) {
  ProfScope prof_("image_stats", st);
  if (a.dl8 && !a.dlogit && a.dbias && (a.C == 3 || a.C == 1) && a.P % a.C == 0 && ((uintptr_t)a.dl8 & 15) == 0) {
    const int64_t HW = a.P / a.C;
    int blocks = cdiv(HW, 256);
    if (blocks > kStatBlocks) blocks = kStatBlocks;
    ++g_launches;
    if (a.C == 3) KC_LAUNCH(image_stats_pix_kernel<3>, blocks, 256, 0, st, a, HW);
    else KC_LAUNCH(image_stats_pix_kernel<1>, blocks, 256, 0, st, a, HW);
    return;
  }
  int blocks = cdiv(a.P, 256);
  if (blocks > kStatBlocks) blocks = kStatBlocks;
  g_launches += 2;
  KC_LAUNCH(image_stats_kernel, blocks, 256, 0, st, a);
  KC_LAUNCH(image_stats_finish_kernel, 1, 256, 0, st, a.partial, blocks, a.want_ce, a.sums, a.minmax,
            a.pos_sums ? nullptr : a.std_acc, a.dbias, a.C);
}

__global__ void __launch_bounds__(256) image_std_pos_kernel(const double* ps, int64_t P, int B, double* partial) {
  __shared__ double scratch[32];
  double t = 0;
  const double inv = 1.0 / B;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
    const double sx = ps[p], sxx = ps[P + p], sh = ps[2 * P + p], shh = ps[3 * P + p];
    const double vx = fmax(sxx * inv - (sx * inv) * (sx * inv), 0.0);
    const double vh = fmax(shh * inv - (sh * inv) * (sh * inv), 0.0);
    const double dd = sqrt(vx) - sqrt(vh);
    t += dd * dd;
  }
  const double r = block_sum(t, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = r;
}
__global__ void sum_doubles_kernel(const double* partial, int n, double* out) {
  __shared__ double scratch[32];
  double t = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += partial[i];
  const double r = block_sum(t, scratch);
  if (threadIdx.x == 0) out[0] = r;
}
void image_std_from_pos_sums(const double* pos_sums, int64_t P, int B_global, double* std_acc,
                             double* partial, cudaStream_t st) {
  ProfScope prof_("image_std", st);
  int blocks = cdiv(P, 256);
  if (blocks > kStatBlocks) blocks = kStatBlocks;
  g_launches += 2;
  KC_LAUNCH(image_std_pos_kernel, blocks, 256, 0, st, pos_sums, P, B_global, partial);
  KC_LAUNCH(sum_doubles_kernel, 1, 256, 0, st, partial, blocks, std_acc);
}

// ====================================================================== latent kernels
__device__ __forceinline__ float gen_normal(uint64_t seed, uint64_t counter, uint32_t stream, int64_t i) {
  uint32_t r[4];
  Philox ph(seed);
  ph(counter + (uint64_t)(i >> 1), stream, r);
  float n0, n1;
  box_muller(r[0], r[1], n0, n1);
  return (i & 1) ? n1 : n0;
}

__global__ void reparam_kernel(const float* head, const float* mean_in, const float* logvar_in,
                               int B, int L, const float* eps, int gen_eps, uint64_t seed,
                               uint64_t counter, float* z, float* mean, float* logvar, float* eps_out) {
  const int64_t n = (int64_t)B * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / L), j = (int)(i % L);
    float m, lv;
    if (head) { m = head[(int64_t)b * 2 * L + j]; lv = head[(int64_t)b * 2 * L + L + j]; }
    else { m = mean_in[i]; lv = logvar_in[i]; }
    float e = 0.0f;
    if (eps) e = eps[i];
    else if (gen_eps) e = gen_normal(seed, counter, 1u, i);
    z[i] = m + lv * 0.5f + e;   // src/abstract_cvae.py:128
    if (mean) mean[i] = m;
    if (logvar) logvar[i] = lv;
    if (eps_out) eps_out[i] = e;
  }
}
void reparameterize(const float* head, int B, int L, const float* eps, int gen_eps, uint64_t seed,
                    uint64_t counter, float* z, float* mean, float* logvar, float* eps_out,
                    cudaStream_t st) {
  ProfScope prof_("reparam", st);
  ++g_launches;
  KC_LAUNCH(reparam_kernel, grid_for((int64_t)B * L, 256), 256, 0, st, head, (const float*)nullptr,
            (const float*)nullptr, B, L, eps, gen_eps, seed, counter, z, mean, logvar, eps_out);
}
void reparam_from_parts(const float* mean, const float* logvar, int B, int L, const float* eps,
                        int gen_eps, uint64_t seed, uint64_t counter, float* z, cudaStream_t st) {
  ProfScope prof_("reparam", st);
  ++g_launches;
  KC_LAUNCH(reparam_kernel, grid_for((int64_t)B * L, 256), 256, 0, st, (const float*)nullptr, mean,
            logvar, B, L, eps, gen_eps, seed, counter, z, (float*)nullptr, (float*)nullptr,
            (float*)nullptr);
}

// one block; Global: totals; Single: per-column power sums (axis=0 moments)
__global__ void __launch_bounds__(256) latent_sums_kernel(const float* z, const float* mean, const float* logvar,
                                                          int B, int L, int model_type, double* sums) {
  __shared__ double scratch[32];
  const int64_t n = (int64_t)B * L;
  double absz = 0, kl = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = z[i];
    absz += fabs(v);
    if (mean) {
      const double m = mean[i], lv = logvar[i];
      kl += fabs(1.0 + lv * lv - m * m - exp(lv * lv));   // src/kurtosis_global_cvae.py:36-38
    }
    if (model_type == 0) { const double v2 = v * v; s1 += v; s2 += v2; s3 += v2 * v; s4 += v2 * v2; }
  }
  double r;
  r = block_sum(absz, scratch); if (threadIdx.x == 0) sums[S_ABSZ] = r;
  r = block_sum(kl, scratch);   if (threadIdx.x == 0) sums[S_KL] = r;
  if (model_type == 0) {
    r = block_sum(s1, scratch); if (threadIdx.x == 0) sums[S_Z1 + 0] = r;
    r = block_sum(s2, scratch); if (threadIdx.x == 0) sums[S_Z1 + 1] = r;
    r = block_sum(s3, scratch); if (threadIdx.x == 0) sums[S_Z1 + 2] = r;
    r = block_sum(s4, scratch); if (threadIdx.x == 0) sums[S_Z1 + 3] = r;
  } else {
    // thread per column (L <= kMaxLatent), rows sequential: deterministic
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
      double c1 = 0, c2 = 0, c3 = 0, c4 = 0;
      for (int b = 0; b < B; ++b) {
        const double v = z[(int64_t)b * L + j], v2 = v * v;
        c1 += v; c2 += v2; c3 += v2 * v; c4 += v2 * v2;
      }
      sums[S_Z1 + 4 * j + 0] = c1; sums[S_Z1 + 4 * j + 1] = c2;
      sums[S_Z1 + 4 * j + 2] = c3; sums[S_Z1 + 4 * j + 3] = c4;
    }
  }
}
void latent_sums(const float* z, const float* mean, const float* logvar, int B, int L, int model_type,
                 double* sums, cudaStream_t st) {
  ProfScope prof_("latent_sums", st);
  ++g_launches;
  KC_LAUNCH(latent_sums_kernel, 1, 256, 0, st, z, mean, logvar, B, L, model_type, sums);
}

struct Moments { double mu, var, sigma, skew, kurt; };
// population central moments from raw power sums (fp64), tf.math.reduce_std / divide_no_nan
__device__ __forceinline__ Moments moments_from_sums(const double* s, double N) {
  Moments m;
  const double e1 = s[0] / N, e2 = s[1] / N, e3 = s[2] / N, e4 = s[3] / N;
  m.mu = e1;
  m.var = fmax(e2 - e1 * e1, 0.0);
  m.sigma = sqrt(m.var);
  const double m3 = e3 - 3.0 * e1 * e2 + 2.0 * e1 * e1 * e1;
  const double m4 = e4 - 4.0 * e1 * e3 + 6.0 * e1 * e1 * e2 - 3.0 * e1 * e1 * e1 * e1;
  if (m.sigma > 0.0) { m.skew = m3 / (m.var * m.sigma); m.kurt = m4 / (m.var * m.var); }
  else { m.skew = 0.0; m.kurt = 0.0; }
  return m;
}
__device__ __forceinline__ double sgn(double v) { return (v > 0.0) - (v < 0.0); }

__global__ void finalize_metrics_kernel(const double* sums, const float* minmax, const double* std_acc,
                                        int Bg, int L, int64_t P, int model_type, LossWeights lw,
                                        int have_ce, float* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double nan_ = nan("");
  const double NP = (double)Bg * (double)P;
  const double mse = sums[S_SE] / NP;
  const double z_l1 = sums[S_ABSZ] / ((double)Bg * L);
  const double rmin = minmax ? minmax[0] : nan_, rmax = minmax ? fmax((double)minmax[1], -(double)minmax[2]) : nan_;
  const double xstd = std_acc ? std_acc[0] / (double)P : nan_;
  for (int i = 0; i < 16; ++i) out[i] = 0.0f;
  if (model_type == 0) {
    const Moments m = moments_from_sums(sums + S_Z1, (double)Bg * L);
    const double var_loss = fabs(1.0 - m.var);
    const double skew_loss = fabs(m.skew);
    const double kurt_loss = fabs((double)lw.kurtosis_target - m.kurt);
    const double loss = lw.w_mse * mse + lw.w_kurtosis * kurt_loss + lw.w_skew * skew_loss + lw.w_z_l1_reg * z_l1;
    // -mean(xhat * (x - log sum exp x))   (src/kurtosis_global_cvae.py:46-47)
    const double ce = have_ce ? -(sums[S_XHX] - log(sums[S_EX]) * sums[S_XH]) / NP : nan_;
    out[0] = (float)loss; out[1] = (float)mse; out[2] = (float)z_l1; out[3] = (float)var_loss;
    out[4] = (float)skew_loss; out[5] = (float)kurt_loss; out[6] = (float)m.kurt;
    out[7] = (float)rmin; out[8] = (float)rmax; out[9] = (float)ce;
    out[10] = (float)(0.5 * sums[S_KL]); out[11] = (float)xstd;
  } else {
    double kl = 0, sl = 0, l2 = 0, k2 = 0;
    for (int j = 0; j < L; ++j) {
      const Moments m = moments_from_sums(sums + S_Z1 + 4 * j, (double)Bg);
      const double dk = m.kurt - (double)lw.kurtosis_target;
      kl += dk * dk; sl += m.skew * m.skew; l2 += m.mu * m.mu; k2 += m.kurt * m.kurt;
    }
    kl /= L; sl /= L; l2 = sqrt(l2);
    const double loss = lw.w_mse * mse + lw.w_kurtosis * kl + lw.w_skew * sl + lw.w_z_l1_reg * l2;
    out[0] = (float)loss; out[1] = (float)mse; out[2] = (float)z_l1; out[3] = (float)l2;
    out[4] = (float)sl; out[5] = (float)kl; out[6] = (float)sqrt(k2 / L);
    out[7] = (float)rmin; out[8] = (float)rmax; out[9] = (float)xstd;
  }
}
void finalize_metrics(const double* sums, const float* minmax, const double* std_acc, int B_global,
                      int L, int64_t P, int model_type, LossWeights lw, int have_ce, float* metrics,
                      cudaStream_t st) {
  ProfScope prof_("finalize_metrics", st);
  ++g_launches;
  KC_LAUNCH(finalize_metrics_kernel, 1, 32, 0, st, sums, minmax, std_acc, B_global, L, P, model_type, lw,
            have_ce, metrics);
}

__global__ void latent_backward_kernel(const float* z, const float* g_z, const double* sums, int Bl,
                                       int Bg, int L, int model_type, LossWeights lw, float* dhead) {
  __shared__ double sh_l2;
  if (model_type == 1) {
    if (threadIdx.x == 0) {
      double l2 = 0;
      for (int j = 0; j < L; ++j) { const double mu = sums[S_Z1 + 4 * j] / Bg; l2 += mu * mu; }
      sh_l2 = sqrt(l2);
    }
    __syncthreads();
  }
  const int64_t n = (int64_t)Bl * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / L), j = (int)(i % L);
    const double zv = z[i];
    double g = g_z ? (double)g_z[i] : 0.0;
    if (model_type == 0) {
      const double N = (double)Bg * L;
      const Moments m = moments_from_sums(sums + S_Z1, N);
      if (m.sigma > 0.0) {
        const double s = (zv - m.mu) / m.sigma;
        const double dK = 4.0 / (N * m.sigma) * (s * s * s - m.skew - m.kurt * s);
        const double dS = 3.0 / (N * m.sigma) * (s * s - 1.0 - m.skew * s);
        g += lw.w_kurtosis * (-sgn((double)lw.kurtosis_target - m.kurt)) * dK + lw.w_skew * sgn(m.skew) * dS;
      }
      g += lw.w_z_l1_reg * sgn(zv) / N;
    } else {
      const double N = (double)Bg;
      const Moments m = moments_from_sums(sums + S_Z1 + 4 * j, N);
      if (m.sigma > 0.0) {
        const double s = (zv - m.mu) / m.sigma;
        const double dK = 4.0 / (N * m.sigma) * (s * s * s - m.skew - m.kurt * s);
        const double dS = 3.0 / (N * m.sigma) * (s * s - 1.0 - m.skew * s);
        g += lw.w_kurtosis * 2.0 * (m.kurt - (double)lw.kurtosis_target) / L * dK + lw.w_skew * 2.0 * m.skew / L * dS;
      }
      if (sh_l2 > 0.0) g += lw.w_z_l1_reg * m.mu / (sh_l2 * N);
    }
    dhead[(int64_t)b * 2 * L + j] = (float)g;             // d/d mean
    dhead[(int64_t)b * 2 * L + L + j] = (float)(0.5 * g); // d/d logvar (additive reparam)
  }
}
void latent_backward(const float* z, const float* g_z, const double* sums, int B_local, int B_global,
                     int L, int model_type, LossWeights lw, float* dhead, cudaStream_t st) {
  ProfScope prof_("latent_backward", st);
  ++g_launches;
  KC_LAUNCH(latent_backward_kernel, grid_for((int64_t)B_local * L, 128), 128, 0, st, z, g_z, sums, B_local,
            B_global, L, model_type, lw, dhead);
}

// ================================================================================ scoring
constexpr int kScorePixPerBlock = 2048;
size_t score_partial_floats(int B, int64_t HW) { return (size_t)3 * B * cdiv(HW, kScorePixPerBlock); }

// partial layout: [B][chunks][3] = (sum, min, max) of the per-pixel error
__global__ void __launch_bounds__(256) score_kernel(const float* x, const float* xhat, int64_t HW, int C,
                                                    float* err, float* partial) {
  __shared__ float scratch[32];
  __shared__ float fs[64];
  const int b = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * kScorePixPerBlock;
  int64_t p1 = p0 + kScorePixPerBlock;
  if (p1 > HW) p1 = HW;
  float t = 0.0f, mn = 3.4e38f, mx = -3.4e38f;
  for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    const int64_t base = ((int64_t)b * HW + p) * C;
    float e = 0.0f;
    for (int c = 0; c < C; ++c) {
      const float d = __ldg(x + base + c) - __ldg(xhat + base + c);
      e = fmaf(d, d, e);
    }
    if (err) err[(int64_t)b * HW + p] = e;
    t += e;
    mn = fminf(mn, e);
    mx = fmaxf(mx, e);
  }
  const float r = block_sum(t, scratch);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  mn = warp_min(mn); mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) { fs[wid] = mn; fs[32 + wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 1; w < nw; ++w) { mn = fminf(mn, fs[w]); mx = fmaxf(mx, fs[32 + w]); }
    float* o = partial + ((int64_t)b * gridDim.x + blockIdx.x) * 3;
    o[0] = r; o[1] = mn; o[2] = mx;
  }
}
__global__ void score_finish_kernel(const float* partial, int chunks, int B, float* score_out, float* err_minmax) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.0f, mn = 3.4e38f, mx = -3.4e38f;
  for (int c = 0; c < chunks; ++c) {
    const float* o = partial + ((int64_t)b * chunks + c) * 3;
    s += o[0]; mn = fminf(mn, o[1]); mx = fmaxf(mx, o[2]);
  }
  score_out[b] = s;
  if (err_minmax) { err_minmax[2 * b] = mn; err_minmax[2 * b + 1] = mx; }
}
void score(const float* x, const float* xhat, int B, int64_t HW, int C, float* err, float* score_out,
           float* err_minmax, float* partial, cudaStream_t st) {
  ProfScope prof_("score", st);
  const int chunks = cdiv(HW, kScorePixPerBlock);
  g_launches += 2;
  KC_LAUNCH(score_kernel, dim3(chunks, B), 256, 0, st, x, xhat, HW, C, err, partial);
  KC_LAUNCH(score_finish_kernel, cdiv(B, 128), 128, 0, st, partial, chunks, B, score_out, err_minmax);
}

__global__ void normalize_scores_kernel(const float* err, const float* score_in, int B, int64_t HW,
                                        float meu, float sigma, float emin, float emax, float thr,
                                        float* norm, float* z, uint8_t* flags) {
  const int64_t n = (int64_t)B * HW;
  const float inv = 1.0f / (emax - emin);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (norm && err) norm[i] = (err[i] - emin) * inv;            // do_anomaly_detection.py:92
    if (i < B) {
      const float zs = (score_in[i] - meu) / sigma;              // :91
      if (z) z[i] = zs;
      if (flags) flags[i] = zs > thr ? 1 : 0;                    // :106
    }
  }
}
void normalize_scores(const float* err, const float* score_in, int B, int64_t HW, float meu, float sigma,
                      float emin, float emax, float thr, float* norm, float* z, uint8_t* flags,
                      cudaStream_t st) {
  ProfScope prof_("normalize_scores", st);
  ++g_launches;
  KC_LAUNCH(normalize_scores_kernel, grid_for((int64_t)B * HW, 256), 256, 0, st, err, score_in, B, HW, meu,
            sigma, emin, emax, thr, norm, z, flags);
}

// =================================================================================== misc
__global__ void add_noise_kernel(const float* x, const float* noise, int64_t n, float stddev,
                                 uint64_t seed, uint64_t counter, float* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float e = noise ? noise[i] : stddev * gen_normal(seed, counter, 2u, i);
    out[i] = x[i] + e;                                            // src/abstract_cvae.py:118
  }
}
void add_noise(const float* x, const float* noise, int64_t n, float stddev, uint64_t seed, uint64_t counter,
               float* out, cudaStream_t st) {
  ProfScope prof_("add_noise", st);
  ++g_launches;
  KC_LAUNCH(add_noise_kernel, grid_for(n, 256), 256, 0, st, x, noise, n, stddev, seed, counter, out);
}

// Keras optimizer_v2 Adam: p -= lr_t * m / (sqrt(v) + eps), lr_t = lr*sqrt(1-b2^t)/(1-b1^t)
__global__ void __launch_bounds__(256) adam_kernel(float* p, const float* g, float* m, float* v, int64_t n,
                                                   float lr_t, float b1, float b2, float eps) {
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float c1 = 1.0f - b1, c2 = 1.0f - b2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    mm.x = b1 * mm.x + c1 * gg.x; vv.x = b2 * vv.x + c2 * gg.x * gg.x; pp.x -= lr_t * mm.x / (sqrtf(vv.x) + eps);
    mm.y = b1 * mm.y + c1 * gg.y; vv.y = b2 * vv.y + c2 * gg.y * gg.y; pp.y -= lr_t * mm.y / (sqrtf(vv.y) + eps);
    mm.z = b1 * mm.z + c1 * gg.z; vv.z = b2 * vv.z + c2 * gg.z * gg.z; pp.z -= lr_t * mm.z / (sqrtf(vv.z) + eps);
    mm.w = b1 * mm.w + c1 * gg.w; vv.w = b2 * vv.w + c2 * gg.w * gg.w; pp.w -= lr_t * mm.w / (sqrtf(vv.w) + eps);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  const int64_t t0 = n4 << 2;
  for (int64_t i = t0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gg = g[i];
    const float mm = b1 * m[i] + c1 * gg, vv = b2 * v[i] + c2 * gg * gg;
    m[i] = mm; v[i] = vv;
    p[i] -= lr_t * mm / (sqrtf(vv) + eps);
  }
}
void adam_update(float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2,
                 float eps, cudaStream_t st) {
  ProfScope prof_("adam", st);
  ++g_launches;
  KC_LAUNCH(adam_kernel, grid_for(n / 4 + 1, 256, 8, 2), 256, 0, st, p, g, m, v, n, lr_t, b1, b2, eps);
}

__global__ void glorot_fill_kernel(float* p, int64_t n, float limit, uint64_t seed, uint32_t stream_id) {
  Philox ph(seed);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r[4];
    ph((uint64_t)(i >> 2), 16u + stream_id, r);
    p[i] = (2.0f * u01(r[i & 3]) - 1.0f) * limit;
  }
}
void glorot_fill(float* p, int64_t n, float limit, uint64_t seed, uint32_t stream_id, cudaStream_t st) {
  ProfScope prof_("glorot_fill", st);
  ++g_launches;
  KC_LAUNCH(glorot_fill_kernel, grid_for(n, 256), 256, 0, st, p, n, limit, seed, stream_id);
}

}  // namespace kc
