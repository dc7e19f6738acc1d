// conv.cu - generic fp32 3x3 convolution family on CUDA cores (any channel counts).
//
// One gather-form kernel covers the three TF layers of the path and their data gradients
// (SURVEY Appendix A1-A4):
//   CONV_S2   Conv2D k3 s2 SAME                 (src/abstract_cvae.py:32)   + dgrad of CONVT_S2
//   CONVT_S2  Conv2DTranspose k3 s2 SAME        (src/abstract_cvae.py:83)   + dgrad of CONV_S2
//   CONV_S1   Conv2DTranspose k3 s1 SAME (flip) (src/abstract_cvae.py:88)   + its dgrad (no flip)
// Mapping: thread = (4 consecutive output columns) x (one output channel), channel fastest,
// so weight reads and output writes are coalesced and input reads are warp-broadcasts.
// These are the precise reference kernels; the tcgen05 kernels in tc_conv.cu take over the
// shapes they cover when precision == BF16_TC.
#include "kernels.h"

namespace kc {

int64_t g_launches = 0;

constexpr int PT = 4;  // output pixels per thread along W

template <int MODE, int EPI>
__global__ void __launch_bounds__(256) conv3x3_kernel(ConvArgs a, int64_t total, int WG) {
  const int CiCo = a.Ci * a.Co;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(idx % a.Co);
    int64_t g = idx / a.Co;
    const int oxg = (int)(g % WG);
    g /= WG;
    const int oy = (int)(g % a.Ho);
    const int n = (int)(g / a.Ho);
    const int ox0 = oxg * PT;

    float acc[PT];
    const float b0 = a.bias ? __ldg(a.bias + co) : 0.0f;
#pragma unroll
    for (int p = 0; p < PT; ++p) acc[p] = b0;

#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      int iy;
      bool vy;
      if (MODE == CONV_S2) {
        iy = 2 * oy + kh - a.pad_t;
        vy = iy >= 0 && iy < a.Hi;
      } else if (MODE == CONV_S1) {
        iy = a.flip ? oy + 1 - kh : oy - 1 + kh;
        vy = iy >= 0 && iy < a.Hi;
      } else {
        const int t = oy + a.pad_t - kh;
        iy = t >> 1;
        vy = t >= 0 && !(t & 1) && iy < a.Hi;
      }
      if (!vy) continue;
      const float* in_row = a.in + ((int64_t)n * a.Hi + iy) * a.Wi * a.Ci;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int ix[PT];
        bool vx[PT];
        bool any = false;
#pragma unroll
        for (int p = 0; p < PT; ++p) {
          const int ox = ox0 + p;
          if (MODE == CONV_S2) {
            ix[p] = 2 * ox + kw - a.pad_l;
            vx[p] = ix[p] >= 0 && ix[p] < a.Wi;
          } else if (MODE == CONV_S1) {
            ix[p] = a.flip ? ox + 1 - kw : ox - 1 + kw;
            vx[p] = ix[p] >= 0 && ix[p] < a.Wi;
          } else {
            const int t = ox + a.pad_l - kw;
            ix[p] = t >> 1;
            vx[p] = t >= 0 && !(t & 1) && ix[p] < a.Wi;
          }
          vx[p] = vx[p] && ox < a.Wo;
          any = any || vx[p];
        }
        if (!any) continue;
        const float* wp = a.w + (int64_t)(kh * 3 + kw) * CiCo + (int64_t)co * a.w_sco;
        for (int ci = 0; ci < a.Ci; ++ci) {
          const float wv = __ldg(wp + (int64_t)ci * a.w_sci);
#pragma unroll
          for (int p = 0; p < PT; ++p)
            if (vx[p]) acc[p] = fmaf(__ldg(in_row + (int64_t)ix[p] * a.Ci + ci), wv, acc[p]);
        }
      }
    }
    const int64_t obase = (((int64_t)n * a.Ho + oy) * a.Wo + ox0) * a.Co + co;
#pragma unroll
    for (int p = 0; p < PT; ++p) {
      if (ox0 + p >= a.Wo) break;
      float v = acc[p];
      const int64_t o = obase + (int64_t)p * a.Co;
      if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.0f);
      if (EPI == EPI_BIAS_SIGMOID) v = 1.0f / (1.0f + expf(-v));
      if (EPI == EPI_MASK) v = __ldg(a.mask + o) > 0.0f ? v : 0.0f;
      a.out[o] = v;
    }
  }
}


// =========================================================================================
// Specialised fp32 kernels for the channel shapes of the path (one side <= 8 channels, the
// other 32): same arithmetic as conv3x3_kernel, but a thread owns whole channel vectors of its
// pixel(s), weights sit in shared memory in [tap][ci][co] order and are read as broadcast
// float4, so every input value feeds 8-32 FMAs instead of 1.
// =========================================================================================
constexpr int MANY = 32;

__device__ __forceinline__ void stage_weights(const ConvArgs& a, float* ws, int cop) {
  // canonical [tap][ci][cop] (zero padded in co) from the strided source layout
  const int n = 9 * a.Ci * cop;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int co = i % cop, ci = (i / cop) % a.Ci, tap = i / (cop * a.Ci);
    ws[i] = co < a.Co ? __ldg(a.w + (int64_t)tap * a.Ci * a.Co + (int64_t)ci * a.w_sci + (int64_t)co * a.w_sco) : 0.f;
  }
}

template <int EPI>
__device__ __forceinline__ void store_many(const ConvArgs& a, int64_t o, const float* acc, const float* sbias) {
  float4* dst = reinterpret_cast<float4*>(a.out + o);
#pragma unroll
  for (int g = 0; g < MANY / 4; ++g) {
    float y[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = acc[g * 4 + e] + sbias[g * 4 + e];
      if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
      if (EPI == EPI_BIAS_SIGMOID) v = 1.0f / (1.0f + expf(-v));
      y[e] = v;
    }
    if (EPI == EPI_MASK) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(a.mask + o) + g);
      y[0] = m.x > 0.f ? y[0] : 0.f; y[1] = m.y > 0.f ? y[1] : 0.f;
      y[2] = m.z > 0.f ? y[2] : 0.f; y[3] = m.w > 0.f ? y[3] : 0.f;
    }
    dst[g] = make_float4(y[0], y[1], y[2], y[3]);
  }
}

// ---- A: Conv2D s2, Ci <= 8 -> Co = 32.  thread = 2 adjacent output pixels x 32 channels ----
template <int EPI>
__global__ void __launch_bounds__(128) conv_s2_few2many_kernel(ConvArgs a, int64_t npairs, int WP) {
  KC_DYN_SMEM(float, ws);
  __shared__ float sbias[MANY];
  stage_weights(a, ws, MANY);
  if (threadIdx.x < MANY) sbias[threadIdx.x] = (a.bias && EPI != EPI_MASK) ? a.bias[threadIdx.x] : 0.f;
  __syncthreads();
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < npairs; idx += (int64_t)gridDim.x * blockDim.x) {
    const int pp = (int)(idx % WP);
    const int oy = (int)((idx / WP) % a.Ho);
    const int n = (int)(idx / ((int64_t)WP * a.Ho));
    const int ox0 = pp * 2;
    float acc[2][MANY];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int c = 0; c < MANY; ++c) acc[p][c] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = 2 * oy + kh - a.pad_t;
      if (iy < 0 || iy >= a.Hi) continue;
      const float* row = a.in + ((int64_t)n * a.Hi + iy) * a.Wi * a.Ci;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix0 = 2 * ox0 + kw - a.pad_l, ix1 = ix0 + 2;
        const bool v0 = ix0 >= 0 && ix0 < a.Wi, v1 = ix1 >= 0 && ix1 < a.Wi && ox0 + 1 < a.Wo;
        const float4* wt = reinterpret_cast<const float4*>(ws + (kh * 3 + kw) * a.Ci * MANY);
        for (int ci = 0; ci < a.Ci; ++ci) {
          const float x0 = v0 ? __ldg(row + (int64_t)ix0 * a.Ci + ci) : 0.f;
          const float x1 = v1 ? __ldg(row + (int64_t)ix1 * a.Ci + ci) : 0.f;
#pragma unroll
          for (int g = 0; g < MANY / 4; ++g) {
            const float4 w = wt[ci * (MANY / 4) + g];
            acc[0][g * 4 + 0] = fmaf(x0, w.x, acc[0][g * 4 + 0]); acc[1][g * 4 + 0] = fmaf(x1, w.x, acc[1][g * 4 + 0]);
            acc[0][g * 4 + 1] = fmaf(x0, w.y, acc[0][g * 4 + 1]); acc[1][g * 4 + 1] = fmaf(x1, w.y, acc[1][g * 4 + 1]);
            acc[0][g * 4 + 2] = fmaf(x0, w.z, acc[0][g * 4 + 2]); acc[1][g * 4 + 2] = fmaf(x1, w.z, acc[1][g * 4 + 2]);
            acc[0][g * 4 + 3] = fmaf(x0, w.w, acc[0][g * 4 + 3]); acc[1][g * 4 + 3] = fmaf(x1, w.w, acc[1][g * 4 + 3]);
          }
        }
      }
    }
    const int64_t o = (((int64_t)n * a.Ho + oy) * a.Wo + ox0) * MANY;
    store_many<EPI>(a, o, acc[0], sbias);
    if (ox0 + 1 < a.Wo) store_many<EPI>(a, o + MANY, acc[1], sbias);
  }
}

// ---- A': Conv2DTranspose s2 gather, Ci <= 8 -> Co = 32.  warp = one output parity phase ----
template <int EPI>
__global__ void __launch_bounds__(128) convT_s2_few2many_kernel(ConvArgs a, int64_t nq, int HQ, int WQ) {
  KC_DYN_SMEM(float, ws);
  __shared__ float sbias[MANY];
  stage_weights(a, ws, MANY);
  if (threadIdx.x < MANY) sbias[threadIdx.x] = (a.bias && EPI != EPI_MASK) ? a.bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int phase = threadIdx.x >> 5, lane = threadIdx.x & 31;   // 4 warps = 4 parity phases
  const int pa = phase >> 1, pb = phase & 1;
  for (int64_t q0 = (int64_t)blockIdx.x * 32; q0 < nq; q0 += (int64_t)gridDim.x * 32) {
    const int64_t q = q0 + lane;
    if (q >= nq) continue;
    const int jq = (int)(q % WQ);
    const int iq = (int)((q / WQ) % HQ);
    const int n = (int)(q / ((int64_t)WQ * HQ));
    const int oy = 2 * iq + pa, ox = 2 * jq + pb;
    if (oy >= a.Ho || ox >= a.Wo) continue;
    float acc[MANY];
#pragma unroll
    for (int c = 0; c < MANY; ++c) acc[c] = 0.f;
    for (int kh = (pa + a.pad_t) & 1; kh < 3; kh += 2) {
      const int ty = oy + a.pad_t - kh;
      if (ty < 0 || (ty >> 1) >= a.Hi) continue;
      const float* row = a.in + ((int64_t)n * a.Hi + (ty >> 1)) * a.Wi * a.Ci;
      for (int kw = (pb + a.pad_l) & 1; kw < 3; kw += 2) {
        const int tx = ox + a.pad_l - kw;
        if (tx < 0 || (tx >> 1) >= a.Wi) continue;
        const float* px = row + (int64_t)(tx >> 1) * a.Ci;
        const float4* wt = reinterpret_cast<const float4*>(ws + (kh * 3 + kw) * a.Ci * MANY);
        for (int ci = 0; ci < a.Ci; ++ci) {
          const float x0 = __ldg(px + ci);
#pragma unroll
          for (int g = 0; g < MANY / 4; ++g) {
            const float4 w = wt[ci * (MANY / 4) + g];
            acc[g * 4 + 0] = fmaf(x0, w.x, acc[g * 4 + 0]);
            acc[g * 4 + 1] = fmaf(x0, w.y, acc[g * 4 + 1]);
            acc[g * 4 + 2] = fmaf(x0, w.z, acc[g * 4 + 2]);
            acc[g * 4 + 3] = fmaf(x0, w.w, acc[g * 4 + 3]);
          }
        }
      }
    }
    store_many<EPI>(a, (((int64_t)n * a.Ho + oy) * a.Wo + ox) * MANY, acc, sbias);
  }
}

constexpr int FEW = 8;
template <int EPI>
__device__ __forceinline__ void store_few(const ConvArgs& a, int64_t o, const float* acc, const float* sbias) {
#pragma unroll
  for (int c = 0; c < FEW; ++c) {
    if (c < a.Co) {
      float v = acc[c] + sbias[c];
      if (EPI == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
      if (EPI == EPI_BIAS_SIGMOID) v = 1.0f / (1.0f + expf(-v));
      if (EPI == EPI_MASK) v = __ldg(a.mask + o + c) > 0.f ? v : 0.f;
      a.out[o + c] = v;
    }
  }
}
__device__ __forceinline__ void fma_4x8(const float4 x, const float4* w, float* acc) {
  // w: 4 input channels x 8 output channels = 8 float4 (ci-major)
  const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 w0 = w[i * 2], w1 = w[i * 2 + 1];
    acc[0] = fmaf(xv[i], w0.x, acc[0]); acc[1] = fmaf(xv[i], w0.y, acc[1]);
    acc[2] = fmaf(xv[i], w0.z, acc[2]); acc[3] = fmaf(xv[i], w0.w, acc[3]);
    acc[4] = fmaf(xv[i], w1.x, acc[4]); acc[5] = fmaf(xv[i], w1.y, acc[5]);
    acc[6] = fmaf(xv[i], w1.z, acc[6]); acc[7] = fmaf(xv[i], w1.w, acc[7]);
  }
}

// ---- B1: Conv2D s2, Ci % 4 == 0 -> Co <= 8.  thread = 2 adjacent output pixels x 8 channels ----
template <int EPI>
__global__ void __launch_bounds__(128) conv_s2_many2few_kernel(ConvArgs a, int64_t npairs, int WP) {
  KC_DYN_SMEM(float, ws);
  __shared__ float sbias[FEW];
  stage_weights(a, ws, FEW);
  if (threadIdx.x < FEW) sbias[threadIdx.x] = (a.bias && EPI != EPI_MASK && (int)threadIdx.x < a.Co) ? a.bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int C4 = a.Ci >> 2;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < npairs; idx += (int64_t)gridDim.x * blockDim.x) {
    const int pp = (int)(idx % WP);
    const int oy = (int)((idx / WP) % a.Ho);
    const int n = (int)(idx / ((int64_t)WP * a.Ho));
    const int ox0 = pp * 2;
    float acc0[FEW], acc1[FEW];
#pragma unroll
    for (int c = 0; c < FEW; ++c) { acc0[c] = 0.f; acc1[c] = 0.f; }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = 2 * oy + kh - a.pad_t;
      if (iy < 0 || iy >= a.Hi) continue;
      const float* row = a.in + ((int64_t)n * a.Hi + iy) * a.Wi * a.Ci;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix0 = 2 * ox0 + kw - a.pad_l, ix1 = ix0 + 2;
        const bool v0 = ix0 >= 0 && ix0 < a.Wi, v1 = ix1 >= 0 && ix1 < a.Wi && ox0 + 1 < a.Wo;
        const float4* p0 = reinterpret_cast<const float4*>(row + (int64_t)(v0 ? ix0 : 0) * a.Ci);
        const float4* p1 = reinterpret_cast<const float4*>(row + (int64_t)(v1 ? ix1 : 0) * a.Ci);
        const float4* wt = reinterpret_cast<const float4*>(ws + (kh * 3 + kw) * a.Ci * FEW);
        for (int c4 = 0; c4 < C4; ++c4) {
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 x0 = v0 ? __ldg(p0 + c4) : z;
          const float4 x1 = v1 ? __ldg(p1 + c4) : z;
          fma_4x8(x0, wt + c4 * 8, acc0);
          fma_4x8(x1, wt + c4 * 8, acc1);
        }
      }
    }
    const int64_t o = (((int64_t)n * a.Ho + oy) * a.Wo + ox0) * a.Co;
    store_few<EPI>(a, o, acc0, sbias);
    if (ox0 + 1 < a.Wo) store_few<EPI>(a, o + a.Co, acc1, sbias);
  }
}

// ---- B2: Conv2DTranspose s2 (pad 0), Ci % 4 == 0 -> Co <= 8.  thread = one input pixel's 2x2 output quad ----
template <int EPI>
__global__ void __launch_bounds__(128) convT_s2_many2few_kernel(ConvArgs a, int64_t nq) {
  KC_DYN_SMEM(float, ws);
  __shared__ float sbias[FEW];
  stage_weights(a, ws, FEW);
  if (threadIdx.x < FEW) sbias[threadIdx.x] = (a.bias && EPI != EPI_MASK && (int)threadIdx.x < a.Co) ? a.bias[threadIdx.x] : 0.f;
  __syncthreads();
  const int C4 = a.Ci >> 2;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(q % a.Wi);
    const int i = (int)((q / a.Wi) % a.Hi);
    const int n = (int)(q / ((int64_t)a.Wi * a.Hi));
    float acc[4][FEW];
#pragma unroll
    for (int ph = 0; ph < 4; ++ph)
#pragma unroll
      for (int c = 0; c < FEW; ++c) acc[ph][c] = 0.f;
#pragma unroll
    for (int di = 0; di < 2; ++di) {
      if (i - di < 0) continue;
#pragma unroll
      for (int dj = 0; dj < 2; ++dj) {
        if (j - dj < 0) continue;
        const float4* px = reinterpret_cast<const float4*>(a.in + (((int64_t)n * a.Hi + (i - di)) * a.Wi + (j - dj)) * a.Ci);
        for (int c4 = 0; c4 < C4; ++c4) {
          const float4 x = __ldg(px + c4);
#pragma unroll
          for (int pa = 0; pa < 2 - di; ++pa)            // kh = 2*di + pa <= 2
#pragma unroll
            for (int pb = 0; pb < 2 - dj; ++pb) {        // kw = 2*dj + pb <= 2
              const int tap = (2 * di + pa) * 3 + (2 * dj + pb);
              fma_4x8(x, reinterpret_cast<const float4*>(ws + tap * a.Ci * FEW) + c4 * 8, acc[pa * 2 + pb]);
            }
        }
      }
    }
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
      const int oy = 2 * i + (ph >> 1), ox = 2 * j + (ph & 1);
      if (oy < a.Ho && ox < a.Wo) store_few<EPI>(a, (((int64_t)n * a.Ho + oy) * a.Wo + ox) * a.Co, acc[ph], sbias);
    }
  }
}

#ifndef KCVAE_EMU
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
// 4-byte async copy; !valid writes a zero without touching the source
__device__ __forceinline__ void cp_async4_zfill(float* smem_dst, const float* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
// 16-byte async copy; !valid writes zeros without touching the source
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#else
static inline void cp_async16(void* d, const void* s) { memcpy(d, s, 16); }
static inline void cp_async4_zfill(float* d, const float* s, bool valid) { *d = valid ? *s : 0.f; }
static inline void cp_async16_zfill(void* d, const void* s, bool valid) { if (valid) memcpy(d, s, 16); else memset(d, 0, 16); }
static inline void cp_async_commit() {}
template <int N> static inline void cp_async_wait() {}
#endif


// =========================================================================================
// Lane-mapped kernels for the few -> 32-channel layers.  The 32-channel side is spread
// over the lanes of a warp, so every global access of that side is one coalesced 128-byte line
// and no shared-memory bank is hit twice:
//   *_lane_co : few -> many.  lane = output channel; its 9*CI weights live in registers; the
//               few-channel input rows are staged in shared memory (zero padded) and read as
//               warp-broadcast float4.
// (The mirrored many -> few mapping, lane = input channel with a warp butterfly per pixel, was
// measured slower than the pixel-pair kernels below on B200 and is not kept.)
// =========================================================================================
// ---- Conv2D s2, CI in {3,4,5,8} -> Co = 32*k.  block = one output row, warp = groups of 4 pixels
template <int CI>
__global__ void __launch_bounds__(256) conv_s2_lane_co_kernel(ConvArgs a, int rows, int epi) {
  KC_DYN_SMEM(float, srow);
  constexpr int NV = (9 * CI + 3) / 4;           // float4 loads per input row per group
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int co = blockIdx.y * 32 + lane;
  float w[9][CI];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < CI; ++ci) w[t][ci] = __ldg(a.w + (int64_t)t * CI * a.Co + (int64_t)ci * a.w_sci + (int64_t)co * a.w_sco);
  const float bias = (a.bias && epi != EPI_MASK) ? __ldg(a.bias + co) : 0.f;
  const int Wo4 = (a.Wo + 3) & ~3;
  const int ncolf = (2 * Wo4 + 1) * CI;           // staged floats per row: column c holds ix = c - pad_l
  const int RW = ((ncolf + 3) & ~3) + 4;
  const int shift = a.pad_l * CI, rowf = a.Wi * CI;
  const bool masked = epi == EPI_MASK;
  auto stage = [&](int row, int sidx) {   // three zero-padded input rows of output row `row`, asynchronously
    const int n = row / a.Ho, oy = row % a.Ho;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = 2 * oy + kh - a.pad_t;
      const bool vrow = iy >= 0 && iy < a.Hi;
      const float* src = a.in + ((int64_t)n * a.Hi + (vrow ? iy : 0)) * rowf;
      float* dst = srow + (sidx * 3 + kh) * RW;
      for (int e = threadIdx.x; e < RW; e += blockDim.x) {
        const int g = e - shift;
        const bool ok = vrow && g >= 0 && g < rowf;
        cp_async4_zfill(dst + e, src + (ok ? g : 0), ok);
      }
    }
    cp_async_commit();
  };
  int it = 0;
  if ((int)blockIdx.x < rows) stage(blockIdx.x, 0);
  for (int row = blockIdx.x; row < rows; row += gridDim.x, ++it) {
    const int n = row / a.Ho, oy = row % a.Ho;
    const int sidx = it & 1;
    if (row + (int)gridDim.x < rows) { stage(row + gridDim.x, sidx ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const float* sbase = srow + sidx * 3 * RW;
    for (int grp = warp; grp < (Wo4 >> 2); grp += nwarps) {
      const int ox0 = grp * 4;
      const int64_t o0 = (((int64_t)n * a.Ho + oy) * a.Wo + ox0) * a.Co + co;
      float mk[4] = {1.f, 1.f, 1.f, 1.f};
      if (masked) {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (ox0 + p < a.Wo) mk[p] = __ldg(a.mask + o0 + (int64_t)p * a.Co);
      }
      float acc[4] = {bias, bias, bias, bias};
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float4* rp = reinterpret_cast<const float4*>(sbase + kh * RW + 2 * ox0 * CI);
        float v[NV * 4];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          const float4 t = rp[q];
          v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int ci = 0; ci < CI; ++ci) acc[p] = fmaf(v[(2 * p + kw) * CI + ci], w[kh * 3 + kw][ci], acc[p]);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (ox0 + p < a.Wo) {
          float y = acc[p];
          if (epi == EPI_BIAS_RELU) y = fmaxf(y, 0.f);
          else if (epi == EPI_BIAS_SIGMOID) y = 1.0f / (1.0f + expf(-y));
          else if (masked) y = mk[p] > 0.f ? y : 0.f;
          a.out[o0 + (int64_t)p * a.Co] = y;
        }
      }
    }
    __syncthreads();
  }
}

// ---- Conv2DTranspose s2 (pad 0, Ho = 2 Hi, Wo = 2 Wi), CI in {3,4,5,8} -> Co = 32*k.
// block = one input row i (output rows 2i, 2i+1), warp = groups of 4 input pixels (8 output columns)
template <int CI>
__global__ void __launch_bounds__(256) convT_s2_lane_co_kernel(ConvArgs a, int rows, int epi) {
  KC_DYN_SMEM(float, srow);
  constexpr int NV = (5 * CI + 3) / 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int co = blockIdx.y * 32 + lane;
  float w[9][CI];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < CI; ++ci) w[t][ci] = __ldg(a.w + (int64_t)t * CI * a.Co + (int64_t)ci * a.w_sci + (int64_t)co * a.w_sco);
  const float bias = (a.bias && epi != EPI_MASK) ? __ldg(a.bias + co) : 0.f;
  const int Wi4 = (a.Wi + 3) & ~3;
  const int ncolf = (Wi4 + 1) * CI;               // column c holds ix = c - 1
  const int RW = ((ncolf + 3) & ~3) + 4;
  const int rowf = a.Wi * CI;
  const bool masked = epi == EPI_MASK;
  auto stage = [&](int row, int sidx) {   // input rows i-1 and i, zero padded, asynchronously
    const int n = row / a.Hi, i = row % a.Hi;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int iy = i - 1 + r;
      const bool vrow = iy >= 0;
      const float* src = a.in + ((int64_t)n * a.Hi + (vrow ? iy : 0)) * rowf;
      float* dst = srow + (sidx * 2 + r) * RW;
      for (int e = threadIdx.x; e < RW; e += blockDim.x) {
        const int g = e - CI;
        const bool ok = vrow && g >= 0 && g < rowf;
        cp_async4_zfill(dst + e, src + (ok ? g : 0), ok);
      }
    }
    cp_async_commit();
  };
  int it = 0;
  if ((int)blockIdx.x < rows) stage(blockIdx.x, 0);
  for (int row = blockIdx.x; row < rows; row += gridDim.x, ++it) {
    const int n = row / a.Hi, i = row % a.Hi;
    const int sidx = it & 1;
    if (row + (int)gridDim.x < rows) { stage(row + gridDim.x, sidx ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const float* sbase = srow + sidx * 2 * RW;
    for (int grp = warp; grp < (Wi4 >> 2); grp += nwarps) {
      const int j0 = grp * 4;
      const int64_t o0 = (((int64_t)n * a.Ho + 2 * i) * a.Wo + 2 * j0) * a.Co + co;
      const int64_t orow = (int64_t)a.Wo * a.Co;
      float mk[2][8];
      if (masked) {      // issued before the arithmetic so the loads overlap it
#pragma unroll
        for (int pa = 0; pa < 2; ++pa)
#pragma unroll
          for (int c = 0; c < 8; ++c) mk[pa][c] = (2 * j0 + c < a.Wo) ? __ldg(a.mask + o0 + pa * orow + (int64_t)c * a.Co) : 0.f;
      }
      float acc[2][8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { acc[0][c] = bias; acc[1][c] = bias; }
      float v[NV * 4];
      {  // input row i-1: taps kh = 2 land on output row 2i
        const float4* rp = reinterpret_cast<const float4*>(sbase + j0 * CI);
#pragma unroll
        for (int q = 0; q < NV; ++q) { const float4 t = rp[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int ci = 0; ci < CI; ++ci) {
            const float xc = v[(q + 1) * CI + ci], xl = v[q * CI + ci];   // input columns j0+q and j0+q-1
            acc[0][2 * q] = fmaf(xc, w[6][ci], fmaf(xl, w[8][ci], acc[0][2 * q]));
            acc[0][2 * q + 1] = fmaf(xc, w[7][ci], acc[0][2 * q + 1]);
          }
      }
      {  // input row i: kh = 0 -> output row 2i, kh = 1 -> output row 2i+1
        const float4* rp = reinterpret_cast<const float4*>(sbase + RW + j0 * CI);
#pragma unroll
        for (int q = 0; q < NV; ++q) { const float4 t = rp[q]; v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int ci = 0; ci < CI; ++ci) {
            const float xc = v[(q + 1) * CI + ci], xl = v[q * CI + ci];
            acc[0][2 * q] = fmaf(xc, w[0][ci], fmaf(xl, w[2][ci], acc[0][2 * q]));
            acc[0][2 * q + 1] = fmaf(xc, w[1][ci], acc[0][2 * q + 1]);
            acc[1][2 * q] = fmaf(xc, w[3][ci], fmaf(xl, w[5][ci], acc[1][2 * q]));
            acc[1][2 * q + 1] = fmaf(xc, w[4][ci], acc[1][2 * q + 1]);
          }
      }
#pragma unroll
      for (int pa = 0; pa < 2; ++pa)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (2 * j0 + c < a.Wo) {
            float y = acc[pa][c];
            if (epi == EPI_BIAS_RELU) y = fmaxf(y, 0.f);
            else if (epi == EPI_BIAS_SIGMOID) y = 1.0f / (1.0f + expf(-y));
            else if (masked) y = mk[pa][c] > 0.f ? y : 0.f;
            a.out[o0 + pa * orow + (int64_t)c * a.Co] = y;
          }
        }
    }
    __syncthreads();
  }
}

// =========================================================================================
// 32 -> few channel layers (encoder Conv2D #1, first decoder Conv2DTranspose): K = 288 per
// output pixel, only CO <= 8 outputs.  Block = two output (resp. input) rows; the input rows are
// staged once in shared memory with 16-byte async copies (coalesced 128-byte pixels) under an
// XOR swizzle of the 16-byte chunk index, so the per-pixel float4 reads of a warp hit 32 distinct
// banks.  The 32 input channels are split over four thread groups (8 channels each) whose
// partial sums are folded through shared memory in a fixed order; weights are warp-broadcast
// float4 reads of a [tap][ci][co] image.
// =========================================================================================
constexpr int M2F_SLOTS = 96;                  // pixel-pair slots per channel group
constexpr int M2F_THREADS = 4 * M2F_SLOTS;

template <int CO>
__device__ __forceinline__ void m2f_stage_weights(const ConvArgs& a, float* ws) {
  for (int i = threadIdx.x; i < 9 * 32 * CO; i += blockDim.x) {
    const int co = i % CO, ci = (i / CO) % 32, tap = i / (CO * 32);
    ws[i] = __ldg(a.w + (int64_t)tap * 32 * CO + (int64_t)ci * a.w_sci + (int64_t)co * a.w_sco);
  }
}
__device__ __forceinline__ float m2f_epi(float v, int epi, float mk) {
  if (epi == EPI_BIAS_RELU) v = fmaxf(v, 0.f);
  else if (epi == EPI_BIAS_SIGMOID) v = 1.0f / (1.0f + expf(-v));
  else if (epi == EPI_MASK) v = mk > 0.f ? v : 0.f;
  return v;
}

// ---- Conv2D s2, Ci = 32 -> CO.  task = (image, pair of output rows); thread = 2 adjacent output pixels
template <int CO>
__global__ void __launch_bounds__(M2F_THREADS) conv_s2_m2f_kernel(ConvArgs a, int n_tasks, int HB, int NCOL, int PPR, int epi) {
  KC_DYN_SMEM(float, sm);
  float* tile = sm;                               // [5 rows][NCOL][32], chunk-swizzled; reused for the fold
  float* ws = sm + 5 * NCOL * 32;                 // [9][32][CO]
  const int tid = threadIdx.x;
  m2f_stage_weights<CO>(a, ws);
  const int p = tid % M2F_SLOTS, cg = tid / M2F_SLOTS;
  const int rr = p / PPR, pp = p % PPR;
  const bool pvalid = p < 2 * PPR;
  const bool masked = epi == EPI_MASK;
  for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
    const int n = task / HB, oy0 = 2 * (task % HB);
    __syncthreads();
    const int chunks = 5 * NCOL * 8;
    for (int idx = tid; idx < chunks; idx += blockDim.x) {
      const int c4 = idx & 7, pc = idx >> 3;
      const int r = pc / NCOL, cs = pc - r * NCOL;
      const int iy = 2 * oy0 + r - a.pad_t, ix = cs - a.pad_l;
      const bool ok = iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi;
      const float* src = a.in + ((((int64_t)n * a.Hi + (ok ? iy : 0)) * a.Wi) + (ok ? ix : 0)) * 32 + c4 * 4;
      cp_async16_zfill(tile + (r * NCOL + cs) * 32 + ((c4 ^ ((cs >> 2) & 7)) << 2), src, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int oy = oy0 + rr;
    const bool live = pvalid && oy < a.Ho;
    float acc[2][CO];
#pragma unroll
    for (int co = 0; co < CO; ++co) { acc[0][co] = 0.f; acc[1][co] = 0.f; }
    if (live) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float* rowp = tile + (2 * rr + kh) * NCOL * 32;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int c4 = 2 * cg + q;
          float x[5][4];
#pragma unroll
          for (int k = 0; k < 5; ++k) {
            const int cs = 4 * pp + k;
            const float4 t = *reinterpret_cast<const float4*>(rowp + cs * 32 + ((c4 ^ ((cs >> 2) & 7)) << 2));
            x[k][0] = t.x; x[k][1] = t.y; x[k][2] = t.z; x[k][3] = t.w;
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4* wp = reinterpret_cast<const float4*>(ws + ((kh * 3 + kw) * 8 + c4) * 4 * CO);
            float wv[4 * CO];
#pragma unroll
            for (int u = 0; u < CO; ++u) { const float4 t = wp[u]; wv[4 * u] = t.x; wv[4 * u + 1] = t.y; wv[4 * u + 2] = t.z; wv[4 * u + 3] = t.w; }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int co = 0; co < CO; ++co) {
                acc[0][co] = fmaf(x[kw][c], wv[c * CO + co], acc[0][co]);
                acc[1][co] = fmaf(x[kw + 2][c], wv[c * CO + co], acc[1][co]);
              }
          }
        }
      }
    }
    __syncthreads();                               // tile reads done: reuse it as [4 groups][2*CO values][M2F_SLOTS]
    float* red = tile;
    constexpr int NV = 2 * CO;
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int co = 0; co < CO; ++co) red[(cg * NV + px * CO + co) * M2F_SLOTS + p] = acc[px][co];
    __syncthreads();
    if (live) {   // the four groups share the final summation: group cg folds values k = cg, cg+4, ...
      const int64_t o0 = (((int64_t)n * a.Ho + oy) * a.Wo + 2 * pp) * CO;
      float mk[(NV + 3) / 4];
#pragma unroll
      for (int u = 0; u < (NV + 3) / 4; ++u) {
        const int k = cg + 4 * u;
        mk[u] = (masked && k < NV && 2 * pp + k / CO < a.Wo) ? __ldg(a.mask + o0 + k) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < (NV + 3) / 4; ++u) {
        const int k = cg + 4 * u;
        if (k < NV && 2 * pp + k / CO < a.Wo) {
          float v = (a.bias && !masked) ? __ldg(a.bias + k % CO) : 0.f;
#pragma unroll
          for (int g = 0; g < 4; ++g) v += red[(g * NV + k) * M2F_SLOTS + p];
          a.out[o0 + k] = m2f_epi(v, epi, mk[u]);
        }
      }
    }
  }
}

// ---- Conv2DTranspose s2 (pad 0, Ho = 2 Hi, Wo = 2 Wi), Ci = 32 -> CO.  task = (image, pair of input
// rows); thread = 2 adjacent input pixels, each producing its 2x2 output quad (9 taps)
template <int CO>
__global__ void __launch_bounds__(M2F_THREADS) convT_s2_m2f_kernel(ConvArgs a, int n_tasks, int HB, int NCOL, int PPR, int epi,
                                                                   int tile_floats) {
  KC_DYN_SMEM(float, sm);
  float* tile = sm;                               // [3 rows][NCOL][32], column c holds ix = c - 1; reused for the fold
  float* ws = sm + tile_floats;
  const int tid = threadIdx.x;
  m2f_stage_weights<CO>(a, ws);
  const int p = tid % M2F_SLOTS, cg = tid / M2F_SLOTS;
  const int rr = p / PPR, pp = p % PPR;
  const bool pvalid = p < 2 * PPR;
  const bool masked = epi == EPI_MASK;
  // tap -> (output parity phase, which of the four input pixels): x00 = (i,j), x01 = (i,j-1), x10 = (i-1,j), x11 = (i-1,j-1)
  for (int task = blockIdx.x; task < n_tasks; task += gridDim.x) {
    const int n = task / HB, i0 = 2 * (task % HB);
    __syncthreads();
    const int chunks = 3 * NCOL * 8;
    for (int idx = tid; idx < chunks; idx += blockDim.x) {
      const int c4 = idx & 7, pc = idx >> 3;
      const int r = pc / NCOL, cs = pc - r * NCOL;
      const int iy = i0 - 1 + r, ix = cs - 1;
      const bool ok = iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi;
      const float* src = a.in + ((((int64_t)n * a.Hi + (ok ? iy : 0)) * a.Wi) + (ok ? ix : 0)) * 32 + c4 * 4;
      cp_async16_zfill(tile + (r * NCOL + cs) * 32 + ((c4 ^ ((cs >> 1) & 7)) << 2), src, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int i = i0 + rr;
    const bool live = pvalid && i < a.Hi;
    float acc[2][4][CO];
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int ph = 0; ph < 4; ++ph)
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[px][ph][co] = 0.f;
    if (live) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c4 = 2 * cg + q;
        float up[3][4], cur[3][4];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int cs = 2 * pp + k;
          const int off = cs * 32 + ((c4 ^ ((cs >> 1) & 7)) << 2);
          const float4 tu = *reinterpret_cast<const float4*>(tile + rr * NCOL * 32 + off);
          const float4 tc = *reinterpret_cast<const float4*>(tile + (rr + 1) * NCOL * 32 + off);
          up[k][0] = tu.x; up[k][1] = tu.y; up[k][2] = tu.z; up[k][3] = tu.w;
          cur[k][0] = tc.x; cur[k][1] = tc.y; cur[k][2] = tc.z; cur[k][3] = tc.w;
        }
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int kh = tap / 3, kw = tap % 3;
          const int ph = (kh & 1) * 2 + (kw & 1);          // output parity phase this tap lands on
          const bool use_up = kh == 2, use_left = kw == 2;  // input pixel (i - kh/2, j - kw/2)
          const float4* wp = reinterpret_cast<const float4*>(ws + (tap * 8 + c4) * 4 * CO);
          float wv[4 * CO];
#pragma unroll
          for (int u = 0; u < CO; ++u) { const float4 t = wp[u]; wv[4 * u] = t.x; wv[4 * u + 1] = t.y; wv[4 * u + 2] = t.z; wv[4 * u + 3] = t.w; }
#pragma unroll
          for (int px = 0; px < 2; ++px) {
            const int k = px + (use_left ? 0 : 1);          // staged column of the input pixel, relative to 2*pp
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float xv = use_up ? up[k][c] : cur[k][c];
#pragma unroll
              for (int co = 0; co < CO; ++co) acc[px][ph][co] = fmaf(xv, wv[c * CO + co], acc[px][ph][co]);
            }
          }
        }
      }
    }
    __syncthreads();                               // tile reads done: reuse it as [4 groups][8*CO values][M2F_SLOTS]
    float* red = tile;
    constexpr int NV = 8 * CO;                     // value k = (px*4 + ph)*CO + co
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int ph = 0; ph < 4; ++ph)
#pragma unroll
        for (int co = 0; co < CO; ++co) red[(cg * NV + (px * 4 + ph) * CO + co) * M2F_SLOTS + p] = acc[px][ph][co];
    __syncthreads();
    if (live) {   // group cg folds the output pixel-phases (px, ph) with (px*4+ph) % 4 == cg
      int64_t oo[2];
      bool ok2[2];
      float mk[2][CO];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pq = cg + 4 * u, px = pq >> 2, ph = pq & 3;
        ok2[u] = 2 * pp + px < a.Wi;
        oo[u] = (((int64_t)n * a.Ho + 2 * i + (ph >> 1)) * a.Wo + 2 * (2 * pp + px) + (ph & 1)) * CO;
#pragma unroll
        for (int co = 0; co < CO; ++co) mk[u][co] = (masked && ok2[u]) ? __ldg(a.mask + oo[u] + co) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (!ok2[u]) continue;
        const int pq = cg + 4 * u;
#pragma unroll
        for (int co = 0; co < CO; ++co) {
          float v = (a.bias && !masked) ? __ldg(a.bias + co) : 0.f;
#pragma unroll
          for (int g = 0; g < 4; ++g) v += red[(g * NV + pq * CO + co) * M2F_SLOTS + p];
          a.out[oo[u] + co] = m2f_epi(v, epi, mk[u][co]);
        }
      }
    }
  }
}

#ifndef KCVAE_EMU
#define KC_SET_SMEM(k, bytes) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#else
#define KC_SET_SMEM(k, bytes)
#endif
#define KC_EPI_SWITCH(KERNEL, ...)                                                               \
  switch (epi) {                                                                                 \
    case EPI_BIAS: { auto k = KERNEL<EPI_BIAS>; KC_SET_SMEM(k, smem); KC_LAUNCH(k, grid, 128, smem, st, __VA_ARGS__); break; }           \
    case EPI_BIAS_RELU: { auto k = KERNEL<EPI_BIAS_RELU>; KC_SET_SMEM(k, smem); KC_LAUNCH(k, grid, 128, smem, st, __VA_ARGS__); break; } \
    case EPI_BIAS_SIGMOID: { auto k = KERNEL<EPI_BIAS_SIGMOID>; KC_SET_SMEM(k, smem); KC_LAUNCH(k, grid, 128, smem, st, __VA_ARGS__); break; } \
    default: { auto k = KERNEL<EPI_MASK>; KC_SET_SMEM(k, smem); KC_LAUNCH(k, grid, 128, smem, st, __VA_ARGS__); break; }                 \
  }

// returns true if a specialised kernel took the launch
// lane-mapped kernels: true if one of them took the launch
static bool conv_forward_lane(int mode, int epi, const ConvArgs& a, cudaStream_t st) {
  const bool up2 = a.pad_t == 0 && a.pad_l == 0 && a.Ho == 2 * a.Hi && a.Wo == 2 * a.Wi;
  if (a.Co % 32 == 0 && (a.Ci == 3 || a.Ci == 4 || a.Ci == 5 || a.Ci == 8) && (mode == CONV_S2 || (mode == CONVT_S2 && up2))) {
    const int rows = mode == CONV_S2 ? a.B * a.Ho : a.B * a.Hi;
    const int Wq4 = ((mode == CONV_S2 ? a.Wo : a.Wi) + 3) & ~3;
    const int ncolf = (mode == CONV_S2 ? 2 * Wq4 + 1 : Wq4 + 1) * a.Ci;
    const int RW = ((ncolf + 3) & ~3) + 4;
    const size_t smem = (size_t)2 * (mode == CONV_S2 ? 3 : 2) * RW * sizeof(float);   // two stages
    if (smem > 160 * 1024) return false;
    const dim3 grid(rows < kNumSMs * 4 ? rows : kNumSMs * 4, a.Co / 32);
    ++g_launches;
#define KC_LANE_CO(KERNEL, CI_)                                 \
  {                                                             \
    auto k = KERNEL<CI_>;                                       \
    if (smem > 48 * 1024) KC_SET_SMEM(k, smem);                 \
    KC_LAUNCH(k, grid, 256, smem, st, a, rows, epi);            \
  }
    if (mode == CONV_S2) {
      if (a.Ci == 3) KC_LANE_CO(conv_s2_lane_co_kernel, 3)
      else if (a.Ci == 4) KC_LANE_CO(conv_s2_lane_co_kernel, 4)
      else if (a.Ci == 5) KC_LANE_CO(conv_s2_lane_co_kernel, 5)
      else KC_LANE_CO(conv_s2_lane_co_kernel, 8)
    } else {
      if (a.Ci == 3) KC_LANE_CO(convT_s2_lane_co_kernel, 3)
      else if (a.Ci == 4) KC_LANE_CO(convT_s2_lane_co_kernel, 4)
      else if (a.Ci == 5) KC_LANE_CO(convT_s2_lane_co_kernel, 5)
      else KC_LANE_CO(convT_s2_lane_co_kernel, 8)
    }
#undef KC_LANE_CO
    return true;
  }
  if (a.Ci == 32 && (a.Co == 5 || a.Co == 8) && ((uintptr_t)a.in % 16 == 0)) {
    if (mode == CONV_S2) {
      const int PPR = cdiv(a.Wo, 2), NCOL = 4 * PPR + 1, HB = cdiv(a.Ho, 2);
      const size_t smem = ((size_t)5 * NCOL * 32 + (size_t)9 * 32 * a.Co) * sizeof(float);
      if (2 * PPR <= M2F_SLOTS && smem <= 110 * 1024 && (size_t)4 * M2F_SLOTS * 2 * a.Co <= (size_t)5 * NCOL * 32) {
        const int n_tasks = a.B * HB;
        const int grid = n_tasks < kNumSMs * 2 ? n_tasks : kNumSMs * 2;
        ++g_launches;
        if (a.Co == 5) { auto k = conv_s2_m2f_kernel<5>; KC_SET_SMEM(k, smem); KC_LAUNCH(k, grid, M2F_THREADS, smem, st, a, n_tasks, HB, NCOL, PPR, epi); }
        else { auto k = conv_s2_m2f_kernel<8>; KC_SET_SMEM(k, smem); KC_LAUNCH(k, grid, M2F_THREADS, smem, st, a, n_tasks, HB, NCOL, PPR, epi); }
        return true;
      }
    }
    if (mode == CONVT_S2 && up2) {
      const int PPR = cdiv(a.Wi, 2), NCOL = 2 * PPR + 1, HB = cdiv(a.Hi, 2);
      size_t tile_floats = (size_t)3 * NCOL * 32;
      const size_t fold = (size_t)4 * M2F_SLOTS * 2 * 4 * a.Co;
      if (fold > tile_floats) tile_floats = fold;
      const size_t smem = (tile_floats + (size_t)9 * 32 * a.Co) * sizeof(float);
      if (2 * PPR <= M2F_SLOTS && smem <= 110 * 1024) {
        const int n_tasks = a.B * HB;
        const int grid = n_tasks < kNumSMs * 2 ? n_tasks : kNumSMs * 2;
        ++g_launches;
        if (a.Co == 5) { auto k = convT_s2_m2f_kernel<5>; KC_SET_SMEM(k, smem); KC_LAUNCH(k, grid, M2F_THREADS, smem, st, a, n_tasks, HB, NCOL, PPR, epi, (int)tile_floats); }
        else { auto k = convT_s2_m2f_kernel<8>; KC_SET_SMEM(k, smem); KC_LAUNCH(k, grid, M2F_THREADS, smem, st, a, n_tasks, HB, NCOL, PPR, epi, (int)tile_floats); }
        return true;
      }
    }
  }
  return false;
}

static bool conv_forward_special(int mode, int epi, const ConvArgs& a, cudaStream_t st) {
  if (a.B <= 0) return false;
  if (conv_forward_lane(mode, epi, a, st)) return true;
  const bool aligned = ((uintptr_t)a.in % 16 == 0) && ((uintptr_t)a.out % 16 == 0) && (!a.mask || (uintptr_t)a.mask % 16 == 0);
  if (!aligned) return false;
  if (a.Co == MANY && a.Ci <= 8) {
    const size_t smem = (size_t)9 * a.Ci * MANY * sizeof(float);
    if (mode == CONV_S2) {
      const int WP = cdiv(a.Wo, 2);
      const int64_t npairs = (int64_t)a.B * a.Ho * WP;
      const int grid = grid_for(npairs, 128, 4, 16);
      ++g_launches;
      KC_EPI_SWITCH(conv_s2_few2many_kernel, a, npairs, WP)
      return true;
    }
    if (mode == CONVT_S2) {
      const int HQ = cdiv(a.Ho, 2), WQ = cdiv(a.Wo, 2);
      const int64_t nq = (int64_t)a.B * HQ * WQ;
      const int grid = grid_for(nq, 32, 4, 16);
      ++g_launches;
      KC_EPI_SWITCH(convT_s2_few2many_kernel, a, nq, HQ, WQ)
      return true;
    }
  }
  if (a.Co <= FEW && a.Ci % 4 == 0 && a.Ci <= 128) {
    const size_t smem = (size_t)9 * a.Ci * FEW * sizeof(float);
    if (mode == CONV_S2) {
      const int WP = cdiv(a.Wo, 2);
      const int64_t npairs = (int64_t)a.B * a.Ho * WP;
      const int grid = grid_for(npairs, 128, 8, 16);
      ++g_launches;
      KC_EPI_SWITCH(conv_s2_many2few_kernel, a, npairs, WP)
      return true;
    }
    if (mode == CONVT_S2 && a.pad_t == 0 && a.pad_l == 0 && a.Ho == 2 * a.Hi && a.Wo == 2 * a.Wi) {
      const int64_t nq = (int64_t)a.B * a.Hi * a.Wi;
      const int grid = grid_for(nq, 128, 4, 16);
      ++g_launches;
      KC_EPI_SWITCH(convT_s2_many2few_kernel, a, nq)
      return true;
    }
  }
  return false;
}

template <int MODE>
static void launch_mode(int epi, const ConvArgs& a, cudaStream_t st) {
  const int WG = cdiv(a.Wo, PT);
  const int64_t total = (int64_t)a.B * a.Ho * WG * a.Co;
  if (total <= 0) return;
  const int grid = grid_for(total, 256, 8, 8);
  ++g_launches;
  switch (epi) {
    case EPI_BIAS: { auto k = conv3x3_kernel<MODE, EPI_BIAS>; KC_LAUNCH(k, grid, 256, 0, st, a, total, WG); break; }
    case EPI_BIAS_RELU: { auto k = conv3x3_kernel<MODE, EPI_BIAS_RELU>; KC_LAUNCH(k, grid, 256, 0, st, a, total, WG); break; }
    case EPI_BIAS_SIGMOID: { auto k = conv3x3_kernel<MODE, EPI_BIAS_SIGMOID>; KC_LAUNCH(k, grid, 256, 0, st, a, total, WG); break; }
    default: { auto k = conv3x3_kernel<MODE, EPI_MASK>; KC_LAUNCH(k, grid, 256, 0, st, a, total, WG); break; }
  }
}

void conv_forward(int mode, int epi, const ConvArgs& a, cudaStream_t st) {
  ProfScope prof_("conv3x3", st);
  if (conv_forward_special(mode, epi, a, st)) return;
  if (mode == CONV_S2) launch_mode<CONV_S2>(epi, a, st);
  else if (mode == CONV_S1) launch_mode<CONV_S1>(epi, a, st);
  else launch_mode<CONVT_S2>(epi, a, st);
}

// ------------------------------------------------------------------------- weight grads
// thread = one dW entry (tap, a, b), b fastest; blockIdx.y = chunk of P rows (n,i).
// Deterministic two-level reduction: per-chunk partials, then wgrad_reduce_kernel.
static int wgrad_chunks(int B, int Hp, int Ca, int Cb) {
  const int E = 9 * Ca * Cb;
  const int eb = cdiv(E, 256);
  int64_t rows = (int64_t)B * Hp;
  int64_t want = (int64_t)kNumSMs * 8 / eb;
  if (want < 1) want = 1;
  if (want > rows) want = rows;
  if (want > 4096) want = 4096;
  return (int)want;
}
size_t wgrad_partial_floats(int B, int Hp, int Ca, int Cb) {
  size_t n = (size_t)wgrad_chunks(B, Hp, Ca, Cb);
  if (n < (size_t)kNumSMs * 4) n = (size_t)kNumSMs * 4;   // the tiled kernel uses <= 4 blocks per SM
  return n * 9 * Ca * Cb;
}

__global__ void __launch_bounds__(256) wgrad_kernel(WgradArgs a, int E, int rows, int rows_per_chunk) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int cb = e % a.Cb;
  const int ca = (e / a.Cb) % a.Ca;
  const int tap = e / (a.Cb * a.Ca);
  const int kh = tap / 3, kw = tap % 3;
  const int r0 = blockIdx.y * rows_per_chunk;
  const int r1 = min(rows, r0 + rows_per_chunk);
  // columns j with 0 <= s*j + d*kw + ox < Wq
  float acc = 0.0f;
  for (int r = r0; r < r1; ++r) {
    const int n = r / a.Hp, i = r % a.Hp;
    const int qy = a.s * i + a.d * kh + a.oy;
    if (qy < 0 || qy >= a.Hq) continue;
    const float* prow = a.P + ((int64_t)r * a.Wp) * a.Ca + ca;
    const float* qrow = a.Q + (((int64_t)n * a.Hq + qy) * a.Wq) * a.Cb + cb;
    const int qx0 = a.d * kw + a.ox;
    for (int j = 0; j < a.Wp; ++j) {
      const int qx = a.s * j + qx0;
      if (qx < 0 || qx >= a.Wq) continue;
      acc = fmaf(__ldg(prow + (int64_t)j * a.Ca), __ldg(qrow + (int64_t)qx * a.Cb), acc);
    }
  }
  a.partial[(int64_t)blockIdx.y * E + e] = acc;
}

__global__ void wgrad_reduce_kernel(const float* partial, int chunks, int E, int Ca, int Cb,
                                    int o_sa, int o_sb, float* out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float s = 0.0f;
  for (int c = 0; c < chunks; ++c) s += partial[(int64_t)c * E + e];
  const int cb = e % Cb;
  const int ca = (e / Cb) % Ca;
  const int tap = e / (Cb * Ca);
  out[(int64_t)tap * Ca * Cb + (int64_t)ca * o_sa + (int64_t)cb * o_sb] = s;
}


// ------------------------------------------------------------ weight grads, tiled version
// Block = row chunk.  For every P row and SEG-pixel segment the P pixels and the three Q rows
// they touch are staged in shared memory (coalesced); a thread owns an RA x RB block of dW
// entries of one tap and one pixel phase, so each shared-memory value feeds RA (or RB) FMAs.
// Phases are folded through shared memory at the end; blocks write partials that
// wgrad_reduce_kernel sums (deterministic).
template <int RA, int RB>
__global__ void __launch_bounds__(256) wgrad_tiled_kernel(WgradArgs a, int rows, int rows_per_chunk, int SEG,
                                                          int n_at, int n_bt, int nph, int cbp) {
  KC_DYN_SMEM(float, sm);
  const int E = 9 * a.Ca * a.Cb;
  const int n_et = 9 * n_at * n_bt;                 // entry tiles
  const int qw = a.s * (SEG - 1) + 3;               // Q pixels per staged row
  float* Ps = sm;                                   // [SEG][Ca]
  float* Qs = sm + ((SEG * a.Ca + 3) & ~3);         // [3][qw][cbp], 16-byte aligned
  const int tid = threadIdx.x;
  const bool active = tid < n_et * nph;
  const int et = active ? tid % n_et : 0, ph = active ? tid / n_et : 0;
  const int bt = et % n_bt, at = (et / n_bt) % n_at, tap = et / (n_bt * n_at);
  const int kh = tap / 3, kw = tap % 3;
  const int a0 = at * RA, b0 = bt * RB;
  const int dmin = a.d < 0 ? 2 * a.d : 0;           // smallest d*kw
  float acc[RA][RB];
#pragma unroll
  for (int i = 0; i < RA; ++i)
#pragma unroll
    for (int j = 0; j < RB; ++j) acc[i][j] = 0.f;

  const int r0 = blockIdx.x * rows_per_chunk;
  const int r1 = min(rows, r0 + rows_per_chunk);
  for (int r = r0; r < r1; ++r) {
    const int n = r / a.Hp, i = r % a.Hp;
    for (int j0 = 0; j0 < a.Wp; j0 += SEG) {
      const int seg = min(SEG, a.Wp - j0);
      const int qx_min = a.s * j0 + dmin + a.ox;
      __syncthreads();
      {  // stage P segment (contiguous in global and in shared memory)
        const float* src = a.P + (((int64_t)r * a.Wp) + j0) * a.Ca;
        const int np = seg * a.Ca;
        if ((np & 3) == 0 && (((uintptr_t)src) & 15) == 0) {
          const float4* s4 = reinterpret_cast<const float4*>(src);
          float4* d4 = reinterpret_cast<float4*>(Ps);
          for (int t = tid; t < (np >> 2); t += blockDim.x) d4[t] = __ldg(s4 + t);
        } else {
          for (int t = tid; t < np; t += blockDim.x) Ps[t] = __ldg(src + t);
        }
        for (int t = np + tid; t < SEG * a.Ca; t += blockDim.x) Ps[t] = 0.f;
      }
      if ((a.Cb & 3) == 0 && (cbp & 3) == 0 && (((uintptr_t)a.Q) & 15) == 0) {
        // vectorised staging of the 3 Q rows: one float4 = 4 channels of a pixel
        const int u_per_px = a.Cb >> 2, cbp4 = cbp >> 2;
        const int step_px = blockDim.x / u_per_px, step_u = blockDim.x % u_per_px;
        for (int k = 0; k < 3; ++k) {
          const int qy = a.s * i + a.d * k + a.oy;
          const bool vrow = qy >= 0 && qy < a.Hq;
          const float4* src = reinterpret_cast<const float4*>(a.Q + (((int64_t)n * a.Hq + (vrow ? qy : 0)) * a.Wq) * a.Cb);
          float4* dst = reinterpret_cast<float4*>(Qs + (int64_t)k * qw * cbp);
          int px = tid / u_per_px, u = tid % u_per_px;
          while (px < qw) {
            const int qx = qx_min + px;
            dst[px * cbp4 + u] = (vrow && qx >= 0 && qx < a.Wq) ? __ldg(src + (int64_t)qx * u_per_px + u) : make_float4(0.f, 0.f, 0.f, 0.f);
            px += step_px; u += step_u;
            if (u >= u_per_px) { u -= u_per_px; ++px; }
          }
        }
      } else {  // scalar staging; (px, c) advanced incrementally, no div/mod
        const int step_px = blockDim.x / a.Cb, step_c = blockDim.x % a.Cb;
        for (int k = 0; k < 3; ++k) {
          const int qy = a.s * i + a.d * k + a.oy;
          const bool vrow = qy >= 0 && qy < a.Hq;
          const float* src = a.Q + (((int64_t)n * a.Hq + (vrow ? qy : 0)) * a.Wq) * a.Cb;
          float* dst = Qs + (int64_t)k * qw * cbp;
          int px = tid / a.Cb, c = tid % a.Cb;
          while (px < qw) {
            const int qx = qx_min + px;
            dst[px * cbp + c] = (vrow && qx >= 0 && qx < a.Wq) ? __ldg(src + (int64_t)qx * a.Cb + c) : 0.f;
            px += step_px; c += step_c;
            if (c >= a.Cb) { c -= a.Cb; ++px; }
          }
        }
      }
      __syncthreads();
      if (active) {
        const float* qrow = Qs + (int64_t)kh * qw * cbp + (a.d * kw - dmin) * cbp + b0;
        const bool pvec = (RA % 4 == 0) && ((a.Ca & 3) == 0) && (a0 + RA <= a.Ca);
        const bool qvec = (RB % 4 == 0) && ((cbp & 3) == 0) && (b0 + RB <= a.Cb);
#pragma unroll 2
        for (int j = ph; j < seg; j += nph) {
          float pv[RA], qv[RB];
          const float* pp = Ps + j * a.Ca + a0;
          if (pvec) {
#pragma unroll
            for (int x = 0; x < RA / 4; ++x) {
              const float4 v = reinterpret_cast<const float4*>(pp)[x];
              pv[x * 4] = v.x; pv[x * 4 + 1] = v.y; pv[x * 4 + 2] = v.z; pv[x * 4 + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int x = 0; x < RA; ++x) pv[x] = (a0 + x < a.Ca) ? pp[x] : 0.f;
          }
          const float* qp = qrow + (int64_t)(a.s * j) * cbp;
          if (qvec) {
#pragma unroll
            for (int y = 0; y < RB / 4; ++y) {
              const float4 v = reinterpret_cast<const float4*>(qp)[y];
              qv[y * 4] = v.x; qv[y * 4 + 1] = v.y; qv[y * 4 + 2] = v.z; qv[y * 4 + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int y = 0; y < RB; ++y) qv[y] = (b0 + y < a.Cb) ? qp[y] : 0.f;
          }
#pragma unroll
          for (int x = 0; x < RA; ++x)
#pragma unroll
            for (int y = 0; y < RB; ++y) acc[x][y] = fmaf(pv[x], qv[y], acc[x][y]);
        }
      }
    }
  }
  // fold the pixel phases: sm reused as [nph][E]
  __syncthreads();
  if (active) {
#pragma unroll
    for (int x = 0; x < RA; ++x)
#pragma unroll
      for (int y = 0; y < RB; ++y)
        if (a0 + x < a.Ca && b0 + y < a.Cb) sm[(int64_t)ph * E + (tap * a.Ca + a0 + x) * a.Cb + b0 + y] = acc[x][y];
  }
  __syncthreads();
  for (int e = tid; e < E; e += blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < nph; ++p) s += sm[(int64_t)p * E + e];
    a.partial[(int64_t)blockIdx.x * E + e] = s;
  }
}


// ------------------------------------------------- weight grads, row-streaming version
// Same thread tiling as wgrad_tiled_kernel, but the operands arrive through a two-stage
// cp.async pipeline: while the block accumulates P row r against the three Q rows it touches,
// row r+1 is already in flight, so the global-load latency that dominated the staged kernel is
// hidden.  Each row (and the contiguous 3-row Q span) is one linear 16-byte-chunk copy; image
// borders are handled by clipping each thread's pixel range for its tap, not by padding.
// Optionally the same pass produces sum_pixels P[., a] (the bias gradient when P is a gradient).
// elements [k_lo, k_hi) of the float array starting at element `base` of `src` -> dst[mis + k],
// mis = address misalignment (in floats) of element `base`, so 16-byte chunks line up
__device__ __forceinline__ int stage_linear(float* dst, const float* src, int64_t base, int k_lo, int k_hi) {
  const int mis = (int)(((reinterpret_cast<uintptr_t>(src) >> 2) + (uint64_t)base) & 3);
  int head_end = k_lo + ((4 - ((mis + k_lo) & 3)) & 3);
  if (head_end > k_hi) head_end = k_hi;
  const int body = (k_hi - head_end) >> 2;
  const int tail0 = head_end + (body << 2);
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int t = tid; t < body; t += nt) cp_async16(dst + mis + head_end + 4 * t, src + base + head_end + 4 * t);
  if (tid < head_end - k_lo) dst[mis + k_lo + tid] = __ldg(src + base + k_lo + tid);
  if (tid >= 32 && tid - 32 < k_hi - tail0) dst[mis + tail0 + tid - 32] = __ldg(src + base + tail0 + tid - 32);
  return mis;
}
// npix pixels of C = 4*C4 channels (16-byte aligned source) -> dst[pixel * cpad + c]
__device__ __forceinline__ void stage_padded(float* dst, const float* src, int npix, int C4, int cpad) {
  const int total = npix * C4;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int px = t / C4, u = t - px * C4;
    cp_async16(dst + px * cpad + 4 * u, src + 4 * (int64_t)t);
  }
}

struct WgradRowsGeom {
  int rows, rows_per_block, n_at, n_bt, nph;
  int p_pad, q_pad;        // 1: channel-padded staging (C % 4 == 0), 0: linear
  int cap, cbp;            // staged channel strides of P and Q
  int p_floats, q_floats;  // per-stage buffer sizes (multiples of 4)
  int want_colsum, pstride;
};

template <int RA, int RB>
__global__ void __launch_bounds__(256) wgrad_rows_kernel(WgradArgs a, WgradRowsGeom gm) {
  KC_DYN_SMEM(float, sm);
  const int E = 9 * a.Ca * a.Cb;
  const int n_et = 9 * gm.n_at * gm.n_bt;
  const int tid = threadIdx.x;
  const bool active = tid < n_et * gm.nph;
  const int et = active ? tid % n_et : 0, ph = active ? tid / n_et : 0;
  const int bt = et % gm.n_bt, at = (et / gm.n_bt) % gm.n_at, tap = et / (gm.n_bt * gm.n_at);
  const int kh = tap / 3, kw = tap % 3;
  const int a0 = at * RA, b0 = bt * RB;
  const int rsel = a.d > 0 ? kh : 2 - kh;          // which row of the staged 3-row span this tap reads
  const int dlo = a.d < 0 ? 2 * a.d : 0;
  const int qoff = a.d * kw + a.ox;                // qx = s*j + qoff
  int jlo = 0, jhi = a.Wp;
  while (jlo < jhi && a.s * jlo + qoff < 0) ++jlo;
  while (jhi > jlo && a.s * (jhi - 1) + qoff >= a.Wq) --jhi;
  const bool cs_thread = gm.want_colsum && active && tap == 4 && bt == 0;
  const bool pvec = (RA % 4 == 0) && gm.p_pad && (a0 + RA <= a.Ca);
  const bool qvec = (RB % 4 == 0) && gm.q_pad && (b0 + RB <= a.Cb);
  const int stage_floats = gm.p_floats + gm.q_floats;
  const int qrow_f = a.Wq * gm.cbp;

  float acc[RA][RB];
  float bsum[RA];
#pragma unroll
  for (int i = 0; i < RA; ++i) {
    bsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < RB; ++j) acc[i][j] = 0.f;
  }

  const int r0 = blockIdx.x * gm.rows_per_block;
  const int r1 = min(gm.rows, r0 + gm.rows_per_block);
  int pmis[2] = {0, 0}, qmis[2] = {0, 0};

  auto stage = [&](int r, int sidx) {
    float* Pb = sm + sidx * stage_floats;
    float* Qb = Pb + gm.p_floats;
    const int n = r / a.Hp, i = r % a.Hp;
    const int64_t pbase = (int64_t)r * a.Wp * a.Ca;
    if (gm.p_pad) stage_padded(Pb, a.P + pbase, a.Wp, a.Ca >> 2, gm.cap);
    else pmis[sidx] = stage_linear(Pb, a.P, pbase, 0, a.Wp * a.Ca);
    const int qb = a.s * i + a.oy + dlo;            // first row of the span (may be outside the image)
    const int lo = qb < 0 ? 0 : qb, hi = qb + 2 >= a.Hq ? a.Hq - 1 : qb + 2;
    if (lo <= hi) {
      if (gm.q_pad) {
        stage_padded(Qb + (lo - qb) * qrow_f, a.Q + (((int64_t)n * a.Hq + lo) * a.Wq) * a.Cb, (hi - lo + 1) * a.Wq,
                     a.Cb >> 2, gm.cbp);
      } else {
        const int rowf = a.Wq * a.Cb;
        qmis[sidx] = stage_linear(Qb, a.Q, ((int64_t)n * a.Hq + qb) * rowf, (lo - qb) * rowf, (hi - qb + 1) * rowf);
      }
    }
    cp_async_commit();
  };

  if (r0 < r1) stage(r0, 0);
  for (int r = r0; r < r1; ++r) {
    const int sidx = (r - r0) & 1;
    if (r + 1 < r1) { stage(r + 1, sidx ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const int i = r % a.Hp;
    const int qy = a.s * i + a.d * kh + a.oy;
    if (active && qy >= 0 && qy < a.Hq) {
      const float* Pb = sm + sidx * stage_floats + (gm.p_pad ? 0 : pmis[sidx]) + a0;
      const float* Qb = sm + sidx * stage_floats + gm.p_floats + (gm.q_pad ? 0 : qmis[sidx]) + rsel * qrow_f + qoff * gm.cbp + b0;
      int j = ph;
      while (j < jlo) j += gm.nph;
#pragma unroll 2
      for (; j < jhi; j += gm.nph) {
        float pv[RA], qv[RB];
        const float* pp = Pb + j * gm.cap;
        if (pvec) {
#pragma unroll
          for (int x = 0; x < RA / 4; ++x) {
            const float4 v = reinterpret_cast<const float4*>(pp)[x];
            pv[x * 4] = v.x; pv[x * 4 + 1] = v.y; pv[x * 4 + 2] = v.z; pv[x * 4 + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int x = 0; x < RA; ++x) pv[x] = (a0 + x < a.Ca) ? pp[x] : 0.f;
        }
        const float* qp = Qb + (a.s * j) * gm.cbp;
        if (qvec) {
#pragma unroll
          for (int y = 0; y < RB / 4; ++y) {
            const float4 v = reinterpret_cast<const float4*>(qp)[y];
            qv[y * 4] = v.x; qv[y * 4 + 1] = v.y; qv[y * 4 + 2] = v.z; qv[y * 4 + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int y = 0; y < RB; ++y) qv[y] = (b0 + y < a.Cb) ? qp[y] : 0.f;
        }
#pragma unroll
        for (int x = 0; x < RA; ++x)
#pragma unroll
          for (int y = 0; y < RB; ++y) acc[x][y] = fmaf(pv[x], qv[y], acc[x][y]);
        if (cs_thread) {
#pragma unroll
          for (int x = 0; x < RA; ++x) bsum[x] += pv[x];
        }
      }
    }
    __syncthreads();
  }
  // fold the pixel phases: sm reused as [nph][E + Ca]
  const int FE = E + (gm.want_colsum ? a.Ca : 0);
  if (active) {
#pragma unroll
    for (int x = 0; x < RA; ++x)
#pragma unroll
      for (int y = 0; y < RB; ++y)
        if (a0 + x < a.Ca && b0 + y < a.Cb) sm[ph * FE + (tap * a.Ca + a0 + x) * a.Cb + b0 + y] = acc[x][y];
    if (cs_thread) {
#pragma unroll
      for (int x = 0; x < RA; ++x)
        if (a0 + x < a.Ca) sm[ph * FE + E + a0 + x] = bsum[x];
    }
  }
  __syncthreads();
  for (int e = tid; e < FE; e += blockDim.x) {
    float t = 0.f;
    for (int p = 0; p < gm.nph; ++p) t += sm[p * FE + e];
    a.partial[(int64_t)blockIdx.x * gm.pstride + e] = t;
  }
}

// out[...] = sum over `nparts` partial vectors (stride `pstride`), 8 threads per entry in a fixed order;
// entries >= E are the fused column sums of P
constexpr int WR_SLICES = 8;
__global__ void __launch_bounds__(256) wgrad_reduce_sliced_kernel(const float* __restrict__ partial, int nparts, int pstride,
                                                                  int E, int FE, int Ca, int Cb, int o_sa, int o_sb,
                                                                  float* __restrict__ out, float* __restrict__ pcolsum) {
  __shared__ float red[256];
  const int oi = threadIdx.x / WR_SLICES, sl = threadIdx.x % WR_SLICES;
  const int e = blockIdx.x * (256 / WR_SLICES) + oi;
  float t = 0.f;
  if (e < FE) {
    const int per = (nparts + WR_SLICES - 1) / WR_SLICES;
    const int c0 = sl * per, c1 = min(nparts, c0 + per);
#pragma unroll 4
    for (int c = c0; c < c1; ++c) t += __ldg(partial + (int64_t)c * pstride + e);
  }
  red[threadIdx.x] = t;
  __syncthreads();
  if (sl == 0 && e < FE) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < WR_SLICES; ++i) v += red[threadIdx.x + i];
    if (e < E) {
      const int cb = e % Cb, ca = (e / Cb) % Ca, tap = e / (Cb * Ca);
      out[(int64_t)tap * Ca * Cb + (int64_t)ca * o_sa + (int64_t)cb * o_sb] = v;
    } else {
      pcolsum[e - E] = v;
    }
  }
}

struct WgradPlan { int ra, rb, n_at, n_bt, nph, seg, cbp, blocks, rpc; size_t smem; bool ok; };
static WgradPlan wgrad_plan(int B, int Hp, int Wp, int Ca, int Cb, int s) {
  WgradPlan p{};
  if (Ca == 3 && Cb % 8 == 0) { p.ra = 3; p.rb = 8; }
  else if (Ca == 5 && Cb % 8 == 0) { p.ra = 5; p.rb = 8; }
  else if (Cb == 3 && Ca % 8 == 0) { p.ra = 8; p.rb = 3; }
  else if (Cb == 5 && Ca % 8 == 0) { p.ra = 8; p.rb = 5; }
  else { p.ra = 4; p.rb = 4; }
  p.n_at = cdiv(Ca, p.ra); p.n_bt = cdiv(Cb, p.rb);
  const int n_et = 9 * p.n_at * p.n_bt;
  p.nph = 256 / n_et;
  p.seg = Wp < 160 ? Wp : 160;                     // whole rows when they fit
  p.cbp = Cb % 4 == 0 ? Cb + 4 : Cb + 1;            // de-phase the taps' bank mapping
  const int qw = s * (p.seg - 1) + 3;
  const size_t stage = (size_t)(((size_t)p.seg * Ca + 3) & ~(size_t)3) + (size_t)3 * qw * p.cbp;
  const size_t fold = (size_t)(p.nph > 0 ? p.nph : 1) * 9 * Ca * Cb;
  p.smem = (stage > fold ? stage : fold) * sizeof(float);
  p.ok = p.nph >= 1 && p.smem <= 200 * 1024;
  const int rows = B * Hp;
  const int per_sm = p.smem <= 48 * 1024 ? 4 : (p.smem <= 70 * 1024 ? 3 : 2);   // co-resident blocks hide the staging
  int blocks = kNumSMs * per_sm;
  if (blocks > rows) blocks = rows;
  p.rpc = cdiv(rows, blocks);
  p.blocks = cdiv(rows, p.rpc);
  return p;
}

void colsum(const float* in, int64_t rows, int C, float* out, float* partial, cudaStream_t st);
struct RowsPlan { WgradRowsGeom g; int ra, rb, blocks; size_t smem; bool ok; };
static RowsPlan wgrad_rows_plan(const WgradArgs& a) {
  RowsPlan p{};
  const int Ca = a.Ca, Cb = a.Cb;
  if (Ca == 3 && Cb % 8 == 0) { p.ra = 3; p.rb = 8; }
  else if (Ca == 5 && Cb % 8 == 0) { p.ra = 5; p.rb = 8; }
  else if (Cb == 3 && Ca % 8 == 0) { p.ra = 8; p.rb = 3; }
  else if (Cb == 5 && Ca % 8 == 0) { p.ra = 8; p.rb = 5; }
  else { p.ra = 4; p.rb = 4; }
  WgradRowsGeom& g = p.g;
  g.n_at = cdiv(Ca, p.ra); g.n_bt = cdiv(Cb, p.rb);
  const int n_et = 9 * g.n_at * g.n_bt;
  g.nph = 256 / n_et;
  if (g.nph < 1 || (a.d != 1 && a.d != -1)) return p;
  g.p_pad = (Ca % 4 == 0) && ((uintptr_t)a.P % 16 == 0);
  g.q_pad = (Cb % 4 == 0) && ((uintptr_t)a.Q % 16 == 0);
  g.cap = g.p_pad ? Ca + 4 : Ca;
  g.cbp = g.q_pad ? Cb + 4 : Cb;
  g.p_floats = g.p_pad ? a.Wp * g.cap : ((a.Wp * Ca + 4 + 3) & ~3);
  g.q_floats = g.q_pad ? 3 * a.Wq * g.cbp : ((3 * a.Wq * Cb + 4 + 3) & ~3);
  // the fused column sum of P rides on the centre tap, which must be in range for every pixel
  auto centre_ok = [&](int np, int nq, int off) { return a.d + off >= 0 && a.s * (np - 1) + a.d + off < nq; };
  g.want_colsum = a.pcolsum && centre_ok(a.Hp, a.Hq, a.oy) && centre_ok(a.Wp, a.Wq, a.ox);
  const int E = 9 * Ca * Cb;
  g.pstride = E + (g.want_colsum ? Ca : 0);
  const size_t stage = (size_t)2 * (g.p_floats + g.q_floats);
  const size_t fold = (size_t)g.nph * g.pstride;
  p.smem = (stage > fold ? stage : fold) * sizeof(float);
  if (p.smem > 200 * 1024) return p;
  g.rows = a.B * a.Hp;
  const int per_sm = p.smem <= 100 * 1024 ? 2 : 1;
  int blocks = kNumSMs * per_sm;
  if (blocks > g.rows) blocks = g.rows;
  g.rows_per_block = cdiv(g.rows, blocks);
  p.blocks = cdiv(g.rows, g.rows_per_block);
  p.ok = p.blocks >= 1;
  return p;
}

// returns true when the column sums of P (a.pcolsum) still have to be computed separately
static bool conv_wgrad_impl(const WgradArgs& a, cudaStream_t st) {
  ProfScope prof_("wgrad", st);
  const int E = 9 * a.Ca * a.Cb;
  const RowsPlan rp = wgrad_rows_plan(a);
  if (rp.ok) {
    g_launches += 2;
#ifndef KCVAE_EMU
#define KC_WR_ATTR(k) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rp.smem)
#else
#define KC_WR_ATTR(k)
#endif
#define KC_WR_LAUNCH(RA_, RB_)                                   \
  {                                                              \
    auto k = wgrad_rows_kernel<RA_, RB_>;                        \
    KC_WR_ATTR(k);                                               \
    KC_LAUNCH(k, rp.blocks, 256, rp.smem, st, a, rp.g);          \
  }
    if (rp.ra == 3 && rp.rb == 8) KC_WR_LAUNCH(3, 8)
    else if (rp.ra == 5 && rp.rb == 8) KC_WR_LAUNCH(5, 8)
    else if (rp.ra == 8 && rp.rb == 3) KC_WR_LAUNCH(8, 3)
    else if (rp.ra == 8 && rp.rb == 5) KC_WR_LAUNCH(8, 5)
    else KC_WR_LAUNCH(4, 4)
#undef KC_WR_LAUNCH
#undef KC_WR_ATTR
    const int FE = rp.g.pstride;
    KC_LAUNCH(wgrad_reduce_sliced_kernel, cdiv(FE, 256 / WR_SLICES), 256, 0, st, a.partial, rp.blocks, rp.g.pstride, E, FE,
              a.Ca, a.Cb, a.o_sa, a.o_sb, a.out, a.pcolsum);
    return a.pcolsum && !rp.g.want_colsum;
  }
  const WgradPlan pl = wgrad_plan(a.B, a.Hp, a.Wp, a.Ca, a.Cb, a.s);
  if (pl.ok) {
    const int rows_t = a.B * a.Hp;
    g_launches += 2;
#ifndef KCVAE_EMU
#define KC_WG_ATTR(k) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)
#else
#define KC_WG_ATTR(k)
#endif
#define KC_WG_LAUNCH(RA_, RB_)                                                                              \
  {                                                                                                         \
    auto k = wgrad_tiled_kernel<RA_, RB_>;                                                                  \
    KC_WG_ATTR(k);                                                                                          \
    KC_LAUNCH(k, pl.blocks, 256, pl.smem, st, a, rows_t, pl.rpc, pl.seg, pl.n_at, pl.n_bt, pl.nph, pl.cbp); \
  }
    if (pl.ra == 3 && pl.rb == 8) KC_WG_LAUNCH(3, 8)
    else if (pl.ra == 5 && pl.rb == 8) KC_WG_LAUNCH(5, 8)
    else if (pl.ra == 8 && pl.rb == 3) KC_WG_LAUNCH(8, 3)
    else if (pl.ra == 8 && pl.rb == 5) KC_WG_LAUNCH(8, 5)
    else KC_WG_LAUNCH(4, 4)
#undef KC_WG_LAUNCH
#undef KC_WG_ATTR
    KC_LAUNCH(wgrad_reduce_kernel, cdiv(E, 256), 256, 0, st, a.partial, pl.blocks, E, a.Ca, a.Cb, a.o_sa, a.o_sb, a.out);
  } else {
    const int chunks = wgrad_chunks(a.B, a.Hp, a.Ca, a.Cb);
    const int rows = a.B * a.Hp;
    const int rpc = cdiv(rows, chunks);
    dim3 grid(cdiv(E, 256), cdiv(rows, rpc));
    g_launches += 2;
    KC_LAUNCH(wgrad_kernel, grid, 256, 0, st, a, E, rows, rpc);
    KC_LAUNCH(wgrad_reduce_kernel, cdiv(E, 256), 256, 0, st, a.partial, (int)grid.y, E, a.Ca, a.Cb,
              a.o_sa, a.o_sb, a.out);
  }
  return a.pcolsum != nullptr;
}

void conv_wgrad(const WgradArgs& a, cudaStream_t st) {
  if (conv_wgrad_impl(a, st)) colsum(a.P, (int64_t)a.B * a.Hp * a.Wp, a.Ca, a.pcolsum, a.partial, st);
}

// ---------------------------------------------------------------------- column sums
// small C (<= 256): block = C*floor(256/C) threads so a thread's channel is fixed while it
// strides linearly (coalesced) through its row chunk.  large C: thread per column.
static int colsum_blocks(int64_t rows, int C) {
  if (C > 256) return 1;
  const int tpb = C * (256 / C);
  const int rows_per_iter = tpb / C;
  int64_t want = (rows + (int64_t)rows_per_iter * 64 - 1) / ((int64_t)rows_per_iter * 64);
  if (want < 1) want = 1;
  if (want > kNumSMs * 8) want = kNumSMs * 8;
  return (int)want;
}
size_t colsum_partial_floats(int64_t rows, int C) { return (size_t)colsum_blocks(rows, C) * (C > 256 ? 1 : C); }

__global__ void colsum_small_kernel(const float* in, int64_t rows, int C, int64_t rows_per_block, float* partial) {
  __shared__ float sm[256];
  const int tpb = blockDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc = 0.0f;
  if (r0 < r1) {
    const int64_t e1 = (r1 - r0) * C;
    const float* base = in + r0 * C;
    for (int64_t e = threadIdx.x; e < e1; e += tpb) acc += __ldg(base + e);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  if ((int)threadIdx.x < C) {
    float s = 0.0f;
    for (int t = threadIdx.x; t < tpb; t += C) s += sm[t];
    partial[(int64_t)blockIdx.x * C + threadIdx.x] = s;
  }
}
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float* partial, int blocks, int C, float* out) {
  __shared__ float red[256];   // 8 threads per column, each a contiguous slice of the blocks (fixed order)
  const int oi = threadIdx.x / WR_SLICES, sl = threadIdx.x % WR_SLICES;
  const int c = blockIdx.x * (256 / WR_SLICES) + oi;
  float t = 0.0f;
  if (c < C) {
    const int per = (blocks + WR_SLICES - 1) / WR_SLICES;
    const int b0 = sl * per, b1 = min(blocks, b0 + per);
#pragma unroll 4
    for (int b = b0; b < b1; ++b) t += __ldg(partial + (int64_t)b * C + c);
  }
  red[threadIdx.x] = t;
  __syncthreads();
  if (sl == 0 && c < C) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < WR_SLICES; ++i) s += red[threadIdx.x + i];
    out[c] = s;
  }
}
__global__ void colsum_large_kernel(const float* in, int64_t rows, int C, float* out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.0f;
  for (int64_t r = 0; r < rows; ++r) s += __ldg(in + r * C + c);
  out[c] = s;
}

void colsum(const float* in, int64_t rows, int C, float* out, float* partial, cudaStream_t st) {
  ProfScope prof_("colsum", st);
  if (C > 256) {
    ++g_launches;
    KC_LAUNCH(colsum_large_kernel, cdiv(C, 256), 256, 0, st, in, rows, C, out);
    return;
  }
  const int tpb = C * (256 / C);
  const int blocks = colsum_blocks(rows, C);
  const int64_t rpb = (rows + blocks - 1) / blocks;
  g_launches += 2;
  KC_LAUNCH(colsum_small_kernel, blocks, tpb, 0, st, in, rows, C, rpb, partial);
  KC_LAUNCH(colsum_finish_kernel, cdiv(C, 256 / WR_SLICES), 256, 0, st, partial, blocks, C, out);
}

}  // namespace kc
