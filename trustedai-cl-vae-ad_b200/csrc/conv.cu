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
  return (size_t)wgrad_chunks(B, Hp, Ca, Cb) * 9 * Ca * Cb;
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

void conv_wgrad(const WgradArgs& a, cudaStream_t st) {
  ProfScope prof_("wgrad", st);
  const int E = 9 * a.Ca * a.Cb;
  const int chunks = wgrad_chunks(a.B, a.Hp, a.Ca, a.Cb);
  const int rows = a.B * a.Hp;
  const int rpc = cdiv(rows, chunks);
  dim3 grid(cdiv(E, 256), cdiv(rows, rpc));
  g_launches += 2;
  KC_LAUNCH(wgrad_kernel, grid, 256, 0, st, a, E, rows, rpc);
  KC_LAUNCH(wgrad_reduce_kernel, cdiv(E, 256), 256, 0, st, a.partial, (int)grid.y, E, a.Ca, a.Cb,
            a.o_sa, a.o_sb, a.out);
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
__global__ void colsum_finish_kernel(const float* partial, int blocks, int C, float* out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.0f;
  for (int b = 0; b < blocks; ++b) s += partial[(int64_t)b * C + c];
  out[c] = s;
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
  KC_LAUNCH(colsum_finish_kernel, cdiv(C, 256), 256, 0, st, partial, blocks, C, out);
}

}  // namespace kc
