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
  size_t n = (size_t)wgrad_chunks(B, Hp, Ca, Cb);
  if (n < (size_t)kNumSMs * 2) n = (size_t)kNumSMs * 2;   // the tiled kernel uses <= 2 blocks per SM
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
  float* Qs = sm + SEG * a.Ca;                      // [3][qw][cbp]
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
      {  // stage P segment (contiguous)
        const float* src = a.P + (((int64_t)r * a.Wp) + j0) * a.Ca;
        for (int t = tid; t < seg * a.Ca; t += blockDim.x) Ps[t] = __ldg(src + t);
        for (int t = seg * a.Ca + tid; t < SEG * a.Ca; t += blockDim.x) Ps[t] = 0.f;
      }
      for (int k = 0; k < 3; ++k) {  // stage the 3 Q rows (zero outside the image)
        const int qy = a.s * i + a.d * k + a.oy;
        const bool vrow = qy >= 0 && qy < a.Hq;
        const float* src = a.Q + (((int64_t)n * a.Hq + (vrow ? qy : 0)) * a.Wq) * a.Cb;
        float* dst = Qs + (int64_t)k * qw * cbp;
        for (int t = tid; t < qw * a.Cb; t += blockDim.x) {
          const int px = t / a.Cb, c = t % a.Cb;
          const int qx = qx_min + px;
          dst[px * cbp + c] = (vrow && qx >= 0 && qx < a.Wq) ? __ldg(src + (int64_t)qx * a.Cb + c) : 0.f;
        }
      }
      __syncthreads();
      if (active) {
        const float* qrow = Qs + (int64_t)kh * qw * cbp + (a.d * kw - dmin) * cbp + b0;
        for (int j = ph; j < seg; j += nph) {
          float pv[RA], qv[RB];
#pragma unroll
          for (int x = 0; x < RA; ++x) pv[x] = (a0 + x < a.Ca) ? Ps[j * a.Ca + a0 + x] : 0.f;
          const float* qp = qrow + (int64_t)(a.s * j) * cbp;
#pragma unroll
          for (int y = 0; y < RB; ++y) qv[y] = (b0 + y < a.Cb) ? qp[y] : 0.f;
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

struct WgradPlan { int ra, rb, n_at, n_bt, nph, seg, cbp, blocks, rpc; size_t smem; bool ok; };
static WgradPlan wgrad_plan(int B, int Hp, int Wp, int Ca, int Cb, int s) {
  WgradPlan p{};
  if (Ca == 3) { p.ra = 3; p.rb = 8; }
  else if (Ca == 5) { p.ra = 5; p.rb = 4; }
  else if (Cb == 3) { p.ra = 8; p.rb = 3; }
  else if (Cb == 5) { p.ra = 4; p.rb = 5; }
  else { p.ra = 4; p.rb = 4; }
  p.n_at = cdiv(Ca, p.ra); p.n_bt = cdiv(Cb, p.rb);
  const int n_et = 9 * p.n_at * p.n_bt;
  p.nph = 256 / n_et;
  p.seg = s == 1 ? 64 : 32;
  if (p.seg > Wp) p.seg = Wp;
  p.cbp = Cb % 4 == 0 ? Cb + 4 : Cb + 1;            // de-phase the taps' bank mapping
  const int qw = s * (p.seg - 1) + 3;
  const size_t stage = (size_t)p.seg * Ca + (size_t)3 * qw * p.cbp;
  const size_t fold = (size_t)(p.nph > 0 ? p.nph : 1) * 9 * Ca * Cb;
  p.smem = (stage > fold ? stage : fold) * sizeof(float);
  p.ok = p.nph >= 1 && p.smem <= 200 * 1024;
  const int rows = B * Hp;
  int blocks = kNumSMs * 2;
  if (blocks > rows) blocks = rows;
  p.rpc = cdiv(rows, blocks);
  p.blocks = cdiv(rows, p.rpc);
  return p;
}

void conv_wgrad(const WgradArgs& a, cudaStream_t st) {
  ProfScope prof_("wgrad", st);
  const int E = 9 * a.Ca * a.Cb;
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
    else if (pl.ra == 5 && pl.rb == 4) KC_WG_LAUNCH(5, 4)
    else if (pl.ra == 8 && pl.rb == 3) KC_WG_LAUNCH(8, 3)
    else if (pl.ra == 4 && pl.rb == 5) KC_WG_LAUNCH(4, 5)
    else KC_WG_LAUNCH(4, 4)
#undef KC_WG_LAUNCH
#undef KC_WG_ATTR
    KC_LAUNCH(wgrad_reduce_kernel, cdiv(E, 256), 256, 0, st, a.partial, pl.blocks, E, a.Ca, a.Cb, a.o_sa, a.o_sb, a.out);
    return;
  }
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
