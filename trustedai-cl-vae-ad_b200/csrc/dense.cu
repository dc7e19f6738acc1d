// dense.cu - strided fp32 GEMM with deterministic split-K for the Dense layers of the path
// (src/abstract_cvae.py:41-45 encoder Dense / head, :76 decoder Dense) and their
// gradients.  C[M,N] = epi( A[M,K] . B[K,N] (+ bias[n]) ), arbitrary element strides so
// every transpose the backward pass needs is the same kernel.
#include "kernels.h"

namespace kc {

constexpr int TM = 32, TN = 32, TK = 32;

static int gemm_splits(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 1;
  const int64_t tiles = (int64_t)cdiv(M, TM) * cdiv(N, TN);
  int64_t want = (int64_t)kNumSMs * 4 / tiles;
  int64_t maxs = K / (4 * TK);
  if (want > maxs) want = maxs;
  if (want > 512) want = 512;
  if (want < 1) want = 1;
  return (int)want;
}
size_t gemm_partial_floats(int M, int N, int K) {
  if (M <= 0 || N <= 0) return 0;
  const int s = gemm_splits(M, N, K);
  return s > 1 ? (size_t)s * M * N : 0;
}

__device__ __forceinline__ float gemm_epi(float v, int m, int n, int N, const float* bias,
                                          const float* mask, int relu) {
  if (bias) v += __ldg(bias + n);
  if (relu) v = fmaxf(v, 0.0f);
  if (mask) v = __ldg(mask + (int64_t)m * N + n) > 0.0f ? v : 0.0f;
  return v;
}

// 256 threads: (ty, tx) = 16x16, each thread a 2x2 micro-tile of the 32x32 block tile
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs a, int splits, int k_per_split) {
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(a.K, k_begin + k_per_split);
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
    // cooperative loads: 1024 elements per tile, 4 per thread
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int li = threadIdx.x + t * 256;
      {  // A tile: choose the faster-varying index to follow the smaller stride
        int mm, kk;
        if (a.a_sk <= a.a_sm) { kk = li % TK; mm = li / TK; } else { mm = li % TM; kk = li / TM; }
        const int gm = m0 + mm, gk = k0 + kk;
        As[kk][mm] = (gm < a.M && gk < k_end) ? __ldg(a.A + (int64_t)gm * a.a_sm + (int64_t)gk * a.a_sk) : 0.0f;
      }
      {
        int nn, kk;
        if (a.b_sn <= a.b_sk) { nn = li % TN; kk = li / TN; } else { kk = li % TK; nn = li / TK; }
        const int gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < a.N && gk < k_end) ? __ldg(a.Bm + (int64_t)gk * a.b_sk + (int64_t)gn * a.b_sn) : 0.0f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float a0 = As[kk][ty], a1 = As[kk][ty + 16];
      const float b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m >= a.M || n >= a.N) continue;
      if (splits > 1) a.partial[((int64_t)blockIdx.z * a.M + m) * a.N + n] = acc[i][j];
      else a.C[(int64_t)m * a.N + n] = gemm_epi(acc[i][j], m, n, a.N, a.bias, a.mask, a.relu);
    }
}

// 8 threads per output entry, each a contiguous slice of the splits, folded in a fixed order
constexpr int GR_SLICES = 8;
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(GemmArgs a, int splits) {
  __shared__ float red[256];
  const int64_t MN = (int64_t)a.M * a.N;
  const int oi = threadIdx.x / GR_SLICES, sl = threadIdx.x % GR_SLICES;
  const int64_t i = (int64_t)blockIdx.x * (256 / GR_SLICES) + oi;
  float t = 0.0f;
  if (i < MN) {
    const int per = (splits + GR_SLICES - 1) / GR_SLICES;
    const int z0 = sl * per, z1 = min(splits, z0 + per);
#pragma unroll 4
    for (int z = z0; z < z1; ++z) t += __ldg(a.partial + (int64_t)z * MN + i);
  }
  red[threadIdx.x] = t;
  __syncthreads();
  if (sl == 0 && i < MN) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < GR_SLICES; ++k) s += red[threadIdx.x + k];
    const int m = (int)(i / a.N), n = (int)(i % a.N);
    a.C[i] = gemm_epi(s, m, n, a.N, a.bias, a.mask, a.relu);
  }
}

void gemm(const GemmArgs& a, cudaStream_t st) {
  ProfScope prof_("gemm", st);
  if (a.M <= 0 || a.N <= 0) return;
  int splits = a.partial ? gemm_splits(a.M, a.N, a.K) : 1;
  int kps = cdiv(cdiv(a.K, splits), TK) * TK;
  if (kps < TK) kps = TK;
  splits = cdiv(a.K, kps);
  if (splits < 1) splits = 1;
  dim3 grid(cdiv(a.N, TN), cdiv(a.M, TM), splits);
  ++g_launches;
  KC_LAUNCH(gemm_kernel, grid, 256, 0, st, a, splits, kps);
  if (splits > 1) {
    ++g_launches;
    KC_LAUNCH(gemm_splitk_reduce_kernel, cdiv((int64_t)a.M * a.N, 256 / GR_SLICES), 256, 0, st, a, splits);
  }
}

// =========================================================================================
// Wide Dense: the decoder Dense (src/abstract_cvae.py:76) has K = latent (tens) and
// N = h0*w0*filters (1e5): pure weight/activation streaming.  Three bandwidth kernels replace
// five generic GEMM launches; every global access is a coalesced float4.
//   forward   C[b,n]  = relu(bias[n] + sum_k A[b,k] W[k,n])
//   wgrad     dW[k,n] = sum_b A[b,k] G[b,n]      db[n] = sum_b G[b,n]      (one read of G)
//   dgrad     dA[b,k] = sum_n G[b,n] W[k,n]      (deterministic two-level reduction over n)
// =========================================================================================
constexpr int DW_BT = 8;          // batch rows per thread (forward); 16 halves the L2 re-reads of W but measured slower (208 registers, 8 warps per SM)
constexpr int DW_KT = 32;         // k rows per block (wgrad)
constexpr int DG_NC = 128;        // columns per staged chunk (dgrad)
constexpr int DG_LD = DG_NC + 4;  // smem row stride == 4 (mod 32): conflict-free float4 reads
constexpr int DG_SLICES = 8;

bool dense_wide_ok(const void* A, const void* W, const void* C, const void* bias, int M, int N, int K) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return M > 0 && K > 0 && N >= 64 && (N & 3) == 0 && K <= 4096 && al(A) && al(W) && al(C) && (!bias || al(bias));
}

// NC = output columns per thread: 4, or 8 when the bf16 chunk-planar copy is written (one whole 16-byte unit per thread
// and row, so that a warp's stores are four contiguous 128-byte runs instead of sixteen 16-byte pieces)
template <int BT, int NC>
__global__ void __launch_bounds__(128) dense_wide_fwd_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                             const float* __restrict__ bias, float* __restrict__ C,
                                                             int M, int N, int K, int relu, uint4* __restrict__ Cp, int Cc,
                                                             int split) {
  KC_DYN_SMEM(float, As);   // [BT][K]
  constexpr int NV = NC / 4;
  const int m0 = blockIdx.y * BT;
  for (int i = threadIdx.x; i < BT * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    As[i] = (m0 + r < M) ? __ldg(A + (int64_t)(m0 + r) * K + k) : 0.f;
  }
  __syncthreads();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) * NC;
  if (n >= N) return;
  float acc[BT][NC];
#pragma unroll
  for (int r = 0; r < BT; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;
  const float4* wp = reinterpret_cast<const float4*>(W + n);
  const int64_t wstride = N >> 2;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float w[NC];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 t = __ldg(wp + (int64_t)k * wstride + v);
      w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
    }
#pragma unroll
    for (int r = 0; r < BT; ++r) {
      const float a = As[r * K + k];
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[r][c] = fmaf(a, w[c], acc[r][c]);
    }
  }
  float bv[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) bv[c] = bias ? __ldg(bias + n + c) : 0.f;
#pragma unroll
  for (int r = 0; r < BT; ++r) {
    if (m0 + r >= M) break;
    float y[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { y[c] = acc[r][c] + bv[c]; if (relu) y[c] = fmaxf(y[c], 0.f); }
    if (C) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
        *reinterpret_cast<float4*>(C + (int64_t)(m0 + r) * N + n + 4 * v) = make_float4(y[4 * v], y[4 * v + 1], y[4 * v + 2], y[4 * v + 3]);
    }
    if (NC == 8 && Cp) {   // bf16 chunk-planar copy [M][Cc/8][N/Cc pixels][8] for a tensor-core consumer (output seen as [pixels][Cc])
      const int px = n / Cc, ch = n - px * Cc;
      uint32_t h2[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t lo = __float_as_uint(y[2 * e]), hi = __float_as_uint(y[2 * e + 1]);     // round to nearest even
        h2[e] = ((lo + 0x7FFFu + ((lo >> 16) & 1u)) >> 16) | ((hi + 0x7FFFu + ((hi >> 16) & 1u)) & 0xFFFF0000u);
      }
      const int KC = Cc >> 3;
      const int64_t hwp = N / Cc;
      const int64_t unit = ((int64_t)(m0 + r) * (split ? 2 * KC : KC) + (ch >> 3)) * hwp + px;
      Cp[unit] = make_uint4(h2[0], h2[1], h2[2], h2[3]);
      if (split) {   // lo planes: what bf16 dropped, again in bf16 (y = hi + lo to 2^-17)
        uint32_t l2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float r0 = y[2 * e] - __uint_as_float(h2[e] << 16), r1 = y[2 * e + 1] - __uint_as_float(h2[e] & 0xFFFF0000u);
          const uint32_t lo = __float_as_uint(r0), hi = __float_as_uint(r1);
          l2[e] = ((lo + 0x7FFFu + ((lo >> 16) & 1u)) >> 16) | ((hi + 0x7FFFu + ((hi >> 16) & 1u)) & 0xFFFF0000u);
        }
        Cp[unit + (int64_t)KC * hwp] = make_uint4(l2[0], l2[1], l2[2], l2[3]);
      }
    }
  }
}

// C (fp32 [M,N]) and / or Cp (bf16 chunk-planar copy, the output seen as [N/Cc pixels][Cc channels], Cc % 8 == 0;
// split: hi planes followed by lo planes per row of C)
void dense_wide_forward(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int relu,
                        cudaStream_t st, void* Cp, int Cc, int split) {
  ProfScope prof_("dense_wide_fwd", st);
  const bool planar = Cp && Cc % 8 == 0 && N % 8 == 0;
  dim3 grid(cdiv(N / (planar ? 8 : 4), 128), cdiv(M, DW_BT));
  const size_t smem = (size_t)DW_BT * K * sizeof(float);
  ++g_launches;
  if (planar) {
#ifndef KCVAE_EMU
    if (smem > 48 * 1024) cudaFuncSetAttribute(dense_wide_fwd_kernel<DW_BT, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    KC_LAUNCH((dense_wide_fwd_kernel<DW_BT, 8>), grid, 128, smem, st, A, W, bias, C, M, N, K, relu, reinterpret_cast<uint4*>(Cp), Cc, split);
  } else {
#ifndef KCVAE_EMU
    if (smem > 48 * 1024) cudaFuncSetAttribute(dense_wide_fwd_kernel<DW_BT, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    KC_LAUNCH((dense_wide_fwd_kernel<DW_BT, 4>), grid, 128, smem, st, A, W, bias, C, M, N, K, relu, static_cast<uint4*>(nullptr), 8, 0);
  }
}

// block (64 column quads) x (4 k-subgroups of 8): dW tile [32 k][256 n]; grid.y = k tiles
__global__ void __launch_bounds__(256) dense_wide_wgrad_kernel(const float* __restrict__ A, const float* __restrict__ G,
                                                               float* __restrict__ dW, float* __restrict__ db,
                                                               int M, int N, int K) {
  KC_DYN_SMEM(float, At);   // [M][DW_KT] slice of A for this k tile (zero padded)
  const int k0 = blockIdx.y * DW_KT;
  for (int i = threadIdx.x; i < M * DW_KT; i += blockDim.x) {
    const int b = i / DW_KT, kk = i - b * DW_KT;
    At[i] = (k0 + kk < K) ? __ldg(A + (int64_t)b * K + k0 + kk) : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int n = (blockIdx.x * 64 + tx) * 4;
  if (n >= N) return;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* gp = reinterpret_cast<const float4*>(G + n);
  const int64_t gstride = N >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(At) + ty * 2;
#pragma unroll 4
  for (int b = 0; b < M; ++b) {
    const float4 g = __ldg(gp + (int64_t)b * gstride);
    const float4 x0 = a4[b * (DW_KT / 4)], x1 = a4[b * (DW_KT / 4) + 1];
    const float av[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i][0] = fmaf(av[i], g.x, acc[i][0]); acc[i][1] = fmaf(av[i], g.y, acc[i][1]);
      acc[i][2] = fmaf(av[i], g.z, acc[i][2]); acc[i][3] = fmaf(av[i], g.w, acc[i][3]);
    }
    bs.x += g.x; bs.y += g.y; bs.z += g.z; bs.w += g.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + ty * 8 + i;
    if (k < K) *reinterpret_cast<float4*>(dW + (int64_t)k * N + n) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  if (db && blockIdx.y == 0 && ty == 0) *reinterpret_cast<float4*>(db + n) = bs;
}

static int dense_wide_dgrad_blocks(int N) {
  int nb = cdiv(N, DG_NC);
  if (nb > kNumSMs * 2) nb = kNumSMs * 2;
  return nb;
}
size_t dense_wide_partial_floats(int M, int N, int K) {
  return (size_t)dense_wide_dgrad_blocks(N) * cdiv(M, 32) * cdiv(K, 32) * 1024;
}

// grid (column blocks, b tiles, k tiles); thread = 4x4 (b,k) micro tile x one of 4 column phases
__global__ void __launch_bounds__(256) dense_wide_dgrad_kernel(const float* __restrict__ G, const float* __restrict__ W,
                                                               float* __restrict__ partial, int M, int N, int K,
                                                               int cols_per_block) {
  __shared__ __align__(16) float Gs[32 * DG_LD];
  __shared__ __align__(16) float Ws[32 * DG_LD];
  const int tid = threadIdx.x;
  const int o = tid & 63, phase = tid >> 6;
  const int tb = o & 7, tk = o >> 3;
  const int b0 = blockIdx.y * 32, k0 = blockIdx.z * 32;
  const int c_begin = blockIdx.x * cols_per_block;
  const int c_end = min(N, c_begin + cols_per_block);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  for (int c0 = c_begin; c0 < c_end; c0 += DG_NC) {
    __syncthreads();
    for (int i = tid; i < 32 * (DG_NC / 4); i += 256) {   // coalesced float4 rows
      const int r = i / (DG_NC / 4), q = i - r * (DG_NC / 4);
      const int c = c0 + q * 4;
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f), wv = gv;
      if (c < c_end) {
        if (b0 + r < M) gv = __ldg(reinterpret_cast<const float4*>(G + (int64_t)(b0 + r) * N + c));
        if (k0 + r < K) wv = __ldg(reinterpret_cast<const float4*>(W + (int64_t)(k0 + r) * N + c));
      }
      *reinterpret_cast<float4*>(Gs + r * DG_LD + q * 4) = gv;
      *reinterpret_cast<float4*>(Ws + r * DG_LD + q * 4) = wv;
    }
    __syncthreads();
#pragma unroll 2
    for (int q = phase; q < DG_NC / 4; q += 4) {
      float4 g[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i] = *reinterpret_cast<const float4*>(Gs + (tb + 8 * i) * DG_LD + q * 4);
        w[i] = *reinterpret_cast<const float4*>(Ws + (tk + 8 * i) * DG_LD + q * 4);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(g[i].x, w[j].x, acc[i][j]); acc[i][j] = fmaf(g[i].y, w[j].y, acc[i][j]);
          acc[i][j] = fmaf(g[i].z, w[j].z, acc[i][j]); acc[i][j] = fmaf(g[i].w, w[j].w, acc[i][j]);
        }
    }
  }
  // fold the 4 column phases (fixed order): Gs reused as [4][32*32]
  __syncthreads();
  float* red = Gs;   // 4096 floats <= 32*DG_LD = 4224
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[phase * 1024 + (tb + 8 * i) * 32 + (tk + 8 * j)] = acc[i][j];
  __syncthreads();
  float* out = partial + (((int64_t)blockIdx.x * gridDim.y + blockIdx.y) * gridDim.z + blockIdx.z) * 1024;
  for (int e = tid; e < 1024; e += 256) out[e] = (red[e] + red[1024 + e]) + (red[2048 + e] + red[3072 + e]);
}

// dA[b,k] = sum over column blocks, DG_SLICES threads per output, fixed summation order
__global__ void __launch_bounds__(256) dense_wide_dgrad_reduce_kernel(const float* __restrict__ partial, int nblk, int M,
                                                                      int K, int bt, int kt, float* __restrict__ dA) {
  __shared__ float red[256];
  const int oi = threadIdx.x / DG_SLICES, sl = threadIdx.x % DG_SLICES;
  const int64_t out_idx = (int64_t)blockIdx.x * (256 / DG_SLICES) + oi;
  const int64_t total = (int64_t)M * K;
  float s = 0.f;
  if (out_idx < total) {
    const int b = (int)(out_idx / K), k = (int)(out_idx % K);
    const int tile = (b >> 5) * kt + (k >> 5), e = (b & 31) * 32 + (k & 31);
    const int per = (nblk + DG_SLICES - 1) / DG_SLICES;
    const int j0 = sl * per, j1 = min(nblk, j0 + per);
    const int64_t stride = (int64_t)bt * kt * 1024;
    const float* p = partial + (int64_t)tile * 1024 + e;
#pragma unroll 4
    for (int j = j0; j < j1; ++j) s += __ldg(p + (int64_t)j * stride);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  if (sl == 0 && out_idx < total) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < DG_SLICES; ++i) t += red[threadIdx.x + i];
    dA[out_idx] = t;
  }
}

// G [M,N] (gradient of the Dense output, ReLU mask already applied), A [M,K], W [K,N]
// st_w: stream of the weight-gradient kernel (independent of the data gradient: may run beside it)
void dense_wide_backward(const float* A, const float* G, const float* W, float* dW, float* db, float* dA, float* partial,
                         int M, int N, int K, cudaStream_t st, cudaStream_t st_w) {
  {
    ProfScope prof_("dense_wide_wgrad", st_w);
    dim3 grid(cdiv(N / 4, 64), cdiv(K, DW_KT));
    const size_t smem = (size_t)M * DW_KT * sizeof(float);
    ++g_launches;
#ifndef KCVAE_EMU
    if (smem > 48 * 1024) cudaFuncSetAttribute(dense_wide_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
#endif
    KC_LAUNCH(dense_wide_wgrad_kernel, grid, 256, smem, st_w, A, G, dW, db, M, N, K);
  }
  if (dA) {
    ProfScope prof_("dense_wide_dgrad", st);
    const int nblk = dense_wide_dgrad_blocks(N);
    const int cpb = cdiv(cdiv(N, nblk), DG_NC) * DG_NC;
    const int nb = cdiv(N, cpb);
    const int bt = cdiv(M, 32), kt = cdiv(K, 32);
    g_launches += 2;
    KC_LAUNCH(dense_wide_dgrad_kernel, dim3(nb, bt, kt), 256, 0, st, G, W, partial, M, N, K, cpb);
    KC_LAUNCH(dense_wide_dgrad_reduce_kernel, cdiv((int64_t)M * K, 256 / DG_SLICES), 256, 0, st, partial, nb, M, K, bt, kt, dA);
  }
}

}  // namespace kc
