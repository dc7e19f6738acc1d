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

__global__ void gemm_splitk_reduce_kernel(GemmArgs a, int splits) {
  const int64_t MN = (int64_t)a.M * a.N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < MN;
       i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int z = 0; z < splits; ++z) s += a.partial[(int64_t)z * MN + i];
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
    KC_LAUNCH(gemm_splitk_reduce_kernel, grid_for((int64_t)a.M * a.N, 256), 256, 0, st, a, splits);
  }
}

}  // namespace kc
