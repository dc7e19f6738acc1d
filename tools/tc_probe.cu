// tc_probe.cu - descriptor lab for the hand-written tcgen05 path (development tool).
//
// Runs single tcgen05.mma sequences with the exact shared-memory layouts / descriptor
// conventions the kcvae tensor-core kernels rely on and compares the TMEM result with a host
// GEMM.  Every wait is bounded, so a wrong descriptor reports FAIL instead of hanging.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I trustedai-cl-vae-ad_b200/csrc tools/tc_probe.cu -o build/tc_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_common.cuh"

using namespace kc::tc;

struct ProbeArgs {
  const unsigned char* a_img; int a_bytes;   // raw shared-memory image of operand A region
  const unsigned char* b_img; int b_bytes;
  int a_start, a_lbo, a_sbo, a_kstep;        // bytes
  int b_start, b_lbo, b_sbo, b_kstep;
  int ksteps, M, N, a_mn, b_mn;
  float* out;                                 // [128 lanes][N] dump
  int* status;                                // 0 ok, 1 timeout
};

__global__ void __launch_bounds__(128) probe_kernel(ProbeArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  unsigned char* sa = smem;
  unsigned char* sb = smem + ((p.a_bytes + 1023) / 1024) * 1024;
  for (int i = threadIdx.x; i < p.a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(p.a_img)[i];
  for (int i = threadIdx.x; i < p.b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(p.b_img)[i];
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<64>(&tmem_slot);
  if (threadIdx.x == 32) { mbar_init(&mbar, 1); fence_mbar_init(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16_f32(p.M, p.N, p.a_mn, p.b_mn);
    for (int ks = 0; ks < p.ksteps; ++ks) {
      const uint64_t da = make_desc_kmajor_noswz(smem_u32(sa) + p.a_start + ks * p.a_kstep, p.a_lbo, p.a_sbo);
      const uint64_t db = make_desc_kmajor_noswz(smem_u32(sb) + p.b_start + ks * p.b_kstep, p.b_lbo, p.b_sbo);
      mma_bf16_ss(tmem, da, db, idesc, ks > 0);
    }
    mma_commit(&mbar);
  }
  const bool ok = mbar_wait(&mbar, 0);
  fence_after_sync();
  if (!ok) { if (threadIdx.x == 0) *p.status = 1; }
  else {
    for (int c0 = 0; c0 < p.N; c0 += 8) {
      float v[8];
      tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int j = 0; j < 8; ++j) p.out[(size_t)threadIdx.x * p.N + c0 + j] = v[j];
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

static unsigned short f2bf(float f) {
  unsigned u; memcpy(&u, &f, 4);
  unsigned r = u + 0x7FFF + ((u >> 16) & 1);
  return (unsigned short)(r >> 16);
}
static float bf2f(unsigned short h) { unsigned u = (unsigned)h << 16; float f; memcpy(&f, &u, 4); return f; }

struct Case {
  const char* name;
  int M, N, K;          // logical GEMM D[M,N] = A[M,K] B[N,K]^T
  int a_mn, b_mn;       // operand majors
  int row_shift;        // K-major A: start shifted by this many rows
  int pair_gap;         // K-major A: rows between the two K chunks of an MMA (0 = normal chunk stride)
  bool dump_layout;
};

int main() {
  std::vector<Case> cases = {
      {"kmajor M128 N16 K32", 128, 16, 32, 0, 0, 0, 0, false},
      {"kmajor M128 N16 K32 shift3", 128, 16, 32, 0, 0, 3, 0, false},
      {"kmajor M128 N16 K32 shift37", 128, 16, 32, 0, 0, 37, 0, false},
      {"kmajor M128 N32 K64", 128, 32, 64, 0, 0, 0, 0, false},
      {"kmajor M128 N48 K16", 128, 48, 16, 0, 0, 1, 0, false},
      {"kmajor paired chunks, LBO = 131 rows", 128, 16, 32, 0, 0, 2, 131, false},
      {"kmajor paired chunks, LBO = 34 rows (conv-like shared image)", 128, 16, 32, 0, 0, 2, -34, false},
      {"mnmajor A,B M128 N32 K64", 128, 32, 64, 1, 1, 0, 0, false},
      {"mnmajor A,B M128 N16 K32", 128, 16, 32, 1, 1, 0, 0, false},
      {"kmajor M64 N8 K16 (layout dump)", 64, 8, 16, 0, 0, 0, 0, true},
      {"kmajor M64 N16 K16 (layout dump)", 64, 16, 16, 0, 0, 0, 0, true},
  };
  int fails = 0;
  for (const Case& c : cases) {
    const int R = 320;  // rows available in the A image (room for shifts)
    std::vector<float> A((size_t)c.M * c.K), B((size_t)c.N * c.K);
    srand(7);
    for (auto& v : A) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.0f));
    for (auto& v : B) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.0f));
    int gap = c.pair_gap;
    if (gap < 0) {  // conv-like: both K chunks of an MMA read ONE image at rows m and m+gap
      gap = -gap;
      std::vector<float> img((size_t)(c.K / 16) * 512 * 8);
      for (auto& v : img) v = bf2f(f2bf((rand() % 2001 - 1000) / 1000.0f));
      for (int m = 0; m < c.M; ++m) for (int k = 0; k < c.K; ++k)
        A[(size_t)m * c.K + k] = img[((size_t)(k / 16) * 512 + m + ((k / 8) & 1 ? gap : 0)) * 8 + (k % 8)];
    }
    if (c.dump_layout) {
      for (int m = 0; m < c.M; ++m) for (int k = 0; k < c.K; ++k) A[(size_t)m * c.K + k] = k == 0 ? (float)(m + 1) : 0.f;
      for (int n = 0; n < c.N; ++n) for (int k = 0; k < c.K; ++k) B[(size_t)n * c.K + k] = k == 0 ? 1.f + n / 64.f : 0.f;
    }
    ProbeArgs p{};
    std::vector<unsigned short> ai, bi;
    const int kchunks = c.K / 8;
    if (!c.a_mn) {
      // [chunk][row] x 16 B ; logical (m, k) -> chunk k/8, row m + shift (+ gap for odd chunks when pairing)
      const int chunk_stride_rows = R;
      ai.assign((size_t)kchunks * chunk_stride_rows * 8 + (size_t)(gap + 8) * 8, 0);
      for (int m = 0; m < c.M; ++m) for (int k = 0; k < c.K; ++k) {
        size_t unit;
        if (gap) {  // chunks (2s, 2s+1) live in chunk slot s: even chunk at row, odd chunk at row + gap
          unit = (size_t)(k / 16) * chunk_stride_rows + m + c.row_shift + ((k / 8) & 1 ? gap : 0);
        } else {
          unit = (size_t)(k / 8) * chunk_stride_rows + m + c.row_shift;
        }
        ai[unit * 8 + (k % 8)] = f2bf(A[(size_t)m * c.K + k]);
      }
      p.a_start = c.row_shift * 16;
      p.a_lbo = gap ? gap * 16 : chunk_stride_rows * 16;
      p.a_sbo = 128;
      p.a_kstep = gap ? chunk_stride_rows * 16 : 2 * chunk_stride_rows * 16;
    } else {
      // MN-major from the same [chunk][pixel] tile: M = channels (chunk g = m/8), K = pixels
      // element (m, k) at g*CH + k*16 + (m%8)*2 ; LBO = 8 pixels * 16 B, SBO = chunk stride
      const int CH = R * 16;
      ai.assign((size_t)(c.M / 8) * R * 8, 0);
      for (int m = 0; m < c.M; ++m) for (int k = 0; k < c.K; ++k)
        ai[((size_t)(m / 8) * R + k) * 8 + (m % 8)] = f2bf(A[(size_t)m * c.K + k]);
      p.a_start = 0; p.a_lbo = 128; p.a_sbo = CH; p.a_kstep = 16 * 16;
    }
    if (!c.b_mn) {
      const int rows = c.N;
      bi.assign((size_t)kchunks * rows * 8, 0);
      for (int n = 0; n < c.N; ++n) for (int k = 0; k < c.K; ++k) bi[((size_t)(k / 8) * rows + n) * 8 + (k % 8)] = f2bf(B[(size_t)n * c.K + k]);
      p.b_start = 0; p.b_lbo = rows * 16; p.b_sbo = 128; p.b_kstep = 2 * rows * 16;
    } else {
      const int CH = R * 16;
      bi.assign((size_t)(c.N / 8) * R * 8, 0);
      for (int n = 0; n < c.N; ++n) for (int k = 0; k < c.K; ++k) bi[((size_t)(n / 8) * R + k) * 8 + (n % 8)] = f2bf(B[(size_t)n * c.K + k]);
      p.b_start = 0; p.b_lbo = 128; p.b_sbo = CH; p.b_kstep = 16 * 16;
    }
    p.a_bytes = (int)(ai.size() * 2); p.b_bytes = (int)(bi.size() * 2);
    p.ksteps = c.K / 16; p.M = c.M; p.N = c.N; p.a_mn = c.a_mn; p.b_mn = c.b_mn;
    unsigned char *da, *db; float* dout; int* dstat;
    cudaMalloc(&da, p.a_bytes); cudaMalloc(&db, p.b_bytes); cudaMalloc(&dout, 128 * c.N * sizeof(float)); cudaMalloc(&dstat, 4);
    cudaMemcpy(da, ai.data(), p.a_bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), p.b_bytes, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, 128 * c.N * sizeof(float)); cudaMemset(dstat, 0, 4);
    p.a_img = da; p.b_img = db; p.out = dout; p.status = dstat;
    const int smem = ((p.a_bytes + 1023) / 1024) * 1024 + p.b_bytes + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out((size_t)128 * c.N);
    int stat = -1;
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&stat, dstat, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", c.name, cudaGetErrorString(e)); return 2; }
    if (stat != 0) { printf("[%s] FAIL: mbarrier timeout\n", c.name); ++fails; continue; }
    if (c.dump_layout) {
      printf("[%s] lane -> row held (col 0), 0 = empty:\n", c.name);
      for (int l = 0; l < 128; ++l) printf("%d%s", (int)lroundf(out[(size_t)l * c.N]), (l % 32 == 31) ? "\n" : " ");
      printf("  lane0 cols: "); for (int n = 0; n < c.N; ++n) printf("%.3f ", out[n]); printf("\n");
    } else {
      double maxerr = 0;
      for (int m = 0; m < c.M; ++m) for (int n = 0; n < c.N; ++n) {
        double s = 0;
        for (int k = 0; k < c.K; ++k) s += (double)A[(size_t)m * c.K + k] * B[(size_t)n * c.K + k];
        maxerr = fmax(maxerr, fabs(s - out[(size_t)m * c.N + n]));
      }
      const bool ok = maxerr < 1e-3;
      printf("[%s] %s max|err| = %.3g\n", c.name, ok ? "PASS" : "FAIL", maxerr);
      if (!ok) ++fails;
    }
    cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dstat);
  }
  printf("tc_probe: %d failing case(s)\n", fails);
  return fails ? 1 : 0;
}
