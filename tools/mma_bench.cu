// mma_bench.cu - issue-rate lab for the tcgen05 shapes the kcvae kernels use (development tool).
//
// One thread issues a train of tcgen05.mma (kind::f16, M=128, K=16, SS operands in the no-swizzle
// canonical layout with linear rows) against a resident shared-memory tile and measures clock64()
// from the first issue to the completion of the commit.  Answers: what does ONE small-N MMA cost
// when its A operand comes from shared memory, and how does that depend on N, on the chunk-plane
// stride (bank mapping of the two K chunks) and on the row shift (3x3 taps)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I trustedai-cl-vae-ad_b200/csrc tools/mma_bench.cu -o build/mma_bench
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace kc::tc;

struct BenchArgs {
  int n;            // UMMA N
  int count;        // MMAs in the train
  int ch_bytes;     // chunk-plane stride (LBO of A)
  int shifts;       // 1: every MMA reads a different row shift (conv taps); 0: same start
  int mtiles;       // distinct 128-row M-tiles cycled through
  int accumulate;   // 1: all MMAs accumulate into one TMEM tile (dependent); 0: rotate over TMEM columns
  int reps;
  long long* cycles;  // [grid] best of reps
  long long* ld_cycles;  // [grid][2]: average / max cycles of one tcgen05.ld.x8 + wait issued by another warp meanwhile
  int* status;
};

__global__ void __launch_bounds__(128) bench_kernel(BenchArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  __shared__ int done_flag;
  if (threadIdx.x == 0) done_flag = 0;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (threadIdx.x == 32) { mbar_init(&mbar, 1); fence_mbar_init(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    // same issue idiom as the product kernels: converged warp, one elected lane, unrolled taps
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16_f32(128, p.n);
    const uint32_t a_base = smem_u32(smem);
    const uint32_t b_base = a_base + 160 * 1024;
    const uint64_t db = make_desc_kmajor_noswz(b_base, p.n * 16, 128);
    const uint64_t da0 = make_desc_kmajor_noswz(a_base, p.ch_bytes, 128);
    long long best = 1ll << 60;
    const int per_mt = p.count / p.mtiles;      // MMAs per M-tile (18 in the product kernel)
    for (int r = 0; r < p.reps; ++r) {
      const long long t0 = clock64();
#pragma unroll 1
      for (int mt = 0; mt < p.mtiles; ++mt) {
        const uint32_t d = p.accumulate ? tmem : tmem + (uint32_t)((mt * p.n) & 511 & ~(p.n - 1));
        const uint64_t da_mt = desc_advance(da0, (uint32_t)(mt * 128));
        if (per_mt == 18) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t shift = p.shifts ? (uint32_t)((2 - tap / 3) * 32 + (2 - tap % 3)) : 0u;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t da = desc_advance(da_mt, (uint32_t)(2 * ks) * (uint32_t)(p.ch_bytes / 16) + shift);
              if (leader) mma_bf16_ss(d, da, db, idesc, 1);
            }
          }
        } else {
#pragma unroll 1
          for (int k = 0; k < per_mt; ++k)
            if (leader) mma_bf16_ss(d, da_mt, db, idesc, 1);
        }
      }
      if (leader) mma_commit(&mbar);
      __syncwarp();
      if (!mbar_wait(&mbar, r & 1)) { if (leader) *p.status = 1; break; }
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (leader) { p.cycles[blockIdx.x] = best; *reinterpret_cast<volatile int*>(&done_flag) = 1; }
  }
  if (warp == 1) {
    // TMEM read latency seen by an epilogue warp while the MMA train runs (columns 448.. are not MMA targets
    // in the rotating cases with N <= 64; the values are irrelevant)
    volatile int* flag = &done_flag;
    long long sum = 0, mx = 0; int cnt = 0;
    float acc = 0.f;
    while (!*flag && cnt < 100000) {
      const long long t0 = clock64();
      float v[8];
      tmem_ld8(tmem + ((uint32_t)32 << 16) + 448, v);
      { uint32_t sink; asm volatile("mov.b32 %0, %1;" : "=r"(sink) : "f"(v[0] + v[7])); acc += __uint_as_float(sink); }   // forces the scoreboard wait before the clock read
      const long long t1 = clock64() + (acc == 123.456f ? 1 : 0);
      sum += t1 - t0; mx = t1 - t0 > mx ? t1 - t0 : mx; ++cnt;
    }
    if ((threadIdx.x & 31) == 0) { p.ld_cycles[blockIdx.x * 2] = cnt ? sum / cnt : 0; p.ld_cycles[blockIdx.x * 2 + 1] = mx; }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  struct Case { const char* name; int n, count, ch, shifts, mtiles, acc; };
  const int CH = 34 * 32 * 16;   // the kernels' chunk plane (17408 B)
  std::vector<Case> cases = {
      {"N16  8 M-tiles x 18 (tail phase B / out conv)", 16, 144, CH, 1, 8, 0},
      {"N16  8 x 18 no shift", 16, 144, CH, 0, 8, 0},
      {"N16  8 x 18 CH+16", 16, 144, CH + 16, 1, 8, 0},
      {"N16  1 M-tile x 144 same operands", 16, 144, CH, 0, 1, 1},
      {"N8   8 x 18", 8, 144, CH, 1, 8, 0},
      {"N32  8 x 18", 32, 144, CH, 1, 8, 0},
      {"N32  9 x 2 (C2I)", 32, 18, CH, 0, 9, 0},
      {"N64  8 x 18", 64, 144, CH, 1, 8, 0},
      {"N128 8 x 18", 128, 144, CH, 1, 8, 0},
      {"N256 8 x 18", 256, 144, CH, 1, 8, 0},
      {"N256 1 x 144 same operands", 256, 144, CH, 0, 1, 1},
  };
  long long *dcyc, *dld; int* dstat;
  cudaMalloc(&dld, 2 * 148 * sizeof(long long));
  cudaMalloc(&dcyc, 148 * sizeof(long long)); cudaMalloc(&dstat, 4);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1}) {
    for (const Case& c : cases) {
      BenchArgs p{c.n, c.count, c.ch, c.shifts, c.mtiles, c.acc, 50, dcyc, dld, dstat};
      cudaMemset(dstat, 0, 4);
      bench_kernel<<<grid, 128, smem>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
      std::vector<long long> cyc(grid);
      int stat = 0;
      cudaMemcpy(cyc.data(), dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      cudaMemcpy(&stat, dstat, 4, cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (long long v : cyc) mx = v > mx ? v : mx;
      long long ld[2];
      cudaMemcpy(ld, dld, sizeof(ld), cudaMemcpyDeviceToHost);
      printf("grid %3d  %-52s %s  %7lld cycles  %6.1f cycles/MMA  (floor %d)   concurrent tcgen05.ld.x8: avg %lld max %lld cycles\n", grid, c.name,
             stat ? "TIMEOUT" : "ok", mx, (double)mx / c.count, 128 * c.n / 256, ld[0], ld[1]);
    }
  }
  return 0;
}
