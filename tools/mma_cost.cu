// Development aid: cycles per tcgen05.mma (kind::f16, bf16 -> fp32, K = 16, operands in shared memory) as a function of the
// shape, the operand majors and the descriptor strides.  One CTA per SM; one thread issues `iters` MMAs back to back into
// the same accumulator, commits, waits; clock64 around it.  The numbers decide how the kernels lay out their operands
// ("MMA cost table").  Build + run (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I trustedai-cl-vae-ad_b200/csrc tools/mma_cost.cu -o gpurun_out/mma_cost && gpurun_out/mma_cost
#include <cstdio>
#include <vector>
#include "tc_common.cuh"

using namespace kc::tc;

struct Case { int M, N, a_mn, b_mn; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; uint32_t a_step, b_step; const char* note; };

__global__ void __launch_bounds__(128, 1) mma_cost_kernel(Case c, int iters, long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16_f32(c.M, c.N, c.a_mn, c.b_mn);
    const uint64_t da0 = make_desc_kmajor_noswz(smem_u32(smem), c.a_lbo, c.a_sbo);
    const uint64_t db0 = make_desc_kmajor_noswz(smem_u32(smem + 96 * 1024), c.b_lbo, c.b_sbo);
    long long t0 = 0;
    for (int rep = 0; rep < 2; ++rep) {           // first pass warms up
      t0 = clock64();
#pragma unroll 4
      for (int i = 0; i < iters; ++i) {
        const uint32_t k = (uint32_t)(i & 15);
        if (leader) mma_bf16_ss(tmem, desc_advance(da0, k * c.a_step), desc_advance(db0, k * c.b_step), idesc, 1);
      }
      if (leader) mma_commit(&bar);
      __syncwarp();
      mbar_wait(&bar, (uint32_t)rep & 1u, 1u << 24);
    }
    const long long t1 = clock64();
    if (leader) cycles[blockIdx.x] = t1 - t0;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

int main() {
  const uint32_t CH = 1088 * 16;        // plane stride of the specialised kernels (34 x 32 halo pixels)
  std::vector<Case> cases = {
      {128, 32, 0, 0, 128 * 16, 128, 32 * 16, 128, 128, 0, "K-major, packed 8-row groups (forward kernels)"},
      {128, 16, 0, 0, 128 * 16, 128, 16 * 16, 128, 128, 0, "K-major N=16"},
      {128, 64, 0, 0, 128 * 16, 128, 64 * 16, 128, 128, 0, "K-major N=64"},
      {128, 128, 0, 0, 128 * 16, 128, 128 * 16, 128, 128, 0, "K-major N=128"},
      {128, 256, 0, 0, 128 * 16, 128, 256 * 16, 128, 128, 0, "K-major N=256"},
      {64, 32, 1, 1, 128, 512, 128, CH, 16, 16, "MN-major M=64, A groups 512 B apart (old out-layer wgrad)"},
      {64, 32, 1, 1, 128, 528, 128, CH, 16, 16, "MN-major M=64, A groups 528 B apart"},
      {64, 32, 1, 1, 128, 528, 128, CH + 16, 16, 16, "MN-major M=64, A 528, B CH+16"},
      {128, 32, 1, 1, 128, 512, 128, CH, 16, 16, "MN-major M=128, A groups 512 B apart (new out-layer wgrad)"},
      {128, 32, 1, 1, 128, 528, 128, CH, 16, 16, "MN-major M=128, A groups 528 B apart"},
      {128, 32, 1, 1, 128, 544, 128, CH, 16, 16, "MN-major M=128, A groups 544 B apart"},
      {128, 32, 1, 1, 128, 640, 128, CH, 16, 16, "MN-major M=128, A groups 640 B apart"},
      {128, 32, 1, 1, 128, 528, 128, CH + 16, 16, 16, "MN-major M=128, A 528, B CH+16"},
      {128, 32, 1, 1, 128, 128, 128, 128, 0, 0, "MN-major M=128, groups packed 128 B apart (16 pixels only)"},
      {128, 32, 1, 1, 128, 256, 128, 256, 16, 16, "MN-major M=128, groups 256 B apart"},
      {128, 32, 1, 0, 128, 528, 32 * 16, 128, 16, 0, "A MN-major 528, B K-major"},
      {128, 32, 0, 1, 128 * 16, 128, 128, CH, 0, 16, "A K-major, B MN-major"},
      {128, 64, 1, 1, 128, 528, 128, 4368, 16, 16, "MN-major M=128 N=64, A 528, B 4368"},
      {128, 128, 1, 1, 128, 528, 128, 4368, 16, 16, "MN-major M=128 N=128, A 528, B 4368"},
      {128, 256, 1, 1, 128, 528, 128, 1040, 16, 16, "MN-major M=128 N=256, A 528, B 1040"},
      {128, 256, 1, 1, 128, 512, 128, 1024, 16, 16, "MN-major M=128 N=256, A 512, B 1024"},
      {64, 256, 1, 1, 128, 528, 128, 1040, 16, 16, "MN-major M=64 N=256, A 528, B 1040"},
  };
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  printf("%-70s %5s %5s %10s\n", "case", "M", "N", "cyc/MMA");
  for (const Case& c : cases) {
    mma_cost_kernel<<<148, 128, 200 * 1024>>>(c, iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", c.note, cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-70s %5d %5d %10.1f\n", c.note, c.M, c.N, (double)mx / iters);
  }
  return 0;
}
