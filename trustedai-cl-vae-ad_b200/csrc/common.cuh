// common.cuh - shared definitions for the kcvae kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cstdio>

#ifdef KCVAE_EMU
#include "cuda_emu.h"  // tests/emu: g++ functional simulation of these kernels (tests only)
#else
#include <cuda_runtime.h>
#define KC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define KC_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char name##_raw_[];    \
  type* name = reinterpret_cast<type*>(name##_raw_)
#endif

namespace kc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// counted by every launcher; kcvae_launch_count() reports it (bench.py "gpu_launches")
extern int64_t g_launches;

// ---- per-launch device timing (bench.py roofline): CUDA events on the launching stream ----
// model.cu sets g_tag to the layer/role before calling a launcher; when profiling is enabled
// each launcher brackets its kernels with two events keyed "<tag>/<kernel>".
extern const char* g_tag;
extern bool g_prof_on;
void prof_begin(const char* kernel, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(const char* kernel, cudaStream_t s) : st(s), on(g_prof_on) { if (on) prof_begin(kernel, st); }
  ~ProfScope() { if (on) prof_end(st); }
};

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
// grid for a grid-stride kernel: enough blocks for `work` items, capped at `waves` full
// waves of `per_sm` resident blocks on the 148 SMs
static inline int grid_for(int64_t work, int block, int per_sm = 8, int waves = 4) {
  int64_t need = (work + block - 1) / block;
  int64_t cap = (int64_t)kNumSMs * per_sm * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ---- warp / block reductions -----------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result valid in thread 0 (blockDim.x multiple of 32, <= 1024).
// `scratch` is a caller-provided __shared__ T[32]; safe to call repeatedly.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  T r = (T)0;
  if (wid == 0) {
    r = lane < nw ? scratch[lane] : (T)0;
    r = warp_sum(r);
  }
  return r;
}

// ---- Philox4x32-10 counter RNG (device eps / image noise / Glorot init) -----------------
struct Philox {
  uint32_t k0, k1;
  __host__ __device__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ void operator()(uint64_t ctr, uint32_t stream, uint32_t out[4]) const {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0x6b637661u;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
      c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};
__device__ __forceinline__ float u01(uint32_t u) { return ((u >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// Box-Muller: two uniforms -> two standard normals
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  float r = sqrtf(-2.0f * logf(u01(a)));
  float s, c;
  sincosf(6.283185307179586f * u01(b), &s, &c);
  n0 = r * c; n1 = r * s;
}

}  // namespace kc
