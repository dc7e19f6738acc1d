// cuda_emu.h - TEST INFRASTRUCTURE ONLY.
//
// A minimal functional simulator of the CUDA execution model so that the very same
// kernel sources under trustedai-cl-vae-ad_b200/csrc/ can be compiled with plain g++
// (-DKCVAE_EMU) and their index arithmetic / reductions / orchestration checked on a
// GPU-less machine (and under ASan).  Every CUDA thread of a block runs as a ucontext
// fiber; __syncthreads() and warp shuffles are cooperative yields.  Blocks run one after
// another.  tcgen05 / TMA kernels are NOT emulated (they are compiled out).
//
// The product never loads the emulated library: trustedai-cl-vae-ad_b200/_lib.py only
// opens libkcvae.so and raises if it or the GPU is missing.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <algorithm>
using std::min;
using std::max;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __shared__ static
#define __constant__ static const
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

namespace kcemu {
struct State {
  uint3 tid, bid;
  dim3 bdim, gdim;
  ucontext_t sched;
  std::vector<ucontext_t> ctx;
  std::vector<int> st;  // 0 running, 1 at syncthreads, 2 done
  char* stacks = nullptr;
  size_t stack_sz = 96 * 1024;
  int cur = 0, n = 0, sync_gen = 0;
  uint64_t shfl[1024];
  unsigned char* dyn_smem = nullptr;
  size_t dyn_cap = 0;
  const std::function<void()>* body = nullptr;
};
inline State& S() { static State s; return s; }

inline void yield_() { State& s = S(); swapcontext(&s.ctx[s.cur], &s.sched); }
inline void trampoline() {
  State& s = S();
  (*s.body)();
  s.st[s.cur] = 2;
  swapcontext(&s.ctx[s.cur], &s.sched);
}
inline void set_tid(int i) {
  State& s = S();
  s.tid.x = i % s.bdim.x;
  s.tid.y = (i / s.bdim.x) % s.bdim.y;
  s.tid.z = i / (s.bdim.x * s.bdim.y);
}
inline void launch(dim3 g, dim3 b, size_t smem, const std::function<void()>& body) {
  State& s = S();
  int n = (int)(b.x * b.y * b.z);
  if (n <= 0 || n > 1024) { fprintf(stderr, "kcemu: bad block size %d\n", n); abort(); }
  if (!s.stacks) s.stacks = (char*)malloc(s.stack_sz * 1024);
  if (smem > s.dyn_cap) { free(s.dyn_smem); s.dyn_smem = (unsigned char*)aligned_alloc(1024, (smem + 1023) / 1024 * 1024); s.dyn_cap = smem; }
  s.bdim = b; s.gdim = g; s.n = n; s.body = &body;
  s.ctx.resize(n); s.st.resize(n);
  for (unsigned bz = 0; bz < g.z; ++bz) for (unsigned by = 0; by < g.y; ++by) for (unsigned bx = 0; bx < g.x; ++bx) {
    s.bid = uint3{bx, by, bz};
    s.sync_gen = 0;
    for (int i = 0; i < n; ++i) {
      getcontext(&s.ctx[i]);
      s.ctx[i].uc_stack.ss_sp = s.stacks + (size_t)i * s.stack_sz;
      s.ctx[i].uc_stack.ss_size = s.stack_sz;
      s.ctx[i].uc_link = &s.sched;
      makecontext(&s.ctx[i], (void (*)())trampoline, 0);
      s.st[i] = 0;
    }
    int alive = n;
    while (alive > 0) {
      for (int i = 0; i < n; ++i) {
        if (s.st[i] == 2) continue;
        s.cur = i; set_tid(i);
        swapcontext(&s.sched, &s.ctx[i]);
        if (s.st[i] == 2) --alive;
      }
      int waiting = 0;
      for (int i = 0; i < n; ++i) waiting += (s.st[i] == 1);
      if (alive > 0 && waiting == alive) { s.sync_gen++; for (int i = 0; i < n; ++i) if (s.st[i] == 1) s.st[i] = 0; }
    }
  }
  s.body = nullptr;
}
inline int lin_tid() { State& s = S(); return s.cur; }
template <typename T> inline T shfl_idx(T v, int src_lane) {
  State& s = S();
  static_assert(sizeof(T) <= 8, "shfl width");
  int me = s.cur, base = me & ~31;
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  s.shfl[me] = raw;
  yield_();
  int src = base + (src_lane & 31);
  T r = v;
  if (src < s.n) { uint64_t q = s.shfl[src]; memcpy(&r, &q, sizeof(T)); }
  yield_();
  return r;
}
}  // namespace kcemu

#define threadIdx (kcemu::S().tid)
#define blockIdx (kcemu::S().bid)
#define blockDim (kcemu::S().bdim)
#define gridDim (kcemu::S().gdim)

static inline void __syncthreads() {
  kcemu::State& s = kcemu::S();
  int g = s.sync_gen;
  s.st[s.cur] = 1;
  while (s.sync_gen == g) kcemu::yield_();
}
static inline void __syncwarp(unsigned = 0xffffffffu) { kcemu::yield_(); }
static inline void __threadfence() {}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return kcemu::shfl_idx(v, (kcemu::lin_tid() & 31) ^ m); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  int l = kcemu::lin_tid() & 31;
  return kcemu::shfl_idx(v, l + d < 32 ? l + d : l);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return kcemu::shfl_idx(v, src); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }

// ---- runtime API subset ---------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }

#define KC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kcemu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })
#define KC_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(kcemu::S().dyn_smem)
