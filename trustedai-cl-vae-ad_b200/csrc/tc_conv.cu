// tc_conv.cu - tensor-core (tcgen05 + TMEM + TMA) 3x3 stride-1 implicit-GEMM convolution
// for the decoder output layer (src/abstract_cvae.py:87-89 Conv2DTranspose k3 s1 'same',
// SURVEY A4) - 51 % of the forward MACs of the README config.
//
//   x_hat[n,y,x,co] = sigmoid(b[co] + sum_{kh,kw,ci} a[n, y+1-kh, x+1-kw, ci] * W[kh,kw,co,ci])
//
// Implicit GEMM without im2col.  A persistent CTA per SM walks 32x30-pixel output tiles:
//   * Cin/8 TMA boxes (cp.async.bulk.tensor.4d) bring the 34x32-pixel bf16 halo tile into
//     shared memory as [8-channel chunk][pixel] x 16 B - TMA's zero fill supplies the SAME padding;
//   * that layout is the no-swizzle K-major UMMA operand with linear rows, so each of the
//     9 taps is the SAME tile read through a descriptor whose start address is shifted by
//     (dy*32 + dx) pixels: 8 M-tiles x 9 taps x (Cin/16) tcgen05.mma (M=128, N=16, K=16)
//     accumulate into 8 TMEM accumulators (128 columns), double buffered (256 columns);
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b), add bias, apply the sigmoid and
//     store x_hat; TMA / MMA / epilogue overlap through mbarrier pipelines.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread),
// warps 2..5 = epilogue (TMEM lane groups (warp & 3)).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <cuda.h>

#include "kernels.h"
#include "tc_common.cuh"
#include "tc_conv.h"

namespace kc {

using namespace tc;

namespace {

constexpr int TW = 30, TR = 32;            // output tile (cols, rows)
constexpr int PW = TW + 2, PR = TR + 2;    // halo tile
constexpr int NPIX = PR * PW;              // 1088 halo pixels
constexpr int MT = (TR * PW) / 128;        // 8 M-tiles of 128 padded-row positions
constexpr int NPAD = 16;                   // UMMA N (Cout padded)
constexpr int kStages = 2;
constexpr int kThreads = 192;
constexpr int kSubD = 2;                   // tc_out_dgrad: epilogue warps per TMEM lane group
constexpr int kThreadsD = 64 + 4 * kSubD * 32;
constexpr int kThreadsT = 64 + 8 * 32 + 4 * 32;   // fused tail: TMA, MMA, 8 phase-A epilogue warps, 4 phase-B epilogue warps
constexpr int kThreadsE = 320;            // kernels with a per-tile epilogue: 2 + 8 warps (two epilogue warps per TMEM lane group)
static_assert(TR * PW == MT * 128, "tile must be a whole number of M=128 tiles");

// 3-D tiled TMA load from a [planes][H][W*8] bf16 view (chunk-planar activations, 8-channel NHWC tensors): box (pixels x 8
// channels = contiguous bytes, rows, 1 plane) -> a [pixel] x 16 B chunk plane in shared memory.  The box start must be
// 16-byte aligned in global memory (an 8-byte aligned inner coordinate raises "illegal instruction").
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* mbar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(mbar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(mbar)) : "memory");
}

static void sum_partials(const float* partial, int nparts, int E, float* out, cudaStream_t st);

// Development aid (-DKCVAE_TAIL_TIMING): every warp of CTA 0 accumulates the cycles it spends in each mbarrier
// wait and its total time; the launcher prints them.  Compiled out of the product.
#ifdef KCVAE_TAIL_TIMING
__device__ long long g_tail_dbg[24 * 16];
#define TAIL_TIMING_DECL long long tw_[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long tstart_ = clock64();
#define TSPAN_BEGIN const long long ts_ = clock64();
#define TSPAN_END(idx) tw_[idx] += clock64() - ts_;
#define TWAIT(idx, bar, par) [&] { const long long t_ = clock64(); const bool r_ = mbar_wait(bar, par); tw_[idx] += clock64() - t_; return r_; }()
#define TAIL_TIMING_DUMP if (blockIdx.x == 0 && lane == 0) { for (int i_ = 0; i_ < 14; ++i_) g_tail_dbg[warp * 16 + i_] = tw_[i_]; g_tail_dbg[warp * 16 + 14] = clock64() - tstart_; }
#else
#define TAIL_TIMING_DECL
#define TSPAN_BEGIN
#define TSPAN_END(idx)
#define TWAIT(idx, bar, par) mbar_wait(bar, par)
#define TAIL_TIMING_DUMP
#endif


struct OutConvParams {
  const __nv_bfloat16* wimg;  // [9 taps][CIN/16][2 chunks][NPAD][8] bf16 (tc_prep_out_weights)
  const float* bias;          // [Cout]
  float* xhat;                // [B,H,W,Cout] fp32
  int B, H, W, Cout;
  int tiles_y, tiles_x, num_tiles;
  int apply_sigmoid;
  int* error_flag;            // set to 1 if a bounded barrier wait expires
};

template <int CIN>
__global__ void __launch_bounds__(kThreadsE, 1)
tc_out_conv_kernel(const __grid_constant__ CUtensorMap tmap, OutConvParams p) {
  constexpr int KC = CIN / 8;                        // 16-byte channel chunks
  constexpr int KS = CIN / 16;                       // K=16 steps per tap
  constexpr uint32_t CH = NPIX * 16;                 // bytes per chunk plane
  constexpr uint32_t TILE_BYTES = KC * CH;           // one halo tile
  constexpr uint32_t WB = 9 * KS * 2 * NPAD * 16;    // weight image bytes
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* s_tile = smem;                                      // kStages x (TILE_BYTES + 128)
  unsigned char* s_w = smem + kStages * (TILE_BYTES + 128);          // weights
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[NPAD];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup -------------------------------------------------------------------
  for (int i = threadIdx.x; i < (int)(WB / 16); i += kThreadsE)
    reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(p.wimg)[i];
  if (threadIdx.x < NPAD) s_bias[threadIdx.x] = (int)threadIdx.x < p.Cout ? p.bias[threadIdx.x] : 0.f;
  // the 2 units past each stage's tile are only ever read into discarded columns: zero them
  if (threadIdx.x < kStages * 8) {
    const int s = threadIdx.x / 8, j = threadIdx.x % 8;
    reinterpret_cast<uint4*>(s_tile + s * (TILE_BYTES + 128) + TILE_BYTES)[j] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
    fence_mbar_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ========================================
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        if (!mbar_wait(&empty_bar[s], ph ^ 1)) { *p.error_flag = 1; break; }
        const int n = t / (p.tiles_y * p.tiles_x);
        const int rem = t % (p.tiles_y * p.tiles_x);
        const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
        mbar_expect_tx(&full_bar[s], TILE_BYTES);
#pragma unroll
        for (int c = 0; c < KC; ++c)   // one box per 8-channel chunk: smem = [chunk][row][col] x 16 B
          tma_load_3d(s_tile + s * (TILE_BYTES + 128) + c * CH, &tmap, &full_bar[s], (tx * TW - 1) * 8, ty * TR - 1, n * KC + c);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==========================================
    // the whole warp runs the (uniform) loop; only the elected lane issues tcgen05 instructions
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16_f32(128, NPAD);
    const uint64_t db0 = make_desc_kmajor_noswz(smem_u32(s_w), NPAD * 16, 128);
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages, a = it & 1;
      const uint32_t ph = (it / kStages) & 1, aph = (it >> 1) & 1;
      if (!mbar_wait(&tempty_bar[a], aph ^ 1)) { if (leader) *p.error_flag = 1; break; }   // epilogue drained TMEM stage
      if (!mbar_wait(&full_bar[s], ph)) { if (leader) *p.error_flag = 1; break; }          // TMA landed
      fence_after_sync();
      const uint64_t da0 = make_desc_kmajor_noswz(smem_u32(s_tile + s * (TILE_BYTES + 128)), CH, 128);
#pragma unroll 2
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t d_tmem = tmem + (uint32_t)(a * MT * NPAD + mt * NPAD);
        const uint64_t da_mt = desc_advance(da0, (uint32_t)(mt * 128));
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t shift = (uint32_t)((2 - tap / 3) * PW + (2 - tap % 3));   // flipped taps (A4)
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t da = desc_advance(da_mt, (uint32_t)(2 * ks) * (CH / 16) + shift);
            const uint64_t db = desc_advance(db0, (uint32_t)((tap * KS + ks) * 2 * NPAD));
            if (leader) mma_bf16_ss(d_tmem, da, db, idesc, (tap | ks) != 0);
          }
        }
      }
      if (leader) {
        mma_commit(&empty_bar[s]);    // smem stage reusable once these MMAs have read it
        mma_commit(&tfull_bar[a]);    // accumulators ready for the epilogue
      }
      __syncwarp();
    }
  } else {
    // ================================ epilogue warps ======================================
    const int lg = warp & 3;                      // TMEM lane group this warp may access
    const int half = (warp - 2) >> 2;             // two warps per lane group take alternate M-tiles
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int n = t / (p.tiles_y * p.tiles_x);
      const int rem = t % (p.tiles_y * p.tiles_x);
      const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
      if (!mbar_wait(&tfull_bar[a], aph)) { if (lane == 0) *p.error_flag = 1; break; }
      fence_after_sync();
#pragma unroll 1
      for (int mt = half; mt < MT; mt += 2) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(a * MT * NPAD + mt * NPAD), v);
        const int q = mt * 128 + lg * 32 + lane;  // padded-row position
        const int r = q / PW, c = q % PW;
        const int oy = ty * TR + r, ox = tx * TW + c;
        if (c < TW && oy < p.H && ox < p.W) {
          float* o = p.xhat + (((int64_t)n * p.H + oy) * p.W + ox) * p.Cout;
#pragma unroll
          for (int co = 0; co < 8; ++co) {
            if (co < p.Cout) {
              float y = v[co] + s_bias[co];
              if (p.apply_sigmoid) y = 1.0f / (1.0f + __expf(-y));
              o[co] = y;
            }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// W [3,3,Cout,Cin] fp32 -> bf16 UMMA B-operand image [tap][ks][chunk(2)][n(NPAD)][8]
__device__ __forceinline__ void tc_prep_out_weights_body(const float* w, int Cout, int Cin, __nv_bfloat16* img) {
  const int KS = Cin / 16;
  const int total = 9 * KS * 2 * NPAD * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % 8;
    const int n = (i / 8) % NPAD;
    const int kc = (i / (8 * NPAD)) % 2;
    const int ks = (i / (16 * NPAD)) % KS;
    const int tap = i / (16 * NPAD * KS);
    const int ci = ks * 16 + kc * 8 + j;
    const float v = n < Cout ? w[((int64_t)tap * Cout + n) * Cin + ci] : 0.f;
    img[i] = __float2bfloat16(v);
  }
}

// Fused tail, nine-tap form with the three horizontal taps in the N dimension (Cin = 32): B operand
// [kh 3][ks 2][chunk 2][n = kw*8 + co (32)][8]; one MMA then yields, per pixel q, the three partial sums
// T_kw[q][co] = sum_ci a[q + (2-kh) rows][ci] W[kh][kw][co][ci], and the epilogue adds T_2[q] + T_1[q+1] + T_0[q+2]
// with two warp shuffles (an M-tile row is one 32-pixel halo row = one warp).
constexpr int TAILB_N = 32;
__device__ __forceinline__ void tc_prep_tail_kw_weights_body(const float* w, int Cout, int Cin, __nv_bfloat16* img) {
  const int total = 3 * 2 * 2 * TAILB_N * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % 8, n = (i / 8) % TAILB_N, kc = (i / (8 * TAILB_N)) % 2, ks = (i / (16 * TAILB_N)) % 2, kh = i / (32 * TAILB_N);
    const int kw = n >> 3, co = n & 7, ci = ks * 16 + kc * 8 + j;
    const float v = (kw < 3 && co < Cout && ci < Cin) ? w[((int64_t)(kh * 3 + kw) * Cout + co) * Cin + ci] : 0.f;
    img[i] = __float2bfloat16(v);
  }
}

// C2I image of the same weights (fused tail, Cout <= 3): B operand [N = 32 rows n = tap*Cout + co][K = 32 ci],
// K-major canonical units [kchunk 4][n 32][8]; n = tap * 3 + co whatever Cout is (missing channels are zero columns)
__device__ __forceinline__ void tc_prep_tail_c2i_weights_body(const float* w, int Cout, int Cin, __nv_bfloat16* img) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * 32 * 8; i += gridDim.x * blockDim.x) {
    const int j = i % 8, n = (i / 8) % 32, kchunk = i / 256;
    const int ci = kchunk * 8 + j, tap = n / 3, co = n % 3;
    img[i] = __float2bfloat16(tap < 9 && co < Cout && ci < Cin ? w[((int64_t)tap * Cout + co) * Cin + ci] : 0.f);
  }
}

// ============================================================================================
// Output-layer data gradient on tensor cores:
//   g_a[n,y,x,ci] = (a[n,y,x,ci] > 0) * sum_{kh,kw,co} dl[n, y-1+kh, x-1+kw, co] * W[kh,kw,co,ci]
// dl arrives as bf16 NHWC padded to 8 channels (one 16-byte chunk per pixel), so K per tap is 8
// and two taps are PAIRED into each K=16 MMA: the descriptor's leading byte offset is the
// distance between the two taps' pixels.  M=128 pixels, N=32 (Cin of the forward layer).
struct OutDgradParams {
  const __nv_bfloat16* wimg;   // [5 pairs][2][NPAD_D][8]
  const __nv_bfloat16* mask;   // forward activation a (chunk-planar bf16, Cin channels)
  const uint32_t* relu_bits;   // [B,H,W] bit c = (a[.., c] > 0); when set it replaces the reads of `mask`
  float* g_out;                // [B,H,W,Cin] fp32 (or nullptr)
  __nv_bfloat16* g_s2d;        // bf16 space-to-depth, chunk-planar [B][parity 4][Cin/8][H/2][W/2][8] (or nullptr): parity (y&1)*2+(x&1)
  float* chan_partial;         // [grid*4][32] per-warp channel sums of g (bias gradient of the producer) or nullptr
  int B, H, W, Cin;
  int tiles_y, tiles_x, num_tiles;
  int* error_flag;
};
constexpr int NPAD_D = 32;

template <bool BITS>   // BITS: ReLU mask = one 32-bit word per pixel (relu_bits); otherwise read the bf16 activation
__global__ void __launch_bounds__(kThreadsD, 1)
tc_out_dgrad_kernel(const __grid_constant__ CUtensorMap tmap, OutDgradParams p) {
  constexpr uint32_t CH = NPIX * 16;
  constexpr uint32_t TILE_BYTES = CH;                 // one 8-channel chunk plane
  constexpr uint32_t STAGE = TILE_BYTES + 128;
  constexpr uint32_t WB = 5 * 2 * NPAD_D * 16;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* s_tile = smem;
  unsigned char* s_w = smem + kStages * STAGE;
  unsigned char* s_tr = s_w + WB;                     // 2 KB per epilogue warp: transposes the bf16 stores
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TAIL_TIMING_DECL

  for (int i = threadIdx.x; i < (int)(WB / 16); i += kThreadsD)
    reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(p.wimg)[i];
  if (threadIdx.x < kStages * 8) {
    const int s = threadIdx.x / 8, j = threadIdx.x % 8;
    reinterpret_cast<uint4*>(s_tile + s * STAGE + TILE_BYTES)[j] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4 * kSubD); }
    fence_mbar_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        if (!TWAIT(0, &empty_bar[s], ph ^ 1)) { *p.error_flag = 1; break; }
        const int n = t / (p.tiles_y * p.tiles_x);
        const int rem = t % (p.tiles_y * p.tiles_x);
        const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
        mbar_expect_tx(&full_bar[s], TILE_BYTES);
        tma_load_3d(s_tile + s * STAGE, &tmap, &full_bar[s], (tx * TW - 1) * 8, ty * TR - 1, n);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16_f32(128, NPAD_D);
    const uint32_t w_base = smem_u32(s_w);
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages, a = it & 1;
      const uint32_t ph = (it / kStages) & 1, aph = (it >> 1) & 1;
      if (!TWAIT(1, &tempty_bar[a], aph ^ 1)) { if (leader) *p.error_flag = 1; break; }
      if (!TWAIT(2, &full_bar[s], ph)) { if (leader) *p.error_flag = 1; break; }
      fence_after_sync();
      const uint32_t tile_base = smem_u32(s_tile + s * STAGE);
#pragma unroll 2
      for (int mt = 0; mt < MT; ++mt) {
        const uint32_t d_tmem = tmem + (uint32_t)(a * MT * NPAD_D + mt * NPAD_D);
#pragma unroll
        for (int pr = 0; pr < 5; ++pr) {
          const int t0 = 2 * pr, t1 = 2 * pr + 1;
          const uint32_t sh0 = (uint32_t)((t0 / 3) * PW + (t0 % 3));           // un-flipped taps
          const uint32_t sh1 = pr < 4 ? (uint32_t)((t1 / 3) * PW + (t1 % 3)) : sh0 + 1;  // pair 4: zero weights
          const uint64_t da = make_desc_kmajor_noswz(tile_base + (uint32_t)(mt * 128 + sh0) * 16, (sh1 - sh0) * 16, 128);
          const uint64_t db = make_desc_kmajor_noswz(w_base + (uint32_t)(pr * 2 * NPAD_D * 16), NPAD_D * 16, 128);
          if (leader) mma_bf16_ss(d_tmem, da, db, idesc, pr != 0);
        }
      }
      if (leader) {
        mma_commit(&empty_bar[s]);
        mma_commit(&tfull_bar[a]);
      }
      __syncwarp();
    }
  } else {
    // ============ epilogue: kSubD warps per TMEM lane group, each takes every kSubD-th M-tile ============
    // The ReLU mask does not depend on the MMAs.  With the fused forward it is one 32-bit word per
    // pixel (relu_bits); the words of ALL M-tiles this warp owns in the tile are requested before the
    // accumulator wait.  Without it the four 16-byte units of the activation are read instead.
    const int lg = warp & 3;
    const int sub = (warp - 2) >> 2;
    constexpr int NMT = (MT + kSubD - 1) / kSubD;      // M-tiles per warp and tile
    float csum[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) csum[c] = 0.f;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int n = t / (p.tiles_y * p.tiles_x);
      const int rem = t % (p.tiles_y * p.tiles_x);
      const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
      uint32_t bits[NMT];
      if (BITS) {
#pragma unroll
        for (int k = 0; k < NMT; ++k) {
          const int mt = sub + k * kSubD;
          const int q = mt * 128 + lg * 32 + lane;
          const int r = q / PW, c = q % PW;
          const int oy = ty * TR + r, ox = tx * TW + c;
          const bool live = mt < MT && c < TW && oy < p.H && ox < p.W;
          bits[k] = live ? __ldg(p.relu_bits + ((int64_t)n * p.H + oy) * p.W + ox) : 0u;
        }
      }
      bool waited = false;
#pragma unroll
      for (int k = 0; k < NMT; ++k) {
        const int mt = sub + k * kSubD;
        if (mt >= MT) break;
        const int q = mt * 128 + lg * 32 + lane;
        const int r = q / PW, c = q % PW;
        const int oy = ty * TR + r, ox = tx * TW + c;
        const bool live = c < TW && oy < p.H && ox < p.W;
        const int64_t pix = ((int64_t)n * p.H + oy) * p.W + ox;
        uint4 m[4];
        if (!BITS) {
#pragma unroll
          for (int g = 0; g < 4; ++g) m[g] = make_uint4(0, 0, 0, 0);
          if (live) {
            // chunk-planar activation: plane (n, g) holds channels 8g..8g+7 of every pixel as one 16-byte unit
            const int KCm = p.Cin >> 3;
            const int64_t plane = (int64_t)p.H * p.W;
            const uint4* mk = reinterpret_cast<const uint4*>(p.mask) + (int64_t)n * KCm * plane + (int64_t)oy * p.W + ox;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (g < KCm) m[g] = __ldg(mk + g * plane);
          }
        }
        if (!waited) {
          if (!TWAIT(3, &tfull_bar[a], aph)) { if (lane == 0) *p.error_flag = 1; }
          fence_after_sync();
          waited = true;
        }
        const uint32_t ta = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(a * MT * NPAD_D + mt * NPAD_D);
        float4* o = (live && p.g_out) ? reinterpret_cast<float4*>(p.g_out + pix * p.Cin) : nullptr;
        // bf16 space-to-depth output: every lane holds the 64 bytes of ITS pixel; stored as they are, one STG.128
        // touches 32 half-filled sectors in 16 lines and the L1 store path (about one partial sector every two
        // cycles) becomes the bottleneck of the whole kernel.  The four 16-byte chunks go through a per-warp
        // shared-memory scratch instead, so that lane L stores chunk L%4 of pixel L/4 + 8k: four full 128-byte lines
        // per instruction.
        uint4* tr = reinterpret_cast<uint4*>(s_tr + (warp - 2) * 2048);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float v[16];
          { TSPAN_BEGIN
          tmem_ld16(ta + 16 * hh, v);       // whole warp (sync.aligned), dead lanes discard
          { uint32_t sink_; asm volatile("mov.b32 %0, %1;" : "=r"(sink_) : "f"(v[0] + v[15])); (void)sink_; }
          TSPAN_END(4) }
          TSPAN_BEGIN
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int g = hh * 2 + gg;
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              bool pos;
              if (BITS) {
                pos = (bits[k] >> (g * 8 + j)) & 1u;
              } else {
                // bf16 > 0  <=>  sign bit clear and magnitude bits non-zero (dead lanes: mask 0 -> y = 0)
                const uint32_t mw = j < 2 ? m[g].x : (j < 4 ? m[g].y : (j < 6 ? m[g].z : m[g].w));
                const uint32_t h16 = (mw >> ((j & 1) * 16)) & 0xFFFFu;
                pos = (h16 & 0x8000u) == 0 && (h16 & 0x7FFFu) != 0;
              }
              y[j] = pos ? v[gg * 8 + j] : 0.f;
              csum[g * 8 + j] += y[j];
            }
            if (g * 8 < p.Cin) {
              if (o) {
                o[g * 2] = make_float4(y[0], y[1], y[2], y[3]);
                o[g * 2 + 1] = make_float4(y[4], y[5], y[6], y[7]);
              }
              if (p.g_s2d) {
                uint32_t w4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(y[2 * e], y[2 * e + 1]);
                  w4[e] = *reinterpret_cast<uint32_t*>(&b2);
                }
                tr[lane * 4 + (g ^ ((lane >> 1) & 3))] = make_uint4(w4[0], w4[1], w4[2], w4[3]);   // conflict-free both ways
              }
            }
          }
          TSPAN_END(5)
        }
        if (p.g_s2d) {
          __syncwarp();
          // chunk-planar space-to-depth: plane (parity, 8-channel chunk) holds one 16-byte unit per low-res pixel, so
          // the pixels of one column parity form a contiguous run; lanes 0-15 / 16-31 store the even / odd run of chunk kk
          const int xp = lane >> 4, P = 2 * (lane & 15) + xp;   // column of this row whose chunk this lane stores
          const int oxp = tx * TW + P;
          const bool ok = P < TW && oy < p.H && oxp < p.W;
          const int64_t hw = (int64_t)(p.H >> 1) * (p.W >> 1);
          uint4* gdst = reinterpret_cast<uint4*>(p.g_s2d) + ((int64_t)n * 16 + ((oy & 1) * 2 + (oxp & 1)) * 4) * hw +
                        (int64_t)(oy >> 1) * (p.W >> 1) + (oxp >> 1);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint4 u = tr[P * 4 + (kk ^ ((P >> 1) & 3))];
            if (ok) gdst[kk * hw] = u;
          }
          __syncwarp();                                       // scratch is reused by the next M-tile
        }
      }
      if (!waited) {
        if (!mbar_wait(&tfull_bar[a], aph)) { if (lane == 0) *p.error_flag = 1; }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
    }
    if (p.chan_partial) {   // deterministic: fixed per-thread order, then a fixed shuffle tree
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float t = csum[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) p.chan_partial[((int64_t)blockIdx.x * (4 * kSubD) + sub * 4 + lg) * 32 + c] = t;
      }
    }
  }
  TAIL_TIMING_DUMP
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// W [3,3,Cout,Cin] fp32 -> paired-tap B image [pair][chunk(2)][n = ci (32)][8 = co]
__device__ __forceinline__ void tc_prep_dgrad_weights_body(const float* w, int Cout, int Cin, __nv_bfloat16* img) {
  const int total = 5 * 2 * NPAD_D * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % 8;
    const int n = (i / 8) % NPAD_D;
    const int kc = (i / (8 * NPAD_D)) % 2;
    const int pr = i / (16 * NPAD_D);
    const int tap = 2 * pr + kc;
    const float v = (tap < 9 && j < Cout && n < Cin) ? w[((int64_t)tap * Cout + j) * Cin + n] : 0.f;
    img[i] = __float2bfloat16(v);
  }
}

// ============================================================================================
// Output-layer weight gradient on tensor cores (MN-major operands, K = pixels):
//   dW[kh,kw,co,ci] = sum_{n,y,x} dl[n,y,x,co] * a[n, y+1-kh, x+1-kw, ci]
// Per 32x30 tile ONE MMA (M=128, N=32, K=16) per 16 halo pixels covers all nine taps:
//   B operand = the bf16 halo tile of `a` (TMA, [chunk plane][row][32 pixels] x 16 B, read MN-major: N = 32 channels =
//     4 chunk planes, K = 16 consecutive pixels of a halo row);
//   A operand = THREE zero-padded copies of the tile of dl (8 channels per pixel = one 16-byte unit), copy kw shifted
//     by kw columns, stored row-interleaved: unit ((R*3 + kw)*32 + c) = dl[R - 2][c - 2 + kw] (tile-local, zero
//     outside the tile).  Read MN-major with M-groups strided by one copy row (512 B), M-group G = kh*3 + kw of the
//     descriptor that starts at halo pixel (r, c) reads unit ((r + kh)*3 + kw)*32 + c = dl[r - 2 + kh][c - 2 + kw]:
//     exactly the dl pixel that halo pixel (r, c) of `a` meets under tap (kh, kw).
// What bounds such a kernel is the ISSUE of the MMAs, not the tensor pipe (tools/mma_cost.cu: 42 cycles per M=128,
// N=32 MMA whatever the majors and strides): the previous version kept kw in the B start address - three M=64 MMAs
// per K step, 204 per tile - and ran exactly at the rate its issuing thread could build descriptors and issue
// (33 cycles per MMA).  Here the issuing loop adds constants to two 32-bit descriptor words per MMA and nothing else.
// `a` tiles are double-buffered whole (big TMA boxes stream at 7 TB/s, tools/tma_stream.cu; 4 KB boxes do not: 2.6
// TB/s); the tile of dl arrives by TMA in a two-deep staging ring (zero fill outside the image) and the loader
// warps copy it shared -> shared into the three copies, which are single-buffered: the staged tile waits in
// registers while the previous tile's MMAs run, only the shared-memory stores wait for them.  The accumulator (128
// lanes x 32 columns) stays in TMEM over all tiles of the persistent CTA; each CTA writes one partial dW that a
// reduction kernel sums (deterministic).
struct OutWgradParams {
  float* partial;              // [grid][9*Cout*Cin]
  int B, H, W, Cout, Cin;
  int tiles_y, tiles_x, num_tiles;
  int* error_flag;
};
constexpr int DLROWS = TR + 4;                      // dl copy rows (-2 .. TR+1)
constexpr uint32_t DL_BYTES = DLROWS * 3 * PW * 16; // three row-interleaved copies
constexpr uint32_t OW_DL_STAGE = TR * TW * 16;      // staged dl tile, dense rows of TW pixels
constexpr uint32_t OW_STAGE = 4 * NPIX * 16;        // Cin = 32: 4 chunk planes of the halo tile
constexpr uint32_t OW_B_OFF = DL_BYTES + 2 * OW_DL_STAGE;
constexpr uint32_t OW_SMEM = OW_B_OFF + kStages * OW_STAGE;
static_assert(PW == 32, "K steps of 16 pixels must not cross a halo row");
static_assert(OW_B_OFF % 1024 == 0 && OW_SMEM <= 227 * 1024 - 1024, "tc_out_wgrad shared memory");

// tcgen05.mma with the descriptors given as (low word, high word): the low words carry the start address and are what
// an issuing loop advances
__device__ __forceinline__ void mma_bf16_ss_words(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
tc_out_wgrad_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_dl, OutWgradParams p) {
  constexpr int KC = 4;                               // Cin = 32
  constexpr uint32_t CH = NPIX * 16;
  extern __shared__ __align__(1024) unsigned char smem[];        // [dl copies][dl stage 0][dl stage 1][a tile 0][a tile 1]
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], dfull_bar[2], dempty_bar[2], afull_bar, aempty_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // border cells of the copies are zero for every tile: cleared once, the loaders only ever write the interior
  for (uint32_t i = threadIdx.x; i < DL_BYTES / 16; i += kThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc<32>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&dfull_bar[s], 1); mbar_init(&dempty_bar[s], 4); }
    mbar_init(&afull_bar, 4); mbar_init(&aempty_bar, 1);
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int my_tiles = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  unsigned char* dstage = smem + DL_BYTES;
  unsigned char* btiles = smem + OW_B_OFF;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int s = it % kStages, d = it & 1;
        const int n = t / (p.tiles_y * p.tiles_x);
        const int rem = t % (p.tiles_y * p.tiles_x);
        const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
        if (!mbar_wait(&dempty_bar[d], (uint32_t)((it >> 1) & 1) ^ 1u)) { *p.error_flag = 1; break; }
        mbar_expect_tx(&dfull_bar[d], OW_DL_STAGE);
        tma_load_3d(dstage + d * OW_DL_STAGE, &tmap_dl, &dfull_bar[d], tx * TW * 8, ty * TR, n);
        if (!mbar_wait(&empty_bar[s], (uint32_t)((it / kStages) & 1) ^ 1u)) { *p.error_flag = 1; break; }
        mbar_expect_tx(&full_bar[s], OW_STAGE);
#pragma unroll
        for (int c = 0; c < KC; ++c)
          tma_load_3d(btiles + s * OW_STAGE + c * CH, &tmap, &full_bar[s], (tx * TW - 1) * 8, ty * TR - 1, n * KC + c);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16_f32(128, 32, 1, 1);   // both operands MN-major
    // descriptor words (tc_common.cuh): low = start >> 4 | (leading byte offset >> 4) << 16, high = (stride byte offset >> 4) | version
    const uint32_t a_hi = (uint32_t)((PW * 16) >> 4) | (1u << 14), b_hi = (uint32_t)(CH >> 4) | (1u << 14);
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | ((128u >> 4) << 16);
    const uint32_t b_lo0 = (smem_u32(btiles) >> 4) | ((128u >> 4) << 16);
    int it = 0;
    bool ok = true;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages;
      if (!mbar_wait(&full_bar[s], (uint32_t)(it / kStages) & 1u) || !mbar_wait(&afull_bar, (uint32_t)(it & 1))) {
        if (leader) *p.error_flag = 1;
        ok = false;
        break;
      }
      fence_after_sync();
      if (leader) {
        uint32_t a_lo = a_lo0, b_lo = b_lo0 + (uint32_t)s * (OW_STAGE >> 4);
        mma_bf16_ss_words(tmem, a_lo, a_hi, b_lo, b_hi, idesc, (uint32_t)(it != 0));
        mma_bf16_ss_words(tmem, a_lo + 16, a_hi, b_lo + 16, b_hi, idesc, 1u);
#pragma unroll 3
        for (int r = 1; r < PR; ++r) {                 // halo row r: two K steps; the copies advance three copy rows per halo row
          a_lo += 3 * PW; b_lo += PW;
          mma_bf16_ss_words(tmem, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
          mma_bf16_ss_words(tmem, a_lo + 16, a_hi, b_lo + 16, b_hi, idesc, 1u);
        }
        mma_commit(&empty_bar[s]);
        mma_commit(&aempty_bar);
      }
      __syncwarp();
    }
    if (ok && leader) mma_commit(&done_bar);
  } else {
    // ============================ dl-copy loaders (4 warps) ===============================
    const int lt = threadIdx.x - 64;              // 0..127
    constexpr int NU = (TR * TW + 127) / 128;     // interior pixels per thread
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int d = it & 1;
      if (!mbar_wait(&dfull_bar[d], (uint32_t)(it >> 1) & 1u)) { if (lane == 0) *p.error_flag = 1; break; }
      // the staged tile goes to registers (and its ring slot back to the producer) before the wait for the previous
      // tile's MMAs
      const uint4* src = reinterpret_cast<const uint4*>(dstage + d * OW_DL_STAGE);
      uint4 vv[NU];
#pragma unroll
      for (int k = 0; k < NU; ++k) {
        const int u = lt + k * 128;
        vv[k] = u < TR * TW ? src[u] : make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&dempty_bar[d]);
      if (!mbar_wait(&aempty_bar, (uint32_t)((it & 1) ^ 1))) { if (lane == 0) *p.error_flag = 1; break; }
      uint4* dst = reinterpret_cast<uint4*>(smem);
#pragma unroll
      for (int k = 0; k < NU; ++k) {
        const int u = lt + k * 128;
        const int rho = u / TW, c = u % TW;
        if (u < TR * TW) {
          uint4* row = dst + (rho + 2) * (3 * PW) + c + 2;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) row[kw * PW - kw] = vv[k];       // copy kw, column c + 2 - kw
        }
      }
      fence_async_smem();       // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull_bar);
    }
    // ================================ final epilogue =====================================
    const int lg = warp & 3;
    if (my_tiles > 0) {
      if (mbar_wait(&done_bar, 0)) {
        fence_after_sync();
        const int E = 9 * p.Cout * p.Cin;
        float* out = p.partial + (int64_t)blockIdx.x * E;
        float v[32];
        const uint32_t ta = tmem + ((uint32_t)(lg * 32) << 16);
        tmem_ld16(ta, v);
        tmem_ld16(ta + 16, v + 16);
        const int m = lg * 32 + lane;                // M row held by this TMEM lane: tap group G = kh*3 + kw, channel co
        const int g = m >> 3, co = m & 7;
        if (g < 9 && co < p.Cout) {
          for (int ci = 0; ci < p.Cin; ++ci) out[(g * p.Cout + co) * p.Cin + ci] = v[ci];
        }
      } else if (lane == 0) {
        *p.error_flag = 1;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem);
}

// out[e] = sum_i partial[i*E + e]: 8 threads per entry, each a contiguous slice of the partials,
// folded in a fixed order (deterministic); launch with sum_partials().
constexpr int SP_SLICES = 8;
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* partial, int nparts, int E, float* out) {
  __shared__ float red[256];
  const int oi = threadIdx.x / SP_SLICES, sl = threadIdx.x % SP_SLICES;
  const int e = blockIdx.x * (256 / SP_SLICES) + oi;
  float t = 0.f;
  if (e < E) {
    const int per = (nparts + SP_SLICES - 1) / SP_SLICES;
    const int c0 = sl * per, c1 = min(nparts, c0 + per);
#pragma unroll 4
    for (int c = c0; c < c1; ++c) t += __ldg(partial + (int64_t)c * E + e);
  }
  red[threadIdx.x] = t;
  __syncthreads();
  if (sl == 0 && e < E) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < SP_SLICES; ++i) v += red[threadIdx.x + i];
    out[e] = v;
  }
}
// few entries, many partials (channel sums): one block per entry, 256 threads stride over the partials,
// then a fixed shared-memory tree (deterministic)
__global__ void __launch_bounds__(256) sum_partials_wide_kernel(const float* partial, int nparts, int E, float* out) {
  __shared__ float red[256];
  const int e = blockIdx.x;
  float t = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) t += __ldg(partial + (int64_t)i * E + e);
  red[threadIdx.x] = t;
  __syncthreads();
#pragma unroll
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[e] = red[0];
}
static void sum_partials(const float* partial, int nparts, int E, float* out, cudaStream_t st) {
  ++g_launches;
  if (E <= 64 && nparts >= 512) sum_partials_wide_kernel<<<E, 256, 0, st>>>(partial, nparts, E, out);
  else sum_partials_kernel<<<cdiv(E, 256 / SP_SLICES), 256, 0, st>>>(partial, nparts, E, out);
}

// ============================================================================================
// Conv2DTranspose k3 s2 'same' + bias + ReLU on tensor cores (src/abstract_cvae.py:81-84, A3),
// Cin <= 8 -> Cout = 32, by sub-pixel phase decomposition: output parity (pa,pb) of
//   y[2i+pa, 2j+pb, co] = b[co] + sum_{kh = pa (mod 2), kw = pb (mod 2)} x[i-(kh-pa)/2, j-(kw-pb)/2, ci] W[kh,kw,co,ci]
// is a dense 2x2 / 1x2 / 2x1 / 1x1 convolution on the LOW-resolution grid.  Input: bf16 NHWC
// padded to 8 channels (one 16-byte unit per pixel, one TMA plane).  Per M-tile of 128 low-res
// pixels five K=16 MMAs (tap pairs through the leading byte offset) fill four N=32
// accumulators (128 TMEM columns, double buffered); the epilogue writes the 2x2 output pixels
// as bf16 NHWC - the activation the output-layer kernels consume.
struct ConvTParams {
  const __nv_bfloat16* wimg;   // [5 MMAs][2 chunks][32][8]
  const float* bias;           // [32]
  __nv_bfloat16* out;          // chunk-planar bf16 [B][4][2h][2w][8]
  int B, h, w;                 // low-res size
  int tiles_y, tiles_x, num_tiles;
  int* error_flag;
};

__global__ void __launch_bounds__(kThreadsE, 1)
tc_convT_fwd_kernel(const __grid_constant__ CUtensorMap tmap, ConvTParams p) {
  constexpr uint32_t TILE_BYTES = NPIX * 16;          // one 8-channel plane, 34 x 32 pixels
  constexpr uint32_t STAGE = TILE_BYTES + 128;
  constexpr uint32_t WB = 5 * 2 * 32 * 16;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* s_tile = smem;
  unsigned char* s_w = smem + kStages * STAGE;
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < (int)(WB / 16); i += kThreadsE)
    reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(p.wimg)[i];
  if (threadIdx.x < 32) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (threadIdx.x < kStages * 8) {
    const int s = threadIdx.x / 8, j = threadIdx.x % 8;
    reinterpret_cast<uint4*>(s_tile + s * STAGE + TILE_BYTES)[j] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc<256>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
    fence_mbar_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        if (!mbar_wait(&empty_bar[s], ph ^ 1)) { *p.error_flag = 1; break; }
        const int n = t / (p.tiles_y * p.tiles_x);
        const int rem = t % (p.tiles_y * p.tiles_x);
        const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
        mbar_expect_tx(&full_bar[s], TILE_BYTES);
        tma_load_3d(s_tile + s * STAGE, &tmap, &full_bar[s], (tx * TW - 1) * 8, ty * TR - 1, n);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16_f32(128, 32);
    const uint32_t w_base = smem_u32(s_w);
    // (start shift, leading offset) in pixels and destination phase of the five MMAs
    //   MMA0: taps (2,2)@0 , (2,0)@1        -> phase (0,0)     MMA1: taps (0,2)@PW , (0,0)@PW+1 -> phase (0,0)
    //   MMA2: taps (2,1)@1 , (0,1)@PW+1     -> phase (0,1)     MMA3: taps (1,2)@PW , (1,0)@PW+1 -> phase (1,0)
    //   MMA4: tap  (1,1)@PW+1 , zero weights -> phase (1,1)
    int it = 0, mcount = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      if (!mbar_wait(&full_bar[s], ph)) { if (leader) *p.error_flag = 1; break; }
      fence_after_sync();
      const uint32_t tile_base = smem_u32(s_tile + s * STAGE);
      bool ok = true;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt, ++mcount) {
        const int a = mcount & 1;
        const uint32_t aph = (mcount >> 1) & 1;
        if (!mbar_wait(&tempty_bar[a], aph ^ 1)) { if (leader) *p.error_flag = 1; ok = false; break; }
        fence_after_sync();
        const uint32_t d0 = tmem + (uint32_t)(a * 128);
        const uint32_t q0 = tile_base + (uint32_t)(mt * 128) * 16;
        const uint64_t a0 = make_desc_kmajor_noswz(q0, 16, 128);
        const uint64_t a1 = make_desc_kmajor_noswz(q0 + PW * 16, 16, 128);
        const uint64_t a2 = make_desc_kmajor_noswz(q0 + 16, PW * 16, 128);
        const uint64_t a4 = make_desc_kmajor_noswz(q0 + (PW + 1) * 16, 16, 128);
        const uint64_t b0 = make_desc_kmajor_noswz(w_base, 32 * 16, 128);
        if (leader) {
          mma_bf16_ss(d0 + 0, a0, b0, idesc, 0);
          mma_bf16_ss(d0 + 0, a1, desc_advance(b0, 64), idesc, 1);
          mma_bf16_ss(d0 + 32, a2, desc_advance(b0, 128), idesc, 0);
          mma_bf16_ss(d0 + 64, a1, desc_advance(b0, 192), idesc, 0);
          mma_bf16_ss(d0 + 96, a4, desc_advance(b0, 256), idesc, 0);
          mma_commit(&tfull_bar[a]);
        }
        __syncwarp();
      }
      if (!ok) break;
      if (leader) mma_commit(&empty_bar[s]);
      __syncwarp();
    }
  } else {
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;             // half 0: output rows 2i (phases 0,1); half 1: rows 2i+1
    const int H2 = 2 * p.h, W2 = 2 * p.w;
    int it = 0, mcount = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int n = t / (p.tiles_y * p.tiles_x);
      const int rem = t % (p.tiles_y * p.tiles_x);
      const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
      bool ok = true;
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt, ++mcount) {
        const int a = mcount & 1;
        const uint32_t aph = (mcount >> 1) & 1;
        if (!mbar_wait(&tfull_bar[a], aph)) { if (lane == 0) *p.error_flag = 1; ok = false; break; }
        fence_after_sync();
        const int q = mt * 128 + lg * 32 + lane;
        const int r = q / PW, c = q % PW;
        const int i = ty * TR + r, j = tx * TW + c;
        const bool valid = c < TW && i < p.h && j < p.w;
#pragma unroll
        for (int phs = 2 * half; phs < 2 * half + 2; ++phs) {
          float v[32];
          const uint32_t ta = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(a * 128 + phs * 32);
          tmem_ld16(ta, v);
          tmem_ld16(ta + 16, v + 16);
          if (valid) {
            const int oy = 2 * i + (phs >> 1), ox = 2 * j + (phs & 1);
            const int64_t plane = (int64_t)H2 * W2;    // chunk-planar output [B][4][H2][W2][8]
            uint4* o = reinterpret_cast<uint4*>(p.out) + (int64_t)n * 4 * plane + (int64_t)oy * W2 + ox;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float y0 = fmaxf(v[g * 8 + 2 * e] + s_bias[g * 8 + 2 * e], 0.f);
                const float y1 = fmaxf(v[g * 8 + 2 * e + 1] + s_bias[g * 8 + 2 * e + 1], 0.f);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(y0, y1);
                w4[e] = *reinterpret_cast<uint32_t*>(&b2);
              }
              o[g * plane] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
          }
        }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[a]);
      }
      if (!ok) break;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// W [3,3,Cout=32,Cin<=8] fp32 -> five paired-tap B images [mma][chunk(2)][n = co][8 = ci]
__device__ __forceinline__ void tc_prep_convT_weights_body(const float* w, int Cout, int Cin, __nv_bfloat16* img) {
  // (kh,kw) of chunk 0 / chunk 1 of each MMA; -1 = zero weights
  const int taps[5][2] = {{8, 6}, {2, 0}, {7, 1}, {5, 3}, {4, -1}};
  const int total = 5 * 2 * 32 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % 8;
    const int n = (i / 8) % 32;
    const int kc = (i / 256) % 2;
    const int m = i / 512;
    const int tap = taps[m][kc];
    const float v = (tap >= 0 && j < Cin && n < Cout) ? w[((int64_t)tap * Cout + n) * Cin + j] : 0.f;
    img[i] = __float2bfloat16(v);
  }
}

// ============================================================================================
// Conv2DTranspose s2 forward, MANY -> few channels (Cin = 32, Cout <= 8), on tensor cores:
//   out[n, 2i+a, 2j+b, co] = relu(bias[co] + sum_{di,dj in {0,1}} sum_ci in[n, i-di, j-dj, ci] * W[a+2di, b+2dj, co, ci])
// (taps with a+2di > 2 or b+2dj > 2 do not exist).  One GEMM per low-res tile: M = 128 low-res pixels,
// N = 32 = (output parity a,b) x 8 padded channels, K = 32 channels x 4 input shifts = 8 MMAs per M-tile; the shifts
// are descriptor start offsets into a halo tile (one row above, one column left) of the chunk-planar bf16 input.
// The epilogue writes the bf16 8-channel units the next (few -> 32) tensor-core layer reads and, for training,
// the fp32 activation the backward kernels use.
constexpr int TRF = 8;                                // low-res rows per tile
constexpr int FROWS = TRF + 1;
constexpr uint32_t CHF = FROWS * PW * 16;             // bytes per chunk plane of the input tile
constexpr int MTF = (TRF * PW) / 128;                 // 2 M-tiles
struct ConvTFewParams {
  const __nv_bfloat16* wimg;   // [4 shifts][2 ks][2 chunks][32][8]
  const float* bias;           // [Cout]
  uint4* out8;                 // bf16 [B,2h,2w,8] (one 16-byte unit per pixel) or nullptr
  float* out_f32;              // fp32 [B,2h,2w,Cout] or nullptr
  int B, h, w, Cout;
  int tiles_y, tiles_x, num_tiles;
  int* error_flag;
};

// SPLIT: operands arrive as bf16 hi + lo pairs (x = hi + lo to 2^-17) and every product is xh*wh + xl*wh + xh*wl with
// fp32 accumulation - fp32-grade results (the training forward, whose ReLU masks and gradients are held to the fp32 bars)
// for three times the (still negligible) MMA count.  Input planes per frame: [hi 4][lo 4]; weight image: [hi][lo].
template <bool SPLIT>
__global__ void __launch_bounds__(kThreadsE, 1)
tc_convT_few_fwd_kernel(const __grid_constant__ CUtensorMap tmap, ConvTFewParams p) {
  constexpr int NPL = SPLIT ? 8 : 4;                  // chunk planes per tile
  constexpr uint32_t TILE_BYTES = NPL * CHF;
  constexpr uint32_t STAGE = TILE_BYTES + 128;
  constexpr uint32_t WB1 = 4 * 2 * 2 * 32 * 16;       // one weight image
  constexpr uint32_t WB = (SPLIT ? 2 : 1) * WB1;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* s_tile = smem;
  unsigned char* s_w = smem + kStages * STAGE;
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < (int)(WB / 16); i += kThreadsE)
    reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(p.wimg)[i];
  if (threadIdx.x < 8) s_bias[threadIdx.x] = threadIdx.x < p.Cout ? p.bias[threadIdx.x] : 0.f;
  if (threadIdx.x < kStages * 8) {
    const int s = threadIdx.x / 8, j = threadIdx.x % 8;
    reinterpret_cast<uint4*>(s_tile + s * STAGE + TILE_BYTES)[j] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc<64>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
    fence_mbar_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        if (!mbar_wait(&empty_bar[s], ph ^ 1)) { *p.error_flag = 1; break; }
        const int n = t / (p.tiles_y * p.tiles_x);
        const int rem = t % (p.tiles_y * p.tiles_x);
        const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
        mbar_expect_tx(&full_bar[s], TILE_BYTES);
#pragma unroll
        for (int c = 0; c < NPL; ++c)
          tma_load_3d(s_tile + s * STAGE + c * CHF, &tmap, &full_bar[s], (tx * TW - 1) * 8, ty * TRF - 1, n * NPL + c);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16_f32(128, 32);
    const uint64_t b0 = make_desc_kmajor_noswz(smem_u32(s_w), 32 * 16, 128);
    int it = 0, mcount = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      if (!mbar_wait(&full_bar[s], ph)) { if (leader) *p.error_flag = 1; break; }
      fence_after_sync();
      const uint64_t a0 = make_desc_kmajor_noswz(smem_u32(s_tile + s * STAGE), CHF, 128);
      bool ok = true;
#pragma unroll 1
      for (int mt = 0; mt < MTF; ++mt, ++mcount) {
        const int a = mcount & 1;
        const uint32_t aph = (mcount >> 1) & 1;
        if (!mbar_wait(&tempty_bar[a], aph ^ 1)) { if (leader) *p.error_flag = 1; ok = false; break; }
        fence_after_sync();
        const uint32_t d0 = tmem + (uint32_t)(a * 32);
#pragma unroll
        for (int sh = 0; sh < 4; ++sh) {
          // input pixel (i - di, j - dj) of output low-res pixel q sits at tile position q + (1-di)*PW + (1-dj)
          const uint32_t shift = (uint32_t)((1 - (sh >> 1)) * PW + (1 - (sh & 1)));
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t da = desc_advance(a0, (uint32_t)(2 * ks) * (CHF / 16) + (uint32_t)(mt * 128) + shift);
            const uint64_t db = desc_advance(b0, (uint32_t)((sh * 2 + ks) * 2 * 32));
            if (leader) mma_bf16_ss(d0, da, db, idesc, (sh | ks) != 0);
            if constexpr (SPLIT) {
              const uint64_t da_lo = desc_advance(da, 4 * (CHF / 16));      // lo planes sit four planes behind the hi planes
              const uint64_t db_lo = desc_advance(db, WB1 / 16);            // lo weight image behind the hi image
              if (leader) {
                mma_bf16_ss(d0, da_lo, db, idesc, 1);
                mma_bf16_ss(d0, da, db_lo, idesc, 1);
              }
            }
          }
        }
        if (leader) mma_commit(&tfull_bar[a]);
        __syncwarp();
      }
      if (!ok) break;
      if (leader) mma_commit(&empty_bar[s]);
      __syncwarp();
    }
  } else {
    const int lg = warp & 3;
    const int pa = (warp - 2) >> 2;               // output row parity this warp writes
    const int H2 = 2 * p.h, W2 = 2 * p.w;
    int it = 0, mcount = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int n = t / (p.tiles_y * p.tiles_x);
      const int rem = t % (p.tiles_y * p.tiles_x);
      const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
      bool ok = true;
#pragma unroll 1
      for (int mt = 0; mt < MTF; ++mt, ++mcount) {
        const int a = mcount & 1;
        const uint32_t aph = (mcount >> 1) & 1;
        if (!mbar_wait(&tfull_bar[a], aph)) { if (lane == 0) *p.error_flag = 1; ok = false; break; }
        fence_after_sync();
        const int q = mt * 128 + lg * 32 + lane;
        const int r = q / PW, c = q % PW;
        const int i = ty * TRF + r, j = tx * TW + c;
        const bool valid = c < TW && i < p.h && j < p.w;
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(a * 32 + pa * 16), v);
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[a]);          // accumulators are in registers: the slot is free
        if (valid) {
          const int oy = 2 * i + pa;
          const int64_t pix = ((int64_t)n * H2 + oy) * W2 + 2 * j;       // pixels (oy, 2j) and (oy, 2j+1) are adjacent
          float y[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) y[k] = (k & 7) < p.Cout ? fmaxf(v[k] + s_bias[k & 7], 0.f) : 0.f;
          if (p.out8) {
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              uint32_t w4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(y[b * 8 + 2 * e], y[b * 8 + 2 * e + 1]);
                w4[e] = *reinterpret_cast<uint32_t*>(&b2);
              }
              p.out8[pix + b] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
          }
          if (p.out_f32) {     // pixels (oy, 2j) and (oy, 2j+1): 2 * Cout consecutive floats
            float* o = p.out_f32 + pix * p.Cout;
#pragma unroll
            for (int b = 0; b < 2; ++b)
#pragma unroll
              for (int co = 0; co < 8; ++co)
                if (co < p.Cout) o[b * p.Cout + co] = y[b * 8 + co];
          }
        }
      }
      if (!ok) break;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

// W [3,3,Cout<=8,Cin=32] fp32 -> B images [shift (di,dj)][ks][chunk(2)][n = (a*2+b)*8 + co][8 = ci]
__device__ __forceinline__ void tc_prep_convT_few_weights_body(const float* w, int Cout, int Cin, __nv_bfloat16* img) {
  const int total = 4 * 2 * 2 * 32 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % 8;
    const int n = (i / 8) % 32;
    const int kc = (i / 256) % 2;
    const int ks = (i / 512) % 2;
    const int sh = i / 1024;
    const int di = sh >> 1, dj = sh & 1, a = n >> 4, b = (n >> 3) & 1, co = n & 7;
    const int kh = a + 2 * di, kw = b + 2 * dj, ci = ks * 16 + kc * 8 + j;
    const float v = (kh <= 2 && kw <= 2 && co < Cout && ci < Cin) ? w[((int64_t)(kh * 3 + kw) * Cout + co) * Cin + ci] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16(v);
    img[i] = hi;
    img[total + i] = __float2bfloat16(v - __bfloat162float(hi));      // lo image (SPLIT kernels)
  }
}
__global__ void tc_prep_convT_few_weights_kernel(const float* w, int Cout, int Cin, __nv_bfloat16* img) { tc_prep_convT_few_weights_body(w, Cout, Cin, img); }

// fp32 NHWC with C <= 8 channels -> bf16 NHWC padded to 8 channels (16 bytes per pixel)
__global__ void pack_c8_bf16_kernel(const float* in, int64_t npix, int C, uint4* out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = c < C ? in[i * C + c] : 0.f;
    uint32_t w4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      w4[e] = *reinterpret_cast<uint32_t*>(&b2);
    }
    out[i] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

// ============================================================================================
// Backward of the last Conv2DTranspose s2 (Cin <= 8 -> 32) on tensor cores.  The incoming
// gradient G = d loss / d a_last arrives from tc_out_dgrad as bf16 *space-to-depth*
// [B,h,w,4 parities,32]: tap (kh,kw) of the stride-2 gather  G[2i+kh, 2j+kw]  is then parity
// plane (kh&1, kw&1) at the low-resolution shift (kh>>1, kw>>1) - again plain descriptor start
// offsets into one TMA-loaded tile of 16 chunk planes (rows 0..TRD, one halo row/column at the
// bottom/right; TMA zero fill = the cropped transposed-conv border).
constexpr int TRD = 8;                           // low-res rows per tile
constexpr int GROWS = TRD + 1;
constexpr uint32_t CHD = GROWS * PW * 16;        // bytes per chunk plane of the G tile
constexpr uint32_t GT_BYTES = 16 * CHD;          // 16 chunk planes
constexpr int MTD = (TRD * PW) / 128;            // 2 M-tiles

struct ConvTBwdParams {
  const __nv_bfloat16* wimg;     // dgrad: [9 taps][2 ks][2 chunks][16][8]
  const float* mask;             // dgrad: previous activation fp32 [B,h,w,Cin]
  float* g_prev;                 // dgrad: [B,h,w,Cin] fp32, or nullptr
  uint4* g_planes;               // dgrad: the same gradient as bf16 space-to-depth planes [B][4*KC][h/2][w/2][8] (chunk 0 of every
  int g_KC;                      //        parity is written; the general engine reads them), or nullptr
  const __nv_bfloat16* a_prev8;  // wgrad: previous activation bf16 [B,h,w,8]
  float* partial;                // wgrad: [grid][9*32*Cin]
  int B, h, w, Cin;
  int tiles_y, tiles_x, num_tiles;
  int* error_flag;
};

__global__ void __launch_bounds__(kThreads, 1)
tc_convT_dgrad_kernel(const __grid_constant__ CUtensorMap tmap, ConvTBwdParams p) {
  constexpr uint32_t STAGE = GT_BYTES + 128;
  constexpr uint32_t WB = 9 * 2 * 2 * 16 * 16;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* s_w = smem + kStages * STAGE;
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < (int)(WB / 16); i += kThreads)
    reinterpret_cast<uint4*>(s_w)[i] = reinterpret_cast<const uint4*>(p.wimg)[i];
  if (threadIdx.x < kStages * 8) {
    const int s = threadIdx.x / 8, j = threadIdx.x % 8;
    reinterpret_cast<uint4*>(smem + s * STAGE + GT_BYTES)[j] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc<64>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    fence_mbar_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        if (!mbar_wait(&empty_bar[s], ph ^ 1)) { *p.error_flag = 1; break; }
        const int n = t / (p.tiles_y * p.tiles_x);
        const int rem = t % (p.tiles_y * p.tiles_x);
        const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
        mbar_expect_tx(&full_bar[s], GT_BYTES);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          tma_load_3d(smem + s * STAGE + c * CHD, &tmap, &full_bar[s], tx * TW * 8, ty * TRD, n * 16 + c);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16_f32(128, 16);
    const uint64_t db0 = make_desc_kmajor_noswz(smem_u32(s_w), 16 * 16, 128);
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages, a = it & 1;
      const uint32_t ph = (it / kStages) & 1, aph = (it >> 1) & 1;
      if (!mbar_wait(&tempty_bar[a], aph ^ 1)) { if (leader) *p.error_flag = 1; break; }
      if (!mbar_wait(&full_bar[s], ph)) { if (leader) *p.error_flag = 1; break; }
      fence_after_sync();
      const uint64_t da0 = make_desc_kmajor_noswz(smem_u32(smem + s * STAGE), CHD, 128);
#pragma unroll
      for (int mt = 0; mt < MTD; ++mt) {
        const uint32_t d_tmem = tmem + (uint32_t)(a * MTD * 16 + mt * 16);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int kh = tap / 3, kw = tap % 3;
          const uint32_t par = (uint32_t)((kh & 1) * 2 + (kw & 1));
          const uint32_t shift = (uint32_t)((kh >> 1) * PW + (kw >> 1));
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t da = desc_advance(da0, (par * 4 + 2 * ks) * (CHD / 16) + (uint32_t)(mt * 128) + shift);
            const uint64_t db = desc_advance(db0, (uint32_t)((tap * 2 + ks) * 2 * 16));
            if (leader) mma_bf16_ss(d_tmem, da, db, idesc, (tap | ks) != 0);
          }
        }
      }
      if (leader) {
        mma_commit(&empty_bar[s]);
        mma_commit(&tfull_bar[a]);
      }
      __syncwarp();
    }
  } else {
    const int lg = warp & 3;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int a = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int n = t / (p.tiles_y * p.tiles_x);
      const int rem = t % (p.tiles_y * p.tiles_x);
      const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
      // the ReLU mask of both M-tiles is requested before the accumulator wait (and before any store:
      // g_prev and mask may alias as far as the compiler knows, which would serialise load/store pairs)
      float mk[MTD][8];
      int64_t pixs[MTD];
      bool lives[MTD];
#pragma unroll
      for (int mt = 0; mt < MTD; ++mt) {
        const int q = mt * 128 + lg * 32 + lane;
        const int r = q / PW, c = q % PW;
        const int i = ty * TRD + r, j = tx * TW + c;
        lives[mt] = c < TW && i < p.h && j < p.w;
        pixs[mt] = ((int64_t)n * p.h + i) * p.w + j;
#pragma unroll
        for (int ci = 0; ci < 8; ++ci) mk[mt][ci] = (lives[mt] && ci < p.Cin) ? __ldg(p.mask + pixs[mt] * p.Cin + ci) : 0.f;
      }
      if (!mbar_wait(&tfull_bar[a], aph)) { if (lane == 0) *p.error_flag = 1; break; }
      fence_after_sync();
#pragma unroll
      for (int mt = 0; mt < MTD; ++mt) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(a * MTD * 16 + mt * 16), v);
        if (lives[mt]) {
#pragma unroll
          for (int ci = 0; ci < 8; ++ci) v[ci] = (ci < p.Cin && mk[mt][ci] > 0.f) ? v[ci] : 0.f;
          if (p.g_prev) {
#pragma unroll
            for (int ci = 0; ci < 8; ++ci)
              if (ci < p.Cin) p.g_prev[pixs[mt] * p.Cin + ci] = v[ci];
          }
          if (p.g_planes) {      // one 16-byte unit per pixel straight into the planes the next layer's backward reads
            const int q = mt * 128 + lg * 32 + lane;
            const int i = ty * TRD + q / PW, j = tx * TW + q % PW;
            const int par = ((i & 1) << 1) | (j & 1);
            const int64_t u = (((int64_t)n * (4 * p.g_KC) + par * p.g_KC) * (p.h >> 1) + (i >> 1)) * (p.w >> 1) + (j >> 1);
            uint32_t w4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
              w4[e] = *reinterpret_cast<uint32_t*>(&b2);
            }
            p.g_planes[u] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[a]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

// W [3,3,Cout=32,Cin<=8] fp32 -> dgrad B image [tap][ks][chunk(2)][n = ci (16)][8 = co]
__device__ __forceinline__ void tc_prep_convT_dgrad_weights_body(const float* w, int Cout, int Cin, __nv_bfloat16* img) {
  const int total = 9 * 2 * 2 * 16 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % 8;
    const int n = (i / 8) % 16;
    const int kc = (i / 128) % 2;
    const int ks = (i / 256) % 2;
    const int tap = i / 512;
    const int co = ks * 16 + kc * 8 + j;
    const float v = (n < Cin && co < Cout) ? w[((int64_t)tap * Cout + co) * Cin + n] : 0.f;
    img[i] = __float2bfloat16(v);
  }
}

__global__ void tc_prep_out_weights_kernel(const float* w, int Cout, int Cin, __nv_bfloat16* img) { tc_prep_out_weights_body(w, Cout, Cin, img); }
__global__ void tc_prep_tail_c2i_weights_kernel(const float* w, int Cout, int Cin, __nv_bfloat16* img) { tc_prep_tail_c2i_weights_body(w, Cout, Cin, img); }
__global__ void tc_prep_tail_kw_weights_kernel(const float* w, int Cout, int Cin, __nv_bfloat16* img) { tc_prep_tail_kw_weights_body(w, Cout, Cin, img); }
__global__ void tc_prep_dgrad_weights_kernel(const float* w, int Cout, int Cin, __nv_bfloat16* img) { tc_prep_dgrad_weights_body(w, Cout, Cin, img); }
__global__ void tc_prep_convT_weights_kernel(const float* w, int Cout, int Cin, __nv_bfloat16* img) { tc_prep_convT_weights_body(w, Cout, Cin, img); }
__global__ void tc_prep_convT_dgrad_weights_kernel(const float* w, int Cout, int Cin, __nv_bfloat16* img) { tc_prep_convT_dgrad_weights_body(w, Cout, Cin, img); }

// all four weight images of a training step in ONE launch (grid.y selects the image): the last Conv2DTranspose
// forward, the fused tail's output convolution, the output-layer dgrad and the Conv2DTranspose dgrad
struct PrepAllArgs {
  const float* w_convT;   // [3,3,Clast,Cprev]
  const float* w_out;     // [3,3,Cout,Clast]
  int Cprev, Clast, Cout, tail_c2i;
  __nv_bfloat16 *img_convT, *img_tail, *img_dgrad, *img_convT_dgrad;
};
__global__ void tc_prep_all_kernel(PrepAllArgs a) {
  switch (blockIdx.y) {
    case 0: tc_prep_convT_weights_body(a.w_convT, a.Clast, a.Cprev, a.img_convT); break;
    case 1:
      if (a.tail_c2i) tc_prep_tail_c2i_weights_body(a.w_out, a.Cout, a.Clast, a.img_tail);
      else tc_prep_tail_kw_weights_body(a.w_out, a.Cout, a.Clast, a.img_tail);
      break;
    case 2: tc_prep_dgrad_weights_body(a.w_out, a.Cout, a.Clast, a.img_dgrad); break;
    default: tc_prep_convT_dgrad_weights_body(a.w_convT, a.Clast, a.Cprev, a.img_convT_dgrad); break;
  }
}

// dW[kh,kw,co,ci] = sum a_prev[i,j,ci] * G[2i+kh, 2j+kw, co] : MN-major, K = low-res pixels, ONE MMA (M=64, N=128)
// per 16 pixels of the G tile:
//   B = the whole space-to-depth G tile, N = 128 = 4 parities x 32 channels = its 16 chunk planes (constant stride);
//   A = TWO zero-padded copies of the a_prev plane (rows -1..TRD), copy dh shifted right by dh columns, stored
//     row-interleaved: unit ((R*2 + dh)*PW + c) = a[R - 1][c - dh].  M-groups are strided by one copy row, so M-group
//     2*g + dh of the descriptor that starts at G pixel (gr, gc) reads a[gr + g - 1][gc - dh]: group g = 0 <-> vertical
//     low-res shift 1, g = 1 <-> shift 0; dh = the horizontal shift.
// (Before: the horizontal shift was the B start address and every parity its own MMA - six M=64, N=32 MMAs per K
// step at 38 cycles each, tools/mma_cost.cu - and the kernel ran at their rate.)  The accumulator (M rows (g, dh, ci),
// columns (parity, co)) persists in TMEM over the CTA's tiles; combinations that are no tap of the 3x3 kernel are
// dropped by the epilogue.
constexpr int APROWS = TRD + 2;
constexpr uint32_t AP_BYTES = APROWS * 2 * PW * 16;
static_assert(PW == 32, "two 16-pixel K steps per G row");

__global__ void __launch_bounds__(kThreads, 1)
tc_convT_wgrad_kernel(const __grid_constant__ CUtensorMap tmap, ConvTBwdParams p) {
  constexpr uint32_t STAGE = AP_BYTES + GT_BYTES + 128;    // [a_prev copies][G tile][pad]
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // column 0 of the shifted copy is never written by the loaders: zero for every tile
  for (int i = threadIdx.x; i < kStages * APROWS; i += kThreads) {
    const int s = i / APROWS, R = i % APROWS;
    reinterpret_cast<uint4*>(smem + s * STAGE)[(R * 2 + 1) * PW] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc<128>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1 + 4); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int my_tiles = p.num_tiles > (int)blockIdx.x ? (p.num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        if (!mbar_wait(&empty_bar[s], ph ^ 1)) { *p.error_flag = 1; break; }
        const int n = t / (p.tiles_y * p.tiles_x);
        const int rem = t % (p.tiles_y * p.tiles_x);
        const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
        mbar_expect_tx(&full_bar[s], GT_BYTES);
#pragma unroll
        for (int c = 0; c < 16; ++c)
          tma_load_3d(smem + s * STAGE + AP_BYTES + c * CHD, &tmap, &full_bar[s], tx * TW * 8, ty * TRD, n * 16 + c);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16_f32(64, 128, 1, 1);
    // descriptor words (tc_common.cuh): low = start >> 4 | (leading byte offset >> 4) << 16, high = (stride byte offset >> 4) | version
    const uint32_t a_hi = (uint32_t)((PW * 16) >> 4) | (1u << 14), b_hi = (uint32_t)(CHD >> 4) | (1u << 14);
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | ((128u >> 4) << 16);
    int it = 0;
    bool ok = true;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      if (!mbar_wait(&full_bar[s], ph)) { if (leader) *p.error_flag = 1; ok = false; break; }
      fence_after_sync();
      if (leader) {
        uint32_t a_lo = a_lo0 + (uint32_t)s * (STAGE >> 4), b_lo = a_lo + (AP_BYTES >> 4);
        mma_bf16_ss_words(tmem, a_lo, a_hi, b_lo, b_hi, idesc, (uint32_t)(it != 0));
        mma_bf16_ss_words(tmem, a_lo + 16, a_hi, b_lo + 16, b_hi, idesc, 1u);
#pragma unroll 4
        for (int gr = 1; gr < GROWS; ++gr) {           // G row gr: two K steps; the copies advance two copy rows per G row
          a_lo += 2 * PW; b_lo += PW;
          mma_bf16_ss_words(tmem, a_lo, a_hi, b_lo, b_hi, idesc, 1u);
          mma_bf16_ss_words(tmem, a_lo + 16, a_hi, b_lo + 16, b_hi, idesc, 1u);
        }
        mma_commit(&empty_bar[s]);
      }
      __syncwarp();
    }
    if (ok && leader) mma_commit(&done_bar);
  } else {
    const int lt = threadIdx.x - 64;
    int it = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      if (!mbar_wait(&empty_bar[s], ph ^ 1)) { if (lane == 0) *p.error_flag = 1; break; }
      const int n = t / (p.tiles_y * p.tiles_x);
      const int rem = t % (p.tiles_y * p.tiles_x);
      const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
      uint4* dst = reinterpret_cast<uint4*>(smem + s * STAGE);
      {  // loads first, stores after (see tc_out_wgrad_kernel)
        constexpr int NU = (APROWS * PW + 127) / 128;
        uint4 vv[NU];
#pragma unroll
        for (int k = 0; k < NU; ++k) {
          const int u = lt + k * 128;
          const int rho = u / PW - 1, c = u % PW;
          const int i = ty * TRD + rho, j = tx * TW + c;
          vv[k] = make_uint4(0, 0, 0, 0);
          if (u < APROWS * PW && rho >= 0 && rho < TRD && c < TW && i < p.h && j < p.w)
            vv[k] = __ldg(reinterpret_cast<const uint4*>(p.a_prev8) + ((int64_t)n * p.h + i) * p.w + j);
        }
#pragma unroll
        for (int k = 0; k < NU; ++k) {
          const int u = lt + k * 128;
          if (u < APROWS * PW) {
            const int R = u / PW, c = u % PW;
            dst[(R * 2) * PW + c] = vv[k];                              // copy 0
            if (c + 1 < PW) dst[(R * 2 + 1) * PW + c + 1] = vv[k];      // copy 1: one column to the right
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }
    const int lg = warp & 3;
    if (lg < 2 && my_tiles > 0) {     // M = 64: rows 0..15 sit in TMEM lanes 0..15, rows 16..31 in lanes 32..47
      if (mbar_wait(&done_bar, 0)) {
        fence_after_sync();
        const int E = 9 * 32 * p.Cin;
        float* out = p.partial + (int64_t)blockIdx.x * E;
        const int row = lg * 16 + lane;            // (g, dh, ci)
        const int g = row >> 4, dh = (row >> 3) & 1, ci = row & 7;
#pragma unroll 1
        for (int par = 0; par < 4; ++par) {
          float v[32];
          const uint32_t ta = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(par * 32);
          tmem_ld16(ta, v);
          tmem_ld16(ta + 16, v + 16);
          const int pa = par >> 1, pb = par & 1;
          // vertical: parity row 0 -> kh = 2 (group 0) or 0 (group 1); parity row 1 -> kh = 1 (group 1 only)
          const int kh = pa == 0 ? (g == 0 ? 2 : 0) : 1;
          const int kw = pb == 0 ? 2 * dh : 1;
          const bool use = lane < 16 && ci < p.Cin && (pa == 0 || g == 1) && (pb == 0 || dh == 0);
          if (use) {
            for (int co = 0; co < 32; ++co) out[((kh * 3 + kw) * 32 + co) * p.Cin + ci] = v[co];
          }
        }
      } else if (lane == 0) {
        *p.error_flag = 1;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

// ============================================================================================
// Fused decoder tail (forward / scoring):  a_prev (<= 8 ch, low-res)  --Conv2DTranspose s2,
// bias, ReLU-->  a_last (32 ch)  --Conv2DTranspose s1, bias, sigmoid-->  x_hat  [--> err, score]
// The 32-channel full-resolution activation (73 % of all activation bytes of the model) only
// ever exists as a bf16 halo tile in SHARED MEMORY:
//   phase A  3 M-tiles x 5 paired-tap MMAs (N=32) on the low-res a_prev tile (TMA) -> TMEM
//   epi A    TMEM -> bias, ReLU, bf16 -> the [chunk][pixel] halo tile of a_last in smem
//            (zeros outside the image = SAME padding of the next layer)
//   phase B  8 M-tiles x 18 MMAs (N=16) on that tile -> TMEM          (as tc_out_conv_kernel)
//   epi B    TMEM -> bias, sigmoid -> x_hat and/or err = sum_c (x - x_hat)^2, score partials
// TMEM: 2 x 128 columns (phase A slots) + 2 x 128 columns (phase B buffers) = 512.
// The a_last smem tile and the phase B accumulators are double buffered, so the tensor pipe
// works on tile t+1 (phase A) while the epilogue warps finish tile t.
constexpr int PA = 20;                     // pitch of the low-res a_prev tile
constexpr int A3ROWS = 21;
constexpr uint32_t A3_BYTES = A3ROWS * PA * 16;
constexpr uint32_t A3_STAGE_BYTES = ((A3_BYTES + 128 + 127) / 128) * 128;   // TMA destinations must stay 128-byte aligned
constexpr int LR_ROWS = 18, LR_COLS = 17;  // low-res pixels whose 2x2 phases cover the 34x32 halo tile
constexpr int MTA = 3;                     // phase A M-tiles (LR_ROWS * PA = 360 <= 384)

struct TailParams {
  const __nv_bfloat16* wimgA;  // convT images [5][2][32][8]
  const __nv_bfloat16* wimgB;  // out-conv images [9][2][2][16][8]
  const float* biasA;          // [32]
  const float* biasB;          // [Cout]
  const float* x;              // [B,H,W,Cout] fp32 (needed for err / score) or nullptr
  float* xhat;                 // [B,H,W,Cout] or nullptr
  uint4* a_last;               // chunk-planar bf16 [B][4][H][W][8] copy of the intermediate activation (training) or nullptr
  uint32_t* relu_bits;         // [B,H,W]: bit c = (a_last[.., c] > 0), the ReLU mask the output-layer dgrad needs (training) or nullptr
  float* err;                  // [B,H,W] or nullptr
  float* score_partial;        // [num_tiles][4][3] (sum, min, max of err per epilogue warp) or nullptr
  int B, H, W, Cout;
  int tiles_y, tiles_x, num_tiles;
  int apply_sigmoid;
  int xrow;                    // floats per shared-memory row of the frame tile (TMA box: xrow x TR), 0 without x
  uint32_t x_stage;            // bytes per frame-tile stage (multiple of 128)
  int* error_flag;
};

// C2I ("col2im") form of the output convolution, Cout <= 3: instead of nine shifted K=32 products per output
// tile (144 MMAs whose A operand is re-read from shared memory nine times - the tensor pipe then idles on
// operand fetch), ONE product T[q][(tap,co)] = sum_ci a[q][ci] W[tap][co][ci] per halo position q (18 MMAs,
// N = 32), and the nine taps are summed as shifted reads of T:  out[r,c,co] = sum_tap T[(r+2-kh, c+2-kw)][tap,co].
// T travels TMEM -> shared memory through a 3-deep ring of M-tiles (4 halo rows each) in [n][q] order, so
// both the transposing stores and the shifted loads are conflict-free.
constexpr int C2I_MT = 9;       // M-tiles of 128 halo positions (34 x 32 = 1088 -> 8.5)
constexpr int C2I_RING = 3;     // shared-memory ring depth (M-tiles)
constexpr int C2I_TSLOTS = 4;   // TMEM ring of 32-column accumulators
constexpr int C2I_NV = 27;      // 9 taps x 3 channels
constexpr uint32_t C2I_T_BYTES = C2I_NV * C2I_RING * 128 * 4;

template <bool C2I>
__global__ void __launch_bounds__(kThreadsT, 1)
tc_tail_fused_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapx, TailParams p) {
  constexpr uint32_t CH = NPIX * 16;
  constexpr uint32_t A4_BYTES = 4 * CH;
  constexpr uint32_t A4_STAGE = A4_BYTES + 128;
  constexpr uint32_t A3_STAGE = A3_STAGE_BYTES;
  constexpr uint32_t WA_BYTES = 5 * 2 * 32 * 16, WB_BYTES = C2I ? 2 * 2 * 32 * 16 : 3 * 2 * 2 * TAILB_N * 16;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* s_a4 = smem;                               // 2 stages
  unsigned char* s_a3 = smem + 2 * A4_STAGE;                // 2 stages
  unsigned char* s_x = s_a3 + 2 * A3_STAGE;                 // 2 stages of the frame tile the error is taken against (scoring)
  unsigned char* s_wA = s_x + 2 * p.x_stage;
  unsigned char* s_wB = s_wA + WA_BYTES;
  float* s_T = reinterpret_cast<float*>(s_wB + WB_BYTES);   // C2I: [C2I_NV][C2I_RING * 128]
  __shared__ uint64_t a3_full[2], a3_empty[2], Afull[2], Aempty[2], a4_ready[2], a4_free[2], Bfull[2], Bempty[2];
  __shared__ uint64_t Tfull[C2I_TSLOTS], Tempty[C2I_TSLOTS], x_full[2], x_empty[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TAIL_TIMING_DECL

  for (int i = threadIdx.x; i < (int)(WA_BYTES / 16); i += kThreadsT) reinterpret_cast<uint4*>(s_wA)[i] = reinterpret_cast<const uint4*>(p.wimgA)[i];
  for (int i = threadIdx.x; i < (int)(WB_BYTES / 16); i += kThreadsT) reinterpret_cast<uint4*>(s_wB)[i] = reinterpret_cast<const uint4*>(p.wimgB)[i];
  if (threadIdx.x < 32) {   // pads behind the tiles: only ever read into discarded rows/columns
    const int s = (threadIdx.x >> 3) & 1, j = threadIdx.x & 7;
    if (threadIdx.x < 16) reinterpret_cast<uint4*>(s_a4 + s * A4_STAGE + A4_BYTES)[j] = make_uint4(0, 0, 0, 0);
    else reinterpret_cast<uint4*>(s_a3 + s * A3_STAGE + A3_BYTES)[j] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a3_full[s], 1); mbar_init(&a3_empty[s], 1);
      mbar_init(&Afull[s], 1);   mbar_init(&Aempty[s], 8);
      mbar_init(&a4_ready[s], 8); mbar_init(&a4_free[s], 1);
      mbar_init(&Bfull[s], 1);   mbar_init(&Bempty[s], 4);
      mbar_init(&x_full[s], 1);  mbar_init(&x_empty[s], 4);
    }
    for (int s = 0; s < C2I_TSLOTS; ++s) { mbar_init(&Tfull[s], 1); mbar_init(&Tempty[s], 4); }
    fence_mbar_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer: low-res a_prev tiles =====================
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int s = it & 1;
        if (!TWAIT(0, &a3_empty[s], ((it >> 1) & 1) ^ 1)) { *p.error_flag = 1; break; }
        const int n = t / (p.tiles_y * p.tiles_x);
        const int rem = t % (p.tiles_y * p.tiles_x);
        const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
        mbar_expect_tx(&a3_full[s], A3_BYTES);
        tma_load_3d(s_a3 + s * A3_STAGE, &tmap, &a3_full[s], ((tx * TW) / 2 - 2) * 8, (ty * TR) / 2 - 2, n);
        if (p.x) {   // the frame tile lands a whole tile ahead of the epilogue that compares against it
          if (!TWAIT(11, &x_empty[s], ((it >> 1) & 1) ^ 1)) { *p.error_flag = 1; break; }
          mbar_expect_tx(&x_full[s], (uint32_t)(p.xrow * TR * 4));
          tma_load_3d(s_x + s * p.x_stage, &tmapx, &x_full[s], (tx * TW * p.Cout) & ~3, ty * TR, n);   // 16-byte aligned box start
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (both phases) ===============================
    // Issue order  A(0), A(1), B(0), A(2), B(1), ...  : phase A of the NEXT tile is queued in
    // front of phase B of the current one, so while the tensor pipe grinds through the 144 phase-B
    // MMAs the epilogue warps already turn A(t+1) into the next shared-memory tile.
    const bool leader = elect_one();
    const uint32_t idescA = make_idesc_bf16_f32(128, 32), idescB = make_idesc_bf16_f32(128, TAILB_N);
    const uint64_t wA0 = make_desc_kmajor_noswz(smem_u32(s_wA), 32 * 16, 128);
    const uint64_t wB0 = make_desc_kmajor_noswz(smem_u32(s_wB), TAILB_N * 16, 128);
    int ma = 0, mb = 0;
    bool ok = true;
    // M-tiles [mt_lo, mt_hi) of phase A of tile itA
    auto issue_A = [&](int itA, int mt_lo, int mt_hi) {
      const int s = itA & 1;
      const uint32_t ph = (itA >> 1) & 1;
      if (mt_lo == 0) {
        if (!TWAIT(1, &a3_full[s], ph)) { if (leader) *p.error_flag = 1; ok = false; return; }
        fence_after_sync();
      }
      const uint32_t a3b = smem_u32(s_a3 + s * A3_STAGE);
#pragma unroll 1
      for (int mt = mt_lo; mt < mt_hi; ++mt, ++ma) {
        const int slot = ma & 1;
        if (!TWAIT(2, &Aempty[slot], ((ma >> 1) & 1) ^ 1)) { if (leader) *p.error_flag = 1; ok = false; return; }
        fence_after_sync();
        const uint32_t d0 = tmem + (uint32_t)(slot * 128);
        const uint32_t q0 = a3b + (uint32_t)(mt * 128) * 16;
        const uint64_t a0 = make_desc_kmajor_noswz(q0, 16, 128);
        const uint64_t a1 = make_desc_kmajor_noswz(q0 + PA * 16, 16, 128);
        const uint64_t a2 = make_desc_kmajor_noswz(q0 + 16, PA * 16, 128);
        const uint64_t a4d = make_desc_kmajor_noswz(q0 + (PA + 1) * 16, 16, 128);
        if (leader) {
          mma_bf16_ss(d0 + 0, a0, wA0, idescA, 0);
          mma_bf16_ss(d0 + 0, a1, desc_advance(wA0, 64), idescA, 1);
          mma_bf16_ss(d0 + 32, a2, desc_advance(wA0, 128), idescA, 0);
          mma_bf16_ss(d0 + 64, a1, desc_advance(wA0, 192), idescA, 0);
          mma_bf16_ss(d0 + 96, a4d, desc_advance(wA0, 256), idescA, 0);
          mma_commit(&Afull[slot]);
        }
        __syncwarp();
      }
      if (mt_hi == MTA) {
        if (leader) mma_commit(&a3_empty[s]);
        __syncwarp();
      }
    };
    int it = 0;
    if ((int)blockIdx.x < p.num_tiles) issue_A(0, 0, MTA);
    for (int t = blockIdx.x; t < p.num_tiles && ok; t += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t ph = (it >> 1) & 1;
      // Only two TMEM slots hold phase-A accumulators, so the third M-tile of A(it+1) needs the first one drained
      // by the epilogue warps: it is issued from the MIDDLE of phase B below, where that wait no longer idles the
      // tensor pipe.
      const bool more = t + (int)gridDim.x < p.num_tiles;
      if (more) { issue_A(it + 1, 0, C2I ? MTA : 2); if (!ok) break; }
      // ---- phase B on the smem tile the epilogue warps produced from A(it)
      if (!TWAIT(4, &a4_ready[s], ph)) { if (leader) *p.error_flag = 1; break; }
      if constexpr (C2I) {
        fence_after_sync();
        const uint64_t da0 = make_desc_kmajor_noswz(smem_u32(s_a4 + s * A4_STAGE), CH, 128);
        const uint64_t wT0 = make_desc_kmajor_noswz(smem_u32(s_wB), 32 * 16, 128);
#pragma unroll 1
        for (int mt = 0; mt < C2I_MT; ++mt, ++mb) {
          const int slot = mb & (C2I_TSLOTS - 1);
          if (!TWAIT(3, &Tempty[slot], ((mb / C2I_TSLOTS) & 1) ^ 1)) { if (leader) *p.error_flag = 1; ok = false; break; }
          fence_after_sync();
          const uint32_t d_tmem = tmem + (uint32_t)(256 + slot * 32);
          const uint64_t da_mt = desc_advance(da0, (uint32_t)(mt * 128));
          if (leader) {
            mma_bf16_ss(d_tmem, da_mt, wT0, idescA, 0);                                                   // channels 0..15
            mma_bf16_ss(d_tmem, desc_advance(da_mt, 2 * (CH / 16)), desc_advance(wT0, 2 * 32), idescA, 1);  // channels 16..31
            mma_commit(&Tfull[slot]);
          }
          __syncwarp();
        }
        if (!ok) break;
        if (leader) mma_commit(&a4_free[s]);
        __syncwarp();
      } else {
        // Two TMEM buffers of four M-tiles (32 columns each: 3 horizontal taps x 8 channels); the epilogue drains one
        // half while the other is computed.  Per M-tile: 3 vertical taps x 2 K steps.
        const uint64_t da0 = make_desc_kmajor_noswz(smem_u32(s_a4 + s * A4_STAGE), CH, 128);
        for (int hh = 0; hh < 2 && ok; ++hh) {
          if (!TWAIT(5, &Bempty[hh], (uint32_t)((it & 1) ^ 1))) { if (leader) *p.error_flag = 1; ok = false; break; }
          fence_after_sync();
#pragma unroll 2
          for (int mi = 0; mi < MT / 2; ++mi) {
            const int mt = hh * (MT / 2) + mi;
            const uint32_t d_tmem = tmem + (uint32_t)(256 + hh * 128 + mi * TAILB_N);
            const uint64_t da_mt = desc_advance(da0, (uint32_t)(mt * 128));
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t da = desc_advance(da_mt, (uint32_t)(2 * ks) * (CH / 16) + (uint32_t)((2 - kh) * PW));
                const uint64_t db = desc_advance(wB0, (uint32_t)((kh * 2 + ks) * 2 * TAILB_N));
                if (leader) mma_bf16_ss(d_tmem, da, db, idescB, (kh | ks) != 0);
              }
            }
          }
          if (leader) mma_commit(&Bfull[hh]);
          __syncwarp();
          if (hh == 0 && more) { issue_A(it + 1, 2, MTA); if (!ok) break; }   // (needs the first phase-A slot drained: see above)
        }
        if (!ok) break;
        if (leader) mma_commit(&a4_free[s]);
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue warps ========================================
    // Two independent groups: warps 2..9 (two per TMEM lane group, one per output-row parity) turn
    // the phase-A accumulators into the next shared-memory tile, warps 10..13 drain phase B.
    // Neither waits for the other, so both overlap the phase-B MMAs of the tile in flight.  Biases
    // live in registers: with one or two warps per scheduler every shared-memory round trip in the
    // per-element chain is exposed latency.
    const int lg = warp & 3;
    const bool groupA = warp < 10;
    const int halfA = (warp - 2) >> 2;
    bool ok = true;
    auto tile_origin = [&](int t, int& n, int& ty0, int& tx0) {
      n = t / (p.tiles_y * p.tiles_x);
      const int rem = t % (p.tiles_y * p.tiles_x);
      ty0 = (rem / p.tiles_x) * TR; tx0 = (rem % p.tiles_x) * TW;
    };
    if (groupA) {
      // ---- epilogue A: low-res phases -> bf16 halo tile of a_last in shared memory
      float bA[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) bA[c] = __ldg(p.biasA + c);
      int ma = 0, it = 0;
      for (int t = blockIdx.x; t < p.num_tiles && ok; t += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        int n, ty0, tx0;
        tile_origin(t, n, ty0, tx0);
        if (!TWAIT(6, &a4_free[s], ph ^ 1)) { if (lane == 0) *p.error_flag = 1; break; }   // B(it-2) done with this stage
        unsigned char* a4s = s_a4 + s * A4_STAGE;
#pragma unroll 1
        for (int mt = 0; mt < MTA; ++mt, ++ma) {
          const int slot = ma & 1;
          if (!TWAIT(7, &Afull[slot], (ma >> 1) & 1)) { if (lane == 0) *p.error_flag = 1; ok = false; break; }
          fence_after_sync();
          const int q = mt * 128 + lg * 32 + lane;
          const int r = q / PA, c = q % PA;
          const bool live = r < LR_ROWS && c < LR_COLS;
#pragma unroll 1
          for (int pb = 0; pb < 2; ++pb) {
            const int pa = halfA, phs = halfA * 2 + pb;
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(slot * 128 + phs * 32), v);
            const int hr = 2 * r + pa - 1, hc = 2 * c + pb - 1;
            if (live && hr >= 0 && hr < PR && hc >= 0 && hc < PW) {
              const int Y = ty0 - 1 + hr, X = tx0 - 1 + hc;
              const bool inside = Y >= 0 && Y < p.H && X >= 0 && X < p.W;
              const bool keep = p.a_last && inside && hr >= 1 && hr <= TR && hc >= 1 && hc <= TW;
              uint4* dst = reinterpret_cast<uint4*>(a4s) + (hr * PW + hc);
              uint4* gdst = keep ? p.a_last + (int64_t)n * 4 * ((int64_t)p.H * p.W) + (int64_t)Y * p.W + X : nullptr;
              // bias + ReLU once per channel; positions outside the image become zero words (SAME padding of the next layer)
#pragma unroll
              for (int c2 = 0; c2 < 32; ++c2) v[c2] = fmaxf(v[c2] + bA[c2], 0.f);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint32_t w4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __nv_bfloat162 b2 = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                  w4[e] = inside ? *reinterpret_cast<uint32_t*>(&b2) : 0u;
                }
                const uint4 u = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                dst[g * (CH / 16)] = u;
                // training: the backward needs this activation; every tile stores its own 32x30 interior
                if (keep) gdst[g * ((int64_t)p.H * p.W)] = u;
              }
              if (keep && p.relu_bits) {   // one 32-bit word per pixel instead of a 64-byte re-read in the dgrad
                uint32_t bits = 0;
#pragma unroll
                for (int c2 = 0; c2 < 32; ++c2) bits |= (__float_as_int(v[c2]) > 0 ? 1u : 0u) << c2;   // v >= +0 after the ReLU
                p.relu_bits[((int64_t)n * p.H + Y) * p.W + X] = bits;
              }
            }
          }
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&Aempty[slot]);
        }
        if (!ok) break;
        fence_async_smem();          // st.shared of the tile -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&a4_ready[s]);
      }
    } else {
      // ---- epilogue B: bias, sigmoid, reconstruction error, score partials
      float bB[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) bB[c] = c < p.Cout ? __ldg(p.biasB + c) : 0.f;
      if constexpr (C2I) {
        constexpr int RW = C2I_RING * 128;          // ring pixels per value plane
        const int u = lg * 32 + lane;               // position inside an M-tile this thread moves to shared memory
        int mb = 0, it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
          int n, ty0, tx0;
          tile_origin(t, n, ty0, tx0);
          float esum = 0.f, emin = 3.4e38f, emax = -3.4e38f;
          const float* xs = reinterpret_cast<const float*>(s_x + (it & 1) * p.x_stage) + ((tx0 * p.Cout) & 3);
          if (p.x && !TWAIT(10, &x_full[it & 1], (it >> 1) & 1)) { if (lane == 0) *p.error_flag = 1; break; }
#pragma unroll 1
          for (int mt = 0; mt < C2I_MT; ++mt, ++mb) {
            // output rows that become complete with this M-tile (halo rows 4mt..4mt+3): 4mt-2 .. 4mt+1
            const int r0 = (mt == 0 ? 0 : 4 * mt - 2) + lg;
            const int r_hi = 4 * mt + 1 < TR - 1 ? 4 * mt + 1 : TR - 1;
            const int oy = ty0 + r0, ox = tx0 + lane;
            const bool live = r0 <= r_hi && lane < TW && oy < p.H && ox < p.W;
            const int64_t pix = ((int64_t)n * p.H + oy) * p.W + ox;
            float xv[3];
#pragma unroll
            for (int co = 0; co < 3; ++co) xv[co] = (p.x && live && co < p.Cout) ? xs[r0 * p.xrow + lane * p.Cout + co] : 0.f;
            const int slot = mb & (C2I_TSLOTS - 1);
            if (!TWAIT(8, &Tfull[slot], (mb / C2I_TSLOTS) & 1)) { if (lane == 0) *p.error_flag = 1; }
            fence_after_sync();
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(256 + slot * 32), v);
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&Tempty[slot]);
            float* dstT = s_T + (mt % C2I_RING) * 128 + u;
#pragma unroll
            for (int k = 0; k < C2I_NV; ++k) dstT[k * RW] = v[k];
            asm volatile("bar.sync 1, 128;\n" ::: "memory");     // the four phase-B warps
            // nine shifted reads per channel; dead lanes read in-range garbage and drop it
            float acc[3] = {bB[0], bB[1], bB[2]};
            const int rr = r0 <= r_hi ? r0 : r_hi, cc = lane < TW ? lane : TW - 1;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const int qg = (rr + 2 - kh) * PW + (cc + 2 - kw);
                const float* src = s_T + ((qg >> 7) % C2I_RING) * 128 + (qg & 127) + (kh * 3 + kw) * 3 * RW;
#pragma unroll
                for (int co = 0; co < 3; ++co) acc[co] += src[co * RW];
              }
            float e = 0.f, y[3];
#pragma unroll
            for (int co = 0; co < 3; ++co) {
              y[co] = p.apply_sigmoid ? __fdividef(1.0f, 1.0f + __expf(-acc[co])) : acc[co];
              const float d = co < p.Cout ? xv[co] - y[co] : 0.f;
              e = fmaf(d, d, e);
            }
            if (live) {
              if (p.xhat) {
#pragma unroll
                for (int co = 0; co < 3; ++co)
                  if (co < p.Cout) p.xhat[pix * p.Cout + co] = y[co];
              }
              if (p.err) p.err[pix] = e;
              esum += e; emin = fminf(emin, e); emax = fmaxf(emax, e);
            }
          }
          if (p.score_partial) {     // fixed shuffle tree -> deterministic
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              esum += __shfl_xor_sync(0xffffffffu, esum, o);
              emin = fminf(emin, __shfl_xor_sync(0xffffffffu, emin, o));
              emax = fmaxf(emax, __shfl_xor_sync(0xffffffffu, emax, o));
            }
            if (lane == 0) {
              float* o3 = p.score_partial + ((int64_t)t * 4 + lg) * 3;
              o3[0] = esum; o3[1] = emin; o3[2] = emax;
            }
          }
          if (p.x) { __syncwarp(); if (lane == 0) mbar_arrive(&x_empty[it & 1]); }
        }
      } else {
      int it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        const int s = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        int n, ty0, tx0;
        tile_origin(t, n, ty0, tx0);
        const float* xs = reinterpret_cast<const float*>(s_x + s * p.x_stage) + ((tx0 * p.Cout) & 3);
        if (p.x && !TWAIT(10, &x_full[s], ph)) { if (lane == 0) *p.error_flag = 1; break; }
        float esum = 0.f, emin = 3.4e38f, emax = -3.4e38f;
        // Straight-line code over a fixed channel count (channels >= Cout are computed and dropped): the serial
        // branchy per-channel form left this single warp per scheduler waiting on one sigmoid chain at a time.
        auto finish = [&](auto cn, int mt, const float (&v)[8]) {
          constexpr int CN = decltype(cn)::value;
          const int q = mt * 128 + lg * 32 + lane;
          const int r = q / PW, c = q % PW;
          const int oy = ty0 + r, ox = tx0 + c;
          const bool live = c < TW && oy < p.H && ox < p.W;
          const int64_t pix = ((int64_t)n * p.H + oy) * p.W + ox;
          const float* xr = xs + r * p.xrow + c * p.Cout;
          float xv[CN], y[CN];
#pragma unroll
          for (int co = 0; co < CN; ++co) xv[co] = (p.x && live && co < p.Cout) ? xr[co] : 0.f;
#pragma unroll
          for (int co = 0; co < CN; ++co) {
            y[co] = v[co] + bB[co];
            if (p.apply_sigmoid) y[co] = __fdividef(1.0f, 1.0f + __expf(-y[co]));
          }
          float e = 0.f;
#pragma unroll
          for (int co = 0; co < CN; ++co) {
            const float d = co < p.Cout ? xv[co] - y[co] : 0.f;
            e = fmaf(d, d, e);
          }
          if (live) {
            if (p.xhat) {
#pragma unroll
              for (int co = 0; co < CN; ++co)
                if (co < p.Cout) p.xhat[pix * p.Cout + co] = y[co];
            }
            if (p.err) p.err[pix] = e;
            esum += e; emin = fminf(emin, e); emax = fmaxf(emax, e);
          }
        };
        bool okb = true;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {            // the two TMEM buffers of four M-tiles (see the MMA issuer)
          if (!TWAIT(9, &Bfull[hh], (uint32_t)(it & 1))) { if (lane == 0) *p.error_flag = 1; okb = false; break; }
          fence_after_sync();
          const uint32_t tb = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(256 + hh * 128);
#pragma unroll 1
          for (int mi = 0; mi < MT / 2; mi += 2) {
            // columns kw*8 + co: partial sums over the vertical taps and channels for horizontal tap kw, taken at THIS
            // halo position; output pixel c needs tap 2 from lane c, tap 1 from lane c+1, tap 0 from lane c+2 (lanes = the
            // 32 pixels of one halo row; c < 30 never reaches past lane 31).  Two M-tiles per pass: their loads, shuffles
            // and sigmoid chains interleave (this role has one warp per scheduler).
            float v0[32], v1[32], acc0[8], acc1[8];
            tmem_ld32x2(tb + (uint32_t)(mi * TAILB_N), tb + (uint32_t)((mi + 1) * TAILB_N), v0, v1);
            if (p.Cout <= 4) {
#pragma unroll
              for (int co = 0; co < 4; ++co) {
                acc0[co] = v0[16 + co] + __shfl_down_sync(0xffffffffu, v0[8 + co], 1) + __shfl_down_sync(0xffffffffu, v0[co], 2);
                acc1[co] = v1[16 + co] + __shfl_down_sync(0xffffffffu, v1[8 + co], 1) + __shfl_down_sync(0xffffffffu, v1[co], 2);
              }
#pragma unroll
              for (int co = 4; co < 8; ++co) { acc0[co] = 0.f; acc1[co] = 0.f; }
              finish(std::integral_constant<int, 4>{}, hh * (MT / 2) + mi, acc0);
              finish(std::integral_constant<int, 4>{}, hh * (MT / 2) + mi + 1, acc1);
            } else {
#pragma unroll
              for (int co = 0; co < 8; ++co) {
                acc0[co] = v0[16 + co] + __shfl_down_sync(0xffffffffu, v0[8 + co], 1) + __shfl_down_sync(0xffffffffu, v0[co], 2);
                acc1[co] = v1[16 + co] + __shfl_down_sync(0xffffffffu, v1[8 + co], 1) + __shfl_down_sync(0xffffffffu, v1[co], 2);
              }
              finish(std::integral_constant<int, 8>{}, hh * (MT / 2) + mi, acc0);
              finish(std::integral_constant<int, 8>{}, hh * (MT / 2) + mi + 1, acc1);
            }
          }
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&Bempty[hh]);
        }
        if (!okb) break;
        if (p.x) { __syncwarp(); if (lane == 0) mbar_arrive(&x_empty[s]); }
        if (p.score_partial) {     // fixed shuffle tree -> deterministic
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            esum += __shfl_xor_sync(0xffffffffu, esum, o);
            emin = fminf(emin, __shfl_xor_sync(0xffffffffu, emin, o));
            emax = fmaxf(emax, __shfl_xor_sync(0xffffffffu, emax, o));
          }
          if (lane == 0) {
            float* o3 = p.score_partial + ((int64_t)t * 4 + lg) * 3;
            o3[0] = esum; o3[1] = emin; o3[2] = emax;
          }
        }
      }
      }   // !C2I
    }
  }
  TAIL_TIMING_DUMP
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// per-frame reduction of the per-tile/per-warp partials written by tc_tail_fused_kernel:
// one block per frame, fixed strided order + fixed shuffle/shared-memory tree (deterministic)
__global__ void __launch_bounds__(128) tail_score_finish_kernel(const float* partial, int entries_per_frame, float* score,
                                                                float* err_minmax) {
  __shared__ float red[3][4];
  const int b = blockIdx.x;
  const float* p0 = partial + (int64_t)b * entries_per_frame * 3;
  float s = 0.f, mn = 3.4e38f, mx = -3.4e38f;
  for (int i = threadIdx.x; i < entries_per_frame; i += 128) {
    s += p0[i * 3]; mn = fminf(mn, p0[i * 3 + 1]); mx = fmaxf(mx, p0[i * 3 + 2]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = mn; red[2][threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    score[b] = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    if (err_minmax) {
      err_minmax[2 * b] = fminf(fminf(red[1][0], red[1][1]), fminf(red[1][2], red[1][3]));
      err_minmax[2 * b + 1] = fmaxf(fmaxf(red[2][0], red[2][1]), fmaxf(red[2][2], red[2][3]));
    }
  }
}

// fp32 NHWC [B,HW,C] (C % 8 == 0) -> chunk-planar bf16 [B][C/8][HW][8]: thread = one 16-byte unit
__global__ void cast_f32_bf16_planar_kernel(const float* in, uint4* out, int B, int64_t HW, int KC) {
  const int64_t total = (int64_t)B * KC * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t px = i % HW;
    const int g = (int)((i / HW) % KC);
    const int64_t n = i / (HW * KC);
    const float4* src = reinterpret_cast<const float4*>(in + ((n * HW + px) * KC + g) * 8);
    const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
    __nv_bfloat162 a = __floats2bfloat162_rn(v0.x, v0.y), b = __floats2bfloat162_rn(v0.z, v0.w);
    __nv_bfloat162 c = __floats2bfloat162_rn(v1.x, v1.y), d = __floats2bfloat162_rn(v1.z, v1.w);
    out[i] = make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                        *reinterpret_cast<uint32_t*>(&d));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

#ifdef KCVAE_TAIL_TIMING
// KCVAE_TAIL_DBG=<kernel tag>: print the per-warp wait cycles of CTA 0 after that launch (synchronises the stream)
static void tc_timing_dump(const char* tag, const char* const* names, int nwarps, double tiles_per_cta, cudaStream_t st) {
  const char* e = std::getenv("KCVAE_TAIL_DBG");
  if (!e || std::strcmp(e, tag) != 0) return;
  cudaStreamSynchronize(st);
  long long hd[24 * 16];
  cudaMemcpyFromSymbol(hd, g_tail_dbg, sizeof(hd));
  std::fprintf(stderr, "%s timing (CTA 0, cycles; tiles/CTA %.1f):\n", tag, tiles_per_cta);
  for (int wv = 0; wv < nwarps; ++wv) {
    std::fprintf(stderr, "  warp %2d total %9lld |", wv, hd[wv * 16 + 14]);
    for (int i = 0; i < 14; ++i) if (hd[wv * 16 + i]) std::fprintf(stderr, " %s %lld", names[i], hd[wv * 16 + i]);
    std::fprintf(stderr, "\n");
  }
}
#endif

bool tc_out_conv_supported(int Cin, int Cout) { return (Cin == 16 || Cin == 32) && Cout >= 1 && Cout <= 8; }

size_t tc_out_weight_image_elems(int Cin) { return (size_t)9 * (Cin / 16) * 2 * NPAD * 8; }

void cast_f32_to_bf16_planar(const float* in, void* out, int B, int64_t HW, int C, cudaStream_t st) {
  ProfScope prof_("cast_bf16", st);
  ++g_launches;
  cast_f32_bf16_planar_kernel<<<grid_for((int64_t)B * HW * (C / 8), 256, 8, 2), 256, 0, st>>>(in, reinterpret_cast<uint4*>(out), B, HW, C / 8);
}
// tensor map of a chunk-planar bf16 activation [B][Cin/8][H][W][8]: dims (W*8, H, B*Cin/8), box = one halo plane
static CUresult make_planar_tmap(CUtensorMap* tmap, const void* act, int B, int H, int W, int Cin, int rows = PR) {
  EncodeTiledFn enc = get_encode_fn();
  const cuuint64_t gdim[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)B * (Cin / 8)};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  const cuuint32_t box[3] = {PW * 8, (cuuint32_t)rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(act), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// tensor map of a bf16 NHWC tensor padded to 8 channels (one 16-byte unit per pixel) seen as [B][H][W*8]:
// box = `rows` rows of `cols` pixels, each row cols*16 contiguous bytes (a 4-D {8, W, H, B} map moves the same box as
// rows*cols separate 16-byte pieces)
static CUresult make_c8_tmap(CUtensorMap* tmap, const void* t8, int B, int H, int W, int cols, int rows) {
  EncodeTiledFn enc = get_encode_fn();
  const cuuint64_t gdim[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  const cuuint32_t box[3] = {(cuuint32_t)cols * 8, (cuuint32_t)rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(t8), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

void tc_prep_out_weights(const float* w, int Cout, int Cin, void* img, cudaStream_t st) {
  ProfScope prof_("tc_prep_weights", st);
  ++g_launches;
  tc_prep_out_weights_kernel<<<8, 256, 0, st>>>(w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(img));
}


bool tc_out_dgrad_supported(int Cin, int Cout) { return Cin == 32 && Cout >= 1 && Cout <= 8; }
size_t tc_dgrad_weight_image_elems() { return (size_t)5 * 2 * NPAD_D * 8; }

void tc_prep_dgrad_weights(const float* w, int Cout, int Cin, void* img, cudaStream_t st) {
  ProfScope prof_("tc_prep_weights", st);
  ++g_launches;
  tc_prep_dgrad_weights_kernel<<<4, 256, 0, st>>>(w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(img));
}

int tc_out_dgrad(const void* dl8_bf16, const void* wimg, const void* mask_bf16, const uint32_t* relu_bits, float* g_out,
                 void* g_s2d_bf16, float* chan_sum, float* chan_partial, int B, int H, int W, int Cin, int* error_flag,
                 cudaStream_t st) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return 1;
  CUtensorMap tmap;
  CUresult r = make_c8_tmap(&tmap, dl8_bf16, B, H, W, PW, PR);
  if (r != CUDA_SUCCESS) return 2;
  OutDgradParams p{};
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.mask = reinterpret_cast<const __nv_bfloat16*>(mask_bf16);
  p.relu_bits = relu_bits;
  p.g_out = g_out; p.g_s2d = reinterpret_cast<__nv_bfloat16*>(g_s2d_bf16); p.B = B; p.H = H; p.W = W; p.Cin = Cin;
  p.tiles_y = cdiv(H, TR); p.tiles_x = cdiv(W, TW);
  p.num_tiles = B * p.tiles_y * p.tiles_x;
  p.error_flag = error_flag;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  const size_t smem = (size_t)kStages * ((size_t)NPIX * 16 + 128) + (size_t)5 * 2 * NPAD_D * 16 + (size_t)4 * kSubD * 2048;
  ProfScope prof_("tc_out_dgrad", st);
  ++g_launches;
  p.chan_partial = chan_sum ? chan_partial : nullptr;
  if (relu_bits) {
    cudaFuncSetAttribute(tc_out_dgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_out_dgrad_kernel<true><<<grid, kThreadsD, smem, st>>>(tmap, p);
  } else {
    cudaFuncSetAttribute(tc_out_dgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_out_dgrad_kernel<false><<<grid, kThreadsD, smem, st>>>(tmap, p);
  }
#ifdef KCVAE_TAIL_TIMING
  {
    static const char* nm[14] = {"smem_empty", "tmem_empty", "tma_full", "tmem_full", "tmem_ld", "math_store", "", "", "", "", "", "", "", ""};
    tc_timing_dump("out_dgrad", nm, kThreadsD / 32, (double)p.num_tiles / grid, st);
  }
#endif
  if (chan_sum) {
    sum_partials(chan_partial, grid * 4 * kSubD, Cin, chan_sum, st);
  }
  return 0;
}

bool tc_out_wgrad_supported(int Cin, int Cout) { return Cin == 32 && Cout >= 1 && Cout <= 8; }
size_t tc_out_wgrad_partial_floats(int Cin, int Cout) { return (size_t)kNumSMs * 9 * Cout * Cin; }

// dW [3,3,Cout,Cin] (Keras Conv2DTranspose layout) from dl8 [B,H,W,8] bf16 and act [B,H,W,Cin] bf16
int tc_out_wgrad(const void* dl8_bf16, const void* act_bf16, float* dW, float* partial, int B, int H, int W, int Cin,
                 int Cout, int* error_flag, cudaStream_t st) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return 1;
  CUtensorMap tmap;
  CUresult r = make_planar_tmap(&tmap, act_bf16, B, H, W, Cin);
  if (r != CUDA_SUCCESS) return 2;
  OutWgradParams p{};
  p.partial = partial; p.B = B; p.H = H; p.W = W; p.Cout = Cout; p.Cin = Cin;
  p.tiles_y = cdiv(H, TR); p.tiles_x = cdiv(W, TW);
  p.num_tiles = B * p.tiles_y * p.tiles_x;
  p.error_flag = error_flag;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  CUtensorMap tmap_dl;
  r = make_c8_tmap(&tmap_dl, dl8_bf16, B, H, W, TW, TR);
  if (r != CUDA_SUCCESS) return 2;
  const size_t smem = OW_SMEM;
  const int E = 9 * Cout * Cin;
  ProfScope prof_("tc_out_wgrad", st);
  ++g_launches;
  cudaFuncSetAttribute(tc_out_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_out_wgrad_kernel<<<grid, kThreads, smem, st>>>(tmap, tmap_dl, p);
  sum_partials(partial, grid, E, dW, st);
  return 0;
}

bool tc_convT_fwd_supported(int Cin, int Cout) { return Cin >= 1 && Cin <= 8 && Cout == 32; }
size_t tc_convT_weight_image_elems() { return (size_t)5 * 2 * 32 * 8; }

void pack_c8_bf16(const float* in, int64_t npix, int C, void* out, cudaStream_t st) {
  ProfScope prof_("pack_c8", st);
  ++g_launches;
  pack_c8_bf16_kernel<<<grid_for(npix, 256, 8, 2), 256, 0, st>>>(in, npix, C, reinterpret_cast<uint4*>(out));
}

void tc_prep_convT_weights(const float* w, int Cout, int Cin, void* img, cudaStream_t st) {
  ProfScope prof_("tc_prep_weights", st);
  ++g_launches;
  tc_prep_convT_weights_kernel<<<4, 256, 0, st>>>(w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(img));
}

// out[B,2h,2w,32] bf16 = relu(bias + convT_s2(in8[B,h,w,8] bf16))
int tc_convT_fwd(const void* in8_bf16, const void* wimg, const float* bias, void* out_bf16, int B, int h, int w,
                 int* error_flag, cudaStream_t st) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return 1;
  CUtensorMap tmap;
  CUresult r = make_c8_tmap(&tmap, in8_bf16, B, h, w, PW, PR);
  if (r != CUDA_SUCCESS) return 2;
  ConvTParams p{};
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.B = B; p.h = h; p.w = w;
  p.tiles_y = cdiv(h, TR); p.tiles_x = cdiv(w, TW);
  p.num_tiles = B * p.tiles_y * p.tiles_x;
  p.error_flag = error_flag;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  const size_t smem = (size_t)kStages * ((size_t)NPIX * 16 + 128) + (size_t)5 * 2 * 32 * 16;
  ProfScope prof_("tc_convT_fwd", st);
  ++g_launches;
  cudaFuncSetAttribute(tc_convT_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_convT_fwd_kernel<<<grid, kThreadsE, smem, st>>>(tmap, p);
  return 0;
}

bool tc_convT_few_fwd_supported(int Cin, int Cout) { return Cin == 32 && Cout >= 1 && Cout <= 8; }
size_t tc_convT_few_weight_image_elems() { return (size_t)2 * 4 * 2 * 2 * 32 * 8; }   // hi image + lo image
void tc_prep_convT_few_weights(const float* w, int Cout, int Cin, void* img, cudaStream_t st) {
  ProfScope prof_("tc_prep_weights", st);
  ++g_launches;
  tc_prep_convT_few_weights_kernel<<<8, 256, 0, st>>>(w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(img));
}
// in: chunk-planar bf16 [B][4][h][w][8] (split: [B][hi 4 | lo 4][h][w][8]); out8: bf16 [B,2h,2w,8] and/or
// out_f32: fp32 [B,2h,2w,Cout] = relu(bias + convT_s2(in))
int tc_convT_few_fwd(const void* in_planar_bf16, const void* wimg, const float* bias, void* out8_bf16, float* out_f32, int B, int h,
                     int w, int Cout, int split, int* error_flag, cudaStream_t st) {
  if (!get_encode_fn()) return 1;
  CUtensorMap tmap;
  if (make_planar_tmap(&tmap, in_planar_bf16, B, h, w, split ? 64 : 32, FROWS) != CUDA_SUCCESS) return 2;
  ConvTFewParams p{};
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.out8 = reinterpret_cast<uint4*>(out8_bf16); p.out_f32 = out_f32;
  p.B = B; p.h = h; p.w = w; p.Cout = Cout;
  p.tiles_y = cdiv(h, TRF); p.tiles_x = cdiv(w, TW);
  p.num_tiles = B * p.tiles_y * p.tiles_x;
  p.error_flag = error_flag;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  const size_t smem = (size_t)kStages * ((size_t)(split ? 8 : 4) * CHF + 128) + (size_t)(split ? 2 : 1) * 4 * 2 * 2 * 32 * 16;
  ProfScope prof_("tc_convT_few_fwd", st);
  ++g_launches;
  if (split) {
    cudaFuncSetAttribute(tc_convT_few_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_convT_few_fwd_kernel<true><<<grid, kThreadsE, smem, st>>>(tmap, p);
  } else {
    cudaFuncSetAttribute(tc_convT_few_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_convT_few_fwd_kernel<false><<<grid, kThreadsE, smem, st>>>(tmap, p);
  }
  return 0;
}

bool tc_convT_bwd_supported(int Cin, int Cout, int h, int w) { return Cin >= 1 && Cin <= 8 && Cout == 32 && h > 0 && w > 0; }
size_t tc_convT_dgrad_weight_image_elems() { return (size_t)9 * 2 * 2 * 16 * 8; }
size_t tc_convT_wgrad_partial_floats(int Cin) { return (size_t)kNumSMs * 9 * 32 * Cin; }

// tensor map of the chunk-planar space-to-depth gradient [B*16 planes][h][w*8]: box = GROWS rows of PW pixels
// (512 contiguous bytes each) of one plane -> the [pixel] x 16 B chunk plane the MMAs read
static int make_s2d_map(CUtensorMap* tmap, const void* g_s2d, int B, int h, int w) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return 1;
  const cuuint64_t gdim[3] = {(cuuint64_t)w * 8, (cuuint64_t)h, (cuuint64_t)B * 16};
  const cuuint64_t gstr[2] = {(cuuint64_t)w * 16, (cuuint64_t)h * w * 16};
  const cuuint32_t box[3] = {PW * 8, GROWS, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(g_s2d), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

void tc_prep_convT_dgrad_weights(const float* w, int Cout, int Cin, void* img, cudaStream_t st) {
  ProfScope prof_("tc_prep_weights", st);
  ++g_launches;
  tc_prep_convT_dgrad_weights_kernel<<<8, 256, 0, st>>>(w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(img));
}

// g_prev[B,h,w,Cin] = (mask > 0) * stride-2 gather of G (space-to-depth bf16) with W
int tc_convT_dgrad(const void* g_s2d, const void* wimg, const float* mask, float* g_prev, int B, int h, int w, int Cin,
                   int* error_flag, cudaStream_t st, void* g_planes, int g_KC) {
  CUtensorMap tmap;
  if (int rc = make_s2d_map(&tmap, g_s2d, B, h, w)) return rc;
  ConvTBwdParams p{};
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.mask = mask; p.g_prev = g_prev; p.B = B; p.h = h; p.w = w; p.Cin = Cin;
  p.g_planes = reinterpret_cast<uint4*>(g_planes); p.g_KC = g_KC;
  if (g_planes && ((h | w) & 1)) return 3;
  p.tiles_y = cdiv(h, TRD); p.tiles_x = cdiv(w, TW);
  p.num_tiles = B * p.tiles_y * p.tiles_x;
  p.error_flag = error_flag;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  const size_t smem = (size_t)kStages * ((size_t)GT_BYTES + 128) + (size_t)9 * 2 * 2 * 16 * 16;
  ProfScope prof_("tc_convT_dgrad", st);
  ++g_launches;
  cudaFuncSetAttribute(tc_convT_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_convT_dgrad_kernel<<<grid, kThreads, smem, st>>>(tmap, p);
  return 0;
}

// dW [3,3,32,Cin] from a_prev8 [B,h,w,8] bf16 and G space-to-depth bf16
int tc_convT_wgrad(const void* g_s2d, const void* a_prev8, float* dW, float* partial, int B, int h, int w, int Cin,
                   int* error_flag, cudaStream_t st) {
  CUtensorMap tmap;
  if (int rc = make_s2d_map(&tmap, g_s2d, B, h, w)) return rc;
  ConvTBwdParams p{};
  p.a_prev8 = reinterpret_cast<const __nv_bfloat16*>(a_prev8);
  p.partial = partial; p.B = B; p.h = h; p.w = w; p.Cin = Cin;
  p.tiles_y = cdiv(h, TRD); p.tiles_x = cdiv(w, TW);
  p.num_tiles = B * p.tiles_y * p.tiles_x;
  p.error_flag = error_flag;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  const size_t smem = (size_t)kStages * ((size_t)AP_BYTES + GT_BYTES + 128);
  const int E = 9 * 32 * Cin;
  ProfScope prof_("tc_convT_wgrad", st);
  ++g_launches;
  cudaFuncSetAttribute(tc_convT_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_convT_wgrad_kernel<<<grid, kThreads, smem, st>>>(tmap, p);
  sum_partials(partial, grid, E, dW, st);
  return 0;
}

bool tc_tail_fused_supported(int Cprev, int Clast, int Cout, int H, int W) {
  return Cprev >= 1 && Cprev <= 8 && Clast == 32 && Cout >= 1 && Cout <= 8 && H % 2 == 0 && W % 2 == 0;
}
// KCVAE_TAIL_C2I=1 selects the col2im form of the output convolution (Cout <= 3).  Measured on B200 (128 frames,
// README shape): nine-tap form 0.314 ms (tensor pipe busy issuing 159 small-N MMAs per tile at ~48 cycles each),
// col2im form 0.398 ms (18 MMAs per tile, but its four epilogue warps then carry 27 shared-memory round trips per
// pixel) - so the nine-tap form is the default.
static bool tail_c2i(int Cout) {
  static const int env = [] { const char* e = std::getenv("KCVAE_TAIL_C2I"); return (e && e[0] == '1') ? 1 : 0; }();
  return env && Cout <= 3;
}
void tc_prep_tail_weights(const float* w, int Cout, int Cin, void* img, cudaStream_t st) {
  ProfScope prof_("tc_prep_weights", st);
  ++g_launches;
  if (!tail_c2i(Cout)) tc_prep_tail_kw_weights_kernel<<<8, 256, 0, st>>>(w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(img));
  else tc_prep_tail_c2i_weights_kernel<<<4, 256, 0, st>>>(w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(img));
}
void tc_prep_all_weights(const float* w_convT, const float* w_out, int Cprev, int Clast, int Cout, void* img_convT,
                         void* img_tail, void* img_dgrad, void* img_convT_dgrad, cudaStream_t st) {
  ProfScope prof_("tc_prep_weights", st);
  ++g_launches;
  PrepAllArgs a{w_convT, w_out, Cprev, Clast, Cout, tail_c2i(Cout) ? 1 : 0, reinterpret_cast<__nv_bfloat16*>(img_convT),
                reinterpret_cast<__nv_bfloat16*>(img_tail), reinterpret_cast<__nv_bfloat16*>(img_dgrad),
                reinterpret_cast<__nv_bfloat16*>(img_convT_dgrad)};
  tc_prep_all_kernel<<<dim3(8, 4), 256, 0, st>>>(a);
}
size_t tc_tail_score_partial_floats(int B, int H, int W) { return (size_t)B * cdiv(H, TR) * cdiv(W, TW) * 4 * 3; }

// fused Conv2DTranspose s2 -> Conv2DTranspose s1 (+ sigmoid, error map, per-frame score).
// in8: bf16 [B,H/2,W/2,8]; any of xhat / err / score may be nullptr (x is required for err / score).
int tc_tail_fused(const void* in8_bf16, const void* wimgA, const void* wimgB, const float* biasA, const float* biasB,
                  const float* x, float* xhat, void* a_last_planar, uint32_t* relu_bits, float* err, float* score,
                  float* err_minmax, float* score_partial, int B, int H, int W, int Cout, int apply_sigmoid, int* error_flag,
                  cudaStream_t st) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return 1;
  const int h = H / 2, w = W / 2;
  CUtensorMap tmap;
  CUresult r = make_c8_tmap(&tmap, in8_bf16, B, h, w, PA, A3ROWS);
  if (r != CUDA_SUCCESS) return 2;
  TailParams p{};
  p.wimgA = reinterpret_cast<const __nv_bfloat16*>(wimgA);
  p.wimgB = reinterpret_cast<const __nv_bfloat16*>(wimgB);
  p.biasA = biasA; p.biasB = biasB; p.x = x; p.xhat = xhat; p.err = err;
  p.a_last = reinterpret_cast<uint4*>(a_last_planar);
  p.relu_bits = a_last_planar ? relu_bits : nullptr;
  p.score_partial = score ? score_partial : nullptr;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.tiles_y = cdiv(H, TR); p.tiles_x = cdiv(W, TW);
  p.num_tiles = B * p.tiles_y * p.tiles_x;
  p.apply_sigmoid = apply_sigmoid;
  p.error_flag = error_flag;
  CUtensorMap tmapx = tmap;   // unused without x
  if (x) {
    // frame tile [TR rows][TW pixels x Cout floats, padded to 16 bytes] by TMA; rows of x must be 16-byte multiples
    if ((W * Cout) % 4 != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return 3;
    p.xrow = ((TW * Cout + 3 + 3) / 4) * 4;   // + up to 3 floats in front: the box starts 16-byte aligned
    p.x_stage = (uint32_t)(((size_t)p.xrow * TR * 4 + 127) / 128 * 128);
    const cuuint64_t xdim[3] = {(cuuint64_t)W * Cout, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t xstr[2] = {(cuuint64_t)W * Cout * 4, (cuuint64_t)H * W * Cout * 4};
    const cuuint32_t xbox[3] = {(cuuint32_t)p.xrow, TR, 1};
    const cuuint32_t xes[3] = {1, 1, 1};
    if (enc(&tmapx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), xdim, xstr, xbox, xes, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return 2;
  }
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  const size_t smem0 = (size_t)2 * ((size_t)4 * NPIX * 16 + 128) + (size_t)2 * A3_STAGE_BYTES + (size_t)5 * 2 * 32 * 16 +
                       (size_t)2 * p.x_stage;
  ProfScope prof_("tc_tail_fused", st);
  ++g_launches;
  if (tail_c2i(Cout)) {
    const size_t smem = smem0 + (size_t)2 * 2 * 32 * 16 + C2I_T_BYTES;
    cudaFuncSetAttribute(tc_tail_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_tail_fused_kernel<true><<<grid, kThreadsT, smem, st>>>(tmap, tmapx, p);
  } else {
    const size_t smem = smem0 + (size_t)3 * 2 * 2 * TAILB_N * 16;
    cudaFuncSetAttribute(tc_tail_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_tail_fused_kernel<false><<<grid, kThreadsT, smem, st>>>(tmap, tmapx, p);
  }
#ifdef KCVAE_TAIL_TIMING
  {
    static const char* nm[14] = {"a3_empty", "a3_full", "Aempty", "Tempty", "a4_ready", "Bempty", "a4_free", "Afull", "Tfull", "Bfull", "x_full", "x_empty", "tmem_ld", "epi_math"};
    tc_timing_dump("tail", nm, kThreadsT / 32, (double)p.num_tiles / grid, st);
  }
#endif
  if (score) {
    ++g_launches;
    tail_score_finish_kernel<<<B, 128, 0, st>>>(score_partial, p.tiles_y * p.tiles_x * 4, score, err_minmax);
  }
  return 0;
}

// returns 0 on success, nonzero if the tensor map could not be built
int tc_out_conv(const void* act_bf16, const void* wimg, const float* bias, float* xhat, int B, int H, int W, int Cin,
                int Cout, int apply_sigmoid, int* error_flag, cudaStream_t st) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return 1;
  CUtensorMap tmap;
  CUresult r = make_planar_tmap(&tmap, act_bf16, B, H, W, Cin);
  if (r != CUDA_SUCCESS) return 2;
  OutConvParams p{};
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.xhat = xhat; p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  p.tiles_y = cdiv(H, TR); p.tiles_x = cdiv(W, TW);
  p.num_tiles = B * p.tiles_y * p.tiles_x;
  p.apply_sigmoid = apply_sigmoid;
  p.error_flag = error_flag;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  ProfScope prof_("tc_out_conv", st);
  ++g_launches;
  auto launch = [&](auto kernel, int cin) {
    const size_t smem = (size_t)kStages * ((size_t)(cin / 8) * NPIX * 16 + 128) + (size_t)9 * (cin / 16) * 2 * NPAD * 16;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, kThreadsE, smem, st>>>(tmap, p);
  };
  if (Cin == 16) launch(tc_out_conv_kernel<16>, 16);
  else launch(tc_out_conv_kernel<32>, 32);
  return 0;
}

}  // namespace kc
