// tc_gen.cu - general-shape tcgen05 / TMEM / TMA convolution engine: see tc_gen.h.
//
// Forward-type products (tc_gconv_kernel).  One persistent CTA per SM walks tiles of 4*MT rows x TW columns of the
// GEMM pixel grid.  A tile's K dimension is cut into slabs of chunk planes; per slab the producer warp issues one TMA
// box per plane (rows x 32 pixels x 16 B: the no-swizzle K-major operand with linear rows, tc_common.cuh) plus one bulk
// copy of the slab's B-operand block from the prepared weight image.  The MMA warp runs a host-built list: every entry
// is one tcgen05.mma M128 x N x K16 whose A descriptor starts at (plane, tap shift) inside the halo tile - taps are start
// addresses, SAME padding is TMA zero fill, stride-2 layers become stride-1 products through the space-to-depth element
// map.  Accumulators live in TMEM (double buffered when they fit); 8 epilogue warps apply bias / ReLU / sigmoid / the
// ReLU mask of a data gradient and store bf16 planes (PLAIN or S2D, hi + lo) and / or fp32 NHWC.
//
// Weight-gradient products (tc_gwgrad_kernel).  Both operands are read MN-major from the same kind of tiles with
// K = 16 pixels per MMA; the unshifted operand comes through a windowed tensor map whose innermost extent is the tile
// width, so TMA zero fill clears the tile's padding columns; accumulators persist in TMEM over all tiles of the CTA;
// a bias gradient is one more accumulator against a plane of ones.  Each CTA dumps its accumulators once; a gather
// kernel folds the CTAs in a fixed order (deterministic) straight into the Keras weight layout.
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>

#include "kernels.h"
#include "tc_common.cuh"
#include "tc_gen.h"

namespace kc {

using namespace tc;

namespace {

constexpr int GP = 32;                 // tile pitch: pixels per shared-memory row
constexpr int G1_MAXMMA = 768;
constexpr int G1_MAXSLAB = 32;
constexpr int G1_MAXPL = 32;           // planes per slab (hi + lo)
constexpr int G1_EPI_WARPS = 16;       // four epilogue warps per TMEM lane group
constexpr int G1_THREADS = 64 + 32 * G1_EPI_WARPS;   // TMA warp, MMA warp, epilogue warps
constexpr int G1_MAXSTAGES = 6;       // shared-memory stages of both engines: as many as fit (plan field n_stages), at least 2
constexpr size_t SMEM_BUDGET = 225 * 1024;

// one tcgen05.mma of the list, ready to issue: low descriptor words (start address and leading byte offset, 16-byte units,
// relative to the stage base), the instruction descriptor, and the accumulator column with the accumulate flag in bit 31.
// The high descriptor words are the same for every entry (stride byte offset 128, descriptor version 1).
struct G1Mma { uint32_t a_lo, b_lo, idesc, dcol_acc; };
constexpr uint32_t G1_DESC_HI = (128u >> 4) | (1u << 14);
struct G1Slab { int mma0, mma_n, nplanes, b_src, b_bytes, pad0, pad1, pad2; int plane[G1_MAXPL]; };
struct G1Group { int slab0, slab_n, a_par, pad; };
struct G1PlanDev {
  int n_groups, MT, NB, acc_cols, R_in, row0, col0, TW, in_PL, n_slabs, n_mma, type;
  uint32_t CHb, a_region, stage_bytes, Cop;
  int n_stages, pad_a, pad_b, pad_c;
  G1Group groups[2];
  G1Slab slabs[G1_MAXSLAB];
  G1Mma mma[G1_MAXMMA];
};
constexpr uint32_t G1_PLAN_BYTES = ((sizeof(G1PlanDev) + 1023) / 1024) * 1024;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* mbar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(mbar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* mbar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(mbar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(mbar)) : "memory");
}

// plane tensor addressing (16-byte units)
struct PlaneRef {
  uint4* base;
  int layout, KC, PLimg, H, W;     // PLimg = planes per image (hi + lo); H, W = plane dims
};
__device__ __forceinline__ int64_t unit_index(const PlaneRef& t, int n, int oy, int ox, int chunk, int lo) {
  if (t.layout == GEN_S2D) {
    const int par = ((oy & 1) << 1) | (ox & 1);
    const int plane = (lo * 4 + par) * t.KC + chunk;
    return (((int64_t)n * t.PLimg + plane) * t.H + (oy >> 1)) * t.W + (ox >> 1);
  }
  const int plane = lo * t.KC + chunk;
  return (((int64_t)n * t.PLimg + plane) * t.H + oy) * t.W + ox;
}

struct G1Params {
  const G1PlanDev* plan;
  const unsigned char* wimg;
  int B, Hg, Wg, tiles_y, tiles_x, num_tiles;
  int in_PL;         // planes per image of the K-side tensor (hi + lo)
  int plan_units;    // 16-byte units of the plan that are in use
  int pre;
  const float* bias;
  PlaneRef mask; int has_mask;
  const float* mask_f32;
  PlaneRef out; int has_out, out_lo;
  float* out_f32;
  int Cn;            // real N-side channels
  int dense_n, dense_ld, dense_cc, dense_tr;   // GEN_DENSE epilogue (see GenEpilogue)
  int Ho, Wo;        // logical output dims
  int* error_flag;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// epilogue variants (template flags): only the code a product needs is compiled into its kernel
constexpr int E_PRE = 1, E_MASK = 2, E_OUTP = 4, E_LO = 8, E_F32 = 16, E_MASKF = 32, E_ALL = 63, E_DENSE = 64;

// Steps of a plane tensor in 16-byte units, computed once per kernel: the epilogue then addresses "pixel base + b * db +
// chunk * cs (+ lo)" with 32-bit arithmetic instead of evaluating unit_index() per 8-channel group.
struct PlaneStep { int cs, db, lo; };     // next chunk plane; output column parity b = 1 (transposed convolution); lo planes
__device__ __forceinline__ PlaneStep plane_step(const PlaneRef& t) {
  const int hw = t.H * t.W;
  PlaneStep s;
  s.cs = hw;
  s.db = t.layout == GEN_S2D ? t.KC * hw : 1;
  s.lo = (t.layout == GEN_S2D ? 4 : 1) * t.KC * hw;
  return s;
}

// walks tiles t = blockIdx.x, + gridDim.x, ... as mixed-radix digits (group, tile column, tile row, image): one division
// chain at the start, additions with carry per tile
struct TileWalk {
  int grp, tx, ty, n, dgrp, dtx, dty, dn, n_groups, tiles_x, tiles_y;
  __device__ __forceinline__ void init(int t0, int step, int n_groups_, int tiles_x_, int tiles_y_) {
    n_groups = n_groups_; tiles_x = tiles_x_; tiles_y = tiles_y_;
    grp = t0 % n_groups; t0 /= n_groups; tx = t0 % tiles_x; t0 /= tiles_x; ty = t0 % tiles_y; n = t0 / tiles_y;
    dgrp = step % n_groups; step /= n_groups; dtx = step % tiles_x; step /= tiles_x; dty = step % tiles_y; dn = step / tiles_y;
  }
  __device__ __forceinline__ void next() {
    grp += dgrp; int c = grp >= n_groups; grp -= c ? n_groups : 0;
    tx += dtx + c; c = tx >= tiles_x; tx -= c ? tiles_x : 0;
    ty += dty + c; c = ty >= tiles_y; ty -= c ? tiles_y : 0;
    n += dn + c;
  }
};

// one 8-channel group of one output pixel: pre-op, mask, stores.  optr = the group's hi unit in the output planes,
// fpix = the pixel's index in the NHWC fp32 tensors
template <int EPI>
__device__ __forceinline__ void g1_emit8(const G1Params& p, const float* s_bias, float* v, int chunk, uint4 m, uint4* optr, int lo_off,
                                         int64_t fpix) {
  const int c0 = chunk * 8;
  if ((EPI & E_PRE) && p.pre != GEN_PRE_NONE) {
    const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c0), b1 = *reinterpret_cast<const float4*>(s_bias + c0 + 4);
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    // channels past Cn: zero weights (structural zeros of the gather table) and zero bias, so bias / ReLU leave 0 there
    if (p.pre == GEN_PRE_BIAS_RELU) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k] + bb[k], 0.f);
    } else if (p.pre == GEN_PRE_BIAS_SIGMOID) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = c0 + k < p.Cn ? __fdividef(1.0f, 1.0f + __expf(-(v[k] + bb[k]))) : 0.f;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += bb[k];
    }
  }
  if ((EPI & E_MASK) && p.has_mask) {
    // bf16 pairs of the activation: "> 0" <=> sign clear and not zero; as integers: low half (w << 16) > 0, high half w > 0xFFFF
    const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] = (int32_t)(w[e] << 16) > 0 ? v[2 * e] : 0.f;
      v[2 * e + 1] = (int32_t)w[e] > 0xFFFF ? v[2 * e + 1] : 0.f;
    }
  }
  if ((EPI & E_MASKF) && p.mask_f32) {
    const float* mk = p.mask_f32 + fpix * p.Cn + c0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (c0 + k < p.Cn && !(__ldg(mk + k) > 0.f)) v[k] = 0.f;
  }
  if ((EPI & E_OUTP) && p.has_out) {
    uint32_t h4[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h4[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
    *optr = make_uint4(h4[0], h4[1], h4[2], h4[3]);
    if ((EPI & E_LO) && p.out_lo) {
      uint32_t l4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float r0 = v[2 * e] - __uint_as_float(h4[e] << 16);
        const float r1 = v[2 * e + 1] - __uint_as_float(h4[e] & 0xFFFF0000u);
        l4[e] = pack_bf16x2(r0, r1);
      }
      optr[lo_off] = make_uint4(l4[0], l4[1], l4[2], l4[3]);
    }
  }
  if ((EPI & E_F32) && p.out_f32) {
    float* o = p.out_f32 + fpix * p.Cn + c0;
    if ((p.Cn & 3) == 0 && c0 + 8 <= p.Cn) {           // rows of Cn floats stay 16-byte aligned: two vector stores
      reinterpret_cast<float4*>(o)[0] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(o)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (c0 + k < p.Cn) o[k] = v[k];
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(G1_THREADS, 1)
tc_gconv_kernel(const __grid_constant__ CUtensorMap tmap, G1Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  G1PlanDev* plan = reinterpret_cast<G1PlanDev*>(smem);
  unsigned char* stages = smem + G1_PLAN_BYTES;
  __shared__ uint64_t full_bar[G1_MAXSTAGES], empty_bar[G1_MAXSTAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float s_bias[256 + 8];
  __shared__ uint32_t s_unit[8][2];       // epilogue unit u = (M-tile, 32-column block): {mt | cb << 8 | ncols << 16, 4 x (b << 7 | chunk)}
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 256 + 8; i += G1_THREADS) s_bias[i] = (!(EPI & E_DENSE) && p.bias && i < p.Cn) ? __ldg(p.bias + i) : 0.f;
  if (threadIdx.x < 8) {                  // tile-invariant index arithmetic of the epilogue, done once
    const int acc = __ldg(&p.plan->acc_cols), ncb = (acc + 31) / 32, u = threadIdx.x;
    const int cop = (int)__ldg(&p.plan->Cop), convt = __ldg(&p.plan->type) == GEN_CONVT_S2;
    const int mt = u / ncb, cb = u % ncb, ncols = min(32, acc - cb * 32);
    uint32_t bc = 0;
    for (int j8 = 0; j8 < 4; ++j8) {
      const int col0 = cb * 32 + j8 * 8;
      const int b = convt ? col0 / cop : 0;
      bc |= (uint32_t)((b << 7) | ((col0 - b * cop) >> 3)) << (8 * j8);
    }
    s_unit[u][0] = (uint32_t)(mt | (cb << 8) | (ncols << 16));
    s_unit[u][1] = bc;
  }

  {  // plan -> shared memory (header + the used slab / MMA entries)
    const uint4* src = reinterpret_cast<const uint4*>(p.plan);
    uint4* dst = reinterpret_cast<uint4*>(plan);
    const int nv = p.plan_units;                       // header + slabs + the used MMA entries, in 16-byte units
    constexpr int PER = (int)((sizeof(G1PlanDev) / 16 + G1_THREADS - 1) / G1_THREADS);
    uint4 tmp[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) { const int i = threadIdx.x + k * G1_THREADS; if (i < nv) tmp[k] = __ldg(src + i); }   // all loads in flight
#pragma unroll
    for (int k = 0; k < PER; ++k) { const int i = threadIdx.x + k * G1_THREADS; if (i < nv) dst[i] = tmp[k]; }
    // the pad units behind a slab's last plane are read by dummy K chunks (zero weights) and by discarded rows, and
    // 0 * NaN would poison an accumulator: they start as zeros
    // (only the 8 units right behind every plane position can be such a pad; planes landing there later overwrite them)
    const uint32_t sb = __ldg(&p.plan->stage_bytes), chb = __ldg(&p.plan->CHb), areg = __ldg(&p.plan->a_region);
    const int npos = (int)((areg - 128) / chb);               // plane positions of a stage
    const int nst = __ldg(&p.plan->n_stages);
    for (int i = threadIdx.x; i < nst * npos * 8; i += G1_THREADS) {
      const int s_ = i / (npos * 8), r_ = i % (npos * 8);
      reinterpret_cast<uint4*>(stages + (size_t)s_ * sb + (size_t)(r_ / 8 + 1) * chb)[r_ % 8] = make_uint4(0, 0, 0, 0);
    }
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < G1_MAXSTAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], G1_EPI_WARPS); }
    fence_mbar_init();
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int n_groups = plan->n_groups, MT = plan->MT, NB = plan->NB, acc_cols = plan->acc_cols;
  const int TW = plan->TW, TRr = 4 * MT;
  const uint32_t stage_bytes = plan->stage_bytes;
  const int n_stages = plan->n_stages;

  if (warp == 0) {
    // ================================ TMA producer ========================================
    if (lane == 0) {
      const uint32_t CHb = plan->CHb, a_region = plan->a_region;
      const int row0 = plan->row0, col0 = plan->col0, in_PL = p.in_PL;
      int s = 0;                       // ring position and its phase
      uint32_t ph = 0;
      bool ok = true;
      TileWalk tw;
      tw.init(blockIdx.x, gridDim.x, n_groups, p.tiles_x, p.tiles_y);
      for (int t = blockIdx.x; t < p.num_tiles && ok; t += gridDim.x, tw.next()) {
        const int n = tw.n, ty = tw.ty, tx = tw.tx;
        const G1Group g = plan->groups[tw.grp];
        for (int sl = g.slab0; sl < g.slab0 + g.slab_n; ++sl) {
          if (!mbar_wait(&empty_bar[s], ph ^ 1)) { *p.error_flag = 1; ok = false; break; }
          const G1Slab* S = &plan->slabs[sl];
          unsigned char* dst = stages + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_bar[s], (uint32_t)S->nplanes * CHb + (uint32_t)S->b_bytes);
          for (int pl = 0; pl < S->nplanes; ++pl)
            tma_load_3d(dst + (size_t)pl * CHb, &tmap, &full_bar[s], (tx * TW + col0) * 8, ty * TRr + row0, n * in_PL + S->plane[pl]);
          bulk_load(dst + a_region, p.wimg + S->b_src, (uint32_t)S->b_bytes, &full_bar[s]);
          if (++s == n_stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==========================================
    const bool leader = elect_one();
    int s = 0, tcount = 0;
    uint32_t ph = 0;
    bool ok = true;
    int grp = blockIdx.x % n_groups;
    const int dgrp = gridDim.x % n_groups;
    for (int t = blockIdx.x; t < p.num_tiles && ok; t += gridDim.x, ++tcount) {
      const G1Group g = plan->groups[grp];
      grp += dgrp; grp -= grp >= n_groups ? n_groups : 0;
      const int buf = NB == 2 ? (tcount & 1) : 0;
      const uint32_t bph = (uint32_t)(NB == 2 ? (tcount >> 1) : tcount) & 1u;
      if (!mbar_wait(&tempty_bar[buf], bph ^ 1)) { if (leader) *p.error_flag = 1; break; }
      fence_after_sync();
      for (int sl = g.slab0; sl < g.slab0 + g.slab_n; ++sl) {
        if (!mbar_wait(&full_bar[s], ph)) { if (leader) *p.error_flag = 1; ok = false; break; }
        fence_after_sync();
        const uint32_t base = smem_u32(stages + (size_t)s * stage_bytes);
        const int m0 = plan->slabs[sl].mma0, m1 = m0 + plan->slabs[sl].mma_n;
        // list entry outer, M-tile inner: one 16-byte shared-memory read and two adds per entry, one add per issued MMA
        const uint32_t base16 = base >> 4;
        const uint32_t dbase = tmem + (uint32_t)(buf * MT * acc_cols);
        const uint4* list = reinterpret_cast<const uint4*>(plan->mma);
#pragma unroll 4
        for (int i = m0; i < m1; ++i) {
          const uint4 m = list[i];
          const uint64_t hi = (uint64_t)G1_DESC_HI << 32;
          uint64_t da = hi | (uint64_t)(m.x + base16);
          const uint64_t db = hi | (uint64_t)(m.y + base16);
          uint32_t d = dbase + (m.w & 0x7FFFFFFFu);
          const uint32_t acc = m.w >> 31;
          for (int mt = 0; mt < MT; ++mt) {
            if (leader) mma_bf16_ss(d, da, db, m.z, acc);
            da += 128;                        // next M-tile: 128 pixels x 16 B
            d += (uint32_t)acc_cols;
          }
        }
        if (leader) mma_commit(&empty_bar[s]);
        __syncwarp();
        if (++s == n_stages) { s = 0; ph ^= 1; }
      }
      if (ok && leader) mma_commit(&tfull_bar[buf]);
      __syncwarp();
    }
  } else {
    // ================================ epilogue warps ======================================
    const int e = warp - 2;
    const int lg = warp & 3;                      // TMEM lane group this warp may access
    const int half = e >> 2;                      // the warps of a lane group share the (M-tile, column block) units round robin
    constexpr int WPG = G1_EPI_WARPS / 4;         // warps per lane group
    constexpr int UPW = 8 / WPG;                  // units per warp and tile (a tile has at most 8 units)
    const bool convt = plan->type == GEN_CONVT_S2;
    const int NCB = (acc_cols + 31) / 32;
    const int nunits = MT * NCB;
    const PlaneStep os = plane_step(p.out), ms = plane_step(p.mask);
    TileWalk tw;
    tw.init(blockIdx.x, gridDim.x, n_groups, p.tiles_x, p.tiles_y);
    int tcount = 0;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++tcount, tw.next()) {
      const int n = tw.n, ty = tw.ty, tx = tw.tx;
      const int a_par = plan->groups[tw.grp].a_par;
      const int buf = NB == 2 ? (tcount & 1) : 0;
      const uint32_t bph = (uint32_t)(NB == 2 ? (tcount >> 1) : tcount) & 1u;
      // A warp owns at most UPW (M-tile, 32-column block) units of a tile.  The ReLU-mask units of ALL of them are requested
      // before the accumulator wait: they do not depend on the MMAs, and their global-memory latency then overlaps the
      // tile's MMAs instead of being paid once per unit by a warp with nothing else to run.
      uint4 mk[UPW][4];
      if (EPI & E_MASK) {
#pragma unroll
        for (int k = 0; k < UPW; ++k) {
          const int u = half + WPG * k;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) mk[k][j8] = make_uint4(0, 0, 0, 0);
          if (u < nunits && p.has_mask) {
            const uint32_t u0 = s_unit[u][0], bc = s_unit[u][1];
            const int mt = u0 & 0xFF, ncols = u0 >> 16;
            const int gy = ty * TRr + mt * 4 + lg, gx = tx * TW + lane;
            if (lane < TW && gy < p.Hg && gx < p.Wg) {
              const uint4* mptr = p.mask.base + unit_index(p.mask, n, convt ? 2 * gy + a_par : gy, convt ? 2 * gx : gx, 0, 0);
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                if (j8 * 8 < ncols) {
                  const uint32_t q = (bc >> (8 * j8)) & 0xFFu;
                  mk[k][j8] = __ldg(mptr + (int)(q >> 7) * ms.db + (int)(q & 0x7Fu) * ms.cs);
                }
              }
            }
          }
        }
      }
      if (!mbar_wait(&tfull_bar[buf], bph)) { if (lane == 0) *p.error_flag = 1; break; }
      fence_after_sync();
      // (unrolled only where the prefetched mask registers need static indices: the other variants keep one copy of the
      // unit body, which then stays resident in the instruction cache)
#pragma unroll(((EPI & E_MASK) ? UPW : 1))
      for (int k = 0; k < UPW; ++k) {
        const int u = half + WPG * k;
        if (u >= nunits) break;
        const uint32_t u0 = s_unit[u][0], bc = s_unit[u][1];
        const int mt = u0 & 0xFF, cb = (u0 >> 8) & 0xFF, ncols = u0 >> 16;
        const int gy = ty * TRr + mt * 4 + lg, gx = tx * TW + lane;       // GEMM pixel of this thread: row mt * 4 + lg, column lane
        const bool valid = lane < TW && gy < p.Hg && gx < p.Wg;
        const uint32_t ta = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)((buf * MT + mt) * acc_cols + cb * 32);
        float v[32];
        if (ncols == 32) tmem_ld32(ta, v);
        else tmem_ld16(ta, v);
        if constexpr ((EPI & E_DENSE) != 0) {
          // Dense product: this thread holds GEMM row n (an output unit of the Dense layer / a column of its weight matrix)
          // for 32 columns (batch rows / latent columns).  Bias is per row; fp32 output is [column][n] (coalesced over the
          // warp's 32 consecutive n); the bf16 image of the Dense output is [column = frame][chunk][pixel][8].
          const int n_row = gy * GP + gx;
          if (valid && n_row < p.dense_n) {
            const float rb = p.bias ? __ldg(p.bias + n_row) : 0.f;
            const int pc = n_row / p.dense_cc, cc = n_row - pc * p.dense_cc;
            const int64_t hw = (int64_t)p.out.H * p.out.W;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = cb * 32 + j;
              if (j < ncols && col < p.Cn) {
                float y = v[j] + rb;
                if (p.pre == GEN_PRE_BIAS_RELU) y = fmaxf(y, 0.f);
                const int64_t oi = p.dense_tr ? (int64_t)n_row * p.dense_ld + col : (int64_t)col * p.dense_ld + n_row;
                if (p.mask_f32 && !(__ldg(p.mask_f32 + oi) > 0.f)) y = 0.f;
                if (p.out_f32) p.out_f32[oi] = y;
                if (p.has_out) {
                  __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(p.out.base);
                  const __nv_bfloat16 hi = __float2bfloat16(y);
                  const int64_t uu = ((int64_t)col * p.out.PLimg + (cc >> 3)) * hw + pc;
                  ob[uu * 8 + (cc & 7)] = hi;
                  if (p.out_lo) ob[(uu + (int64_t)p.out.KC * hw) * 8 + (cc & 7)] = __float2bfloat16(y - __bfloat162float(hi));
                }
              }
            }
          }
        } else
        if (valid) {
          const int oy = convt ? 2 * gy + a_par : gy, ox0 = convt ? 2 * gx : gx;
          uint4* optr = ((EPI & E_OUTP) && p.has_out) ? p.out.base + unit_index(p.out, n, oy, ox0, 0, 0) : nullptr;
          const int64_t fpix0 = (EPI & (E_F32 | E_MASKF)) ? ((int64_t)n * p.Ho + oy) * p.Wo + ox0 : 0;
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            if (j8 * 8 < ncols) {
              const uint32_t q = (bc >> (8 * j8)) & 0xFFu;
              const int b = (int)(q >> 7), chunk = (int)(q & 0x7Fu);
              g1_emit8<EPI>(p, s_bias, v + j8 * 8, chunk, (EPI & E_MASK) ? mk[k][j8] : make_uint4(0, 0, 0, 0),
                            optr + (b * os.db + chunk * os.cs), os.lo, fpix0 + b);
            }
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// fp32 weights -> bf16 image through a gather table: idx >= 0: hi part of w[idx]; idx | LO_FLAG: lo part; -1: zero
constexpr int32_t LO_FLAG = GEN_LO_FLAG;
__global__ void gen_gather_weights_kernel(const float* __restrict__ w, const int32_t* __restrict__ idx, int64_t n, __nv_bfloat16* img) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t e = idx[i];
    float v = 0.f;
    if (e >= 0) {
      const float x = __ldg(w + (e & (LO_FLAG - 1)));
      const __nv_bfloat16 hi = __float2bfloat16(x);
      v = (e & LO_FLAG) ? x - __bfloat162float(hi) : x;
    }
    img[i] = __float2bfloat16(v);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn gen_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// plane tensor [B*PL][H][W*8] bf16: box = `rows` rows of 32 pixels of one plane
CUresult make_planes_map(CUtensorMap* tmap, const void* base, int64_t nplanes, int H, int W, int rows) {
  EncodeTiledFn enc = gen_encode_fn();
  const cuuint64_t gdim[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)nplanes};
  const cuuint64_t gstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  const cuuint32_t box[3] = {GP * 8, (cuuint32_t)rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}
// the same tensor cut into windows of TW pixels: dims (TW*8, W/TW, H, planes); a 32-pixel box starting at element 0 of a
// window gets its last 32-TW pixels zero filled (they lie outside the innermost extent)
CUresult make_window_map(CUtensorMap* tmap, const void* base, int64_t nplanes, int H, int W, int TW, int rows) {
  EncodeTiledFn enc = gen_encode_fn();
  const cuuint64_t gdim[4] = {(cuuint64_t)TW * 8, (cuuint64_t)(W / TW), (cuuint64_t)H, (cuuint64_t)nplanes};
  const cuuint64_t gstr[3] = {(cuuint64_t)TW * 16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  const cuuint32_t box[4] = {GP * 8, 1, (cuuint32_t)rows, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// (parity, channel) of K element (plane, slot) of a plane tensor; channel = -1: unused slot
void elem_of(int layout, int KC, int C, int plane, int slot, int* par, int* ch) {
  if (layout == GEN_X3) {
    const int e = plane * 8 + slot;
    *par = e / 3; *ch = e < 12 ? e % 3 : -1;
    if (e >= 12) *par = 0;
  } else if (layout == GEN_S2D) {
    *par = plane / KC;
    const int c = (plane % KC) * 8 + slot;
    *ch = c < C ? c : -1;
  } else {
    *par = -1;
    const int c = plane * 8 + slot;
    *ch = c < C ? c : -1;
  }
}
int hi_planes(int layout, int KC) { return layout == GEN_S2D ? 4 * KC : (layout == GEN_X3 ? 2 : (layout == GEN_X27 ? 4 : KC)); }

}  // namespace

// ============================================================================================
//                                  forward-type plan
// ============================================================================================
struct GenConvPlan {
  GenConvSpec spec;
  G1PlanDev host;
  G1PlanDev* dev = nullptr;
  std::vector<int32_t> table;     // weight gather table, one entry per bf16 element of the image
  int32_t* table_dev = nullptr;
  size_t smem = 0;
};

GenConvPlan* gen_conv_plan_create(const GenConvSpec& s, const char** why_not, bool upload) {
  static const char* msg = "";
  auto no = [&](const char* m) -> GenConvPlan* { msg = m; if (why_not) *why_not = msg; return nullptr; };
  if (s.Ck <= 0 || s.Cn <= 0 || s.Hg <= 0 || s.Wg <= 0) return no("empty shape");
  const int Cop = (s.Cn + 15) / 16 * 16;
  const int n_groups = s.kind == GEN_CONVT_S2 ? 2 : 1;
  const int acc_cols = s.kind == GEN_CONVT_S2 ? 2 * Cop : Cop;
  if (s.kind == GEN_DENSE && s.Wg != GP) return no("a Dense product is laid out 32 pixels wide");
  if (acc_cols > 256) return no("more than 256 accumulator columns per M-tile");
  const int PLh = hi_planes(s.in_layout, s.KCk);
  if (s.kind == GEN_CONV_S2 && s.in_layout == GEN_PLAIN) return no("stride-2 product needs an S2D / X3 input");
  if (s.kind != GEN_CONV_S2 && s.in_layout != GEN_PLAIN) return no("product needs a PLAIN input");

  GenConvPlan* P = new GenConvPlan();
  P->spec = s;
  G1PlanDev& D = P->host;
  memset(&D, 0, sizeof(D));
  D.n_groups = n_groups; D.type = s.kind; D.Cop = (uint32_t)Cop; D.acc_cols = acc_cols;
  D.MT = (8 * acc_cols <= 512) ? 4 : ((4 * acc_cols <= 512) ? 2 : 1);    // rows per tile = 4 * MT; two TMEM buffers
  D.NB = 2;
  int halo_r, halo_c;
  if (s.kind == GEN_CONV_S2) { D.row0 = 0; D.col0 = 0; halo_r = 1; halo_c = 1; }
  else if (s.kind == GEN_CONVT_S2) { D.row0 = -1; D.col0 = -1; halo_r = 1; halo_c = 1; }
  else if (s.kind == GEN_DENSE) { D.row0 = 0; D.col0 = 0; halo_r = 0; halo_c = 0; }
  else { D.row0 = -1; D.col0 = -1; halo_r = 2; halo_c = 2; }
  D.TW = GP - halo_c;
  D.R_in = 4 * D.MT + halo_r;
  D.CHb = (uint32_t)D.R_in * GP * 16;
  D.in_PL = PLh * (s.split ? 2 : 1);

  struct Tap { int t0, t1, shift; };     // (di,dj) or (kh,kw)
  std::vector<Tap> taps;
  if (s.kind == GEN_CONV_S2) for (int di = 0; di < 2; ++di) for (int dj = 0; dj < 2; ++dj) taps.push_back({di, dj, di * GP + dj});
  else if (s.kind == GEN_CONVT_S2) for (int di = 0; di < 2; ++di) for (int dj = 0; dj < 2; ++dj) taps.push_back({di, dj, (1 - di) * GP + (1 - dj)});
  else if (s.kind == GEN_DENSE) taps.push_back({0, 0, 0});
  else for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw)
    taps.push_back({kh, kw, s.flip ? (2 - kh) * GP + (2 - kw) : kh * GP + kw});

  // weight source of (group a, tap, K element (plane, slot), column n), -1 = structural zero
  auto wsrc = [&](int a, const Tap& tp, int plane, int slot, int n) -> int32_t {
    int par, ch;
    elem_of(s.in_layout, s.KCk, s.Ck, plane, slot, &par, &ch);
    if (ch < 0) return -1;
    if (s.kind == GEN_DENSE) {
      if (n >= s.Cn) return -1;
      if (s.ones_col1 && n == s.ones_col1 - 1) return s.ones_src;
      return s.w_mode == 0 ? ch * (s.w_stride ? s.w_stride : s.Cn) + n + s.w_col0 : (n + s.w_col0) * (s.w_stride ? s.w_stride : s.Ck) + ch;
    }
    int kh, kw, cn;
    if (s.kind == GEN_CONV_S2) {
      kh = 2 * tp.t0 + (par >> 1); kw = 2 * tp.t1 + (par & 1); cn = n;
    } else if (s.kind == GEN_CONVT_S2) {
      const int b = n / Cop;
      kh = a + 2 * tp.t0; kw = b + 2 * tp.t1; cn = n - b * Cop;
    } else {
      kh = tp.t0; kw = tp.t1; cn = n;
    }
    if (kh > 2 || kw > 2 || cn >= s.Cn) return -1;
    const int tap = kh * 3 + kw;
    return s.w_mode == 0 ? (tap * s.Ck + ch) * s.Cn + cn : (tap * s.Cn + cn) * s.Ck + ch;
  };
  // is K chunk (plane) of this tap / group structurally non-zero, and how many columns does the tap reach
  auto tap_n = [&](int a, const Tap& tp) -> int {
    if (s.kind != GEN_CONVT_S2) return Cop;
    if (a + 2 * tp.t0 > 2) return 0;
    return tp.t1 == 1 ? Cop : 2 * Cop;
  };
  auto chunk_live = [&](int a, const Tap& tp, int plane, int N) -> bool {
    for (int slot = 0; slot < 8; ++slot)
      for (int n = 0; n < N; n += 1)
        if (wsrc(a, tp, plane, slot, n) >= 0) return true;
    return false;
  };

  // planes per slab: the largest even count whose stage (A planes + B block) fits twice in shared memory; room that is
  // left becomes further stages (n_stages).  Measured: cutting slabs smaller to get a deeper ring loses more to the
  // per-slab fixed costs (TMA boxes of 2-4 KB, commits) than the depth wins - conv1's weight gradient 0.112 -> 0.154 ms
  // with 4-row tiles and three stages - while extra stages of the SAME size are free: conv0's 0.167 -> 0.139 ms.
  auto slab_b_bytes = [&](int a, int p0, int p1) -> size_t {     // hi image bytes of planes [p0,p1)
    size_t bytes = 0;
    for (const Tap& tp : taps) {
      const int N = tap_n(a, tp);
      if (!N) continue;
      int live = 0;
      for (int pl = p0; pl < p1; ++pl) live += chunk_live(a, tp, pl, N) ? 1 : 0;
      bytes += (size_t)((live + 1) / 2) * N * 32;
    }
    return bytes;
  };
  const int mult = s.split ? 2 : 1;
  auto planes_per_slab = [&](size_t stage_cap) -> int {          // 0: not even one plane fits
    int ps = PLh;
    for (;; ) {
      size_t worst = 0;
      for (int a = 0; a < n_groups; ++a)
        for (int p0 = 0; p0 < PLh; p0 += ps) worst = std::max(worst, slab_b_bytes(a, p0, std::min(PLh, p0 + ps)) * mult);
      const size_t need = (size_t)ps * mult * D.CHb + 128 + worst;
      if (need <= stage_cap && ps * mult <= G1_MAXPL) return ps;
      if (ps <= 1) return 0;
      ps = (ps + 1) / 2;
    }
  };
  const int PS = planes_per_slab((SMEM_BUDGET - G1_PLAN_BYTES) / 2);
  if (!PS) { delete P; return no("one K chunk plane does not fit a shared-memory stage"); }
  D.a_region = (uint32_t)(PS * mult) * D.CHb + 128;

  int n_slabs = 0, n_mma = 0;
  size_t img_bytes = 0, worst_b = 0;
  std::vector<int32_t>& T = P->table;
  for (int a = 0; a < n_groups; ++a) {
    D.groups[a].slab0 = n_slabs; D.groups[a].a_par = a;
    bool first = true;
    for (int p0 = 0; p0 < PLh; p0 += PS) {
      const int p1 = std::min(PLh, p0 + PS);
      if (n_slabs >= G1_MAXSLAB) { delete P; return no("too many K slabs"); }
      G1Slab& S = D.slabs[n_slabs];
      S.mma0 = n_mma; S.nplanes = 0;
      for (int pl = p0; pl < p1; ++pl) S.plane[S.nplanes++] = pl;
      if (s.split) for (int pl = p0; pl < p1; ++pl) S.plane[S.nplanes++] = PLh + pl;
      const int np_hi = p1 - p0;
      const size_t hi_bytes = slab_b_bytes(a, p0, p1);
      S.b_src = (int)img_bytes; S.b_bytes = (int)(hi_bytes * mult);
      size_t boff = 0;       // offset inside the slab's hi block
      for (const Tap& tp : taps) {
        const int N = tap_n(a, tp);
        if (!N) continue;
        std::vector<int> live;
        for (int pl = p0; pl < p1; ++pl) if (chunk_live(a, tp, pl, N)) live.push_back(pl);
        for (size_t i = 0; i < live.size(); i += 2) {
          const int pa = live[i], pb = i + 1 < live.size() ? live[i + 1] : -1;
          // B block of this MMA: [2 chunks][N rows][8]
          const size_t e0 = (img_bytes + boff) / 2;
          if (T.size() < (img_bytes + hi_bytes * mult) / 2) T.resize((img_bytes + hi_bytes * mult) / 2, -1);
          for (int c = 0; c < 2; ++c) {
            const int pl = c == 0 ? pa : pb;
            for (int n = 0; n < N; ++n)
              for (int slot = 0; slot < 8; ++slot) {
                const int32_t src = pl >= 0 ? wsrc(a, tp, pl, slot, n) : -1;
                T[e0 + ((size_t)c * N + n) * 8 + slot] = src;
                if (s.split) T[e0 + hi_bytes / 2 + ((size_t)c * N + n) * 8 + slot] = src >= 0 ? (src | LO_FLAG) : -1;
              }
          }
          const uint32_t a_off = (uint32_t)(pa - p0) * D.CHb + (uint32_t)tp.shift * 16;
          const uint32_t a_lbo = pb >= 0 ? (uint32_t)(pb - pa) * D.CHb : 16u;      // dummy second chunk: the next pixel, zero weights
          const uint32_t b_hi = D.a_region + (uint32_t)boff, b_lo = b_hi + (uint32_t)hi_bytes;
          const uint32_t lo_planes = (uint32_t)np_hi * D.CHb;                       // lo planes sit behind the slab's hi planes
          const int variants = s.split ? 3 : 1;
          for (int vnt = 0; vnt < variants; ++vnt) {
            if (n_mma >= G1_MAXMMA) { delete P; return no("MMA list too long"); }
            G1Mma& M = D.mma[n_mma++];
            const uint32_t ao = a_off + (vnt == 1 ? lo_planes : 0), bo = vnt == 2 ? b_lo : b_hi;
            M.a_lo = (ao >> 4) | ((a_lbo >> 4) << 16);
            M.b_lo = (bo >> 4) | (((uint32_t)N * 16u >> 4) << 16);
            M.idesc = make_idesc_bf16_f32(128, N);
            M.dcol_acc = 0u | (first ? 0u : 0x80000000u);
            first = false;
          }
          boff += (size_t)N * 32;
        }
      }
      S.mma_n = n_mma - S.mma0;
      img_bytes += hi_bytes * mult;
      worst_b = std::max(worst_b, hi_bytes * mult);
      ++n_slabs;
    }
    D.groups[a].slab_n = n_slabs - D.groups[a].slab0;
  }
  T.resize(img_bytes / 2, -1);
  D.n_slabs = n_slabs; D.n_mma = n_mma;
  D.stage_bytes = (uint32_t)(((size_t)D.a_region + worst_b + 1023) / 1024 * 1024);
  D.n_stages = (int)std::min<size_t>(G1_MAXSTAGES, (SMEM_BUDGET - G1_PLAN_BYTES) / D.stage_bytes);
  if (D.n_stages < 2) { delete P; return no("shared-memory plan too large"); }
  P->smem = G1_PLAN_BYTES + (size_t)D.n_stages * D.stage_bytes;
  if (P->smem > 227 * 1024) { delete P; return no("shared-memory plan too large"); }
  if (!upload) return P;
  if (cudaMalloc(reinterpret_cast<void**>(&P->dev), sizeof(G1PlanDev)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&P->table_dev), std::max<size_t>(T.size(), 1) * sizeof(int32_t)) != cudaSuccess) {
    gen_conv_plan_free(P);
    return no("cudaMalloc failed");
  }
  cudaMemcpy(P->dev, &D, sizeof(G1PlanDev), cudaMemcpyHostToDevice);
  cudaMemcpy(P->table_dev, T.data(), T.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  return P;
}

int64_t gen_conv_plan_dump(const GenConvPlan* p, int32_t* out, int64_t capacity) {
  const G1PlanDev& D = p->host;
  std::vector<int32_t> v = {1, D.n_groups, D.MT, D.NB, D.acc_cols, D.R_in, D.row0, D.col0, D.TW, D.in_PL, D.n_slabs, D.n_mma, D.type,
                            (int32_t)D.CHb, (int32_t)D.a_region, (int32_t)D.stage_bytes, (int32_t)D.Cop};
  for (int g = 0; g < 2; ++g) { v.push_back(D.groups[g].slab0); v.push_back(D.groups[g].slab_n); v.push_back(D.groups[g].a_par); }
  for (int i = 0; i < D.n_slabs; ++i) {
    const G1Slab& S = D.slabs[i];
    v.insert(v.end(), {S.mma0, S.mma_n, S.nplanes, S.b_src, S.b_bytes});
    for (int k = 0; k < G1_MAXPL; ++k) v.push_back(S.plane[k]);
  }
  for (int i = 0; i < D.n_mma; ++i) {
    const G1Mma& M = D.mma[i];
    v.insert(v.end(), {(int32_t)M.a_lo, (int32_t)M.b_lo, (int32_t)M.idesc, (int32_t)M.dcol_acc});
  }
  v.push_back((int32_t)p->table.size());
  v.insert(v.end(), p->table.begin(), p->table.end());
  if (out && capacity >= (int64_t)v.size()) memcpy(out, v.data(), v.size() * sizeof(int32_t));
  return (int64_t)v.size();
}

void gen_conv_plan_free(GenConvPlan* p) {
  if (!p) return;
  if (p->dev) cudaFree(p->dev);
  if (p->table_dev) cudaFree(p->table_dev);
  delete p;
}
size_t gen_conv_weight_image_bytes(const GenConvPlan* p) { return p->table.size() * 2; }
int gen_conv_Cop(const GenConvPlan* p) { return (int)p->host.Cop; }

const int32_t* gen_conv_table(const GenConvPlan* p, size_t* n) { *n = p->table.size(); return p->table.data(); }
void gen_gather_weights(const float* w, const int32_t* table_dev, int64_t n, void* img, cudaStream_t st) {
  ProfScope prof_("gen_prep_weights", st);
  ++g_launches;
  gen_gather_weights_kernel<<<grid_for(n, 256, 8, 2), 256, 0, st>>>(w, table_dev, n, reinterpret_cast<__nv_bfloat16*>(img));
}
void gen_conv_prep_weights(const GenConvPlan* p, const float* w, void* img, cudaStream_t st) {
  ProfScope prof_("gen_prep_weights", st);
  ++g_launches;
  const int64_t n = (int64_t)p->table.size();
  gen_gather_weights_kernel<<<grid_for(n, 256, 8, 1), 256, 0, st>>>(w, p->table_dev, n, reinterpret_cast<__nv_bfloat16*>(img));
}

static PlaneRef plane_ref(const GenPlanes& t) {
  PlaneRef r;
  r.base = reinterpret_cast<uint4*>(t.base);
  r.layout = t.layout; r.KC = t.KC; r.PLimg = t.planes(); r.H = t.H; r.W = t.W;
  return r;
}

int gen_conv_run(const GenConvPlan* P, const GenPlanes& in, const void* wimg, const GenEpilogue& e, int B, int* error_flag,
                 const char* name, cudaStream_t st) {
  if (!gen_encode_fn() || !P->dev) return 1;
  const GenConvSpec& s = P->spec;
  const G1PlanDev& D = P->host;
  // a non-split plan may read the hi planes of a split tensor; a split plan needs the lo planes right behind the hi planes
  if (s.split ? in.planes() != D.in_PL : in.planes() < D.in_PL) return 3;
  CUtensorMap tmap;
  if (make_planes_map(&tmap, in.base, (int64_t)B * in.planes(), in.H, in.W, D.R_in) != CUDA_SUCCESS) return 2;
  G1Params p{};
  p.plan = P->dev;
  p.wimg = reinterpret_cast<const unsigned char*>(wimg);
  p.B = B; p.Hg = s.Hg; p.Wg = s.Wg; p.in_PL = in.planes();
  p.plan_units = (int)((offsetof(G1PlanDev, mma) + (size_t)D.n_mma * sizeof(G1Mma) + 15) / 16);
  p.tiles_y = cdiv(s.Hg, 4 * D.MT); p.tiles_x = cdiv(s.Wg, D.TW);
  p.num_tiles = B * p.tiles_y * p.tiles_x * D.n_groups;
  p.pre = e.pre; p.bias = e.bias;
  p.has_mask = e.mask != nullptr;
  if (e.mask) p.mask = plane_ref(*e.mask);
  p.mask_f32 = e.mask_f32;
  p.has_out = e.out != nullptr;
  if (e.out) { p.out = plane_ref(*e.out); p.out_lo = e.out->split; }
  p.out_f32 = e.out_f32;
  p.Cn = s.Cn;
  p.dense_n = e.dense_n; p.dense_ld = e.dense_ld; p.dense_cc = e.dense_cc > 0 ? e.dense_cc : 1; p.dense_tr = e.dense_tr;
  p.Ho = s.kind == GEN_CONVT_S2 ? 2 * s.Hg : s.Hg;
  p.Wo = s.kind == GEN_CONVT_S2 ? 2 * s.Wg : s.Wg;
  p.error_flag = error_flag;
  const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
  ProfScope prof_(name, st);
  ++g_launches;
  int flags = 0;
  if (e.pre != GEN_PRE_NONE) flags |= E_PRE;
  if (e.mask) flags |= E_MASK;
  if (e.mask_f32) flags |= E_MASKF;
  if (e.out) flags |= E_OUTP | (e.out->split ? E_LO : 0);
  if (e.out_f32) flags |= E_F32;
#define KC_G1_LAUNCH(F)                                                                                      \
  do {                                                                                                       \
    cudaFuncSetAttribute(tc_gconv_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem);   \
    tc_gconv_kernel<F><<<grid, G1_THREADS, P->smem, st>>>(tmap, p);                                          \
  } while (0)
  if (s.kind == GEN_DENSE) { KC_G1_LAUNCH(E_DENSE); return 0; }
  switch (flags) {     // the products the model issues get their own lean kernels; anything else runs the catch-all
    case E_PRE | E_OUTP | E_LO: KC_G1_LAUNCH(E_PRE | E_OUTP | E_LO); break;
    case E_PRE | E_OUTP: KC_G1_LAUNCH(E_PRE | E_OUTP); break;
    case E_PRE | E_F32: KC_G1_LAUNCH(E_PRE | E_F32); break;
    case E_MASK | E_OUTP: KC_G1_LAUNCH(E_MASK | E_OUTP); break;
    case E_MASK | E_F32: KC_G1_LAUNCH(E_MASK | E_F32); break;
    default: KC_G1_LAUNCH(E_ALL); break;
  }
#undef KC_G1_LAUNCH
  return 0;
}

// ============================================================================================
//                                  weight-gradient kernel
// ============================================================================================
namespace {

constexpr int G2_THREADS = 192;        // TMA warp, MMA warp, 4 dump warps
constexpr int G2_MAXMMA = 96;
constexpr int G2_MAXROLE = 4;
constexpr int G2_MAXPL = 64;
constexpr int G2_COLS = 512;

// low / high descriptor words relative to the stage base (B relative to the plane of ones when b_ones), instruction
// descriptor, accumulator column
struct G2Mma { uint32_t a_lo, a_hi, b_lo, b_hi, idesc, d_col, b_ones, pad; };
struct G2Role { int mma0, mma_n, nS, nU, ncols, pad0, pad1, pad2; int s_plane[G2_MAXPL]; int u_plane[G2_MAXPL]; };
struct G2PlanDev {
  int n_roles, TRr, R_s, row0, col0, TW, s_PL, u_PL;
  uint32_t CHs, CHu, s_region, u_region, stage_bytes, ones_off, n_stages, pad1;
  G2Role roles[G2_MAXROLE];
  G2Mma mma[G2_MAXMMA];
};
constexpr uint32_t G2_PLAN_BYTES = ((sizeof(G2PlanDev) + 1023) / 1024) * 1024;

struct G2Params {
  const G2PlanDev* plan;
  float* partial;              // [grid][G2_COLS][128]
  int B, Hg, Wg, tiles_y, tiles_x, num_tiles;
  int s_PL, u_PL;              // planes per image of the two tensors (hi + lo; only hi planes are read)
  int* error_flag;
};

__global__ void __launch_bounds__(G2_THREADS, 1)
tc_gwgrad_kernel(const __grid_constant__ CUtensorMap tmap_s, const __grid_constant__ CUtensorMap tmap_u, G2Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  G2PlanDev* plan = reinterpret_cast<G2PlanDev*>(smem);
  unsigned char* stages = smem + G2_PLAN_BYTES;
  __shared__ uint64_t full_bar[G1_MAXSTAGES], empty_bar[G1_MAXSTAGES], done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.plan);
    uint32_t* dst = reinterpret_cast<uint32_t*>(plan);
    for (int i = threadIdx.x; i < (int)(sizeof(G2PlanDev) / 4); i += G2_THREADS) dst[i] = __ldg(src + i);
    // the pad behind a role's last S plane is read as K elements against zero-filled gradient pixels, and 0 * NaN would
    // poison an accumulator: those units start as zeros
    // (K elements = pixels: every pixel of a loaded plane is written by TMA; only the 8 units behind each S plane position can
    // be read without having been written.  Plane positions a role never loads feed discarded rows / columns only.)
    const uint32_t sb = __ldg(&p.plan->stage_bytes), chs = __ldg(&p.plan->CHs), sreg = __ldg(&p.plan->s_region);
    const int npos = (int)((sreg - 128) / chs);
    const int nst = (int)__ldg(&p.plan->n_stages);
    for (int i = threadIdx.x; i < nst * npos * 8; i += G2_THREADS) {
      const int s_ = i / (npos * 8), r_ = i % (npos * 8);
      reinterpret_cast<uint4*>(stages + (size_t)s_ * sb + (size_t)(r_ / 8 + 1) * chs)[r_ % 8] = make_uint4(0, 0, 0, 0);
    }
  }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  if (threadIdx.x == 32) {
    for (int s = 0; s < G1_MAXSTAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  {  // two planes of ones (bias gradient operand) behind the stages
    uint4* ones = reinterpret_cast<uint4*>(smem + plan->ones_off);
    const int nu = (int)(2 * plan->CHu / 16);
    for (int i = threadIdx.x; i < nu; i += G2_THREADS) ones[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int n_roles = plan->n_roles;
  const int role = blockIdx.x % n_roles;
  const int rank_in_role = blockIdx.x / n_roles;
  const int ctas_in_role = (gridDim.x - role + n_roles - 1) / n_roles;
  const int TRr = plan->TRr, TW = plan->TW;
  const uint32_t stage_bytes = plan->stage_bytes;
  const int n_stages = plan->n_stages;
  const int per_img = p.tiles_y * p.tiles_x;
  const int my_tiles = p.num_tiles > rank_in_role ? (p.num_tiles - 1 - rank_in_role) / ctas_in_role + 1 : 0;
  const G2Role* R = &plan->roles[role];

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t CHs = plan->CHs, CHu = plan->CHu, s_region = plan->s_region;
      int s = 0;
      uint32_t ph = 0;
      for (int t = rank_in_role; t < p.num_tiles; t += ctas_in_role) {
        if (!mbar_wait(&empty_bar[s], ph ^ 1)) { *p.error_flag = 1; break; }
        const int n = t / per_img, rem = t % per_img;
        const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
        unsigned char* dst = stages + (size_t)s * stage_bytes;
        mbar_expect_tx(&full_bar[s], (uint32_t)R->nS * CHs + (uint32_t)R->nU * CHu);
        for (int pl = 0; pl < R->nS; ++pl)
          tma_load_3d(dst + (size_t)pl * CHs, &tmap_s, &full_bar[s], (tx * TW + plan->col0) * 8, ty * TRr + plan->row0,
                      n * p.s_PL + R->s_plane[pl]);
        for (int pl = 0; pl < R->nU; ++pl)
          tma_load_4d(dst + s_region + (size_t)pl * CHu, &tmap_u, &full_bar[s], 0, tx, ty * TRr, n * p.u_PL + R->u_plane[pl]);
        if (++s == n_stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t ones_addr = smem_u32(smem + plan->ones_off);
    const int KS = 2 * TRr;                       // 16-pixel K steps per tile
    int it = 0, s = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (int t = rank_in_role; t < p.num_tiles; t += ctas_in_role, ++it) {
      if (!mbar_wait(&full_bar[s], ph)) { if (leader) *p.error_flag = 1; ok = false; break; }
      fence_after_sync();
      const uint32_t base16 = smem_u32(stages + (size_t)s * stage_bytes) >> 4;
      // list entry outer, K step inner (every entry owns its accumulator): descriptors are built once per entry and tile,
      // each issued MMA costs two adds
      for (int i = R->mma0; i < R->mma0 + R->mma_n; ++i) {
        const uint4 w0 = reinterpret_cast<const uint4*>(plan->mma)[2 * i];
        const uint4 w1 = reinterpret_cast<const uint4*>(plan->mma)[2 * i + 1];
        uint64_t da = ((uint64_t)w0.y << 32) | (uint64_t)(w0.x + base16);
        uint64_t db = ((uint64_t)w0.w << 32) | (uint64_t)(w0.z + (w1.z ? (ones_addr >> 4) : base16));
        const uint32_t d = tmem + w1.y;
        uint32_t acc = it != 0;
#pragma unroll 4
        for (int ks = 0; ks < KS; ++ks) {
          if (leader) mma_bf16_ss(d, da, db, w1.x, acc);
          da += 16; db += 16;                // 16 pixels x 16 B
          acc = 1;
        }
      }
      if (leader) mma_commit(&empty_bar[s]);
      __syncwarp();
      if (++s == n_stages) { s = 0; ph ^= 1; }
    }
    if (ok && leader) mma_commit(&done_bar);
    __syncwarp();
  } else if (my_tiles > 0) {
    // ================================ dump: TMEM -> this CTA's partial block =====================================
    const int lg = warp & 3;
    bool landed = false;
    for (uint32_t spin = 0; spin < (1u << 22) && !landed; ++spin) {     // the whole kernel lies before this barrier: back off
      landed = mbar_try_wait(&done_bar, 0) != 0;
      if (!landed) __nanosleep(256);
    }
    if (landed) {
      fence_after_sync();
      float* out = p.partial + (size_t)blockIdx.x * G2_COLS * 128 + (size_t)(lg * 32 + lane);
      for (int cb = 0; cb * 32 < R->ncols; ++cb) {
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(cb * 32), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) out[(size_t)(cb * 32 + j) * 128] = v[j];
      }
    } else if (lane == 0) {
      *p.error_flag = 1;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// out[e] = sum over the CTAs of the entry's role of partial[cta][src] for up to 4 sources per entry.  One warp per entry:
// lane l folds CTAs l, l + 32, ... in order, then a fixed shuffle tree (deterministic for a given grid).
__global__ void __launch_bounds__(256) gen_wgrad_reduce_kernel(const float* __restrict__ partial, const int32_t* __restrict__ src, int E,
                                                               int n_roles, int grid, int tiles, float* dW, float* db, int EW) {
  const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (e >= E) return;
  float acc = 0.f;
  for (int k = 0; k < 4; ++k) {
    const int32_t sidx = __ldg(src + (size_t)e * 4 + k);
    if (sidx < 0) continue;
    const int role = sidx / (G2_COLS * 128), off = sidx % (G2_COLS * 128);
    const int ctas = (grid - role + n_roles - 1) / n_roles;
    const int live = tiles < ctas ? tiles : ctas;              // CTAs beyond the tile count never wrote their block
    for (int c = lane; c < live; c += 32) acc += __ldg(partial + (size_t)(c * n_roles + role) * G2_COLS * 128 + off);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    if (e < EW) dW[e] = acc;
    else if (db) db[e - EW] = acc;
  }
}

}  // namespace

struct GenWgradPlan {
  GenWgradSpec spec;
  G2PlanDev host;
  G2PlanDev* dev = nullptr;
  std::vector<int32_t> src;      // [EW + Cu][4]
  int32_t* src_dev = nullptr;
  int EW = 0;
  size_t smem = 0;
};

GenWgradPlan* gen_wgrad_plan_create(const GenWgradSpec& s, const char** why_not, bool upload) {
  static const char* msg = "";
  auto no = [&](const char* m) -> GenWgradPlan* { msg = m; if (why_not) *why_not = msg; return nullptr; };
  const int nS = hi_planes(s.s_layout, s.s_KC), nU = hi_planes(s.u_layout, s.u_KC);
  int halo_r, halo_c, row0, col0;
  if (s.kind == GEN_CONV_S2) { row0 = 0; col0 = 0; halo_r = 1; halo_c = 1; }
  else if (s.kind == GEN_CONVT_S2) { row0 = -1; col0 = -1; halo_r = 1; halo_c = 1; }
  else if (s.kind == GEN_DENSE) { row0 = 0; col0 = 0; halo_r = 0; halo_c = 0; }
  else { row0 = -1; col0 = -1; halo_r = 2; halo_c = 2; }
  int TW = 0;
  for (int d = GP - halo_c; d >= 8; --d) if (s.Wg % d == 0) { TW = d; break; }
  if (!TW) return no("no tile width in [8, 31] divides the gradient's width");
  struct Tap { int t0, t1, shift; };
  std::vector<Tap> taps;
  if (s.kind == GEN_CONV_S2 && s.s_layout == GEN_X27) taps.push_back({0, 0, 0});     // the patch planes hold all nine taps of a pixel
  else if (s.kind == GEN_CONV_S2) for (int di = 0; di < 2; ++di) for (int dj = 0; dj < 2; ++dj) taps.push_back({di, dj, di * GP + dj});
  else if (s.kind == GEN_CONVT_S2) for (int di = 0; di < 2; ++di) for (int dj = 0; dj < 2; ++dj) taps.push_back({di, dj, (1 - di) * GP + (1 - dj)});
  else if (s.kind == GEN_DENSE) taps.push_back({0, 0, 0});
  else for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw)
    taps.push_back({kh, kw, s.flip ? (2 - kh) * GP + (2 - kw) : kh * GP + kw});

  // An accumulator = one (tap, A plane run, B plane run) product: M = 8 * A planes (64 or 128, padded with whatever planes
  // follow), N = 8 * B planes.  Two ways to cover a layer:
  //   dense   A = 16- (or 8-) plane blocks of the operand with more planes, B = the whole other operand, every tap;
  //   pruned  (stride-2 layers) one accumulator per VALID (tap, parity) pair - 9 of the 16 combinations exist in a 3x3
  //           kernel - over one parity's planes of the space-to-depth operand and the whole plain operand.
  // The cheaper one by MMA cycles per K step (tools/mma_bench.cu: 40 / 49 / 66 / 130 cycles at N <= 32 / 64 / 128 / 256)
  // and by accumulator roles (each role streams its operands again) is taken.
  struct Run { bool is_s; int plane0, nplanes; };
  struct Acc { int tap; Run a, b; int M, N, role, d_col; bool bias; };
  auto cyc = [](int N) { return N <= 32 ? 40 : (N <= 64 ? 49 : (N <= 128 ? 66 : 130)); };
  auto score = [&](const std::vector<Acc>& v) {
    double c = 0; int cols = 0;
    for (const Acc& a : v) { c += cyc(a.N); cols += a.N; }
    const int roles = (cols + 32 + G2_COLS - 1) / G2_COLS;
    return c * (1.0 + 0.35 * (roles - 1));
  };
  auto dense_plan = [&](std::vector<Acc>& out) -> bool {
    const bool a_is_s = nS >= nU;
    const int nA = a_is_s ? nS : nU, nB = a_is_s ? nU : nS;
    const int Mblk = nA >= 16 ? 128 : 64;
    int Nb = nB * 8;
    if (Mblk == 128 && Nb % 16) Nb += 8;                  // reads one more plane: ignored columns
    if (Nb > 256) return false;
    const int blkA = Mblk / 8;
    for (int blk = 0; blk * blkA < nA; ++blk)
      for (size_t tp = 0; tp < taps.size(); ++tp)
        out.push_back({(int)tp, {a_is_s, blk * blkA, std::min(blkA, nA - blk * blkA)}, {!a_is_s, 0, nB}, Mblk, Nb, 0, 0, false});
    return true;
  };
  auto pruned_plan = [&](std::vector<Acc>& out) -> bool {
    if (s.kind == GEN_CONV_S1 || s.kind == GEN_DENSE) return false;
    const bool p_is_s = s.kind == GEN_CONV_S2;            // the space-to-depth operand: the layer input (Conv2D) or the gradient (ConvT)
    if ((p_is_s ? s.s_layout : s.u_layout) != GEN_S2D) return false;
    const int KCp = p_is_s ? s.s_KC : s.u_KC, nQ = p_is_s ? nU : nS;
    int best = -1; double best_cost = 1e30;
    for (int mode = 0; mode < 2; ++mode) {                // 0: A = parity block, B = plain operand; 1: A = plain operand, B = parity block
      int M, N;
      if (mode == 0) { if (KCp != 8 && KCp != 16) continue; M = KCp * 8; N = nQ * 8; }
      else { if (nQ > 16) continue; M = nQ <= 8 ? 64 : 128; N = KCp * 8; }
      if (M == 128 && N % 16) N += 8;
      if (N > 256 || N < 8) continue;
      const double cost = 9.0 * cyc(N) * (1.0 + 0.35 * ((9 * N + 32 + G2_COLS - 1) / G2_COLS - 1));
      if (cost < best_cost) { best_cost = cost; best = mode; }
    }
    if (best < 0) return false;
    for (size_t tp = 0; tp < taps.size(); ++tp)
      for (int par = 0; par < 4; ++par) {
        if (2 * taps[tp].t0 + (par >> 1) > 2 || 2 * taps[tp].t1 + (par & 1) > 2) continue;      // kh, kw of this (tap, parity)
        Run pr{p_is_s, par * KCp, KCp}, qr{!p_is_s, 0, nQ};
        int M, N;
        if (best == 0) { M = KCp * 8; N = nQ * 8; } else { M = nQ <= 8 ? 64 : 128; N = KCp * 8; }
        if (M == 128 && N % 16) N += 8;
        out.push_back({(int)tp, best == 0 ? pr : qr, best == 0 ? qr : pr, M, N, 0, 0, false});
      }
    return true;
  };
  std::vector<Acc> accs, cand;
  const bool have_dense = dense_plan(accs);
  bool pruned = false;
  if (pruned_plan(cand) && (!have_dense || score(cand) < score(accs))) { accs.swap(cand); pruned = true; }
  else if (!have_dense) return no("N side wider than 256");
  // bias accumulators: A = runs of the gradient's planes, B = two planes of ones.  With the pruned plan over a
  // space-to-depth gradient the runs are the parity blocks the product accumulators already load.
  std::vector<Acc> bias;
  if (s.kind == GEN_DENSE) {
    // a Dense data gradient: no bias accumulator
  } else if (pruned && s.kind == GEN_CONVT_S2 && (s.u_KC == 8 || s.u_KC == 16)) {
    for (int par = 0; par < 4; ++par)
      bias.push_back({-1, {false, par * s.u_KC, s.u_KC}, {false, 0, 0}, s.u_KC * 8, s.u_KC == 16 ? 16 : 8, 0, 0, true});
  } else {
    const int MblkU = nU >= 16 ? 128 : 64, blkU = MblkU / 8;
    for (int ub = 0; ub * blkU < nU; ++ub)
      bias.push_back({-1, {false, ub * blkU, std::min(blkU, nU - ub * blkU)}, {false, 0, 0}, MblkU, MblkU == 128 ? 16 : 8, 0, 0, true});
  }
  // roles: product accumulators in order, balanced by columns, at most 512 columns each; every bias accumulator joins a
  // role that already loads its planes (else the emptiest one)
  int total_cols = 0;
  for (const Acc& a : accs) total_cols += a.N;
  for (const Acc& a : bias) total_cols += a.N;
  const int n_roles = (total_cols + G2_COLS - 1) / G2_COLS;
  if (n_roles > G2_MAXROLE) return no("too many accumulator roles");
  {
    const int target = (total_cols + n_roles - 1) / n_roles;
    std::vector<int> cols(n_roles, 0);
    int r = 0;
    for (Acc& a : accs) {
      if (cols[r] + a.N > G2_COLS || (cols[r] + a.N > target && cols[r] > 0 && r + 1 < n_roles)) ++r;
      if (r >= n_roles) return no("accumulators do not pack into the planned roles");
      a.role = r; a.d_col = cols[r]; cols[r] += a.N;
    }
    for (Acc& b : bias) {
      int pick = -1;
      for (int q = 0; q < n_roles && pick < 0; ++q) {
        if (cols[q] + b.N > G2_COLS) continue;
        for (const Acc& a : accs)
          if (a.role == q && ((!a.a.is_s && a.a.plane0 == b.a.plane0 && a.a.nplanes == b.a.nplanes) ||
                              (!a.b.is_s && a.b.plane0 == b.a.plane0 && a.b.nplanes == b.a.nplanes))) { pick = q; break; }
      }
      for (int q = 0; q < n_roles && pick < 0; ++q) if (cols[q] + b.N <= G2_COLS) pick = q;
      if (pick < 0) return no("bias accumulators do not fit");
      b.role = pick; b.d_col = cols[pick]; cols[pick] += b.N;
    }
    for (const Acc& b : bias) accs.push_back(b);
  }

  GenWgradPlan* P = new GenWgradPlan();
  P->spec = s;
  G2PlanDev& D = P->host;
  memset(&D, 0, sizeof(D));
  D.n_roles = n_roles;
  // plane positions per role: every run an accumulator of the role touches is loaded once, runs back to back
  struct Placed { bool is_s; int plane0, nplanes, pos; };
  std::vector<std::vector<Placed>> placed(n_roles);
  std::vector<int> s_cnt(n_roles, 0), u_cnt(n_roles, 0);
  int s_extent = 1, u_extent = 1;                          // planes a region must hold (M / N padding included)
  auto place = [&](int r, const Run& run, int span) -> int {
    if (run.nplanes == 0) return 0;
    for (const Placed& q : placed[r]) if (q.is_s == run.is_s && q.plane0 == run.plane0 && q.nplanes == run.nplanes) {
      int& ext = run.is_s ? s_extent : u_extent; ext = std::max(ext, q.pos + span); return q.pos;
    }
    int& cnt = run.is_s ? s_cnt[r] : u_cnt[r];
    placed[r].push_back({run.is_s, run.plane0, run.nplanes, cnt});
    const int pos = cnt; cnt += run.nplanes;
    int& ext = run.is_s ? s_extent : u_extent; ext = std::max(ext, std::max(cnt, pos + span));
    return pos;
  };
  struct Pos { int a, b; };
  std::vector<Pos> pos(accs.size());
  for (size_t i = 0; i < accs.size(); ++i) {
    pos[i].a = place(accs[i].role, accs[i].a, accs[i].M / 8);
    pos[i].b = accs[i].bias ? 0 : place(accs[i].role, accs[i].b, accs[i].N / 8);
  }
  for (int r = 0; r < n_roles; ++r) if (s_cnt[r] > G2_MAXPL || u_cnt[r] > G2_MAXPL) { delete P; return no("too many planes per role"); }
  // tile rows: the largest of 8, 4, 2 whose two stages fit; spare room becomes extra stages of the same size (with two
  // stages one is always being consumed, so a single tile is in flight per SM; see the note at the forward planner)
  auto stages_that_fit = [&](int trr, size_t* CHs_, size_t* CHu_, size_t* s_reg_, size_t* u_reg_, size_t* stage_) -> int {
    const size_t CHs = (size_t)(trr + halo_r) * GP * 16, CHu = (size_t)trr * GP * 16;
    const size_t s_reg = (size_t)s_extent * CHs + 128, u_reg = (size_t)u_extent * CHu;
    const size_t stage = (s_reg + u_reg + 1023) / 1024 * 1024;
    *CHs_ = CHs; *CHu_ = CHu; *s_reg_ = s_reg; *u_reg_ = u_reg; *stage_ = stage;
    if (G2_PLAN_BYTES + 2 * CHu >= SMEM_BUDGET) return 0;
    return (int)((SMEM_BUDGET - G2_PLAN_BYTES - 2 * CHu) / stage);
  };
  int TRr = 0, n_stages = 0;
  size_t CHs = 0, CHu = 0, s_reg = 0, u_reg = 0, stage = 0;
  for (int trr : {8, 4, 2}) if (!TRr && stages_that_fit(trr, &CHs, &CHu, &s_reg, &u_reg, &stage) >= 2) TRr = trr;
  if (!TRr) { delete P; return no("operand planes do not fit two shared-memory stages"); }
  n_stages = std::min(G1_MAXSTAGES, stages_that_fit(TRr, &CHs, &CHu, &s_reg, &u_reg, &stage));
  D.CHs = (uint32_t)CHs; D.CHu = (uint32_t)CHu; D.s_region = (uint32_t)s_reg; D.u_region = (uint32_t)u_reg;
  D.stage_bytes = (uint32_t)stage; D.n_stages = (uint32_t)n_stages;
  D.TRr = TRr; D.R_s = TRr + halo_r; D.row0 = row0; D.col0 = col0; D.TW = TW; D.s_PL = nS; D.u_PL = nU;
  D.ones_off = (uint32_t)(G2_PLAN_BYTES + (size_t)n_stages * D.stage_bytes);
  P->smem = (size_t)D.ones_off + 2 * D.CHu;
  if (P->smem > 227 * 1024) { delete P; return no("shared-memory plan too large"); }

  int n_mma = 0;
  for (int r = 0; r < n_roles; ++r) {
    G2Role& R = D.roles[r];
    R.mma0 = n_mma; R.nS = 0; R.nU = 0; R.ncols = 0;
    for (const Placed& q : placed[r])
      for (int k = 0; k < q.nplanes; ++k) {
        if (q.is_s) R.s_plane[q.pos + k] = q.plane0 + k; else R.u_plane[q.pos + k] = q.plane0 + k;
      }
    R.nS = s_cnt[r]; R.nU = u_cnt[r];
    for (size_t i = 0; i < accs.size(); ++i) {
      const Acc& a = accs[i];
      if (a.role != r) continue;
      if (n_mma >= G2_MAXMMA) { delete P; return no("MMA list too long"); }
      G2Mma& M = D.mma[n_mma++];
      auto off = [&](const Run& run, int p, bool shifted) -> uint32_t {
        const uint32_t o = run.is_s ? (uint32_t)p * D.CHs : D.s_region + (uint32_t)p * D.CHu;
        return o + ((shifted && run.is_s && a.tap >= 0) ? (uint32_t)taps[a.tap].shift * 16 : 0u);
      };
      // MN-major operands: leading byte offset = 128 (8 pixels), stride byte offset = plane stride
      const uint32_t a_off = off(a.a, pos[i].a, true), a_sbo = a.a.is_s ? D.CHs : D.CHu;
      M.a_lo = (a_off >> 4) | ((128u >> 4) << 16); M.a_hi = (a_sbo >> 4) | (1u << 14);
      if (a.bias) {
        M.b_lo = 0u | ((128u >> 4) << 16); M.b_hi = (D.CHu >> 4) | (1u << 14); M.b_ones = 1;
      } else {
        const uint32_t b_off = off(a.b, pos[i].b, true), b_sbo = a.b.is_s ? D.CHs : D.CHu;
        M.b_lo = (b_off >> 4) | ((128u >> 4) << 16); M.b_hi = (b_sbo >> 4) | (1u << 14); M.b_ones = 0;
      }
      M.idesc = make_idesc_bf16_f32(a.M, a.N, 1, 1);
      M.d_col = (uint32_t)a.d_col;
      R.ncols = std::max(R.ncols, a.d_col + a.N);
    }
    R.mma_n = n_mma - R.mma0;
  }

  // scatter table: dW element -> (role, column, TMEM lane) of the accumulator whose element ranges contain it
  auto lane_of = [](int M, int row) { return M == 128 ? row : (row / 16) * 32 + row % 16; };
  auto find = [&](int tap_i, int es, int eu, bool bias) -> int32_t {
    for (const Acc& a : accs) {
      if (a.bias != bias || (!bias && a.tap != tap_i)) continue;
      const int ea = a.a.is_s ? es : eu;
      if (ea < a.a.plane0 * 8 || ea >= a.a.plane0 * 8 + a.M) continue;
      int col = 0;
      if (!bias) {
        const int eb = a.b.is_s ? es : eu;
        if (eb < a.b.plane0 * 8 || eb >= a.b.plane0 * 8 + a.b.nplanes * 8) continue;
        col = eb - a.b.plane0 * 8;
      }
      if (ea - a.a.plane0 * 8 >= a.a.nplanes * 8) continue;       // rows of the M padding belong to other planes
      return a.role * (G2_COLS * 128) + (a.d_col + col) * 128 + lane_of(a.M, ea - a.a.plane0 * 8);
    }
    return -1;
  };
  const int Cs = s.Cs, Cu = s.Cu;
  P->EW = (s.kind == GEN_DENSE ? 1 : 9) * Cs * Cu;
  P->src.assign((size_t)(P->EW + Cu) * 4, -1);
  if (s.kind == GEN_DENSE) {       // out[(cs, cu)]: w_mode 0 -> cs * Cu + cu ; 1 -> cu * Cs + cs
    const int sh = s.split_dense ? s.s_KC / 2 * 8 : 0, uh = s.split_dense ? s.u_KC / 2 * 8 : 0;   // element offsets of the lo halves
    for (int cs = 0; cs < Cs; ++cs) for (int cu = 0; cu < Cu; ++cu) {
      int32_t* e = &P->src[(size_t)(s.w_mode == 0 ? cs * Cu + cu : cu * Cs + cs) * 4];
      e[0] = find(0, cs, cu, false);
      if (s.split_dense) { e[1] = find(0, cs + sh, cu, false); e[2] = find(0, cs, cu + uh, false); }
    }
  } else
  for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) for (int cs = 0; cs < Cs; ++cs) for (int cu = 0; cu < Cu; ++cu) {
    int tap_i, es, eu;
    if (s.kind == GEN_CONV_S2 && s.s_layout == GEN_X27) {
      tap_i = 0; es = (kh * 3 + kw) * 3 + cs; eu = cu;
    } else if (s.kind == GEN_CONV_S2) {
      const int di = kh >> 1, a = kh & 1, dj = kw >> 1, b = kw & 1, par = a * 2 + b;
      tap_i = di * 2 + dj;
      es = s.s_layout == GEN_X3 ? par * 3 + cs : par * s.s_KC * 8 + cs;
      eu = cu;
    } else if (s.kind == GEN_CONVT_S2) {
      const int di = kh >> 1, a = kh & 1, dj = kw >> 1, b = kw & 1, par = a * 2 + b;
      tap_i = di * 2 + dj;
      es = cs;
      eu = par * s.u_KC * 8 + cu;
    } else {
      tap_i = kh * 3 + kw; es = cs; eu = cu;
    }
    const int tap9 = kh * 3 + kw;
    const int e = s.w_mode == 0 ? (tap9 * Cs + cs) * Cu + cu : (tap9 * Cu + cu) * Cs + cs;
    P->src[(size_t)e * 4] = find(tap_i, es, eu, false);
  }
  for (int cu = 0; cu < Cu && s.kind != GEN_DENSE; ++cu) {
    const int npar = s.u_layout == GEN_S2D ? 4 : 1;
    for (int par = 0; par < npar; ++par) {
      const int eu = s.u_layout == GEN_S2D ? par * s.u_KC * 8 + cu : cu;
      P->src[(size_t)(P->EW + cu) * 4 + par] = find(-1, -1, eu, true);
    }
  }
  for (size_t e = 0; e < (size_t)P->EW; ++e) if (P->src[e * 4] < 0) { delete P; return no("internal: a weight-gradient element has no accumulator"); }
  if (!upload) return P;
  if (cudaMalloc(reinterpret_cast<void**>(&P->dev), sizeof(G2PlanDev)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&P->src_dev), P->src.size() * sizeof(int32_t)) != cudaSuccess) {
    gen_wgrad_plan_free(P);
    return no("cudaMalloc failed");
  }
  cudaMemcpy(P->dev, &D, sizeof(G2PlanDev), cudaMemcpyHostToDevice);
  cudaMemcpy(P->src_dev, P->src.data(), P->src.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  return P;
}

int64_t gen_wgrad_plan_dump(const GenWgradPlan* p, int32_t* out, int64_t capacity) {
  const G2PlanDev& D = p->host;
  std::vector<int32_t> v = {2, D.n_roles, D.TRr, D.R_s, D.row0, D.col0, D.TW, D.s_PL, D.u_PL, (int32_t)D.CHs, (int32_t)D.CHu,
                            (int32_t)D.s_region, (int32_t)D.u_region, (int32_t)D.stage_bytes, (int32_t)D.ones_off, p->EW, p->spec.Cu};
  int n_mma = 0;
  for (int r = 0; r < D.n_roles; ++r) {
    const G2Role& R = D.roles[r];
    v.insert(v.end(), {R.mma0, R.mma_n, R.nS, R.nU, R.ncols});
    for (int k = 0; k < G2_MAXPL; ++k) v.push_back(R.s_plane[k]);
    for (int k = 0; k < G2_MAXPL; ++k) v.push_back(R.u_plane[k]);
    n_mma = std::max(n_mma, R.mma0 + R.mma_n);
  }
  v.push_back(n_mma);
  for (int i = 0; i < n_mma; ++i) {
    const G2Mma& M = D.mma[i];
    v.insert(v.end(), {(int32_t)M.a_lo, (int32_t)M.a_hi, (int32_t)M.b_lo, (int32_t)M.b_hi, (int32_t)M.idesc, (int32_t)M.d_col, (int32_t)M.b_ones});
  }
  v.push_back((int32_t)p->src.size());
  v.insert(v.end(), p->src.begin(), p->src.end());
  if (out && capacity >= (int64_t)v.size()) memcpy(out, v.data(), v.size() * sizeof(int32_t));
  return (int64_t)v.size();
}

void gen_wgrad_plan_free(GenWgradPlan* p) {
  if (!p) return;
  if (p->dev) cudaFree(p->dev);
  if (p->src_dev) cudaFree(p->src_dev);
  delete p;
}
size_t gen_wgrad_partial_floats(const GenWgradPlan*) { return (size_t)kNumSMs * G2_COLS * 128; }

int gen_wgrad_run(const GenWgradPlan* P, const GenPlanes& S, const GenPlanes& U, float* dW, float* db, float* partial, int B,
                  int* error_flag, const char* name, cudaStream_t st) {
  if (!gen_encode_fn() || !P->dev) return 1;
  const GenWgradSpec& s = P->spec;
  const G2PlanDev& D = P->host;
  CUtensorMap ms, mu;
  // hi planes only: lo planes (split tensors) sit behind them and are skipped through the per-image plane count
  if (make_planes_map(&ms, S.base, (int64_t)B * S.planes(), S.H, S.W, D.R_s) != CUDA_SUCCESS) return 2;
  if (make_window_map(&mu, U.base, (int64_t)B * U.planes(), U.H, U.W, D.TW, D.TRr) != CUDA_SUCCESS) return 2;
  G2Params p{};
  p.plan = P->dev;
  p.s_PL = S.planes(); p.u_PL = U.planes();        // planes per image INCLUDING lo planes (only the hi planes are read)
  p.partial = partial;
  p.B = B; p.Hg = s.Hg; p.Wg = s.Wg;
  p.tiles_y = cdiv(s.Hg, D.TRr); p.tiles_x = s.Wg / D.TW;
  p.num_tiles = B * p.tiles_y * p.tiles_x;
  p.error_flag = error_flag;
  if (S.planes() < D.s_PL || U.planes() < D.u_PL) return 3;
  int grid = p.num_tiles * D.n_roles < kNumSMs ? p.num_tiles * D.n_roles : kNumSMs / D.n_roles * D.n_roles;
  if (grid < D.n_roles) grid = D.n_roles;
  ProfScope prof_(name, st);
  g_launches += 2;
  cudaFuncSetAttribute(tc_gwgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P->smem);
  tc_gwgrad_kernel<<<grid, G2_THREADS, P->smem, st>>>(ms, mu, p);
  const int E = P->EW + (db ? s.Cu : 0);
  gen_wgrad_reduce_kernel<<<cdiv(E, 8), 256, 0, st>>>(partial, P->src_dev, E, D.n_roles, grid, p.num_tiles, dW, db, P->EW);
  return 0;
}

// ============================================================================================
//                                         packers
// ============================================================================================
namespace {

__global__ void gen_pack_x3_kernel(const float* __restrict__ x, int B, int H, int W, int split, uint4* out, uint4* x27) {
  const int h2 = H / 2, w2 = W / 2;
  const int64_t total = (int64_t)B * h2 * w2;
  const int PL = split ? 4 : 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % w2), r = (int)((i / w2) % h2);
    const int64_t n = i / ((int64_t)w2 * h2);
    float v[16];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const float* src = x + (((int64_t)n * H + 2 * r + a) * W + 2 * j) * 3;      // 6 contiguous floats: (b = 0,1) x 3 channels
#pragma unroll
      for (int k = 0; k < 6; ++k) v[a * 6 + k] = __ldg(src + k);
    }
    v[12] = v[13] = v[14] = v[15] = 0.f;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      hi[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
      lo[e] = pack_bf16x2(v[2 * e] - __uint_as_float(hi[e] << 16), v[2 * e + 1] - __uint_as_float(hi[e] & 0xFFFF0000u));
    }
    const int64_t plane = (int64_t)h2 * w2;
    uint4* o = out + (int64_t)n * PL * plane + (int64_t)r * w2 + j;
    o[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    o[plane] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    if (split) {
      o[2 * plane] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      o[3 * plane] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    }
    if (x27) {
      // the 3x3 stride-2 patch of output pixel (r, j): element (kh*3 + kw)*3 + c = x[2r + kh][2j + kw][c], zero outside the image
      // (TF SAME on even sizes pads bottom / right only).  Rows 2r, 2r+1 and columns 2j, 2j+1 are the values loaded above.
      float pv[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) pv[e] = 0.f;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int y = 2 * r + kh;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int xx = 2 * j + kw;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float val;
            if (kh < 2 && kw < 2) val = v[kh * 6 + kw * 3 + c];
            else val = (y < H && xx < W) ? __ldg(x + (((int64_t)n * H + y) * W + xx) * 3 + c) : 0.f;
            pv[(kh * 3 + kw) * 3 + c] = val;
          }
        }
      }
      uint4* o27 = x27 + (int64_t)n * 4 * plane + (int64_t)r * w2 + j;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        o27[q * plane] = make_uint4(pack_bf16x2(pv[8 * q], pv[8 * q + 1]), pack_bf16x2(pv[8 * q + 2], pv[8 * q + 3]),
                                    pack_bf16x2(pv[8 * q + 4], pv[8 * q + 5]), pack_bf16x2(pv[8 * q + 6], pv[8 * q + 7]));
    }
  }
}

// one thread per (image, output pixel, chunk)
__global__ void gen_pack_nhwc_kernel(const float* __restrict__ in, int B, int H, int W, int C, PlaneRef out, int split) {
  const int KC = out.KC;
  const int64_t total = (int64_t)B * H * W * KC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), chunk = (int)((i / ((int64_t)W * H)) % KC);
    const int n = (int)(i / ((int64_t)W * H * KC));
    const float* src = in + (((int64_t)n * H + y) * W + x) * C + chunk * 8;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = chunk * 8 + k < C ? __ldg(src + k) : 0.f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
      lo[e] = pack_bf16x2(v[2 * e] - __uint_as_float(hi[e] << 16), v[2 * e + 1] - __uint_as_float(hi[e] & 0xFFFF0000u));
    }
    out.base[unit_index(out, n, y, x, chunk, 0)] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (split) out.base[unit_index(out, n, y, x, chunk, 1)] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void gen_unpack_nhwc_kernel(PlaneRef in, int split, int B, int H, int W, int C, float* out) {
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), x = (int)((i / C) % W), y = (int)((i / ((int64_t)C * W)) % H);
    const int n = (int)(i / ((int64_t)C * W * H));
    const __nv_bfloat16* u = reinterpret_cast<const __nv_bfloat16*>(in.base + unit_index(in, n, y, x, c >> 3, 0));
    float v = __bfloat162float(u[c & 7]);
    if (split) {
      const __nv_bfloat16* l = reinterpret_cast<const __nv_bfloat16*>(in.base + unit_index(in, n, y, x, c >> 3, 1));
      v += __bfloat162float(l[c & 7]);
    }
    out[i] = v;
  }
}

}  // namespace

namespace {
// one thread per (chunk, n): 8 loads with stride N (coalesced across the warp), one 16-byte store (two when split)
__global__ void gen_pack_rows_T_kernel(const float* __restrict__ in, int R, int N, int split, uint4* out, int Np, int ones_n) {
  const int KC = (R + 7) / 8;
  const int64_t total = (int64_t)KC * Np;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % Np), chunk = (int)(i / Np);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = chunk * 8 + k < R ? (n < N ? __ldg(in + (int64_t)(chunk * 8 + k) * N + n) : (n == ones_n ? 1.f : 0.f)) : 0.f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
      lo[e] = pack_bf16x2(v[2 * e] - __uint_as_float(hi[e] << 16), v[2 * e + 1] - __uint_as_float(hi[e] & 0xFFFF0000u));
    }
    out[i] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (split) out[total + i] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
}  // namespace
void gen_pack_rows_T(const float* in, int R, int N, int split, void* out, cudaStream_t st, int Np, int ones_n) {
  ProfScope prof_("gen_pack_rows_T", st);
  ++g_launches;
  if (Np < N) Np = N;
  gen_pack_rows_T_kernel<<<grid_for((int64_t)((R + 7) / 8) * Np, 256, 8, 4), 256, 0, st>>>(in, R, N, split, reinterpret_cast<uint4*>(out), Np, ones_n);
}
namespace {
__global__ void gen_pack_cols_kernel(const float* __restrict__ in, const float* __restrict__ bias, int N, int C, int split, uint4* out,
                                     int Np, int ones_n) {
  const int KC = (C + 7) / 8;
  const int64_t total = (int64_t)KC * Np;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % Np), chunk = (int)(i / Np);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = chunk * 8 + k;
      v[k] = c < C ? (n < N ? __ldg(in + (int64_t)n * C + c) : ((n == ones_n && bias) ? __ldg(bias + c) : 0.f)) : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
      lo[e] = pack_bf16x2(v[2 * e] - __uint_as_float(hi[e] << 16), v[2 * e + 1] - __uint_as_float(hi[e] & 0xFFFF0000u));
    }
    out[i] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (split) out[total + i] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
}  // namespace
void gen_pack_cols(const float* in, const float* bias, int N, int C, int split, void* out, int Np, int ones_n, cudaStream_t st) {
  ProfScope prof_("gen_pack_cols", st);
  ++g_launches;
  gen_pack_cols_kernel<<<grid_for((int64_t)((C + 7) / 8) * Np, 256, 8, 4), 256, 0, st>>>(in, bias, N, C, split, reinterpret_cast<uint4*>(out), Np, ones_n);
}
void gen_pack_x3(const float* x, int B, int H, int W, int split, void* out, cudaStream_t st, void* x27_out) {
  ProfScope prof_("gen_pack_x3", st);
  ++g_launches;
  gen_pack_x3_kernel<<<grid_for((int64_t)B * (H / 2) * (W / 2), 256, 8, 4), 256, 0, st>>>(x, B, H, W, split, reinterpret_cast<uint4*>(out),
                                                                                          reinterpret_cast<uint4*>(x27_out));
}
void gen_pack_nhwc(const float* in, int B, int H, int W, int C, const GenPlanes& out, cudaStream_t st) {
  ProfScope prof_("gen_pack", st);
  ++g_launches;
  gen_pack_nhwc_kernel<<<grid_for((int64_t)B * H * W * out.KC, 256, 8, 4), 256, 0, st>>>(in, B, H, W, C, plane_ref(out), out.split);
}
void gen_unpack_nhwc(const GenPlanes& in, int B, int H, int W, int C, float* out, cudaStream_t st) {
  ++g_launches;
  gen_unpack_nhwc_kernel<<<grid_for((int64_t)B * H * W * C, 256, 8, 4), 256, 0, st>>>(plane_ref(in), in.split, B, H, W, C, out);
}

}  // namespace kc
