// Development aid: what a persistent one-CTA-per-SM kernel gets out of TMA tile loads of a chunk-planar activation
// [B][4][H][W][8] bf16 (the decoder tail's 32-channel tensors) as a function of the box shape, the column offset of the box
// (halo tiles start one pixel left of a 30-pixel tile grid: 16-byte but not 32-byte aligned), the ring depth and the number
// of planes per box.  No compute: a consumer thread releases every slot as soon as it has landed.  Build + run (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I trustedai-cl-vae-ad_b200/csrc tools/tma_stream.cu -o /tmp/tma_stream && /tmp/tma_stream
#include <cuda.h>
#include <cstdio>
#include <vector>
#include "tc_common.cuh"

using namespace kc::tc;

__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* map, uint64_t* mbar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(mbar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma4(void* dst, const CUtensorMap* map, uint64_t* mbar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(mbar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void arrive(uint64_t* mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(mbar)) : "memory"); }

struct Cfg {
  int four_d;      // 1: one 4-D box (all 4 planes) per slab; 0: four 3-D boxes (one per plane)
  int box_w;       // pixels per box row
  int rows;        // rows per slab
  int slabs;       // slabs per tile (rows * slabs = tile rows incl. halo)
  int row0;        // first row of a tile relative to ty * tile_h
  int tile_h, tile_w, xoff;
  int slots;       // ring depth
  const char* note;
};

__global__ void __launch_bounds__(64, 1) tma_stream_kernel(const __grid_constant__ CUtensorMap map, Cfg c, int B, int H, int W, int* err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_bar[64], empty_bar[64];
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.slots; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const int tiles_y = (H + c.tile_h - 1) / c.tile_h, tiles_x = (W + c.tile_w - 1) / c.tile_w;
  const int num_tiles = B * tiles_y * tiles_x;
  const uint32_t plane_bytes = (uint32_t)c.box_w * 16 * c.rows, slab_bytes = 4 * plane_bytes;
  if (threadIdx.x == 0) {
    int g = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n = t / (tiles_y * tiles_x), rem = t % (tiles_y * tiles_x), ty = rem / tiles_x, tx = rem % tiles_x;
      for (int j = 0; j < c.slabs; ++j, ++g) {
        const int slot = g % c.slots;
        if (!mbar_wait(&empty_bar[slot], (uint32_t)((g / c.slots) & 1) ^ 1u, 1u << 22)) { *err = 1; return; }
        expect_tx(&full_bar[slot], slab_bytes);
        unsigned char* dst = smem + (size_t)slot * slab_bytes;
        const int x = (tx * c.tile_w + c.xoff) * 8, y = ty * c.tile_h + c.row0 + j * c.rows;
        if (c.four_d) tma4(dst, &map, &full_bar[slot], x, y, 0, n);
        else
          for (int pl = 0; pl < 4; ++pl) tma3(dst + pl * plane_bytes, &map, &full_bar[slot], x, y, n * 4 + pl);
      }
    }
  } else if (threadIdx.x == 32) {
    int g = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x)
      for (int j = 0; j < c.slabs; ++j, ++g) {
        const int slot = g % c.slots;
        if (!mbar_wait(&full_bar[slot], (uint32_t)(g / c.slots) & 1u, 1u << 22)) { *err = 2; return; }
        arrive(&empty_bar[slot]);
      }
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int B = 256, H = 224, W = 300;
  const size_t bytes = (size_t)B * 4 * H * W * 16;
  void* act;
  cudaMalloc(&act, bytes);
  cudaMemset(act, 0, bytes);
  int* err;
  cudaMalloc(&err, 4);
  cudaMemset(err, 0, 4);
  void* big;                       // L2 flush between runs
  cudaMalloc(&big, 256u << 20);
  EncFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
  std::vector<Cfg> cfgs = {
      {0, 32, 34, 1, -1, 32, 30, -1, 2, "3-D boxes 32 px x 34 rows, 2 tile stages, x = 30 tx - 1 (old out-layer wgrad)"},
      {0, 32, 34, 1, -1, 32, 30, 0, 2, "  same, x = 30 tx"},
      {0, 32, 34, 1, -1, 32, 30, -2, 2, "  same, x = 30 tx - 2"},
      {0, 32, 34, 1, -1, 32, 32, 0, 2, "  same, x = 32 tx (512-byte aligned rows)"},
      {0, 32, 34, 1, -1, 32, 30, -1, 3, "  x = 30 tx - 1, 3 tile stages"},
      {1, 32, 2, 17, -1, 32, 30, -1, 21, "4-D boxes 32 px x 2 rows x 4 planes, 21 slots, x = 30 tx - 1 (slab ring)"},
      {1, 32, 2, 17, -1, 32, 30, -2, 21, "  same, x = 30 tx - 2"},
      {1, 32, 2, 17, -1, 32, 30, 0, 21, "  same, x = 30 tx"},
      {1, 32, 2, 17, -1, 32, 32, 0, 21, "  same, x = 32 tx"},
      {1, 32, 2, 17, -1, 32, 30, -1, 42, "  x = 30 tx - 1, 42 slots"},
      {1, 32, 17, 2, -1, 32, 30, -1, 5, "4-D boxes 32 px x 17 rows x 4 planes (35 KB), 5 slots, x = 30 tx - 1"},
      {1, 32, 17, 2, -1, 32, 30, -2, 5, "  same, x = 30 tx - 2"},
      {1, 32, 17, 2, -1, 32, 32, 0, 5, "  same, x = 32 tx"},
      {0, 32, 17, 2, -1, 32, 30, -1, 5, "3-D boxes 32 px x 17 rows (8.7 KB x 4), 5 slots, x = 30 tx - 1"},
      {0, 32, 17, 2, -1, 32, 30, -2, 5, "  same, x = 30 tx - 2"},
      {0, 32, 17, 2, -1, 32, 32, 0, 5, "  same, x = 32 tx"},
  };
  cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("%-100s %9s %9s\n", "pattern", "ms", "TB/s");
  for (const Cfg& c : cfgs) {
    CUtensorMap map;
    CUresult r;
    if (c.four_d) {
      const cuuint64_t gdim[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, 4, (cuuint64_t)B};
      const cuuint64_t gstr[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)4 * H * W * 16};
      const cuuint32_t box[4] = {(cuuint32_t)c.box_w * 8, (cuuint32_t)c.rows, 4, 1};
      const cuuint32_t es[4] = {1, 1, 1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, act, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const cuuint64_t gdim[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)B * 4};
      const cuuint64_t gstr[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
      const cuuint32_t box[3] = {(cuuint32_t)c.box_w * 8, (cuuint32_t)c.rows, 1};
      const cuuint32_t es[3] = {1, 1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, act, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.note, (int)r); continue; }
    const size_t smem = (size_t)c.slots * 4 * c.box_w * 16 * c.rows;
    if (smem > 220 * 1024) { printf("%s: smem %zu\n", c.note, smem); continue; }
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaMemsetAsync(big, rep, 256u << 20);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      tma_stream_kernel<<<148, 64, smem>>>(map, c, B, H, W, err);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.note, cudaGetErrorString(e)); return 1; }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
    int herr = 0;
    cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
    const int tiles = B * ((H + c.tile_h - 1) / c.tile_h) * ((W + c.tile_w - 1) / c.tile_w);
    const double moved = (double)tiles * c.slabs * 4 * c.box_w * 16 * c.rows;
    printf("%-100s %9.3f %9.2f%s\n", c.note, best, moved / (best * 1e-3) / 1e12, herr ? "  (TIMEOUT)" : "");
    cudaMemset(err, 0, 4);
  }
  return 0;
}
