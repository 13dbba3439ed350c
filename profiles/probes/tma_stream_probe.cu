// Probe: how fast can the SMs of a B200 stream an L2-resident bf16 weight slab through a TMA + mbarrier ring?
// Decides the round-2 design of the clip kernel's weight stream:
//   mode 0  unicast        every CTA loads every [128 x 64] tile (16 KB)                  — the round-1 kernel's pattern
//   mode 1  pair multicast cluster of 2; CTA r loads rows [64 r, +64) of the tile and multicasts it to both CTAs
//   mode 2  pair half      CTA r of a pair loads only its 64-row half (8 KB)               — the cta_group::2 pattern
//   mode 3  quad multicast cluster of 4; CTA r loads rows [32 r, +32) and multicasts to all four
// Output: time per pass over the slab, bytes landed per SM per second, L2 bytes read per second (chip).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream_probe tma_stream_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int TILE = 16384, MAXS = 12;

// CS = cluster size (1, 2, 4); MC = multicast.  Each CTA of a cluster loads rows [128 / CS * rank, +128 / CS) of a tile.
template <int CS, bool MC>
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int tiles, int passes, int nst) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + MAXS * TILE);
  uint64_t* empty = full + MAXS;
  const uint32_t rank = (CS > 1) ? cluster_ctarank() : 0;
  constexpr int PART = TILE / CS, PROWS = 128 / CS;
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAXS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], MC ? CS : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CS > 1) cluster_sync();
  const int start = (int)(((blockIdx.x / CS) * 37u) % (unsigned)tiles);
  const long long total = (long long)tiles * passes;
  if (threadIdx.x == 0) {                     // producer
    int s = 0; uint32_t ph = 0;
    int t = start;
    for (long long i = 0; i < total; ++i) {
      mbar_wait(&empty[s], ph ^ 1u);
      const int row = (t >> 2) * 128 + PROWS * (int)rank, kc = (t & 3) * 64;
      if (MC) {
        mbar_expect_tx(&full[s], TILE);
        tma_load_2d_mc(smem + s * TILE + rank * PART, &tm, &full[s], kc, row, (uint16_t)((1u << CS) - 1));
      } else {
        mbar_expect_tx(&full[s], PART);
        tma_load_2d(smem + s * TILE, &tm, &full[s], kc, row);
      }
      if (++t == tiles) t = 0;
      if (++s == nst) { s = 0; ph ^= 1u; }
    }
  } else if (threadIdx.x == 32) {             // consumer
    int s = 0; uint32_t ph = 0;
    for (long long i = 0; i < total; ++i) {
      mbar_wait(&full[s], ph);
      if (MC) { for (uint32_t c = 0; c < (uint32_t)CS; ++c) mbar_arrive_remote(&empty[s], c); }
      else mbar_arrive(&empty[s]);
      if (++s == nst) { s = 0; ph ^= 1u; }
    }
  }
  __syncthreads();
  if (CS > 1) cluster_sync();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CS, bool MC>
static void run(const char* name, EncodeTiledFn enc, void* slab, int rows, int grid, int nst, int passes) {
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {256, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {512};
  const cuuint32_t box[2] = {64, (cuuint32_t)(128 / CS)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, slab, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  const int tiles = rows / 128 * 4;
  const int smem = MAXS * TILE + 1024 + 512;
  CK(cudaFuncSetAttribute(probe<CS, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaLaunchKernelEx(&cfg, probe<CS, MC>, tm, tiles, 2, nst));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  CK(cudaLaunchKernelEx(&cfg, probe<CS, MC>, tm, tiles, passes, nst));
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  const double slab_bytes = (double)tiles * TILE;
  const double us_pass = ms * 1e3 / passes;
  const double landed_per_sm = (MC ? slab_bytes : slab_bytes / CS) / (us_pass * 1e-6) / 1e9;      // bytes written into one SM's smem
  const double l2_read = slab_bytes / CS * grid / (us_pass * 1e-6) / 1e12;                         // bytes requested from L2 (chip)
  printf("%-16s grid %3d stages %2d : %8.1f us per slab pass | %6.1f GB/s landed per SM | %5.2f TB/s read from L2 (chip)\n",
         name, grid, nst, us_pass, landed_per_sm, l2_read);
}

int main(int argc, char** argv) {
  const int passes = argc > 1 ? atoi(argv[1]) : 20;
  CK(cudaSetDevice(0));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int rows = 26112;                      // 13.4 MB bf16 [rows, 256]: the clip kernel's per-step weight set
  void* slab;
  CK(cudaMalloc(&slab, (size_t)rows * 512));
  CK(cudaMemset(slab, 1, (size_t)rows * 512));
  const int grids[] = {1, 2, 4, 16, 74, 148};
  for (int nst : {4, 8, 12})
    for (int g : grids) run<1, false>("unicast", enc, slab, rows, g, nst, passes);
  for (int nst : {4, 8, 12})
    for (int g : {2, 4, 16, 74, 148}) run<2, true>("pair multicast", enc, slab, rows, g, nst, passes);
  for (int nst : {4, 8, 12})
    for (int g : {2, 4, 16, 74, 148}) run<2, false>("pair half", enc, slab, rows, g, nst, passes);
  for (int nst : {4, 8, 12})
    for (int g : {4, 16, 72, 132, 148}) run<4, true>("quad multicast", enc, slab, rows, g, nst, passes);
  for (int nst : {8})
    for (int g : {4, 16, 72, 132, 148}) run<4, false>("quad quarter", enc, slab, rows, g, nst, passes);
  return 0;
}
