// Persistent tcgen05 GEMM for the large flat GEMMs (WavLM in_proj / out_proj / fc1 / fc2, the multi-kernel denoiser's in_proj and
// linear1):  D[TMEM, fp32] = A[smem, bf16] * W[smem, bf16]^T with the epilogue of tile i overlapped with the main loop of tile i+1.
//
//   grid = min(#tiles, #SMs); CTA c runs tiles c, c + grid, ... (m fastest: CTAs that run together share the weight tile in L2).
//   warp 0 = TMA producer (runs ahead across tile boundaries through the STAGES-deep ring), warp 1 = MMA issuer, warp 2 allocates
//   the 512 TMEM columns = TWO 128 x 256 fp32 accumulators, warps 2..9 = epilogue (two warps per TMEM lane quarter, interleaved
//   32-column chunks).  Barriers: full / empty per stage, tmem_full / tmem_empty per accumulator.
//   The non-persistent kernel (dsg_tc_gemm.cuh) serialises setup + main loop + epilogue per tile: ~12 us per 128 x 256 x 1024 tile
//   of which ~6 us is the main loop; measured 690 TFLOP/s on the WavLM layer GEMMs.
//
// Epilogues: EPI_BF16 / EPI_GELU (bias, optional GELU, bf16 row-major stores) and EPI_RESID (fp32 out += acc + bias): the
// read-modify-write is done by the TMA engine in L2 — every warp stages its 32 x 32 block (+ bias) in a SWIZZLE_128B shared-memory
// tile and issues one `cp.reduce.async.bulk.tensor.2d ... .add` per block: no reads of the old value by the SM, one 4 KB bulk
// operation instead of 32 scattered 16-byte loads + 32 stores per instruction, rows beyond M clipped by the tensor map.
#pragma once
#include "dsg_tc_gemm.cuh"

namespace tc {

template <int STAGES, int EPI>
struct TcPersistSmem {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TILE_OFF = STAGES * STAGE_BYTES;                       // 8 warps x 4 KB (EPI_RESID only)
  static constexpr int TILE_BYTES = (EPI == EPI_RESID) ? 8 * 4096 : 0;
  static constexpr int PRM_OFF = TILE_OFF + TILE_BYTES;                       // bias of the tile being written out, double-buffered
  static constexpr int BAR_OFF = PRM_OFF + 2 * BN * 4;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
  static_assert(TOTAL <= 232448, "shared memory budget");
};

DSG_DEVINL void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
DSG_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DSG_DEVINL void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
DSG_DEVINL void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int STAGES, int EPI>
__global__ void __launch_bounds__(320, 1)
tc_gemm_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmOut, const TcEpiArgs ep, int m_tiles, int n_tiles) {
  static_assert(EPI == EPI_BF16 || EPI == EPI_GELU || EPI == EPI_RESID, "persistent GEMM: bf16 / GELU / residual epilogues");
  using SM = TcPersistSmem<STAGES, EPI>;
  constexpr int BN = SM::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* prm = reinterpret_cast<float*>(smem + SM::PRM_OFF);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int num_kb = (ep.K + BK - 1) / BK;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (EPI == EPI_RESID) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===== TMA producer: one continuous stream of k-blocks over all tiles of this CTA =====
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int m0 = (t % m_tiles) * BM, n0 = (t / m_tiles) * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
        if (elect_one_lane()) {
          mbar_expect_tx(&full_bar[s], SM::STAGE_BYTES);
          tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
          tma_load_2d(a_dst + SM::A_BYTES, &tmB, &full_bar[s], kb * BK, n0);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator i & 1 for the i-th tile of this CTA =====
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    int s = 0, i = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int buf = i & 1;
      mbar_wait(&tmem_empty[buf], (((uint32_t)i >> 1) & 1u) ^ 1u);      // the epilogue has drained this accumulator
      tcgen05_fence_after();
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * SM::STAGE_BYTES), b_addr = a_addr + SM::A_BYTES;
        if (elect_one_lane()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16(tmem_base + buf * BN, make_sw128_desc(a_addr + k * UMMA_K * 2), make_sw128_desc(b_addr + k * UMMA_K * 2), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          tcgen05_commit(&empty_bar[s]);
          if (kb == num_kb - 1) tcgen05_commit(&tmem_full[buf]);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps 2..9: lane quarter = warp % 4, column half = (warp - 2) / 4 =====
    const int wq = warp & 3, half = (warp - 2) >> 2, et = threadIdx.x - 64;      // et: 0..255
    const int rloc = wq * 32 + lane;
    float v[32];
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int buf = i & 1;
      const int m0 = (t % m_tiles) * BM, n0 = (t / m_tiles) * BN;
      const int row = m0 + rloc;
      const bool row_ok = row < ep.M;
      // this tile's bias -> shared memory (double-buffered by accumulator), under a barrier of the 256 epilogue threads
      float* pb = prm + buf * BN;
      pb[et] = (ep.bias != nullptr && n0 + et < ep.N) ? __ldg(ep.bias + n0 + et) : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tmem_full[buf], ((uint32_t)i >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + buf * BN;
#pragma unroll 1
      for (int c = half * 32; c < BN; c += 64) {
        tmem_ld32(taddr + c, v);
        const int n = n0 + c;
        if (n >= ep.N) continue;                                     // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(pb + c + j);
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
        if constexpr (EPI == EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
        }
        if constexpr (EPI == EPI_RESID) {
          // 32 x 32 fp32 block -> SWIZZLE_128B tile (row = 128 B, 16-byte chunk j of row r at chunk j ^ (r & 7)) -> TMA reduce-add
          uint8_t* tile = smem + SM::TILE_OFF + (warp - 2) * 4096;
          if (lane == 0) bulk_wait_read0();                          // the previous block of this warp has been read out
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) { tma_reduce_add_2d(&tmOut, tile, n, m0 + wq * 32); bulk_commit(); }
        } else {
          if (!row_ok) continue;
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row * ep.ldc + n;
          if (n + 32 <= ep.N) {
            store_bf16x32(o, v);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (n + j < ep.N) o[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tmem_empty[buf]);
    }
    if constexpr (EPI == EPI_RESID) { if (lane == 0) bulk_wait0(); }    // every reduce has reached global memory before the CTA exits
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace tc
