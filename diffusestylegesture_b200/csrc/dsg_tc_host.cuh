// Host-side helpers shared by the tensor-core translation units (dsg_tc.cu, dsg_wavlm.cu): tensor-map encoding through
// the driver entry point (no link-time libcuda dependency), GEMM launch, bf16 weight packing.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "dsg_engine.h"
#include "dsg_tc_gemm.cuh"
#include "dsg_tc_kernels.cuh"
#include "dsg_tc_gemm_persistent.cuh"

using bf16 = __nv_bfloat16;
using namespace tc;

// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int get_encode() {
  if (g_encode) return DSG_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return dsg_fail(DSG_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  g_encode = (EncodeTiledFn)fn;
  return DSG_OK;
}

// bf16 row-major [rows, cols] -> 2-D tensor map with a (box_rows x 64) box and 128-byte swizzle
static int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  TRY(get_encode());
  if (cols % 8) return dsg_fail(DSG_ERR_BAD_SHAPE, "tensor map: row length %llu not a multiple of 8 bf16", (unsigned long long)cols);
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * sizeof(bf16)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return dsg_fail(DSG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r,
                                         (unsigned long long)rows, (unsigned long long)cols);
  return DSG_OK;
}

// 3-D variant for batched GEMMs whose A rows overlap (im2col view of a channels-last conv input): dims
// {row_len, rows, batches}, strides in ELEMENTS between consecutive rows / batches.
static int make_tmap3(CUtensorMap* m, const void* ptr, uint64_t row_len, uint64_t rows, uint64_t batches, uint64_t row_stride,
                      uint64_t batch_stride, uint32_t box_rows) {
  TRY(get_encode());
  if ((row_stride % 8) || (batch_stride % 8)) return dsg_fail(DSG_ERR_BAD_SHAPE, "tensor map: strides must be multiples of 8 bf16");
  const cuuint64_t gdim[3] = {row_len, rows, batches};
  const cuuint64_t gstride[2] = {row_stride * sizeof(bf16), batch_stride * sizeof(bf16)};
  const cuuint32_t box[3] = {(cuuint32_t)BK, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return dsg_fail(DSG_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed (%d)", (int)r);
  return DSG_OK;
}

template <int BN, int STAGES, int EPI>
static int launch_tc(dsg_engine* e, const CUtensorMap& a, const CUtensorMap& b, const TcEpiArgs& ep, int n_tiles, cudaStream_t st,
                     int gz = 1) {
  static bool configured = false;
  constexpr int smem = TcSmem<BN, STAGES>::TOTAL;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(tc_gemm_kernel<BN, STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(((ep.rows_per_z > 0 ? ep.rows_per_z : ep.M) + BM - 1) / BM, n_tiles, gz);
  tc_gemm_kernel<BN, STAGES, EPI><<<grid, 256, smem, st>>>(a, b, ep);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

// fp32 row-major [rows, cols] output -> 2-D tensor map with a 32 x 32 box and 128-byte swizzle (the persistent GEMM's
// cp.reduce.async.bulk.tensor epilogue: one reduce-add per warp block; rows / columns beyond the tensor are clipped)
static int make_tmap_f32_out(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  TRY(get_encode());
  if (ld % 4) return dsg_fail(DSG_ERR_BAD_SHAPE, "tensor map: fp32 leading dimension %llu not a multiple of 4", (unsigned long long)ld);
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {ld * sizeof(float)};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return dsg_fail(DSG_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 out) failed (%d)", (int)r);
  return DSG_OK;
}

// Persistent GEMM (dsg_tc_gemm_persistent.cuh): flat 2-D problems with N a multiple of 256-column tiles; `out_map` is the fp32
// output map of EPI_RESID (ignored otherwise).  The weight map must have 256-row boxes.
template <int EPI>
static int launch_tc_persistent(dsg_engine* e, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap* out_map,
                                const TcEpiArgs& ep, cudaStream_t st) {
  constexpr int STAGES = (EPI == EPI_RESID) ? 3 : 4;
  using SM = TcPersistSmem<STAGES, EPI>;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY(cudaFuncSetAttribute(tc_gemm_persistent_kernel<STAGES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
    configured = true;
  }
  const int m_tiles = (ep.M + BM - 1) / BM, n_tiles = (ep.N + 255) / 256;
  const int sms = e->num_sms > 0 ? e->num_sms : 148;
  const int grid = m_tiles * n_tiles < sms ? m_tiles * n_tiles : sms;
  tc_gemm_persistent_kernel<STAGES, EPI><<<grid, 320, SM::TOTAL, st>>>(a, b, out_map ? *out_map : a, ep, m_tiles, n_tiles);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

template <typename T>
static int dalloc0(T** p, size_t n) {
  CUDA_TRY(cudaMalloc((void**)p, n * sizeof(T)));
  CUDA_TRY(cudaMemset(*p, 0, n * sizeof(T)));
  return DSG_OK;
}

static int pack_w(const float* src, bf16* dst, int rows, int cols, long long ld, int rows_pad, int cols_pad) {
  pack_weight_bf16_kernel<<<296, 256>>>(src, dst, rows, cols, ld, rows_pad, cols_pad);
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

