// Helper kernels of the tensor-core path: operand packing, the pre-drawn noise tile, bf16 self-attention.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "dsg_common.cuh"

// fp32 [rows, cols] (leading dim ld_src) -> bf16 [rows_pad, cols_pad], zero padding.
static __global__ void __launch_bounds__(256) pack_weight_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                              int rows, int cols, long long ld_src, int rows_pad, int cols_pad) {
  const long long total = (long long)rows_pad * cols_pad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / cols_pad), c = (int)(e - (long long)r * cols_pad);
    dst[e] = __float2bfloat16_rn((r < rows && c < cols) ? src[(long long)r * ld_src + c] : 0.f);
  }
}
// same, IEEE half (the clip kernel runs linear2 on fp16 operands: 11-bit mantissa for the GELU'd hidden and for W2)
static __global__ void __launch_bounds__(256) pack_weight_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst,
                                                             int rows, int cols, long long ld_src, int rows_pad, int cols_pad) {
  const long long total = (long long)rows_pad * cols_pad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / cols_pad), c = (int)(e - (long long)r * cols_pad);
    dst[e] = __float2half_rn((r < rows && c < cols) ? src[(long long)r * ld_src + c] : 0.f);
  }
}

// linear2 weights of the clip kernel: fp16 [rows, cols] with the K (column) order permuted inside every 128-column chunk to
// the order in which the GELU epilogue leaves the hidden units in tensor memory: K position 2j <- unit j, 2j + 1 <- unit 64 + j.
static __global__ void __launch_bounds__(256) pack_w2_perm_kernel(const float* __restrict__ src, __half* __restrict__ dst, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / cols), c = (int)(e - (long long)r * cols);
    const int p = c & 127, unit = (p & 1) ? 64 + (p >> 1) : (p >> 1);
    dst[e] = __float2half_rn(src[(long long)r * cols + (c & ~127) + unit]);
  }
}

// x fp32 [B, J, T]  ->  xb bf16 [B, S, Jpad] rows 1..T (row 0 and the pad columns stay zero): the K-major A
// operand of the input GEMM.  32x32 tile transpose through shared memory; both sides coalesced.
static __global__ void __launch_bounds__(256) pack_x_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb,
                                                         int J, int T, int S, int Jpad) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, j0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* xs = x + (long long)b * J * T;
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, f = f0 + tx;
    tile[r][tx] = (j < J && f < T) ? xs[(long long)j * T + f] : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* o = xb + (long long)b * S * Jpad;
  for (int r = ty; r < 32; r += 8) {
    const int f = f0 + r, j = j0 + tx;
    if (f < T && j < J) o[(long long)(f + 1) * Jpad + j] = __float2bfloat16_rn(tile[tx][r]);
  }
}

// Pre-drawn posterior noise for one step: z[b, e] = stream normal (draw 1 + k).  It depends on nothing the
// denoiser computes, so it is launched on a forked capture branch and overlaps the transformer GEMMs
// (ALU/MUFU work under tensor-pipe work).  Skipped when the step adds no noise (index 0 / DDIM).
struct NoiseArgs { float* z; const long long* clip_ids; StepRef step; int B; long long per_clip; int sampler; };

static __global__ void __launch_bounds__(256) noise_tile_kernel(const NoiseArgs a) {
  const int index = a.step.index();
  if (index == 0 || a.sampler != 0) return;
  const uint32_t draw = (uint32_t)(1 + a.step.k());
  const uint32_t k0 = a.step.key0(), k1 = a.step.key1(), seg = a.step.segment();
  const long long qpc = a.per_clip >> 2, total = qpc * a.B;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(g / qpc);
    const uint32_t q = (uint32_t)(g - (long long)b * qpc);
    *reinterpret_cast<float4*>(a.z + (long long)b * a.per_clip + 4ll * q) =
        philox_normal4(q, draw, (uint32_t)a.clip_ids[b], seg, k0, k1);
  }
}

static __global__ void bump_step_kernel(LoopParams* lp) { lp->k += 1; }

// ---------------------------------------------------------------------------------------------------
// Global self-attention on tensor cores (nn.TransformerEncoderLayer's SDPA: softmax(q k^T / sqrt(hd)) v, no mask;
// mdm.py:79-86).  One CTA per (clip, head), warp w owns query rows 16w..16w+15; S <= 16*RT keys, head dim 64.
// The problem per CTA is 96 x 96 x 64 — far below a tcgen05 tile (128 x N), so this uses the warp-level
// mma.sync.m16n8k16 bf16 path (HMMA; dynamic shared memory = 3 * 16 RT * (HD + 8) * 2 bytes): Q K^T accumulators stay in registers, are soft-maxed in place and re-used
// as the A fragments of P V (FlashAttention-2 register layout); K/V/Q tiles are staged once in shared memory
// with a 16-byte row pad so that every ldmatrix is bank-conflict free.
// qkv: bf16 [B*S, 3D] (q | k | v), out: bf16 [B*S, D].
// ---------------------------------------------------------------------------------------------------
DSG_DEVINL void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
DSG_DEVINL void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
DSG_DEVINL void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
DSG_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

template <int RT, int HD>   // row tiles of 16: S <= 16 * RT; head dim HD (64: ZEGGS, 96 / 128: the "+" denoisers with S = 151)
__global__ void __launch_bounds__(32 * RT) self_attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                    __nv_bfloat16* __restrict__ out, int S, int D, int heads,
                                                                    float scale_log2e) {
  constexpr int LDS = HD + 8, ROWS = 16 * RT;
  extern __shared__ __align__(16) uint8_t attn_smem[];
  __nv_bfloat16 (*Qs)[LDS] = reinterpret_cast<__nv_bfloat16 (*)[LDS]>(attn_smem);
  __nv_bfloat16 (*Ks)[LDS] = Qs + ROWS;
  __nv_bfloat16 (*Vs)[LDS] = Ks + ROWS;
  const int clip = blockIdx.x / heads, head = blockIdx.x - clip * heads;
  const __nv_bfloat16* base = qkv + (long long)clip * S * 3 * D + head * HD;
  for (int e = threadIdx.x; e < ROWS * 3 * (HD / 8); e += blockDim.x) {
    const int c8 = e % (HD / 8), rest = e / (HD / 8), mat = rest % 3, r = rest / 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < S) v = *reinterpret_cast<const uint4*>(base + (long long)r * 3 * D + mat * D + c8 * 8);
    __nv_bfloat16* dst = mat == 0 ? &Qs[r][c8 * 8] : (mat == 1 ? &Ks[r][c8 * 8] : &Vs[r][c8 * 8]);
    *reinterpret_cast<uint4*>(dst) = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = warp * 16;
  if (r0 >= S) return;

  // ---- scores = Q K^T (raw, unscaled) ----
  float sc[2 * RT][4];
#pragma unroll
  for (int nt = 0; nt < 2 * RT; ++nt) { sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t a[4];
    ldsm_x4(a[0], a[1], a[2], a[3], &Qs[r0 + (lane & 7) + ((lane >> 3) & 1) * 8][kk * 16 + (lane >> 4) * 8]);
#pragma unroll
    for (int np = 0; np < RT; ++np) {       // two key tiles (16 keys) per ldmatrix.x4
      uint32_t b0, b1, b2, b3;
      ldsm_x4(b0, b1, b2, b3, &Ks[np * 16 + (lane & 7) + (lane >> 4) * 8][kk * 16 + ((lane >> 3) & 1) * 8]);
      mma_bf16_16816(sc[2 * np], a, b0, b1);
      mma_bf16_16816(sc[2 * np + 1], a, b2, b3);
    }
  }
  // ---- softmax over keys (rows r0 + lane/4 and + 8) ----
  const int cbase = (lane & 3) * 2;
  float mx0 = -3.0e38f, mx1 = -3.0e38f;
#pragma unroll
  for (int nt = 0; nt < 2 * RT; ++nt) {
    const int c = nt * 8 + cbase;
    if (c >= S) { sc[nt][0] = -3.0e38f; sc[nt][2] = -3.0e38f; }
    if (c + 1 >= S) { sc[nt][1] = -3.0e38f; sc[nt][3] = -3.0e38f; }
    mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 2 * RT; ++nt) {
    sc[nt][0] = exp2f((sc[nt][0] - mx0) * scale_log2e); sc[nt][1] = exp2f((sc[nt][1] - mx0) * scale_log2e);
    sc[nt][2] = exp2f((sc[nt][2] - mx1) * scale_log2e); sc[nt][3] = exp2f((sc[nt][3] - mx1) * scale_log2e);
    sum0 += sc[nt][0] + sc[nt][1]; sum1 += sc[nt][2] + sc[nt][3];
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  // ---- O = P V ----
  float oc[HD / 8][4];
#pragma unroll
  for (int dt = 0; dt < HD / 8; ++dt) { oc[dt][0] = oc[dt][1] = oc[dt][2] = oc[dt][3] = 0.f; }
#pragma unroll
  for (int kt = 0; kt < RT; ++kt) {
    uint32_t a[4];
    a[0] = pack_bf16x2(sc[2 * kt][0], sc[2 * kt][1]);
    a[1] = pack_bf16x2(sc[2 * kt][2], sc[2 * kt][3]);
    a[2] = pack_bf16x2(sc[2 * kt + 1][0], sc[2 * kt + 1][1]);
    a[3] = pack_bf16x2(sc[2 * kt + 1][2], sc[2 * kt + 1][3]);
#pragma unroll
    for (int dp = 0; dp < HD / 16; ++dp) {   // two 8-wide d tiles per ldmatrix.x4.trans
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(b0, b1, b2, b3, &Vs[kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][dp * 16 + (lane >> 4) * 8]);
      mma_bf16_16816(oc[2 * dp], a, b0, b1);
      mma_bf16_16816(oc[2 * dp + 1], a, b2, b3);
    }
  }
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  const int row0 = r0 + (lane >> 2), row1 = row0 + 8;
  __nv_bfloat16* ob = out + (long long)clip * S * D + head * HD;
#pragma unroll
  for (int dt = 0; dt < HD / 8; ++dt) {
    const int c = dt * 8 + cbase;
    if (row0 < S) *reinterpret_cast<uint32_t*>(ob + (long long)row0 * D + c) = pack_bf16x2(oc[dt][0] * inv0, oc[dt][1] * inv0);
    if (row1 < S) *reinterpret_cast<uint32_t*>(ob + (long long)row1 * D + c) = pack_bf16x2(oc[dt][2] * inv1, oc[dt][3] * inv1);
  }
}
