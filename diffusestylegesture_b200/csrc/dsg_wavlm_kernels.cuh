// WavLM-Large conditioning forward: the kernels around the tcgen05 GEMMs (reference main/mydiffusion_zeggs/WavLM/WavLM.py,
// modules_WavLM.py).  Activations are channels-last bf16; the residual stream of the transformer is fp32.
#pragma once
#include <cuda_bf16.h>
#include "dsg_common.cuh"
#include "dsg_tc_gemm.cuh"
#include "dsg_tc_kernels.cuh"

namespace wl {

DSG_DEVINL float ldv(const float* p) { return *p; }
DSG_DEVINL float ldv(const __nv_bfloat16* p) { return __bfloat162float(*p); }
DSG_DEVINL void stv(float* p, float v) { *p = v; }
DSG_DEVINL void stv(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------------
// conv layer 0 (C_in = 1, k = 10, stride 5) + LayerNorm over the 512 channels + GELU  (WavLM.py:404-422, 485-504).
// One warp per output position; lane owns channels lane, lane+32, ...
// ---------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) conv0_ln_gelu_kernel(const float* __restrict__ wav, __nv_bfloat16* __restrict__ out,
                                                           const float* __restrict__ w, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, int B, int N, int L0) {
  __shared__ float ws[10][512];
  __shared__ float gs[512], bs[512];
  for (int e = threadIdx.x; e < 5120; e += blockDim.x) { const int c = e / 10, k = e - c * 10; ws[k][c] = w[e]; }
  for (int e = threadIdx.x; e < 512; e += blockDim.x) { gs[e] = gamma[e]; bs[e] = beta[e]; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)B * L0;
  for (long long pos = (long long)blockIdx.x * 8 + warp; pos < total; pos += (long long)gridDim.x * 8) {
    const int b = (int)(pos / L0), m = (int)(pos - (long long)b * L0);
    const float* x = wav + (long long)b * N + 5 * m;
    float xv[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) xv[k] = __ldg(x + k);
    float v[16], s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) a = fmaf(ws[k][i * 32 + lane], xv[k], a);
      v[i] = a; s += a;
    }
    const float mean = warp_sum(s) * (1.0f / 512.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / 512.0f) + 1e-5f);
    __nv_bfloat16* o = out + pos * 512;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = i * 32 + lane;
      o[c] = __float2bfloat16_rn(tc::gelu_fast((v[i] - mean) * rstd * gs[c] + bs[c]));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Row LayerNorm (+ optional GELU): warp per row, C / 32 values per lane (two-pass variance in registers).
// conv blocks 1..6 (LN + GELU after the GEMM), WavLM.layer_norm, self_attn_layer_norm, final_layer_norm, encoder.layer_norm.
// ---------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut, int C, bool GELU>
static __global__ void __launch_bounds__(256) ln_rows_kernel(const TIn* __restrict__ in, TOut* __restrict__ out,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta, long long rows) {
  constexpr int PER = C / 32;
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const TIn* r = in + row * C;
  float v[PER], s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { v[i] = ldv(r + i * 32 + lane); s += v[i]; }
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + 1e-5f);
  TOut* o = out + row * C;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    float t = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    if (GELU) t = tc::gelu_fast(t);
    stv(o + c, t);
  }
}

// ---------------------------------------------------------------------------------------------------
// Gated relative position bias, the gate (modules_WavLM.py:520-533): per (row, head)
//   (ga, gb) = sigmoid( sum_{r<4} / sum_{r>=4} of grep_linear(h[row, head]) ),  gate = ga * (gb * grep_a[head] - 1) + 2.
// wab = [wa(64) | wb(64) | ba | bb] with the two 4-row sums of grep_linear folded at set-up.
// ---------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) gate_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ wab,
                                                  const float* __restrict__ grep_a, float* __restrict__ gate, int B, int L, int H) {
  __shared__ float ws[130];
  if (threadIdx.x < 130) ws[threadIdx.x] = wab[threadIdx.x];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;        // (row, head)
  if (idx >= (long long)B * L * H) return;
  const int hd = (int)(idx % H);
  const long long row = idx / H;
  const __nv_bfloat16* x = h + row * (H * 64) + hd * 64;
  float a = ws[128], b = ws[129];
#pragma unroll 8
  for (int i = 0; i < 64; i += 2) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + i));
    a = fmaf(f.x, ws[i], fmaf(f.y, ws[i + 1], a));
    b = fmaf(f.x, ws[64 + i], fmaf(f.y, ws[65 + i], b));
  }
  const float ga = 1.0f / (1.0f + __expf(-a)), gb = 1.0f / (1.0f + __expf(-b));
  const int bb = (int)(row / L), i = (int)(row - (long long)bb * L);
  gate[((long long)bb * H + hd) * L + i] = ga * (gb * grep_a[hd] - 1.0f) + 2.0f;
}

// ---------------------------------------------------------------------------------------------------
// Self-attention with the gated relative-position bias as an additive mask (F.multi_head_attention_forward with a float
// attn_mask, modules_WavLM.py:540-563): softmax(q k^T / 8 + gate[b,h,i] * pos_bias[h,i,j]) v, head dim 64, S <= 224.
// Two CTAs per (clip, head) — query rows [0, 112) and [112, 224), 7 warps each (two CTAs fit one SM and hide each other's
// latencies; K / V are staged by both from L2) —, warp w owns 16 query rows; keys are consumed in blocks of 32 with an online
// softmax (running max / sum, FlashAttention-2 register layout) so the score tile never exceeds 16 x 32 per warp.  The bias
// values of the NEXT key block are requested before the current block's MMAs (they are scattered 4-byte reads of the
// [H, S, S] table: one exposed L2 / L1 round trip per block otherwise).
// ---------------------------------------------------------------------------------------------------
constexpr int FA_LD = 72;
static __global__ void __launch_bounds__(224, 2) flash_attn_bias_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                             const float* __restrict__ gate, const float* __restrict__ pos_bias,
                                                             int S, int E, int H) {
  extern __shared__ __align__(16) uint8_t fa_smem[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(fa_smem);
  __nv_bfloat16* Ks = Qs + 224 * FA_LD;
  __nv_bfloat16* Vs = Ks + 224 * FA_LD;
  const int rhalf = blockIdx.x & 1, ch = blockIdx.x >> 1;
  const int clip = ch / H, head = ch - clip * H;
  const __nv_bfloat16* base = qkv + (long long)clip * S * 3 * E + head * 64;
  for (int e = threadIdx.x; e < 224 * 3 * 8; e += blockDim.x) {
    const int c8 = e & 7, rest = e >> 3, mat = rest % 3, r = rest / 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r < S) v = *reinterpret_cast<const uint4*>(base + (long long)r * 3 * E + mat * E + c8 * 8);
    __nv_bfloat16* dst = (mat == 0 ? Qs : (mat == 1 ? Ks : Vs)) + r * FA_LD + c8 * 8;
    *reinterpret_cast<uint4*>(dst) = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = (rhalf * 7 + warp) * 16;
  if (r0 >= S) return;
  const int g = lane >> 2, t2 = (lane & 3) * 2;
  const int ri0 = min(r0 + g, S - 1), ri1 = min(r0 + g + 8, S - 1);
  const float* gt = gate + ((long long)clip * H + head) * S;
  const float g0 = gt[ri0], g1 = gt[ri1];
  const float* pb0 = pos_bias + ((long long)head * S + ri0) * S;
  const float* pb1 = pos_bias + ((long long)head * S + ri1) * S;
  uint32_t qa[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldsm_x4(qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], Qs + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * FA_LD + kk * 16 + (lane >> 4) * 8);
  const float L2E = 1.4426950408889634f;
  float m0 = -3.0e38f, m1 = -3.0e38f, l0 = 0.f, l1 = 0.f;
  float oc[8][4];
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) { oc[dt][0] = oc[dt][1] = oc[dt][2] = oc[dt][3] = 0.f; }
  const int nkb = (S + 31) / 32;
  float pbn[16];                                   // gated bias of the next key block: [nt][e] for row ri0, then for row ri1
  auto load_bias = [&](int kb) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int cc = min(kb * 32 + nt * 8 + t2 + e, S - 1);
        pbn[nt * 2 + e] = __ldg(pb0 + cc);
        pbn[8 + nt * 2 + e] = __ldg(pb1 + cc);
      }
  };
  load_bias(0);
  for (int kb = 0; kb < nkb; ++kb) {
    float pbc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pbc[i] = pbn[i];
    if (kb + 1 < nkb) load_bias(kb + 1);
    float sc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(b0, b1, b2, b3, Ks + (kb * 32 + np * 16 + (lane & 7) + (lane >> 4) * 8) * FA_LD + kk * 16 + ((lane >> 3) & 1) * 8);
        mma_bf16_16816(sc[2 * np], qa[kk], b0, b1);
        mma_bf16_16816(sc[2 * np + 1], qa[kk], b2, b3);
      }
    float bm0 = -3.0e38f, bm1 = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = kb * 32 + nt * 8 + t2 + e;
        const bool ok = c < S;
        sc[nt][e] = ok ? (sc[nt][e] * 0.125f + g0 * pbc[nt * 2 + e]) * L2E : -3.0e38f;
        sc[nt][2 + e] = ok ? (sc[nt][2 + e] * 0.125f + g1 * pbc[8 + nt * 2 + e]) * L2E : -3.0e38f;
        bm0 = fmaxf(bm0, sc[nt][e]); bm1 = fmaxf(bm1, sc[nt][2 + e]);
      }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
    const float a0 = exp2f(m0 - n0), a1 = exp2f(m1 - n1);
    m0 = n0; m1 = n1; l0 *= a0; l1 *= a1;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) { oc[dt][0] *= a0; oc[dt][1] *= a0; oc[dt][2] *= a1; oc[dt][3] *= a1; }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      sc[nt][0] = exp2f(sc[nt][0] - m0); sc[nt][1] = exp2f(sc[nt][1] - m0);
      sc[nt][2] = exp2f(sc[nt][2] - m1); sc[nt][3] = exp2f(sc[nt][3] - m1);
      l0 += sc[nt][0] + sc[nt][1]; l1 += sc[nt][2] + sc[nt][3];
    }
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      uint32_t a[4];
      a[0] = pack_bf16x2(sc[2 * kt][0], sc[2 * kt][1]); a[1] = pack_bf16x2(sc[2 * kt][2], sc[2 * kt][3]);
      a[2] = pack_bf16x2(sc[2 * kt + 1][0], sc[2 * kt + 1][1]); a[3] = pack_bf16x2(sc[2 * kt + 1][2], sc[2 * kt + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(b0, b1, b2, b3, Vs + (kb * 32 + kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * FA_LD + dp * 16 + (lane >> 4) * 8);
        mma_bf16_16816(oc[2 * dp], a, b0, b1);
        mma_bf16_16816(oc[2 * dp + 1], a, b2, b3);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int row0 = r0 + g, row1 = row0 + 8;
  __nv_bfloat16* ob = out + (long long)clip * S * E + head * 64;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    const int c = dt * 8 + t2;
    if (row0 < S) *reinterpret_cast<uint32_t*>(ob + (long long)row0 * E + c) = pack_bf16x2(oc[dt][0] * i0, oc[dt][1] * i0);
    if (row1 < S) *reinterpret_cast<uint32_t*>(ob + (long long)row1 * E + c) = pack_bf16x2(oc[dt][2] * i1, oc[dt][3] * i1);
  }
}

// ---------------------------------------------------------------------------------------------------
// Positional conv input: x fp32 [B, L, 1024] -> group-major bf16 [B*16][Lp][64] with 64 zero rows in front
// (Conv1d padding = 64; the rows after the data stay zero), so that output row t of group g reads the 128 x 64
// contiguous elements starting at row t — a plain K-major GEMM operand (WavLM.py:514-527, 577-579).
// ---------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) pack_posconv_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xg, int B, int L, int Lp) {
  const long long total = (long long)B * L * 1024;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e & 1023);
    const long long bt = e >> 10;
    const int t = (int)(bt % L), b = (int)(bt / L);
    xg[(((long long)b * 16 + (c >> 6)) * Lp + 64 + t) * 64 + (c & 63)] = __float2bfloat16_rn(x[e]);
  }
}

// F.interpolate(..., mode='linear', align_corners=True) along time: [B, L, C] -> [B, P, C]   (sample.py:47)
static __global__ void __launch_bounds__(256) interp_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int L, int P, int C) {
  const float scale = P > 1 ? (float)(L - 1) / (float)(P - 1) : 0.f;
  const long long total = (long long)B * P * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long bp = e / C;
    const int p = (int)(bp % P), b = (int)(bp / P);
    const float real = scale * (float)p;
    const int i0 = (int)real;
    const int i1 = i0 + (i0 < L - 1 ? 1 : 0);
    const float w1 = real - (float)i0, w0 = 1.0f - w1;
    const float* src = in + (long long)b * L * C + c;
    out[e] = w0 * src[(long long)i0 * C] + w1 * src[(long long)i1 * C];
  }
}

// ---------------------------------------------------------------------------------------------------
// weight preparation
// ---------------------------------------------------------------------------------------------------
// Conv1d weight [C_out][C_in][k] fp32 -> bf16 [C_out][k * C_in] with column kk * C_in + ci (im2col order of a
// channels-last input); `scale` (nullable, [k]) multiplies column block kk (weight_norm of the positional conv).
static __global__ void __launch_bounds__(256) pack_conv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin,
                                                         int K, const float* __restrict__ scale) {
  const long long total = (long long)Cout * Cin * K;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(e % K);
    const long long r = e / K;
    const int ci = (int)(r % Cin), co = (int)(r / Cin);
    out[(long long)co * K * Cin + (long long)kk * Cin + ci] = __float2bfloat16_rn(w[e] * (scale ? scale[kk] : 1.0f));
  }
}
// nn.utils.weight_norm(dim=2): w = g[k] * v / ||v[:, :, k]||  ->  scale[k] = g[k] / norm[k]
static __global__ void __launch_bounds__(256) weight_norm_scale_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ scale,
                                                               int Cout, int Cin, int K) {
  const int kk = blockIdx.x;
  __shared__ float red[256];
  float s = 0.f;
  for (long long e = threadIdx.x; e < (long long)Cout * Cin; e += blockDim.x) { const float t = v[e * K + kk]; s = fmaf(t, t, s); }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) scale[kk] = g[kk] / sqrtf(red[0]);
}

}  // namespace wl
