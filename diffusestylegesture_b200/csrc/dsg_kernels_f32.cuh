// fp32 CUDA-core kernels: the validation path of the engine (DSG_PRECISION_FP32) and the non-GEMM kernels
// shared with the tensor-core path.  Each kernel names the reference code it stands for
// (paths relative to /root/reference).
#pragma once
#include <cuda_bf16.h>
#include "dsg_common.cuh"

DSG_DEVINL float ldf(const float* p) { return *p; }
DSG_DEVINL float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
DSG_DEVINL void stf(float* p, float v) { *p = v; }
DSG_DEVINL void stf(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---------------------------------------------------------------------------------------------------
// Generic strided SGEMM  C[m,n] = act( sum_k A[m,k] * B[n,k] + bias[n] + clipvec[clip(m)][n]
//                                       + tvec[trow(clip(m))][n] + addmat[m][n] )
// Rows are addressed "clip-decomposed" so that one kernel covers every Linear on the path:
//   x^T as A (InputProcess, mdm.py:461-467: x[b,:,f] is a strided column of [B,J,T]),
//   rows 1..T of the [B,S,D] token buffer as A and x0[b,j,f] as transposed C (OutputProcess, mdm.py:490-504),
//   weight sub-blocks of input_process2 (mdm.py:141, 202-206) as strided B.
// ---------------------------------------------------------------------------------------------------
struct RowMap {
  int rows_per_clip;        // M = clips * rows_per_clip
  int row0;                 // first row inside a clip
  long long clip_stride;    // elements between clips
  long long row_stride;     // elements between rows
  DSG_DEVINL long long off(int m, int& clip) const {
    clip = m / rows_per_clip;
    const int r = m - clip * rows_per_clip;
    return (long long)clip * clip_stride + (long long)(r + row0) * row_stride;
  }
};

struct GemmF32Args {
  const float* A; RowMap am; long long a_kstride;
  const float* B; long long b_nstride, b_kstride;
  float* C; RowMap cm; long long c_nstride;
  int M, N, K;
  const float* bias;                       // [N] or null
  const float* clipvec; int clipvec_ld;    // [clips][ld] or null
  const float* tvec; int tvec_ld;          // [rows][ld] or null; row chosen by tsel[clip] or tmap[step.index()]
  const int* tsel;                         // per-clip row index (device) or null
  const int* tmap; StepRef step;           // used when tsel == null
  const float* addmat; long long addmat_ld;  // [M][ld] or null
  int act;                                 // 0 none, 1 gelu(erf), 2 silu
};

template <bool A_M_CONTIG, bool SWAP_MN>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmF32Args g) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;
  const int tm = SWAP_MN ? tx : ty, tn = SWAP_MN ? ty : tx;

  // per-thread global row offsets for the A loads (constant over k)
  long long a_off[4]; bool a_ok[4]; int a_mm[4], a_kk[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int e = tid + p * 256;
    int mm, kk;
    if (A_M_CONTIG) { mm = e & 63; kk = e >> 6; } else { kk = e & 15; mm = e >> 4; }
    a_mm[p] = mm; a_kk[p] = kk;
    const int m = m0 + mm;
    a_ok[p] = m < g.M;
    int clip;
    a_off[p] = a_ok[p] ? g.am.off(m, clip) : 0;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int k = k0 + a_kk[p];
      float v = 0.f;
      if (a_ok[p] && k < g.K) v = __ldg(g.A + a_off[p] + (long long)k * g.a_kstride);
      As[a_kk[p]][a_mm[p]] = v;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int e = tid + p * 256;
      const int kk = e & 15, nn = e >> 4;
      const int k = k0 + kk, n = n0 + nn;
      float v = 0.f;
      if (n < g.N && k < g.K) v = __ldg(g.B + (long long)n * g.b_nstride + (long long)k * g.b_kstride);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tm * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tn * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm * 4 + i;
    if (m >= g.M) continue;
    int clip;
    const long long coff = g.cm.off(m, clip);
    const float* cv = g.clipvec ? g.clipvec + (long long)clip * g.clipvec_ld : nullptr;
    const float* tv = nullptr;
    if (g.tvec) {
      const int row = g.tsel ? g.tsel[clip] : g.tmap[g.step.index()];
      tv = g.tvec + (long long)row * g.tvec_ld;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (cv) v += cv[n];
      if (tv) v += tv[n];
      if (g.addmat) v += g.addmat[(long long)m * g.addmat_ld + n];
      if (g.act == 1) v = gelu_erf(v); else if (g.act == 2) v = silu(v);
      g.C[coff + (long long)n * g.c_nstride] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm over rows (F.layer_norm inside nn.TransformerEncoderLayer, eps 1e-5; mdm.py:79-86).
// The residual sum is already in `in` (added by the producing GEMM's epilogue).  One warp per row.
// ---------------------------------------------------------------------------------------------------
template <int MAX_PER_LANE>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int rows, int D) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* r = in + (long long)warp * D;
  float v[MAX_PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < MAX_PER_LANE; ++c) {
    const int d = lane + 32 * c;
    v[c] = d < D ? r[d] : 0.f;
    s += v[c];
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < MAX_PER_LANE; ++c) {
    const int d = lane + 32 * c;
    const float t = d < D ? v[c] - mean : 0.f;
    q += t * t;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
  float* o = out + (long long)warp * D;
#pragma unroll
  for (int c = 0; c < MAX_PER_LANE; ++c) {
    const int d = lane + 32 * c;
    if (d < D) o[d] = (v[c] - mean) * rstd * gamma[d] + beta[d];
  }
}

// ---------------------------------------------------------------------------------------------------
// rope -> windowed causal local attention (q = k = v) -> prepend token -> rope   (mdm.py:207-229,
// local_attention/local_attention.py:91-199, local_attention/rotary.py:6-25).
// One CTA per (clip, local head).  Query frame f attends keys max(0,(f/w-1)*w) .. f  (<= 2w <= 32 keys:
// one key per lane).  Output goes to rows 1..T of the [B,S,D] token buffer with the second rotary (position
// f+1) applied; row 0 = tok = emb_1 + emb_t (rotary at position 0 is the identity).
// rope tables: cs[pos][i] = (cos, sin)(pos * 10000^(-2i/hd)), i < hd/2.
// ---------------------------------------------------------------------------------------------------
struct LocalAttnArgs {
  const float* h;          // input_process2 output: frame f of clip b at h + b*h_clip_stride + (f + h_row0)*D
  long long h_clip_stride; int h_row0;
  float* xs;               // [B,S,D]
  __nv_bfloat16* xsb;      // optional bf16 copy of xs (A operand of the first in_proj GEMM), nullable
  const float* emb1;       // [B,D]    style/seed embedding (step-invariant)
  const float* te;         // [n_t,D]  timestep-embedding table
  const int* tsel; const int* tmap; StepRef step;
  const float2* cs;        // [S][hd/2]
  int T, D, heads, window;
};

__global__ void __launch_bounds__(512) local_attention_kernel(const LocalAttnArgs a) {
  extern __shared__ float smem[];
  const int hd = a.D / a.heads, half = hd >> 1, ldz = hd + 1;
  float* z = smem;                               // [T][hd+1]  rope'd h slice
  float* orow = smem + a.T * ldz;                // [warps][hd]
  const int clip = blockIdx.x / a.heads, head = blockIdx.x - clip * a.heads;
  const int S = a.T + 1;
  const float* hb = a.h + (long long)clip * a.h_clip_stride + (long long)a.h_row0 * a.D + head * hd;
  for (int e = threadIdx.x; e < a.T * half; e += blockDim.x) {
    const int f = e / half, i = e - f * half;
    const float z1 = hb[(long long)f * a.D + i], z2 = hb[(long long)f * a.D + i + half];
    const float2 c = a.cs[f * half + i];
    z[f * ldz + i] = z1 * c.x - z2 * c.y;
    z[f * ldz + i + half] = z2 * c.x + z1 * c.y;
  }
  float* xb = a.xs + (long long)clip * S * a.D + head * hd;
  __nv_bfloat16* xbb = a.xsb ? a.xsb + (long long)clip * S * a.D + head * hd : nullptr;
  if (threadIdx.x < hd) {
    const int row = a.tsel ? a.tsel[clip] : a.tmap[a.step.index()];
    const int col = head * hd + threadIdx.x;
    const float tokv = a.emb1[(long long)clip * a.D + col] + a.te[(long long)row * a.D + col];
    xb[threadIdx.x] = tokv;
    if (xbb) xbb[threadIdx.x] = __float2bfloat16_rn(tokv);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float scale = rsqrtf((float)hd);
  float* ow = orow + warp * hd;
  for (int f = warp; f < a.T; f += nwarps) {
    const int lo = max(0, (f / a.window - 1) * a.window);
    const int nk = f - lo + 1;
    float s = -3.402823466e+38f;
    if (lane < nk) {
      const float* q = z + f * ldz;
      const float* k = z + (lo + lane) * ldz;
      float d = 0.f;
      for (int c = 0; c < hd; ++c) d = fmaf(q[c], k[c], d);
      s = d * scale;
    }
    const float mx = warp_max(s);
    const float p = lane < nk ? expf(s - mx) : 0.f;
    const float inv = 1.0f / warp_sum(p);
    for (int c0 = 0; c0 < hd; c0 += 32) {          // warp-uniform trip count: every lane joins the shuffles
      const int c = c0 + lane, cc = min(c, hd - 1);
      float o = 0.f;
      for (int j = 0; j < nk; ++j) o = fmaf(__shfl_sync(0xffffffffu, p, j), z[(lo + j) * ldz + cc], o);
      if (c < hd) ow[c] = o * inv;
    }
    __syncwarp();
    for (int i = lane; i < half; i += 32) {
      const float2 c = a.cs[(f + 1) * half + i];
      const float o1 = ow[i], o2 = ow[i + half];
      const float r1 = o1 * c.x - o2 * c.y, r2 = o2 * c.x + o1 * c.y;
      xb[(long long)(f + 1) * a.D + i] = r1;
      xb[(long long)(f + 1) * a.D + i + half] = r2;
      if (xbb) {
        xbb[(long long)(f + 1) * a.D + i] = __float2bfloat16_rn(r1);
        xbb[(long long)(f + 1) * a.D + i + half] = __float2bfloat16_rn(r2);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// Global self-attention of nn.TransformerEncoderLayer (F.multi_head_attention_forward: q scaled by hd^-0.5,
// softmax over all S keys, no mask; mdm.py:79-86, 233).  qkv [B*S, 3D] (q | k | v), out [B*S, D].
// One CTA per (clip, head); K padded to hd+1 in shared memory (lane j reads row j: conflict-free).
// ---------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
struct SelfAttnArgs { const TIn* qkv; TOut* out; int S, D, heads; };

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) self_attention_kernel(const SelfAttnArgs<TIn, TOut> a) {
  extern __shared__ float smem[];
  const int hd = a.D / a.heads, ldk = hd + 1;
  const int nwarps = blockDim.x >> 5;
  const int Spad = (a.S + 31) & ~31;
  float* Ks = smem;                      // [S][hd+1]
  float* Vs = Ks + a.S * ldk;            // [S][hd]
  float* qs = Vs + a.S * hd;             // [nwarps][hd]
  float* ps = qs + nwarps * hd;          // [nwarps][Spad]
  const int clip = blockIdx.x / a.heads, head = blockIdx.x - clip * a.heads;
  const TIn* base = a.qkv + (long long)clip * a.S * 3 * a.D + head * hd;
  for (int e = threadIdx.x; e < a.S * hd; e += blockDim.x) {
    const int j = e / hd, d = e - j * hd;
    Ks[j * ldk + d] = ldf(base + (long long)j * 3 * a.D + a.D + d);
    Vs[j * hd + d] = ldf(base + (long long)j * 3 * a.D + 2 * a.D + d);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float scale = rsqrtf((float)hd);
  float* q = qs + warp * hd;
  float* p = ps + warp * Spad;
  for (int i = warp; i < a.S; i += nwarps) {
    for (int d = lane; d < hd; d += 32) q[d] = ldf(base + (long long)i * 3 * a.D + d) * scale;
    __syncwarp();
    float mx = -3.402823466e+38f;
    for (int j = lane; j < a.S; j += 32) {
      const float* k = Ks + j * ldk;
      float s = 0.f;
      for (int d = 0; d < hd; ++d) s = fmaf(q[d], k[d], s);
      p[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < a.S; j += 32) { const float e = expf(p[j] - mx); p[j] = e; sum += e; }
    const float inv = 1.0f / warp_sum(sum);
    __syncwarp();
    TOut* o = a.out + ((long long)clip * a.S + i) * a.D + head * hd;
    for (int d = lane; d < hd; d += 32) {
      float acc = 0.f;
      for (int j = 0; j < a.S; ++j) acc = fmaf(p[j], Vs[j * hd + d], acc);
      stf(o + d, acc * inv);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// Posterior / add-noise step: ONE coalesced, float4-vectorised pass over [B, J*T]  (p_sample:
// gaussian_diffusion.py:264-271 q_posterior mean, :542-557 noise; ddim_sample :768-791).
// Reads x_t and x0 once, writes x_{t-1} once (3 x 4 B per element); noise is generated in-kernel from the
// counter-based stream, one Philox call per 4 elements.  Arithmetic mirrors the reference op order in fp32
// (separate mul / add roundings, no fma contraction).
// ---------------------------------------------------------------------------------------------------
struct PosteriorArgs {
  float* x; const float* x0;
  const float4* coef;            // [nsteps] per-index coefficients (see dsg_set_schedule)
  const long long* clip_ids;     // [B] device
  StepRef step; int index_imm;   // index_imm >= 0 overrides step (unit-test entry)
  int draw_imm;                  // draw number when index_imm >= 0
  int sampler, B; long long per_clip;  // per_clip = J*T (multiple of 4)
  uint32_t k0, k1, segment;
};

DSG_DEVINL float posterior_elem(int sampler, const float4 c, float x0, float xt, float z, bool nz) {
  if (sampler == 0) {
    float r = __fadd_rn(__fmul_rn(c.x, x0), __fmul_rn(c.y, xt));
    if (nz) r = __fadd_rn(r, __fmul_rn(c.z, z));
    return r;
  }
  const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(c.x, xt), x0), c.y);
  return __fadd_rn(__fmul_rn(x0, c.z), __fmul_rn(c.w, eps));
}

__global__ void __launch_bounds__(256) posterior_step_kernel(const PosteriorArgs a) {
  const int index = a.index_imm >= 0 ? a.index_imm : a.step.index();
  const uint32_t draw = a.index_imm >= 0 ? (uint32_t)a.draw_imm : (uint32_t)(1 + a.step.k());
  const float4 c = a.coef[index];
  const bool nz = (index != 0) && (a.sampler == 0);
  const long long quads_per_clip = a.per_clip >> 2;
  const long long total = quads_per_clip * a.B;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(g / quads_per_clip);
    const uint32_t q = (uint32_t)(g - (long long)b * quads_per_clip);
    const long long off = (long long)b * a.per_clip + 4ll * q;
    const float4 xt = *reinterpret_cast<const float4*>(a.x + off);
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(a.x0 + off));
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nz) z = philox_normal4(q, draw, (uint32_t)a.clip_ids[b], a.segment, a.k0, a.k1);
    float4 o;
    o.x = posterior_elem(a.sampler, c, x0.x, xt.x, z.x, nz);
    o.y = posterior_elem(a.sampler, c, x0.y, xt.y, z.y, nz);
    o.z = posterior_elem(a.sampler, c, x0.z, xt.z, z.z, nz);
    o.w = posterior_elem(a.sampler, c, x0.w, xt.w, z.w, nz);
    *reinterpret_cast<float4*>(a.x + off) = o;
  }
}

// PLMS update (plms_sample, gaussian_diffusion.py:1005-1103), fp32 with the reference's op order.  Coefficient rows are the
// DDIM ones: c = { sqrt_recip_abar, sqrt_recipm1_abar, sqrt(abar_prev), sqrt(1 - abar_prev) } at the step's index.
//   mode 0  first step, stage A: eps = (c.x x - x0) / c.y -> hist[0];  tmp = x0 c.z + c.w eps          (:1061-1063)
//   mode 1  first step, stage B: eps2 from (tmp, x0b) with the row of index - 1; eps' = (eps + eps2) / 2 (:1064-1067)
//   mode 2  Adams-Bashforth of order `n` over hist[0] (newest, written here) .. hist[n-1]                 (:1068-1086)
// then x <- index != 0 ? mean_pred : x0                                                                   (:1091-1093)
struct PlmsArgs {
  float* x; const float* x0; const float* x0b; float* hist[4]; float* tmp;
  const float4* coef; int index, mode, n; long long total4;
};
DSG_DEVINL float plms_eps(const float4 c, float xt, float x0) { return __fdiv_rn(__fsub_rn(__fmul_rn(c.x, xt), x0), c.y); }
DSG_DEVINL float plms_out(const float4 c, float xt, float x0, float ep, bool nz) {
  const float pred = __fsub_rn(__fmul_rn(c.x, xt), __fmul_rn(c.y, ep));
  const float mean = __fadd_rn(__fmul_rn(pred, c.z), __fmul_rn(c.w, ep));
  return nz ? mean : x0;
}
__global__ void __launch_bounds__(256) plms_update_kernel(const PlmsArgs a) {
  const float4 c = a.coef[a.index];
  const float4 cp = a.coef[a.index > 0 ? a.index - 1 : 0];
  const bool nz = a.index != 0;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < a.total4; g += (long long)gridDim.x * blockDim.x) {
    const float4 xt4 = reinterpret_cast<const float4*>(a.x)[g];
    const float4 x04 = reinterpret_cast<const float4*>(a.x0)[g];
    const float xt[4] = {xt4.x, xt4.y, xt4.z, xt4.w}, x0[4] = {x04.x, x04.y, x04.z, x04.w};
    float o[4], e[4];
    if (a.mode == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { e[i] = plms_eps(c, xt[i], x0[i]); o[i] = __fadd_rn(__fmul_rn(x0[i], c.z), __fmul_rn(c.w, e[i])); }
      reinterpret_cast<float4*>(a.hist[0])[g] = make_float4(e[0], e[1], e[2], e[3]);
      reinterpret_cast<float4*>(a.tmp)[g] = make_float4(o[0], o[1], o[2], o[3]);
      continue;
    }
    if (a.mode == 1) {
      const float4 t4 = reinterpret_cast<const float4*>(a.tmp)[g], b4 = reinterpret_cast<const float4*>(a.x0b)[g];
      const float4 h4 = reinterpret_cast<const float4*>(a.hist[0])[g];
      const float tm[4] = {t4.x, t4.y, t4.z, t4.w}, xb[4] = {b4.x, b4.y, b4.z, b4.w}, h[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float e2 = plms_eps(cp, tm[i], xb[i]);
        o[i] = plms_out(c, xt[i], x0[i], __fdiv_rn(__fadd_rn(h[i], e2), 2.0f), nz);
      }
    } else {
      float h1[4] = {0, 0, 0, 0}, h2[4] = {0, 0, 0, 0}, h3[4] = {0, 0, 0, 0};
      if (a.n > 1) { const float4 t = reinterpret_cast<const float4*>(a.hist[1])[g]; h1[0] = t.x; h1[1] = t.y; h1[2] = t.z; h1[3] = t.w; }
      if (a.n > 2) { const float4 t = reinterpret_cast<const float4*>(a.hist[2])[g]; h2[0] = t.x; h2[1] = t.y; h2[2] = t.z; h2[3] = t.w; }
      if (a.n > 3) { const float4 t = reinterpret_cast<const float4*>(a.hist[3])[g]; h3[0] = t.x; h3[1] = t.y; h3[2] = t.z; h3[3] = t.w; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        e[i] = plms_eps(c, xt[i], x0[i]);
        float ep;
        if (a.n == 1) ep = e[i];
        else if (a.n == 2) ep = __fdiv_rn(__fsub_rn(__fmul_rn(3.0f, e[i]), h1[i]), 2.0f);
        else if (a.n == 3) ep = __fdiv_rn(__fadd_rn(__fsub_rn(__fmul_rn(23.0f, e[i]), __fmul_rn(16.0f, h1[i])), __fmul_rn(5.0f, h2[i])), 12.0f);
        else ep = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(__fmul_rn(55.0f, e[i]), __fmul_rn(59.0f, h1[i])), __fmul_rn(37.0f, h2[i])),
                                     __fmul_rn(9.0f, h3[i])), 24.0f);
        o[i] = plms_out(c, xt[i], x0[i], ep, nz);
      }
      reinterpret_cast<float4*>(a.hist[0])[g] = make_float4(e[0], e[1], e[2], e[3]);
    }
    reinterpret_cast<float4*>(a.x)[g] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// x_T ~ N(0, I): draw 0 of the stream (th.randn(*shape), gaussian_diffusion.py:704), optionally followed by
// q_sample(init_image, t0, noise) (:236-254, 706-713): x = sa * init + sb * noise.
__global__ void __launch_bounds__(256) init_noise_kernel(float* x, const float* init, int has_noise, float sa, float sb,
                                                        const long long* clip_ids, int B, long long per_clip,
                                                        uint32_t k0, uint32_t k1, uint32_t segment) {
  const long long quads_per_clip = per_clip >> 2, total = quads_per_clip * B;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(g / quads_per_clip);
    const uint32_t q = (uint32_t)(g - (long long)b * quads_per_clip);
    const long long off = (long long)b * per_clip + 4ll * q;
    float4 z;
    if (has_noise) z = *reinterpret_cast<const float4*>(x + off);
    else z = philox_normal4(q, 0u, (uint32_t)clip_ids[b], segment, k0, k1);
    if (init) {
      const float4 i4 = *reinterpret_cast<const float4*>(init + off);
      z.x = __fadd_rn(__fmul_rn(sa, i4.x), __fmul_rn(sb, z.x));
      z.y = __fadd_rn(__fmul_rn(sa, i4.y), __fmul_rn(sb, z.y));
      z.z = __fadd_rn(__fmul_rn(sa, i4.z), __fmul_rn(sb, z.z));
      z.w = __fadd_rn(__fmul_rn(sa, i4.w), __fmul_rn(sb, z.w));
    }
    *reinterpret_cast<float4*>(x + off) = z;
  }
}

// Segment hand-off (sample.py:266-288), batched over clips with the reference's n == 1 semantics:
// root-position shift of channels 0..2 by (sample[:,c,0] - tail[:,c,0]), then frame 0 <- (tail0 + sample0)/2.
__global__ void __launch_bounds__(256) stitch_segment_kernel(const float* tail, float* sample, int J, int T, int n_seed,
                                                            int smoothing) {
  const int b = blockIdx.x;
  const float* tb = tail + (long long)b * J * n_seed;
  float* sb = sample + (long long)b * J * T;
  __shared__ float delta[3];
  if (threadIdx.x < 3) delta[threadIdx.x] = smoothing ? __fsub_rn(sb[threadIdx.x * T], tb[threadIdx.x * n_seed]) : 0.f;
  __syncthreads();
  if (smoothing)
    for (int e = threadIdx.x; e < 3 * T; e += blockDim.x) sb[e] = __fsub_rn(sb[e], delta[e / T]);
  __syncthreads();
  for (int j = threadIdx.x; j < J; j += blockDim.x)
    sb[(long long)j * T] = __fadd_rn(__fmul_rn(tb[(long long)j * n_seed], 0.5f), __fmul_rn(sb[(long long)j * T], 0.5f));
}

__global__ void bump_counter_kernel(int* k) { *k += 1; }
