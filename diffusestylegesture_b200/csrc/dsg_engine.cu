// libdsg.so — engine state and the C ABI declared in include/dsg.h.
// fp32 validation path here; the tcgen05 tensor-core path is in dsg_tc.cu (linked into the same library).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dsg.h"
#include "dsg_engine.h"
#include "dsg_kernels_f32.cuh"

thread_local std::string g_last_error;

int dsg_fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// --------------------------------------------------------------------------------------------------
// small helpers
// --------------------------------------------------------------------------------------------------
static bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static int stage_reserve(Stage& s, size_t bytes) {
  if (s.cap >= bytes) return DSG_OK;
  if (s.p) cudaFree(s.p);
  s.p = nullptr; s.cap = 0;
  CUDA_TRY(cudaMalloc(&s.p, bytes));
  s.cap = bytes;
  return DSG_OK;
}

// Device view of an input buffer: in place when it already lives on the device, else staged.
static int dev_in(dsg_engine* e, int slot, const void* src, size_t bytes, cudaStream_t st, const void** out) {
  if (is_device_ptr(src)) { *out = src; return DSG_OK; }
  int rc = stage_reserve(e->stage[slot], bytes);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->stage[slot].p, src, bytes, cudaMemcpyHostToDevice, st));
  *out = e->stage[slot].p;
  return DSG_OK;
}

int dsg_prof_begin(dsg_engine* e, int tag, cudaStream_t st) {
  if (e->prof_used == e->prof_spans.size()) {
    ProfSpan sp; sp.tag = tag;
    cudaEventCreate(&sp.a); cudaEventCreate(&sp.b);
    e->prof_spans.push_back(sp);
  }
  ProfSpan& sp = e->prof_spans[e->prof_used];
  sp.tag = tag;
  cudaEventRecord(sp.a, st);
  return 0;
}
void dsg_prof_end(dsg_engine* e, cudaStream_t st) { cudaEventRecord(e->prof_spans[e->prof_used++].b, st); }

static void prof_collect(dsg_engine* e) {
  cudaDeviceSynchronize();
  for (size_t i = 0; i < e->prof_used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->prof_spans[i].a, e->prof_spans[i].b) == cudaSuccess) {
      e->prof_ms[e->prof_spans[i].tag] += ms;
      e->prof_n[e->prof_spans[i].tag] += 1;
    }
  }
  e->prof_used = 0;
}

template <typename T>
static int dalloc(T** p, size_t n) {
  CUDA_TRY(cudaMalloc((void**)p, n * sizeof(T)));
  return DSG_OK;
}

static RowMap plain_rows(int M, long long ld) { return RowMap{M > 0 ? M : 1, 0, 0, ld}; }

int launch_gemm_f32(dsg_engine* e, const GemmF32Args& g, bool a_m_contig, bool swap_mn, cudaStream_t st) {
  dim3 grid((g.M + 63) / 64, (g.N + 63) / 64);
  if (a_m_contig && !swap_mn) gemm_f32_kernel<true, false><<<grid, 256, 0, st>>>(g);
  else if (!a_m_contig && swap_mn) gemm_f32_kernel<false, true><<<grid, 256, 0, st>>>(g);
  else if (a_m_contig && swap_mn) gemm_f32_kernel<true, true><<<grid, 256, 0, st>>>(g);
  else gemm_f32_kernel<false, false><<<grid, 256, 0, st>>>(g);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

static GemmF32Args gemm_plain(const float* A, long long lda, const float* B, long long ldb, float* C, long long ldc,
                              int M, int N, int K, const float* bias, int act) {
  GemmF32Args g;
  memset(&g, 0, sizeof g);
  g.A = A; g.am = plain_rows(M, lda); g.a_kstride = 1;
  g.B = B; g.b_nstride = ldb; g.b_kstride = 1;
  g.C = C; g.cm = plain_rows(M, ldc); g.c_nstride = 1;
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.act = act;
  g.step = StepRef{nullptr, 0, 0};
  return g;
}

int launch_layernorm(dsg_engine* e, const float* in, float* out, const float* gamma, const float* beta, int rows,
                     int D, cudaStream_t st) {
  const int blocks = (rows * 32 + 255) / 256;
  if (D <= 256) layernorm_rows_kernel<8><<<blocks, 256, 0, st>>>(in, out, gamma, beta, rows, D);
  else if (D <= 512) layernorm_rows_kernel<16><<<blocks, 256, 0, st>>>(in, out, gamma, beta, rows, D);
  else return dsg_fail(DSG_ERR_BAD_SHAPE, "latent_dim %d > 512 unsupported", D);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

// --------------------------------------------------------------------------------------------------
// create / destroy
// --------------------------------------------------------------------------------------------------
static std::vector<size_t> weight_sizes(const dsg_model_desc& d) {
  const size_t D = d.latent_dim, F = d.ff_size, J = d.njoints, A = d.audio_latent;
  std::vector<size_t> s = {A * d.audio_dim, A, D * J, D, D * (2 * D + A), D, D * D, D, D * D, D,
                           (size_t)d.style_latent * d.style_in, (size_t)d.style_latent};
  if (d.variant == DSG_VARIANT_ATTN3) { s.push_back((D - d.style_latent) * J * d.n_seed); s.push_back(D - d.style_latent); }
  else { s.push_back(A * J); s.push_back(A); }
  s.push_back(J * D); s.push_back(J);
  for (int l = 0; l < d.num_layers; ++l) {
    const size_t per[12] = {3 * D * D, 3 * D, D * D, D, F * D, F, D * F, D, D, D, D, D};
    for (size_t v : per) s.push_back(v);
  }
  if (d.variant == DSG_VARIANT_ATTN5) { s.push_back(A * J); s.push_back(A); }     // embed_text_last (BEAT-TWH-main/model/mdm.py:95)
  return s;
}

static int validate_desc(const dsg_model_desc& d) {
  if (d.variant != DSG_VARIANT_ATTN3 && d.variant != DSG_VARIANT_ATTN4 && d.variant != DSG_VARIANT_ATTN5)
    return dsg_fail(DSG_ERR_UNSUPPORTED, "variant %d: cross_local_attention3 (3), 4 and 5 are implemented", d.variant);
  if (d.njoints <= 0 || d.n_poses <= 0 || d.max_batch <= 0 || d.num_layers <= 0 || d.num_timesteps <= 0)
    return dsg_fail(DSG_ERR_BAD_SHAPE, "non-positive size in descriptor");
  if (d.latent_dim % d.local_heads || d.latent_dim % d.num_heads || (d.latent_dim / d.local_heads) % 2)
    return dsg_fail(DSG_ERR_BAD_SHAPE, "latent_dim %d not divisible by heads", d.latent_dim);
  if (d.n_poses % d.local_window)
    return dsg_fail(DSG_ERR_BAD_SHAPE, "n_poses %d must be a multiple of the local window %d "
                    "(the reference only prints and then fails in einops: local_attention.py:114-126)", d.n_poses, d.local_window);
  if (2 * d.local_window > 32) return dsg_fail(DSG_ERR_BAD_SHAPE, "local window %d > 16 unsupported", d.local_window);
  if (((long long)d.njoints * d.n_poses) % 4) return dsg_fail(DSG_ERR_BAD_SHAPE, "njoints*n_poses must be a multiple of 4");
  if (d.latent_dim > 512 || d.latent_dim % 32) return dsg_fail(DSG_ERR_BAD_SHAPE, "latent_dim must be a multiple of 32, <= 512");
  if (d.variant == DSG_VARIANT_ATTN3 && d.style_latent >= d.latent_dim) return dsg_fail(DSG_ERR_BAD_SHAPE, "style_latent");
  if (d.variant != DSG_VARIANT_ATTN3 && d.style_latent != d.latent_dim) return dsg_fail(DSG_ERR_BAD_SHAPE, "style_latent must equal latent_dim for attn4 / attn5");
  if (d.n_seed <= 0 || d.n_seed >= d.n_poses) return dsg_fail(DSG_ERR_BAD_SHAPE, "n_seed");
  if (d.variant == DSG_VARIANT_ATTN5 && 2 * d.n_seed >= d.n_poses) return dsg_fail(DSG_ERR_BAD_SHAPE, "attn5 needs n_poses > 2 n_seed");
  if (d.precision != DSG_PRECISION_FP32 && d.precision != DSG_PRECISION_BF16) return dsg_fail(DSG_ERR_BAD_SHAPE, "precision");
  return DSG_OK;
}

extern "C" int dsg_engine_create(const dsg_model_desc* desc, const float* const* weights, int32_t n_weights,
                                 const float* pe, dsg_engine** out) {
  if (!desc || !weights || !pe || !out) return dsg_fail(DSG_ERR_BAD_SHAPE, "null argument");
  int rc = validate_desc(*desc);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return dsg_fail(DSG_ERR_BAD_ARCH, "no CUDA device: libdsg has no CPU fallback");
  }
  if (desc->device < 0 || desc->device >= ndev) return dsg_fail(DSG_ERR_BAD_SHAPE, "device ordinal %d of %d", desc->device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, desc->device));
  if (prop.major != 10)
    return dsg_fail(DSG_ERR_BAD_ARCH, "device %d is sm_%d%d; libdsg is built for sm_100a only", desc->device, prop.major, prop.minor);
  DeviceGuard guard(desc->device);

  const std::vector<size_t> sizes = weight_sizes(*desc);
  if ((int)sizes.size() != n_weights)
    return dsg_fail(DSG_ERR_BAD_SHAPE, "expected %d weight tensors, got %d", (int)sizes.size(), n_weights);

  dsg_engine* e = new dsg_engine();
  e->d = *desc;
  e->S = desc->n_poses + 1;
  e->num_sms = prop.multiProcessorCount;
  const int D = desc->latent_dim, J = desc->njoints, T = desc->n_poses, A = desc->audio_latent, F = desc->ff_size;
  const int S = e->S, NT = desc->num_timesteps, MB = desc->max_batch;
  cudaStream_t st = 0;

  // ---- weights: one slab, fp32
  size_t total = 0;
  for (size_t s : sizes) total += (s + 3) & ~size_t(3);
  if ((rc = dalloc(&e->wslab, total))) { delete e; return rc; }
  e->w.resize(sizes.size());
  size_t off = 0;
  for (size_t i = 0; i < sizes.size(); ++i) {
    e->w[i] = e->wslab + off;
    if (cudaMemcpy(e->w[i], weights[i], sizes[i] * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
      rc = dsg_fail(DSG_ERR_CUDA, "copy of weight %d failed: %s", (int)i, cudaGetErrorString(cudaGetLastError()));
      dsg_engine_destroy(e);
      return rc;
    }
    off += (sizes[i] + 3) & ~size_t(3);
  }
#define TRY_CREATE(x) do { int rc__ = (x); if (rc__) { dsg_engine_destroy(e); return rc__; } } while (0)
  // ---- derived tables
  TRY_CREATE(dalloc(&e->pe_d, (size_t)NT * D));
  TRY_CREATE(dalloc(&e->l1, (size_t)NT * D));
  float* pe_d = e->pe_d; float* l1 = e->l1;
  if (cudaMemcpy(pe_d, pe, (size_t)NT * D * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
    rc = dsg_fail(DSG_ERR_CUDA, "copy of pe failed"); dsg_engine_destroy(e); return rc; }
  TRY_CREATE(dalloc(&e->te, (size_t)NT * D));
  TRY_CREATE(dalloc(&e->TW, (size_t)NT * D));
  TRY_CREATE(dalloc(&e->Wxp, (size_t)D * J));
  TRY_CREATE(dalloc(&e->bxp, (size_t)D));
  const int ld2 = 2 * D + A;
  // te = Linear(SiLU(Linear(pe)))   (TimestepEmbedder, mdm.py:441-448) for every original timestep
  TRY_CREATE(launch_gemm_f32(e, gemm_plain(pe_d, D, e->w[W_T0_W], D, l1, D, NT, D, D, e->w[W_T0_B], 2), false, false, st));
  TRY_CREATE(launch_gemm_f32(e, gemm_plain(l1, D, e->w[W_T2_W], D, e->te, D, NT, D, D, e->w[W_T2_B], 0), false, false, st));
  // TW[t] = W_tok * te[t]          (token column block of input_process2, mdm.py:204-206)
  TRY_CREATE(launch_gemm_f32(e, gemm_plain(e->te, D, e->w[W_IN2_W], ld2, e->TW, D, NT, D, D, nullptr, 0), false, false, st));
  // Wxp = W_x * W_pose [D,J]; bxp = W_x * b_pose + b_2   (poseEmbedding folded into input_process2's x block)
  {
    GemmF32Args g = gemm_plain(e->w[W_IN2_W] + D, ld2, e->w[W_POSE_W], 1, e->Wxp, J, D, J, D, nullptr, 0);
    g.b_nstride = 1; g.b_kstride = J;
    TRY_CREATE(launch_gemm_f32(e, g, false, false, st));
    GemmF32Args g2 = gemm_plain(e->w[W_IN2_W] + D, ld2, e->w[W_POSE_B], D, e->bxp, 1, D, 1, D, nullptr, 0);
    g2.addmat = e->w[W_IN2_B]; g2.addmat_ld = 1;
    TRY_CREATE(launch_gemm_f32(e, g2, false, false, st));
  }
  // rotary tables (SinusoidalEmbeddings, rotary.py:6-16): fp32 like the reference
  {
    const int hd = D / desc->local_heads, half = hd / 2;
    std::vector<float2> cs((size_t)S * half);
    for (int p = 0; p < S; ++p)
      for (int i = 0; i < half; ++i) {
        const float inv_freq = 1.0f / powf(10000.0f, (float)(2 * i) / (float)hd);
        const float ang = (float)p * inv_freq;
        cs[(size_t)p * half + i] = make_float2(cosf(ang), sinf(ang));
      }
    TRY_CREATE(dalloc(&e->cs_local, cs.size()));
    cudaMemcpy(e->cs_local, cs.data(), cs.size() * sizeof(float2), cudaMemcpyHostToDevice);
  }
  // ---- conditioning + workspace
  TRY_CREATE(dalloc(&e->emb1, (size_t)MB * D));
  TRY_CREATE(dalloc(&e->cvec, (size_t)MB * D));
  TRY_CREATE(dalloc(&e->enc, (size_t)MB * T * A));
  TRY_CREATE(dalloc(&e->cond, (size_t)MB * T * D));
  TRY_CREATE(dalloc(&e->h, (size_t)MB * T * D));
  TRY_CREATE(dalloc(&e->xs, (size_t)MB * S * D));
  TRY_CREATE(dalloc(&e->qkv, (size_t)MB * S * 3 * D));
  TRY_CREATE(dalloc(&e->att, (size_t)MB * S * D));
  TRY_CREATE(dalloc(&e->ff, (size_t)MB * S * F));
  TRY_CREATE(dalloc(&e->tmp, (size_t)MB * S * D));
  TRY_CREATE(dalloc(&e->x0, (size_t)MB * J * T));
  TRY_CREATE(dalloc(&e->tsel, (size_t)MB));
  TRY_CREATE(dalloc(&e->clip_ids, (size_t)MB));
  TRY_CREATE(dalloc(&e->noise_ids, (size_t)MB));
  TRY_CREATE(dalloc(&e->d_k, (size_t)1));
  TRY_CREATE(dalloc(&e->d_loop, (size_t)1));
  // opt-in shared memory for the attention kernels
  {
    const int hdg = D / desc->num_heads;
    e->smem_self = (size_t)(S * (hdg + 1) + S * hdg + 8 * hdg + 8 * ((S + 31) & ~31)) * sizeof(float);
    const int hdl = D / desc->local_heads;
    e->local_threads = 32 * ((T + 7) / 8 < 16 ? (T + 7) / 8 : 16);        // ~8 query frames per warp
    e->smem_local = (size_t)(T * (hdl + 1) + (e->local_threads / 32) * hdl) * sizeof(float);
    if (e->smem_self > 227 * 1024 || e->smem_local > 227 * 1024) {
      rc = dsg_fail(DSG_ERR_BAD_SHAPE, "sequence too long for the shared-memory attention kernels"); dsg_engine_destroy(e); return rc; }
    cudaFuncSetAttribute(self_attention_kernel<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_self);
    cudaFuncSetAttribute(self_attention_kernel<__nv_bfloat16, __nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_self);
    cudaFuncSetAttribute(local_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_local);
  }
  if (desc->precision == DSG_PRECISION_BF16) TRY_CREATE(dsg_tc_create(e));
  if (cudaDeviceSynchronize() != cudaSuccess) {
    rc = dsg_fail(DSG_ERR_CUDA, "engine setup kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
    dsg_engine_destroy(e); return rc; }
  cudaFree(e->pe_d); cudaFree(e->l1);
  e->pe_d = nullptr; e->l1 = nullptr;
  *out = e;
  return DSG_OK;
}

extern "C" void dsg_engine_destroy(dsg_engine* e) {
  if (!e) return;
  DeviceGuard guard(e->d.device);
  dsg_tc_destroy(e);
  for (float* p : e->plms_buf) if (p) cudaFree(p);
  void* ptrs[] = {e->pe_d, e->l1, e->noise_ids, e->wslab, e->te, e->TW, e->Wxp, e->bxp, e->cs_local, e->emb1, e->cvec, e->enc, e->cond, e->h, e->xs,
                  e->qkv, e->att, e->ff, e->tmp, e->x0, e->tsel, e->clip_ids, e->d_k, e->d_loop, e->coef, e->tmap, e->dbg};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (Stage& s : e->stage) if (s.p) cudaFree(s.p);
  for (ProfSpan& sp : e->prof_spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
  delete e;
}

// --------------------------------------------------------------------------------------------------
// schedule
// --------------------------------------------------------------------------------------------------
extern "C" int dsg_set_schedule(dsg_engine* e, int32_t sampler, int32_t nsteps, const float* coef, const float* qsample,
                                const int32_t* timestep_map) {
  if (!e || !coef || !qsample || !timestep_map) return dsg_fail(DSG_ERR_BAD_SHAPE, "null argument");
  if (sampler != DSG_SAMPLER_DDPM && sampler != DSG_SAMPLER_DDIM && sampler != DSG_SAMPLER_PLMS)
    return dsg_fail(DSG_ERR_UNSUPPORTED, "sampler %d", sampler);
  if (nsteps <= 0 || nsteps > e->d.num_timesteps) return dsg_fail(DSG_ERR_BAD_SHAPE, "nsteps %d", nsteps);
  for (int i = 0; i < nsteps; ++i)
    if (timestep_map[i] < 0 || timestep_map[i] >= e->d.num_timesteps)
      return dsg_fail(DSG_ERR_BAD_SHAPE, "timestep_map[%d] = %d out of range", i, timestep_map[i]);
  DeviceGuard guard(e->d.device);
  if (e->coef) cudaFree(e->coef);
  if (e->tmap) cudaFree(e->tmap);
  e->coef = nullptr; e->tmap = nullptr;
  int rc;
  if ((rc = dalloc(&e->coef, (size_t)nsteps))) return rc;
  if ((rc = dalloc(&e->tmap, (size_t)nsteps))) return rc;
  CUDA_TRY(cudaMemcpy(e->coef, coef, (size_t)nsteps * sizeof(float4), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(e->tmap, timestep_map, (size_t)nsteps * sizeof(int), cudaMemcpyHostToDevice));
  e->qsample.assign(qsample, qsample + 2 * (size_t)nsteps);
  e->sampler = sampler; e->nsteps = nsteps;
  e->graph_valid = false;
  return DSG_OK;
}

// --------------------------------------------------------------------------------------------------
// conditioning (step-invariant part of MDM.forward)
// --------------------------------------------------------------------------------------------------
extern "C" int dsg_set_conditioning(dsg_engine* e, int32_t B, const float* style, const float* seed, const float* audio,
                                    void* stream) {
  return dsg_set_conditioning_ex(e, B, style, seed, audio, nullptr, stream);
}

extern "C" int dsg_set_conditioning_ex(dsg_engine* e, int32_t B, const float* style, const float* seed, const float* audio,
                                       const float* seed_last, void* stream) {
  if (!e || !style || !seed || !audio) return dsg_fail(DSG_ERR_BAD_SHAPE, "null argument");
  if ((e->d.variant == DSG_VARIANT_ATTN5) != (seed_last != nullptr))
    return dsg_fail(DSG_ERR_BAD_SHAPE, "seed_last is required for (and only for) the cross_local_attention5 variant");
  if (B <= 0 || B > e->d.max_batch) return dsg_fail(DSG_ERR_BAD_SHAPE, "batch %d outside 1..%d", B, e->d.max_batch);
  DeviceGuard guard(e->d.device);
  cudaStream_t st = (cudaStream_t)stream;
  const dsg_model_desc& d = e->d;
  const int D = d.latent_dim, J = d.njoints, T = d.n_poses, A = d.audio_latent, NS = d.n_seed;
  const int ld2 = 2 * D + A;
  const int Ta = d.variant == DSG_VARIANT_ATTN3 ? T : (d.variant == DSG_VARIANT_ATTN4 ? T - NS : T - 2 * NS);
  const void *sty_d, *seed_d, *aud_d;
  int rc;
  if ((rc = dev_in(e, 0, style, (size_t)B * d.style_in * sizeof(float), st, &sty_d))) return rc;
  if ((rc = dev_in(e, 1, seed, (size_t)B * J * NS * sizeof(float), st, &seed_d))) return rc;
  if ((rc = dev_in(e, 2, audio, (size_t)B * Ta * d.audio_dim * sizeof(float), st, &aud_d))) return rc;
  const float* sty = (const float*)sty_d; const float* sd = (const float*)seed_d; const float* au = (const float*)aud_d;

  // embed_style -> emb1[:, :style_latent]            (mdm.py:180)
  TRY(launch_gemm_f32(e, gemm_plain(sty, d.style_in, e->w[W_STY_W], d.style_in, e->emb1, D, B, d.style_latent, d.style_in,
                                    e->w[W_STY_B], 0), false, false, st));
  if (d.variant == DSG_VARIANT_ATTN3) {
    // embed_text(seed.reshape(B, J*n_seed)) -> emb1[:, style_latent:]   (mdm.py:182-183)
    TRY(launch_gemm_f32(e, gemm_plain(sd, (long long)J * NS, e->w[W_TXT_W], (long long)J * NS, e->emb1 + d.style_latent, D, B,
                                      D - d.style_latent, J * NS, e->w[W_TXT_B], 0), false, false, st));
    // WavEncoder -> enc [B,T,A]                        (mdm.py:190, 550-552)
    TRY(launch_gemm_f32(e, gemm_plain(au, d.audio_dim, e->w[W_AUD_W], d.audio_dim, e->enc, A, B * T, A, d.audio_dim,
                                      e->w[W_AUD_B], 0), false, false, st));
  } else {
    // embed_text on each seed frame -> enc[:, :n_seed]  (BEAT-TWH-main/model/mdm.py:188): A = seed^T per clip
    GemmF32Args g = gemm_plain(sd, 0, e->w[W_TXT_W], J, e->enc, A, B * NS, A, J, e->w[W_TXT_B], 0);
    g.am = RowMap{NS, 0, (long long)J * NS, 1}; g.a_kstride = NS;
    g.cm = RowMap{NS, 0, (long long)T * A, A};
    TRY(launch_gemm_f32(e, g, true, false, st));
    // WavEncoder -> enc[:, n_seed:]                      (BEAT-TWH-main/model/mdm.py:189-190)
    GemmF32Args g2 = gemm_plain(au, d.audio_dim, e->w[W_AUD_W], d.audio_dim, e->enc, A, B * Ta, A, d.audio_dim, e->w[W_AUD_B], 0);
    g2.cm = RowMap{Ta, NS, (long long)T * A, A};
    TRY(launch_gemm_f32(e, g2, false, false, st));
    if (d.variant == DSG_VARIANT_ATTN5) {
      // embed_text_last on each frame of y['seed_last'] -> enc[:, n_seed + Ta:]   (BEAT-TWH-main/model/mdm.py:229-230)
      const void* last_d;
      if ((rc = dev_in(e, 6, seed_last, (size_t)B * J * NS * sizeof(float), st, &last_d))) return rc;
      const size_t wl = e->w.size() - 2;
      GemmF32Args g3 = gemm_plain((const float*)last_d, 0, e->w[wl], J, e->enc, A, B * NS, A, J, e->w[wl + 1], 0);
      g3.am = RowMap{NS, 0, (long long)J * NS, 1}; g3.a_kstride = NS;
      g3.cm = RowMap{NS, NS + Ta, (long long)T * A, A};
      TRY(launch_gemm_f32(e, g3, true, false, st));
    }
  }
  // cvec[b] = W_tok * emb1[b] + (W_x b_pose + b_2)
  TRY(launch_gemm_f32(e, gemm_plain(e->emb1, D, e->w[W_IN2_W], ld2, e->cvec, D, B, D, D, e->bxp, 0), false, false, st));
  // cond[b,f] = W_a * enc[b,f] + cvec[b]
  {
    GemmF32Args g = gemm_plain(e->enc, A, e->w[W_IN2_W] + 2 * D, ld2, e->cond, D, B * T, D, A, nullptr, 0);
    g.cm = RowMap{T, 0, (long long)T * D, D};
    g.clipvec = e->cvec; g.clipvec_ld = D;
    TRY(launch_gemm_f32(e, g, false, false, st));
  }
  e->cond_batch = B;
  return DSG_OK;
}

// --------------------------------------------------------------------------------------------------
// one denoiser call (fp32 path)
// --------------------------------------------------------------------------------------------------
static int debug_snap(dsg_engine* e, int slot, int B, cudaStream_t st) {
  if (!e->debug) return DSG_OK;
  const size_t n = (size_t)B * e->S * e->d.latent_dim;
  CUDA_TRY(cudaMemcpyAsync(e->dbg + (size_t)slot * e->d.max_batch * e->S * e->d.latent_dim, e->xs, n * sizeof(float),
                           cudaMemcpyDeviceToDevice, st));
  return DSG_OK;
}

int launch_local_attention(dsg_engine* e, int B, const float* h, long long h_clip_stride, int h_row0, float* xs,
                           __nv_bfloat16* xsb, const int* tsel, StepRef step, cudaStream_t st) {
  LocalAttnArgs a;
  a.h = h; a.h_clip_stride = h_clip_stride; a.h_row0 = h_row0; a.xs = xs; a.xsb = xsb; a.emb1 = e->emb1; a.te = e->te; a.tsel = tsel; a.tmap = e->tmap; a.step = step;
  a.cs = e->cs_local; a.T = e->d.n_poses; a.D = e->d.latent_dim; a.heads = e->d.local_heads; a.window = e->d.local_window;
  local_attention_kernel<<<B * e->d.local_heads, e->local_threads, e->smem_local, st>>>(a);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

int launch_self_attention_bf16(dsg_engine* e, int B, const __nv_bfloat16* qkv, __nv_bfloat16* out, cudaStream_t st) {
  SelfAttnArgs<__nv_bfloat16, __nv_bfloat16> sa{qkv, out, e->S, e->d.latent_dim, e->d.num_heads};
  self_attention_kernel<__nv_bfloat16, __nv_bfloat16><<<B * e->d.num_heads, 256, e->smem_self, st>>>(sa);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

int launch_self_attention(dsg_engine* e, int B, const float* qkv, float* out, cudaStream_t st) {
  SelfAttnArgs<float, float> sa{qkv, out, e->S, e->d.latent_dim, e->d.num_heads};
  self_attention_kernel<float, float><<<B * e->d.num_heads, 256, e->smem_self, st>>>(sa);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

static int denoise_f32(dsg_engine* e, int B, const float* x, const int* tsel, StepRef step, float* out, cudaStream_t st) {
  const dsg_model_desc& d = e->d;
  const int D = d.latent_dim, J = d.njoints, T = d.n_poses, F = d.ff_size, S = e->S;
  {  // h = Wxp x_t + cond + TW[t]     (InputProcess + input_process2, mdm.py:196-206)
    GemmF32Args g = gemm_plain(x, 0, e->Wxp, J, e->h, D, B * T, D, J, nullptr, 0);
    g.am = RowMap{T, 0, (long long)J * T, 1}; g.a_kstride = T;
    g.cm = RowMap{T, 0, (long long)T * D, D};
    g.addmat = e->cond; g.addmat_ld = D;
    g.tvec = e->TW; g.tvec_ld = D; g.tsel = tsel; g.tmap = e->tmap; g.step = step;
    PROF(e, PT_GEMM_IN, st, launch_gemm_f32(e, g, true, false, st));
  }
  PROF(e, PT_LOCAL_ATTN, st, launch_local_attention(e, B, e->h, (long long)T * D, 0, e->xs, nullptr, tsel, step, st));
  TRY(debug_snap(e, 0, B, st));
  const int M = B * S;
  for (int l = 0; l < d.num_layers; ++l) {
    float* const* w = &e->w[W_LAYER0 + 12 * l];
    PROF(e, PT_GEMM_QKV, st, launch_gemm_f32(e, gemm_plain(e->xs, D, w[L_INPROJ_W], D, e->qkv, 3 * D, M, 3 * D, D, w[L_INPROJ_B], 0), false, false, st));
    PROF(e, PT_SELF_ATTN, st, launch_self_attention(e, B, e->qkv, e->att, st));
    GemmF32Args go = gemm_plain(e->att, D, w[L_OUTPROJ_W], D, e->tmp, D, M, D, D, w[L_OUTPROJ_B], 0);
    go.addmat = e->xs; go.addmat_ld = D;
    PROF(e, PT_GEMM_OUTPROJ, st, launch_gemm_f32(e, go, false, false, st));
    PROF(e, PT_LAYERNORM, st, launch_layernorm(e, e->tmp, e->xs, w[L_N1_W], w[L_N1_B], M, D, st));
    PROF(e, PT_GEMM_FF1, st, launch_gemm_f32(e, gemm_plain(e->xs, D, w[L_FF1_W], D, e->ff, F, M, F, D, w[L_FF1_B], 1), false, false, st));
    GemmF32Args g2 = gemm_plain(e->ff, F, w[L_FF2_W], F, e->tmp, D, M, D, F, w[L_FF2_B], 0);
    g2.addmat = e->xs; g2.addmat_ld = D;
    PROF(e, PT_GEMM_FF2, st, launch_gemm_f32(e, g2, false, false, st));
    PROF(e, PT_LAYERNORM, st, launch_layernorm(e, e->tmp, e->xs, w[L_N2_W], w[L_N2_B], M, D, st));
    TRY(debug_snap(e, l + 1, B, st));
  }
  {  // x0[b,j,f] = W_out xs[b,f+1] + b_out    (OutputProcess, mdm.py:490-504)
    GemmF32Args g = gemm_plain(e->xs, 0, e->w[W_OUT_W], D, out, 0, B * T, J, D, e->w[W_OUT_B], 0);
    g.am = RowMap{T, 1, (long long)S * D, D};
    g.cm = RowMap{T, 0, (long long)J * T, 1}; g.c_nstride = T;
    PROF(e, PT_GEMM_HEAD, st, launch_gemm_f32(e, g, false, true, st));
  }
  return DSG_OK;
}

int dsg_denoise_step(dsg_engine* e, int B, const float* x, const int* tsel, StepRef step, float* out, cudaStream_t st) {
  if (e->d.precision == DSG_PRECISION_BF16) return dsg_tc_denoise(e, B, x, tsel, step, out, st);
  return denoise_f32(e, B, x, tsel, step, out, st);
}

extern "C" int dsg_denoise(dsg_engine* e, int32_t B, const float* x, const int32_t* timesteps, float* out, void* stream) {
  if (!e || !x || !timesteps || !out) return dsg_fail(DSG_ERR_BAD_SHAPE, "null argument");
  if (B <= 0 || B > e->d.max_batch) return dsg_fail(DSG_ERR_BAD_SHAPE, "batch %d outside 1..%d", B, e->d.max_batch);
  if (e->cond_batch != B) return dsg_fail(DSG_ERR_STATE, "dsg_set_conditioning was called for batch %d, not %d", e->cond_batch, B);
  for (int b = 0; b < B; ++b)
    if (timesteps[b] < 0 || timesteps[b] >= e->d.num_timesteps) return dsg_fail(DSG_ERR_BAD_SHAPE, "timestep %d out of range", timesteps[b]);
  DeviceGuard guard(e->d.device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)B * e->d.njoints * e->d.n_poses * sizeof(float);
  const void* xd;
  int rc;
  if ((rc = dev_in(e, 3, x, bytes, st, &xd))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e->tsel, timesteps, B * sizeof(int), cudaMemcpyHostToDevice, st));
  const bool out_dev = is_device_ptr(out);
  float* od = out_dev ? out : e->x0;
  if ((rc = dsg_denoise_step(e, B, (const float*)xd, e->tsel, StepRef{nullptr, 0, 0}, od, st))) return rc;
  if (!out_dev) {
    CUDA_TRY(cudaMemcpyAsync(out, od, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return DSG_OK;
}

// --------------------------------------------------------------------------------------------------
// posterior step, sampling loop, stitching
// --------------------------------------------------------------------------------------------------
int elementwise_grid(const dsg_engine* e, long long quads) {
  const long long want = (quads + 255) / 256;
  const long long cap = (long long)e->num_sms * 8;       // a multiple of the SM count, grid-stride inside
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

int launch_posterior(dsg_engine* e, int B, float* x, const float* x0, StepRef step, int index_imm, int draw_imm,
                     uint64_t seed, int segment, cudaStream_t st) {
  PosteriorArgs a;
  a.x = x; a.x0 = x0; a.coef = e->coef; a.clip_ids = e->noise_ids; a.step = step; a.index_imm = index_imm; a.draw_imm = draw_imm;
  a.sampler = e->sampler; a.B = B; a.per_clip = (long long)e->d.njoints * e->d.n_poses;
  a.k0 = (uint32_t)(seed & 0xffffffffu); a.k1 = (uint32_t)(seed >> 32); a.segment = (uint32_t)segment;
  posterior_step_kernel<<<elementwise_grid(e, (a.per_clip >> 2) * B), 256, 0, st>>>(a);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

// clip ids key the Philox counter as 32-bit words: ids outside [0, 2^32) would alias another clip's stream and are rejected.
// const_noise: the per-step draws of every clip use clip 0's key (noise[[0]].repeat(...), gaussian_diffusion.py:544-545).
static int upload_clip_ids(dsg_engine* e, int B, const int64_t* clip_ids, bool const_noise, cudaStream_t st) {
  e->h_clip_ids.resize(2 * (size_t)B);
  for (int b = 0; b < B; ++b) {
    const long long id = clip_ids ? (long long)clip_ids[b] : (long long)b;
    if (id < 0 || id > 0xFFFFFFFFll) return dsg_fail(DSG_ERR_BAD_SHAPE, "clip id %lld outside [0, 2^32)", id);
    e->h_clip_ids[b] = id;
  }
  for (int b = 0; b < B; ++b) e->h_clip_ids[B + b] = const_noise ? e->h_clip_ids[0] : e->h_clip_ids[b];
  CUDA_TRY(cudaMemcpyAsync(e->clip_ids, e->h_clip_ids.data(), B * sizeof(long long), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(e->noise_ids, e->h_clip_ids.data() + B, B * sizeof(long long), cudaMemcpyHostToDevice, st));
  return DSG_OK;
}

extern "C" int dsg_posterior_step(dsg_engine* e, int32_t B, float* x, const float* x0, int32_t index, uint64_t seed,
                                  const int64_t* clip_ids, int32_t segment, int32_t draw, void* stream) {
  if (!e || !x || !x0) return dsg_fail(DSG_ERR_BAD_SHAPE, "null argument");
  if (!e->coef) return dsg_fail(DSG_ERR_STATE, "dsg_set_schedule has not been called");
  if (B <= 0 || B > e->d.max_batch) return dsg_fail(DSG_ERR_BAD_SHAPE, "batch %d outside 1..%d", B, e->d.max_batch);
  if (index < 0 || index >= e->nsteps) return dsg_fail(DSG_ERR_BAD_SHAPE, "index %d outside 0..%d", index, e->nsteps - 1);
  DeviceGuard guard(e->d.device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t bytes = (size_t)B * e->d.njoints * e->d.n_poses * sizeof(float);
  int rc;
  if ((rc = upload_clip_ids(e, B, clip_ids, false, st))) return rc;
  const void* x0d;
  if ((rc = dev_in(e, 4, x0, bytes, st, &x0d))) return rc;
  const bool xdev = is_device_ptr(x);
  const void* xd = x;
  if (!xdev && (rc = dev_in(e, 3, x, bytes, st, &xd))) return rc;
  if ((rc = launch_posterior(e, B, (float*)xd, (const float*)x0d, StepRef{nullptr, 0, 0}, index, draw, seed, segment, st))) return rc;
  if (!xdev) {
    CUDA_TRY(cudaMemcpyAsync(x, xd, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return DSG_OK;
}

// PLMS (plms_sample_loop_progressive, gaussian_diffusion.py:1136-1200): one denoiser call per step (two on the first),
// the multistep combination in plms_update_kernel.  The denoiser runs through dsg_denoise_step (either precision).
static int plms_run(dsg_engine* e, int B, float* xd, int k0, int n_run, int first_index, int order, int* hist_len, uint64_t seed,
                    int segment, cudaStream_t st) {
  const long long per_clip = (long long)e->d.njoints * e->d.n_poses;
  const size_t cap = (size_t)e->d.max_batch * per_clip;
  for (float*& p : e->plms_buf) if (!p) TRY(dalloc(&p, cap));
  float* tmp = e->plms_buf[4]; float* x0b = e->plms_buf[5];
  const uint32_t key0 = (uint32_t)(seed & 0xffffffffu), key1 = (uint32_t)(seed >> 32);
  PlmsArgs a;
  a.x = xd; a.x0 = e->x0; a.x0b = x0b; a.tmp = tmp; a.coef = e->coef; a.total4 = (per_clip >> 2) * B;
  const int grid = elementwise_grid(e, a.total4);
  for (int k = k0; k < k0 + n_run; ++k) {
    const int index = first_index - k;
    const StepRef step{nullptr, k, first_index, key0, key1, (uint32_t)segment};
    TRY(dsg_denoise_step(e, B, xd, nullptr, step, e->x0, st));
    a.index = index;
    const int slot = k % order;                      // ring of `order` eps slots: newest at `slot`
    for (int i = 0; i < 4; ++i) a.hist[i] = e->plms_buf[((slot - i) % order + order) % order];
    if (*hist_len == 0) {                            // old_out is None: pseudo improved Euler (:1061-1067)
      if (index == 0) return dsg_fail(DSG_ERR_BAD_SHAPE, "plms: the first step needs a second model call at t - 1 (skip_timesteps leaves one step)");
      a.mode = 0; a.n = 1;
      plms_update_kernel<<<grid, 256, 0, st>>>(a);
      e->launches++;
      const StepRef step2{nullptr, k + 1, first_index, key0, key1, (uint32_t)segment};
      TRY(dsg_denoise_step(e, B, tmp, nullptr, step2, x0b, st));
      a.mode = 1;
      plms_update_kernel<<<grid, 256, 0, st>>>(a);
      e->launches++;
      *hist_len = 1;
    } else {                                         // Adams-Bashforth (:1068-1086), then pop(0) when len >= order (:1088-1089)
      const int len = *hist_len + 1;
      a.mode = 2; a.n = len < order ? len : order;
      plms_update_kernel<<<grid, 256, 0, st>>>(a);
      e->launches++;
      *hist_len = len >= order ? len - 1 : len;
    }
    CUDA_TRY(cudaGetLastError());
  }
  return DSG_OK;
}

extern "C" int dsg_sample_loop(dsg_engine* e, int32_t B, float* x, int32_t noise_given, uint64_t seed,
                               const int64_t* clip_ids, int32_t segment, int32_t skip_timesteps, const float* init_image,
                               void* stream) {
  return dsg_sample_loop_ex(e, B, x, noise_given, seed, clip_ids, segment, skip_timesteps, init_image, nullptr, stream);
}

extern "C" int dsg_sample_loop_ex(dsg_engine* e, int32_t B, float* x, int32_t noise_given, uint64_t seed,
                                  const int64_t* clip_ids, int32_t segment, int32_t skip_timesteps, const float* init_image,
                                  const dsg_loop_opts* opts, void* stream) {
  if (!e || !x) return dsg_fail(DSG_ERR_BAD_SHAPE, "null argument");
  if (!e->coef) return dsg_fail(DSG_ERR_STATE, "dsg_set_schedule has not been called");
  if (B <= 0 || B > e->d.max_batch) return dsg_fail(DSG_ERR_BAD_SHAPE, "batch %d outside 1..%d", B, e->d.max_batch);
  if (e->cond_batch != B) return dsg_fail(DSG_ERR_STATE, "dsg_set_conditioning was called for batch %d, not %d", e->cond_batch, B);
  if (skip_timesteps < 0 || skip_timesteps >= e->nsteps) return dsg_fail(DSG_ERR_BAD_SHAPE, "skip_timesteps %d", skip_timesteps);
  const int n_run = e->nsteps - skip_timesteps;
  const int first_index = n_run - 1;
  const bool const_noise = opts && (opts->flags & DSG_LOOP_CONST_NOISE);
  const int n_dump = opts ? opts->n_dump : 0;
  const int order = (opts && opts->plms_order) ? opts->plms_order : 2;
  if (opts && (opts->flags & ~DSG_LOOP_CONST_NOISE)) return dsg_fail(DSG_ERR_UNSUPPORTED, "unknown loop flags 0x%x", opts->flags);
  if (n_dump < 0 || (n_dump > 0 && (!opts->dump_iters || !opts->dump_out))) return dsg_fail(DSG_ERR_BAD_SHAPE, "dump_iters / dump_out");
  for (int i = 0; i < n_dump; ++i)
    if (opts->dump_iters[i] < 0 || opts->dump_iters[i] >= n_run || (i > 0 && opts->dump_iters[i] <= opts->dump_iters[i - 1]))
      return dsg_fail(DSG_ERR_BAD_SHAPE, "dump_iters must be ascending loop iterations in [0, %d)", n_run);
  if (e->sampler == DSG_SAMPLER_PLMS) {
    if (order < 2 || order > 4) return dsg_fail(DSG_ERR_BAD_SHAPE, "plms order %d: 2..4 (order 1 fails inside the reference, "
                                                "gaussian_diffusion.py:1069, and values outside 1..4 raise there)", order);
    if (const_noise) return dsg_fail(DSG_ERR_UNSUPPORTED, "plms_sample_loop has no const_noise option");
  } else if (e->sampler == DSG_SAMPLER_DDIM && (const_noise || n_dump)) {
    return dsg_fail(DSG_ERR_UNSUPPORTED, "ddim_sample_loop raises NotImplementedError for dump_steps / const_noise "
                    "(gaussian_diffusion.py:913-916)");
  }
  DeviceGuard guard(e->d.device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long per_clip = (long long)e->d.njoints * e->d.n_poses;
  const size_t bytes = (size_t)B * per_clip * sizeof(float);
  int rc;
  if ((rc = upload_clip_ids(e, B, clip_ids, const_noise, st))) return rc;
  const bool xdev = is_device_ptr(x);
  float* xd = x;
  if (!xdev) {
    if ((rc = stage_reserve(e->stage[3], bytes))) return rc;
    xd = (float*)e->stage[3].p;
    if (noise_given) CUDA_TRY(cudaMemcpyAsync(xd, x, bytes, cudaMemcpyHostToDevice, st));
  }
  const void* init_d = nullptr;
  if (init_image && (rc = dev_in(e, 4, init_image, bytes, st, &init_d))) return rc;
  // x_T (th.randn, gaussian_diffusion.py:704) and q_sample for skip_timesteps / init_image (:706-713)
  const bool do_q = init_image != nullptr || skip_timesteps > 0;
  if (!noise_given || do_q) {
    const float sa = do_q ? e->qsample[2 * first_index] : 0.f, sb = do_q ? e->qsample[2 * first_index + 1] : 0.f;
    const float* init_use = (const float*)init_d;
    if (do_q && !init_use) {                     // init_image = zeros_like(img)  (:706-707)
      CUDA_TRY(cudaMemsetAsync(e->x0, 0, bytes, st));
      init_use = e->x0;
    }
    init_noise_kernel<<<elementwise_grid(e, (per_clip >> 2) * B), 256, 0, st>>>(
        xd, do_q ? init_use : nullptr, noise_given, sa, sb, e->clip_ids, B, per_clip, (uint32_t)(seed & 0xffffffffu),
        (uint32_t)(seed >> 32), (uint32_t)segment);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  // the loop, cut at the dump points (dump_steps: `if i in dump_steps: dump.append(deepcopy(sample))`, :664-665)
  int k0 = 0, hist_len = 0;
  for (int piece = 0; piece <= n_dump; ++piece) {
    const int k1 = piece < n_dump ? opts->dump_iters[piece] + 1 : n_run;
    if (k1 > k0) {
      if (e->sampler == DSG_SAMPLER_PLMS) rc = plms_run(e, B, xd, k0, k1 - k0, first_index, order, &hist_len, seed, segment, st);
      else rc = dsg_run_steps(e, B, xd, k0, k1 - k0, first_index, seed, segment, st);
      if (rc) return rc;
    }
    if (piece < n_dump)
      CUDA_TRY(cudaMemcpyAsync(opts->dump_out + (size_t)piece * B * per_clip, xd, bytes, cudaMemcpyDefault, st));
    k0 = k1;
  }
  if (!xdev) {
    CUDA_TRY(cudaMemcpyAsync(x, xd, bytes, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  } else if (n_dump > 0 && !is_device_ptr(opts->dump_out)) {
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return DSG_OK;
}

// The hot loop (gaussian_diffusion.py:721-740): n_run x (denoiser + posterior).  Plain stream launches here;
// dsg_tc.cu overrides with a CUDA-graph replay when the tensor-core path is active.
int dsg_run_steps(dsg_engine* e, int B, float* xd, int k0, int n_run, int first_index, uint64_t seed, int segment, cudaStream_t st) {
  if (e->d.precision == DSG_PRECISION_BF16) return dsg_tc_run_steps(e, B, xd, k0, n_run, first_index, seed, segment, st);
  for (int k = k0; k < k0 + n_run; ++k) {
    const StepRef step{nullptr, k, first_index};
    TRY(denoise_f32(e, B, xd, nullptr, step, e->x0, st));
    PROF(e, PT_POSTERIOR, st, launch_posterior(e, B, xd, e->x0, step, -1, 0, seed, segment, st));
    if (e->profiling && e->prof_used > 4096) prof_collect(e);
  }
  return DSG_OK;
}

int dsg_upload_loop_params(dsg_engine* e, int k0, int first_index, uint64_t seed, int segment, cudaStream_t st) {
  LoopParams lp;
  memset(&lp, 0, sizeof lp);
  lp.k = k0; lp.first_index = first_index;
  lp.key0 = (uint32_t)(seed & 0xffffffffu); lp.key1 = (uint32_t)(seed >> 32); lp.segment = (uint32_t)segment;
  CUDA_TRY(cudaMemcpyAsync(e->d_loop, &lp, sizeof lp, cudaMemcpyHostToDevice, st));   // pageable source: staged before return
  return DSG_OK;
}

extern "C" int dsg_stitch_segment(dsg_engine* e, int32_t B, const float* prev_tail, float* sample, int32_t smoothing,
                                  void* stream) {
  if (!e || !prev_tail || !sample) return dsg_fail(DSG_ERR_BAD_SHAPE, "null argument");
  if (B <= 0 || B > e->d.max_batch) return dsg_fail(DSG_ERR_BAD_SHAPE, "batch %d outside 1..%d", B, e->d.max_batch);
  DeviceGuard guard(e->d.device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sb = (size_t)B * e->d.njoints * e->d.n_poses * sizeof(float);
  const size_t tb = (size_t)B * e->d.njoints * e->d.n_seed * sizeof(float);
  int rc;
  const void* td;
  if ((rc = dev_in(e, 5, prev_tail, tb, st, &td))) return rc;
  const bool sdev = is_device_ptr(sample);
  const void* sd = sample;
  if (!sdev && (rc = dev_in(e, 3, sample, sb, st, &sd))) return rc;
  stitch_segment_kernel<<<B, 256, 0, st>>>((const float*)td, (float*)sd, e->d.njoints, e->d.n_poses, e->d.n_seed, smoothing);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  if (!sdev) {
    CUDA_TRY(cudaMemcpyAsync(sample, sd, sb, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
  }
  return DSG_OK;
}

// --------------------------------------------------------------------------------------------------
// introspection
// --------------------------------------------------------------------------------------------------
extern "C" int64_t dsg_kernel_launch_count(const dsg_engine* e) { return e ? e->launches : -1; }

extern "C" int64_t dsg_debug_read(dsg_engine* e, const char* name, int32_t B, float* dst, int64_t capacity) {
  if (!e || !name) return dsg_fail(DSG_ERR_BAD_SHAPE, "null argument");
  DeviceGuard guard(e->d.device);
  const int D = e->d.latent_dim, S = e->S, T = e->d.n_poses;
  const size_t slot_elems = (size_t)e->d.max_batch * S * D;
  if (!strcmp(name, "enable")) {
    if (!e->dbg) { int rc = dalloc(&e->dbg, slot_elems * (e->d.num_layers + 1)); if (rc) return rc; }
    e->debug = true;
    e->graph_valid = false;
    return 0;
  }
  if (!dst) return dsg_fail(DSG_ERR_BAD_SHAPE, "null dst");
  if (!strcmp(name, "clipprof")) {      // cycle counters of the clip kernel (DSG_CLIP_PROF=1), returned as floats
    if (!dsg_tc_clip_prof(e) || capacity < 32) return dsg_fail(DSG_ERR_STATE, "no clip-kernel profile");
    CUDA_TRY(cudaDeviceSynchronize());
    long long h[32];
    CUDA_TRY(cudaMemcpy(h, dsg_tc_clip_prof(e), sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 32; ++i) dst[i] = (float)h[i];
    return 32;
  }
  if (B <= 0 || B > e->d.max_batch) return dsg_fail(DSG_ERR_BAD_SHAPE, "batch");
  CUDA_TRY(cudaDeviceSynchronize());
  if (!strcmp(name, "h_in")) {
    const int64_t n = (int64_t)B * T * D;
    if (capacity < n) return dsg_fail(DSG_ERR_BAD_SHAPE, "capacity");
    if (e->d.precision == DSG_PRECISION_BF16)    // tensor-core path keeps h as [B,S,D] with the token slot at row 0
      CUDA_TRY(cudaMemcpy2D(dst, (size_t)T * D * sizeof(float), dsg_tc_h(e) + D, (size_t)S * D * sizeof(float),
                            (size_t)T * D * sizeof(float), B, cudaMemcpyDefault));
    else
      CUDA_TRY(cudaMemcpy(dst, e->h, n * sizeof(float), cudaMemcpyDefault));
    return n;
  }
  if (!e->debug) return dsg_fail(DSG_ERR_STATE, "call dsg_debug_read(e, \"enable\", ...) before the denoise call");
  if (!strcmp(name, "tok")) {
    const int64_t n = (int64_t)B * D;
    if (capacity < n) return dsg_fail(DSG_ERR_BAD_SHAPE, "capacity");
    CUDA_TRY(cudaMemcpy2D(dst, D * sizeof(float), e->dbg, (size_t)S * D * sizeof(float), D * sizeof(float), B, cudaMemcpyDefault));
    return n;
  }
  if (!strncmp(name, "xs", 2)) {
    const int l = atoi(name + 2);
    if (l < 0 || l > e->d.num_layers) return dsg_fail(DSG_ERR_BAD_SHAPE, "layer %d", l);
    const int64_t n = (int64_t)B * S * D;
    if (capacity < n) return dsg_fail(DSG_ERR_BAD_SHAPE, "capacity");
    CUDA_TRY(cudaMemcpy(dst, e->dbg + (size_t)l * slot_elems, n * sizeof(float), cudaMemcpyDefault));
    return n;
  }
  return dsg_fail(DSG_ERR_UNSUPPORTED, "unknown tap '%s'", name);
}

extern "C" int dsg_profile(dsg_engine* e, int32_t enable) {
  if (!e) return dsg_fail(DSG_ERR_BAD_SHAPE, "null argument");
  DeviceGuard guard(e->d.device);
  if (e->profiling) prof_collect(e);
  if (enable == 1 && !e->profiling) { for (int i = 0; i < PT_COUNT; ++i) { e->prof_ms[i] = 0; e->prof_n[i] = 0; } }
  e->profiling = enable != 0;
  e->graph_valid = false;
  return DSG_OK;
}

extern "C" int dsg_profile_read(dsg_engine* e, int32_t tag, int64_t* count, double* total_ms) {
  if (!e || !count || !total_ms || tag < 0 || tag >= PT_COUNT) return dsg_fail(DSG_ERR_BAD_SHAPE, "bad profile tag");
  if (e->profiling) prof_collect(e);
  *count = e->prof_n[tag]; *total_ms = e->prof_ms[tag];
  return DSG_OK;
}

extern "C" const char* dsg_profile_tag_name(int32_t tag) {
  static const char* names[PT_COUNT] = {"gemm_in", "local_attention", "gemm_qkv", "self_attention", "gemm_outproj",
                                        "layernorm", "gemm_ff1", "gemm_ff2", "gemm_head_posterior", "posterior", "noise", "other"};
  return (tag >= 0 && tag < PT_COUNT) ? names[tag] : nullptr;
}

extern "C" const char* dsg_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* dsg_version(void) { return "dsg-b200 0.1 (sm_100a)"; }
