// Tensor-core (tcgen05 / TMEM / TMA) path of the engine — DSG_PRECISION_BF16.
//
// One denoiser call = 1 + 1 + 5*L + 1 (+1) kernels (L = 8 -> 44), every Linear a tcgen05 GEMM with its
// consumer fused into the epilogue (dsg_tc_gemm.cuh); activations stay in L2-resident bf16 buffers, the residual
// stream / LayerNorm / posterior stay fp32.  The 1000-step loop replays ONE captured CUDA graph of a step
// (device-side step counter), with the data-independent noise draw on a forked branch.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "dsg_engine.h"
#include "dsg_tc_gemm.cuh"
#include "dsg_tc_kernels.cuh"
#include "dsg_clip_kernel.cuh"

using bf16 = __nv_bfloat16;
using namespace tc;

struct dsg_tc_state {
  int Jpad = 0, Rpad = 0, bn_d = 256;
  bf16 *Wxp = nullptr, *Wout = nullptr;
  std::vector<bf16*> Wqkv, Wo, W1, W2;
  bf16 *xb = nullptr, *xsb = nullptr, *qkvb = nullptr, *attb = nullptr, *ffb = nullptr;
  float *hS = nullptr, *z = nullptr, *xloop = nullptr;
  CUtensorMap tm_xb, tm_xsb, tm_attb, tm_ffb, tm_Wxp, tm_Wout;
  std::vector<CUtensorMap> tm_Wqkv, tm_Wo, tm_W1, tm_W2;
  // graph replay
  cudaStream_t main = nullptr, side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_in = nullptr, ev_out = nullptr;
  cudaGraphExec_t exec = nullptr;
  int graph_B = 0, graph_sampler = -1;
  int nodes_per_step = 0;
  // clip kernel (one persistent CTA per clip): weight slabs + parameter blocks
  bool clip_ok = false;
  bf16 *wK256 = nullptr, *wK1024 = nullptr;
  float *lparams = nullptr, *bout = nullptr;
  uint8_t* xa = nullptr;
  long long* prof = nullptr;
  CUtensorMap tm_cin, tm_c128, tm_c64, tm_cw2;
  bool l2_limit_set = false;
  size_t l2_window_max = 0;
  int pair_clusters = -1;          // CTA pairs (clusters of 2) the device can keep resident at once; -1 = not asked yet
};

#include "dsg_tc_host.cuh"

// ---------------------------------------------------------------------------------------------------
// clip kernel set-up: weight slabs in the row order clip::R_* expects, fp32 parameter blocks, tensor maps
static int clip_setup(dsg_engine* e) {
  dsg_tc_state* t = e->tc;
  const dsg_model_desc& d = e->d;
  using namespace clip;
  if (d.variant != DSG_VARIANT_ATTN3 && d.variant != DSG_VARIANT_ATTN4) return DSG_OK;
  if (d.njoints != J || d.n_poses != T || d.latent_dim != D || d.ff_size != F || d.num_layers != NL || d.num_heads != NH ||
      d.local_heads != LH || d.local_window != WIN) return DSG_OK;           // other geometries run on the multi-kernel path
  const size_t rows256 = (size_t)R_HEAD + JPAD;
  TRY(dalloc0(&t->wK256, rows256 * D));
  TRY(dalloc0(&t->wK1024, (size_t)NL * D * F));
  std::vector<float> lp((size_t)NL * P_SIZE, 0.f), tmp(4096);
  auto fetch = [&](const float* dev, float* dst, size_t n) { return cudaMemcpy(dst, dev, n * sizeof(float), cudaMemcpyDeviceToHost); };
  for (int l = 0; l < NL; ++l) {
    float* const* w = &e->w[W_LAYER0 + 12 * l];
    bf16* base = t->wK256 + (size_t)l * R_LAYER * D;
    for (int h = 0; h < NH; ++h)
      for (int part = 0; part < 3; ++part)
        TRY(pack_w(w[L_INPROJ_W] + (size_t)(part * D + h * HD) * D, base + (size_t)(R_QKV + h * 192 + part * 64) * D, 64, D, D, 64, D));
    TRY(pack_w(w[L_OUTPROJ_W], base + (size_t)R_WO * D, D, D, D, D, D));
    TRY(pack_w(w[L_FF1_W], base + (size_t)R_W1 * D, F, D, D, F, D));
    pack_w2_perm_kernel<<<296, 256>>>(w[L_FF2_W], reinterpret_cast<__half*>(t->wK1024) + (size_t)l * D * F, D, F);   // fp16 slab, hidden-in-TMEM K order
    CUDA_TRY(cudaGetLastError());
    float* p = lp.data() + (size_t)l * P_SIZE;
    CUDA_TRY(fetch(w[L_INPROJ_B], tmp.data(), 3 * D));
    for (int h = 0; h < NH; ++h)
      for (int part = 0; part < 3; ++part)
        for (int i = 0; i < 64; ++i) p[P_BQKV + h * 192 + part * 64 + i] = tmp[part * D + h * HD + i];
    CUDA_TRY(fetch(w[L_OUTPROJ_B], p + P_BO, D)); CUDA_TRY(fetch(w[L_N1_W], p + P_G1, D)); CUDA_TRY(fetch(w[L_N1_B], p + P_BE1, D));
    {   // the v bias passes through the softmax average unchanged: bo' = bo + Wo bv (the k bias cancels in the softmax)
      std::vector<float> wo((size_t)D * D);
      CUDA_TRY(fetch(w[L_OUTPROJ_W], wo.data(), wo.size()));
      for (int o = 0; o < D; ++o) {
        double acc = 0.0;
        for (int i = 0; i < D; ++i) acc += (double)wo[(size_t)o * D + i] * (double)tmp[2 * D + i];
        p[P_BO + o] += (float)acc;
      }
    }
    CUDA_TRY(fetch(w[L_FF1_B], p + P_B1, F));    CUDA_TRY(fetch(w[L_FF2_B], p + P_B2, D));
    CUDA_TRY(fetch(w[L_N2_W], p + P_G2, D));     CUDA_TRY(fetch(w[L_N2_B], p + P_BE2, D));
  }
  TRY(pack_w(e->w[W_OUT_W], t->wK256 + (size_t)R_HEAD * D, J, D, D, JPAD, D));
  TRY(dalloc0(&t->lparams, lp.size()));
  CUDA_TRY(cudaMemcpy(t->lparams, lp.data(), lp.size() * sizeof(float), cudaMemcpyHostToDevice));
  TRY(dalloc0(&t->bout, (size_t)JPAD));
  TRY(dalloc0(&t->prof, (size_t)64));
  TRY(dalloc0(&t->xa, (size_t)d.max_batch * XA_BYTES));
  CUDA_TRY(cudaMemcpy(t->bout, e->w[W_OUT_B], (size_t)J * sizeof(float), cudaMemcpyDeviceToDevice));
  TRY(make_tmap(&t->tm_cin, t->Wxp, D, JPAD, 128));
  TRY(make_tmap(&t->tm_c128, t->wK256, rows256, D, 128));
  TRY(make_tmap(&t->tm_c64, t->wK256, rows256, D, 64));
  TRY(make_tmap(&t->tm_cw2, t->wK1024, (uint64_t)NL * D, F, 128));
  CUDA_TRY(cudaFuncSetAttribute(clip::clip_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(clip::clip_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(clip::clip_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(clip::clip_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  t->clip_ok = true;
  return DSG_OK;
}

static int clip_run(dsg_engine* e, int B, float* xd, int k0, int n_run, int first_index, uint64_t seed, int segment, cudaStream_t st) {
  dsg_tc_state* t = e->tc;
  TRY(dsg_upload_loop_params(e, k0, first_index, seed, segment, st));
  clip::ClipParams p;
  p.x = xd; p.xa = t->xa; p.z = t->z; p.cond = e->cond; p.emb1 = e->emb1; p.te = e->te; p.TW = e->TW; p.cs = e->cs_local;
  p.lparams = t->lparams; p.bout = t->bout; p.coef = e->coef; p.tmap = e->tmap; p.clip_ids = e->noise_ids; p.lp = e->d_loop;
  p.B = B; p.n_run = n_run; p.sampler = e->sampler;
  p.prof = getenv("DSG_CLIP_PROF") ? t->prof : nullptr;
  p.dbg = e->dbg; p.dbg_slot = (long long)e->d.max_batch * e->S * e->d.latent_dim; p.debug = e->debug ? 1 : 0;
  // Fewer clips than half the SMs: a CTA PAIR (cluster of 2) per clip, each streaming half of the weights (dsg_clip_kernel.cuh,
  // CL = 2) — as long as the device keeps that many clusters of this size resident at once (74 on a full B200; fewer, or none, on
  // a partitioned one).  DSG_CLIP_PAIR=0 disables it.
  if (t->pair_clusters < 0) {
    cudaLaunchConfig_t occ;
    memset(&occ, 0, sizeof occ);
    occ.gridDim = dim3(2 * (e->num_sms / 2)); occ.blockDim = dim3(512); occ.dynamicSmemBytes = clip::SMEM_BYTES;
    cudaLaunchAttribute oa[1];
    oa[0].id = cudaLaunchAttributeClusterDimension;
    oa[0].val.clusterDim.x = 2; oa[0].val.clusterDim.y = 1; oa[0].val.clusterDim.z = 1;
    occ.attrs = oa; occ.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, clip::clip_kernel<false, 2>, &occ) != cudaSuccess) { n = 0; cudaGetLastError(); }
    t->pair_clusters = n;
  }
  const char* pair_env = getenv("DSG_CLIP_PAIR");
  const bool pair = B <= t->pair_clusters && 2 * B <= e->num_sms && !(pair_env && !strcmp(pair_env, "0"));
  const int slots = B < e->num_sms ? B : e->num_sms;          // clips in flight
  const int grid = pair ? 2 * B : slots;
  clip::pack_xa_kernel<<<dim3(clip::JPAD / 64, B), 256, 0, st>>>(xd, t->xa);       // x_T as the first step's A k-blocks
  e->launches++;
  // The per-CTA noise scratch (grid x 401 KB) is written by the noise warp and bulk-copied back ~100-300 us later, every step: an
  // access-policy window keeps it resident in L2 (persisting hits), so it costs no HBM traffic.  DSG_L2_PERSIST=0 disables it.
  static const bool l2_persist = !(getenv("DSG_L2_PERSIST") && !strcmp(getenv("DSG_L2_PERSIST"), "0"));
  bool window_set = false;
  if (l2_persist) {
    const size_t zbytes = (size_t)slots * clip::J * clip::T * sizeof(float);
    if (!t->l2_limit_set) {
      cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, e->d.device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
        const size_t want = (size_t)e->num_sms * clip::J * clip::T * sizeof(float);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want < (size_t)prop.persistingL2CacheMaxSize ? want : (size_t)prop.persistingL2CacheMaxSize);
        t->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
      }
      cudaGetLastError();
      t->l2_limit_set = true;
    }
    if (t->l2_window_max > 0) {
      cudaStreamAttrValue av;
      memset(&av, 0, sizeof av);
      av.accessPolicyWindow.base_ptr = t->z;
      av.accessPolicyWindow.num_bytes = zbytes < t->l2_window_max ? zbytes : t->l2_window_max;
      av.accessPolicyWindow.hitRatio = 1.0f;
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      window_set = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
      cudaGetLastError();
    }
  }
  if (pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = clip::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // (a launch error stays in cudaGetLastError, checked below once the access-policy window is off the stream again)
    if (p.prof) cudaLaunchKernelEx(&cfg, clip::clip_kernel<true, 2>, t->tm_cin, t->tm_c128, t->tm_c64, t->tm_cw2, p);
    else cudaLaunchKernelEx(&cfg, clip::clip_kernel<false, 2>, t->tm_cin, t->tm_c128, t->tm_c64, t->tm_cw2, p);
  } else if (p.prof) clip::clip_kernel<true, 1><<<grid, 512, clip::SMEM_BYTES, st>>>(t->tm_cin, t->tm_c128, t->tm_c64, t->tm_cw2, p);
  else clip::clip_kernel<false, 1><<<grid, 512, clip::SMEM_BYTES, st>>>(t->tm_cin, t->tm_c128, t->tm_c64, t->tm_cw2, p);
  e->launches++;
  if (window_set) {                        // the window applies to launches made while it is set: take it off the caller's stream again
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
  }
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

// ---------------------------------------------------------------------------------------------------
int dsg_tc_create(dsg_engine* e) {
  const dsg_model_desc& d = e->d;
  // a LayerNorm row must fit one CTA's accumulator: D = 256 (ZEGGS), 384 (BEAT "+"), 512 (TWH "+")
  if (d.latent_dim != 256 && d.latent_dim != 384 && d.latent_dim != 512)
    return dsg_fail(DSG_ERR_UNSUPPORTED, "DSG_PRECISION_BF16 covers latent_dim 256 / 384 / 512 (the reference's presets); D = %d runs "
                    "on the fp32 engine", d.latent_dim);
  if (d.ff_size % 256) return dsg_fail(DSG_ERR_UNSUPPORTED, "ff_size must be a multiple of 256 for the bf16 path");
  dsg_tc_state* t = new dsg_tc_state();
  e->tc = t;
  const int D = d.latent_dim, F = d.ff_size, J = d.njoints, S = e->S, L = d.num_layers, MB = d.max_batch;
  t->Jpad = (J + 127) / 128 * 128;
  t->Rpad = (MB * S + BM - 1) / BM * BM;
  const int Jpad = t->Jpad, Rpad = t->Rpad;
  TRY(dalloc0(&t->Wxp, (size_t)D * Jpad));
  TRY(dalloc0(&t->Wout, (size_t)Jpad * D));
  TRY(pack_w(e->Wxp, t->Wxp, D, J, J, D, Jpad));
  TRY(pack_w(e->w[W_OUT_W], t->Wout, J, D, D, Jpad, D));
  // tile widths by latent_dim: D-wide outputs (input GEMM, in_proj) in 256- or 128-column tiles; the LayerNorm GEMMs
  // (out_proj, linear2) in ONE D-wide tile whose B operand arrives as boxes of D (256) or D/2 (384, 512) rows
  const int bn_d = (D % 256 == 0) ? 256 : 128, box_ln = (D > 256) ? D / 2 : D;
  t->bn_d = bn_d;
  t->Wqkv.resize(L); t->Wo.resize(L); t->W1.resize(L); t->W2.resize(L);
  t->tm_Wqkv.resize(L); t->tm_Wo.resize(L); t->tm_W1.resize(L); t->tm_W2.resize(L);
  for (int l = 0; l < L; ++l) {
    float* const* w = &e->w[W_LAYER0 + 12 * l];
    TRY(dalloc0(&t->Wqkv[l], (size_t)3 * D * D)); TRY(pack_w(w[L_INPROJ_W], t->Wqkv[l], 3 * D, D, D, 3 * D, D));
    TRY(dalloc0(&t->Wo[l], (size_t)D * D));       TRY(pack_w(w[L_OUTPROJ_W], t->Wo[l], D, D, D, D, D));
    TRY(dalloc0(&t->W1[l], (size_t)F * D));       TRY(pack_w(w[L_FF1_W], t->W1[l], F, D, D, F, D));
    TRY(dalloc0(&t->W2[l], (size_t)D * F));       TRY(pack_w(w[L_FF2_W], t->W2[l], D, F, F, D, F));
    TRY(make_tmap(&t->tm_Wqkv[l], t->Wqkv[l], 3 * D, D, bn_d));
    TRY(make_tmap(&t->tm_Wo[l], t->Wo[l], D, D, box_ln));
    TRY(make_tmap(&t->tm_W1[l], t->W1[l], F, D, 256));
    TRY(make_tmap(&t->tm_W2[l], t->W2[l], D, F, box_ln));
  }
  TRY(dalloc0(&t->xb, (size_t)Rpad * Jpad));
  TRY(dalloc0(&t->xsb, (size_t)Rpad * D));
  TRY(dalloc0(&t->qkvb, (size_t)Rpad * 3 * D));
  TRY(dalloc0(&t->attb, (size_t)Rpad * D));
  TRY(dalloc0(&t->ffb, (size_t)Rpad * F));
  TRY(dalloc0(&t->hS, (size_t)Rpad * D));
  TRY(dalloc0(&t->z, (size_t)MB * J * d.n_poses));
  TRY(dalloc0(&t->xloop, (size_t)MB * J * d.n_poses));
  TRY(make_tmap(&t->tm_xb, t->xb, Rpad, Jpad, BM));
  TRY(make_tmap(&t->tm_xsb, t->xsb, Rpad, D, BM));
  TRY(make_tmap(&t->tm_attb, t->attb, Rpad, D, BM));
  TRY(make_tmap(&t->tm_ffb, t->ffb, Rpad, F, BM));
  TRY(make_tmap(&t->tm_Wxp, t->Wxp, D, Jpad, bn_d));
  TRY(make_tmap(&t->tm_Wout, t->Wout, Jpad, D, 128));
  CUDA_TRY(cudaStreamCreateWithFlags(&t->main, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&t->ev_fork, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&t->ev_join, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&t->ev_in, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&t->ev_out, cudaEventDisableTiming));
  TRY(clip_setup(e));
  return DSG_OK;
}

const float* dsg_tc_h(dsg_engine* e) { return e->tc ? e->tc->hS : nullptr; }
const long long* dsg_tc_clip_prof(dsg_engine* e) { return e->tc ? e->tc->prof : nullptr; }

void dsg_tc_destroy(dsg_engine* e) {
  dsg_tc_state* t = e->tc;
  if (!t) return;
  if (t->exec) cudaGraphExecDestroy(t->exec);
  void* ptrs[] = {t->Wxp, t->Wout, t->xb, t->xsb, t->qkvb, t->attb, t->ffb, t->hS, t->z, t->xloop, t->wK256, t->wK1024, t->lparams, t->bout, t->prof, t->xa};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& v : {t->Wqkv, t->Wo, t->W1, t->W2}) for (bf16* p : v) if (p) cudaFree(p);
  if (t->main) cudaStreamDestroy(t->main);
  if (t->side) cudaStreamDestroy(t->side);
  for (cudaEvent_t ev : {t->ev_fork, t->ev_join, t->ev_in, t->ev_out}) if (ev) cudaEventDestroy(ev);
  delete t;
  e->tc = nullptr;
}

// ---------------------------------------------------------------------------------------------------
static int tc_pack_x(dsg_engine* e, int B, const float* x, cudaStream_t st) {
  dsg_tc_state* t = e->tc;
  dim3 grid((e->d.njoints + 31) / 32, (e->d.n_poses + 31) / 32, B);
  pack_x_bf16_kernel<<<grid, 256, 0, st>>>(x, t->xb, e->d.njoints, e->d.n_poses, e->S, t->Jpad);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

static int tc_noise(dsg_engine* e, int B, StepRef step, cudaStream_t st) {
  NoiseArgs a{e->tc->z, e->noise_ids, step, B, (long long)e->d.njoints * e->d.n_poses, e->sampler};
  noise_tile_kernel<<<elementwise_grid(e, (a.per_clip >> 2) * B), 256, 0, st>>>(a);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

static int tc_debug_snap(dsg_engine* e, int slot, int B, cudaStream_t st) {
  if (!e->debug) return DSG_OK;
  const size_t n = (size_t)B * e->S * e->d.latent_dim;
  CUDA_TRY(cudaMemcpyAsync(e->dbg + (size_t)slot * e->d.max_batch * e->S * e->d.latent_dim, e->xs, n * sizeof(float),
                           cudaMemcpyDeviceToDevice, st));
  return DSG_OK;
}

// self-attention: mma.sync kernel (S <= 96, head dim 64); DSG_ATTN=simt selects the CUDA-core kernel (debugging)
static int tc_attention(dsg_engine* e, int B, cudaStream_t st) {
  dsg_tc_state* t = e->tc;
  static const bool simt = getenv("DSG_ATTN") && !strcmp(getenv("DSG_ATTN"), "simt");
  const int hd = e->d.latent_dim / e->d.num_heads;
  const float scale_log2e = 1.4426950408889634f / sqrtf((float)hd);
  const int grid = B * e->d.num_heads;
  bool done = false;
#define ATTN_CASE(RT, HD)                                                                                                  \
  if (!done && !simt && hd == HD && e->S <= 16 * RT) {                                                                      \
    constexpr int smem = 3 * 16 * RT * (HD + 8) * 2;                                                                        \
    static bool configured = false;                                                                                         \
    if (!configured) {                                                                                                      \
      CUDA_TRY(cudaFuncSetAttribute(self_attention_mma_kernel<RT, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
      configured = true;                                                                                                    \
    }                                                                                                                       \
    self_attention_mma_kernel<RT, HD><<<grid, 32 * RT, smem, st>>>(t->qkvb, t->attb, e->S, e->d.latent_dim, e->d.num_heads, scale_log2e); \
    done = true;                                                                                                            \
  }
  ATTN_CASE(6, 64)        // ZEGGS: S = 89, 4 x 64
  ATTN_CASE(10, 96)       // BEAT "+": S = 151, 4 x 96
  ATTN_CASE(10, 128)      // TWH "+":  S = 151, 4 x 128
#undef ATTN_CASE
  if (!done) return launch_self_attention_bf16(e, B, t->qkvb, t->attb, st);      // any other geometry: CUDA-core kernel
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

// out_proj / linear2 + residual + LayerNorm: one D-wide tile per 128 rows (the whole row's statistics live in one accumulator)
static int launch_ln(dsg_engine* e, const CUtensorMap& a, const CUtensorMap& w, const TcEpiArgs& ep, cudaStream_t st) {
  switch (e->d.latent_dim) {
    case 256: return launch_tc<256, 4, EPI_LN>(e, a, w, ep, 1, st);
    case 384: return launch_tc<384, 3, EPI_LN>(e, a, w, ep, 1, st);
    case 512: return launch_tc<512, 2, EPI_LN>(e, a, w, ep, 1, st);
  }
  return dsg_fail(DSG_ERR_UNSUPPORTED, "LayerNorm GEMM: latent_dim %d", e->d.latent_dim);
}

// everything of one denoiser call up to (not including) the head GEMM
static int tc_body(dsg_engine* e, int B, const int* tsel, StepRef step, cudaStream_t st) {
  dsg_tc_state* t = e->tc;
  const dsg_model_desc& d = e->d;
  const int D = d.latent_dim, F = d.ff_size, S = e->S, T = d.n_poses, M = B * S;
  TcEpiArgs ep;
  memset(&ep, 0, sizeof ep);
  ep.M = M; ep.S = S; ep.T = T; ep.step = step;
  {  // h = Wxp x_t + cond + TW[t]
    TcEpiArgs a = ep;
    a.N = D; a.K = t->Jpad; a.out = t->hS; a.cond = e->cond; a.TW = e->TW; a.tsel = tsel; a.tmap = e->tmap;
    if (t->bn_d == 256) PROF(e, PT_GEMM_IN, st, (launch_tc<256, 4, EPI_IN>(e, t->tm_xb, t->tm_Wxp, a, D / 256, st)));
    else PROF(e, PT_GEMM_IN, st, (launch_tc<128, 4, EPI_IN>(e, t->tm_xb, t->tm_Wxp, a, D / 128, st)));
  }
  PROF(e, PT_LOCAL_ATTN, st, launch_local_attention(e, B, t->hS, (long long)S * D, 1, e->xs, t->xsb, tsel, step, st));
  TRY(tc_debug_snap(e, 0, B, st));
  for (int l = 0; l < d.num_layers; ++l) {
    float* const* w = &e->w[W_LAYER0 + 12 * l];
    {
      TcEpiArgs a = ep;
      a.N = 3 * D; a.K = D; a.bias = w[L_INPROJ_B]; a.out = t->qkvb; a.ldc = 3 * D;
      if (t->bn_d == 256) PROF(e, PT_GEMM_QKV, st, (launch_tc<256, 4, EPI_BF16>(e, t->tm_xsb, t->tm_Wqkv[l], a, 3 * D / 256, st)));
      else PROF(e, PT_GEMM_QKV, st, (launch_tc<128, 4, EPI_BF16>(e, t->tm_xsb, t->tm_Wqkv[l], a, 3 * D / 128, st)));
    }
    PROF(e, PT_SELF_ATTN, st, tc_attention(e, B, st));
    {
      TcEpiArgs a = ep;
      a.N = D; a.K = D; a.bias = w[L_OUTPROJ_B]; a.xs = e->xs; a.xsb = t->xsb; a.gamma = w[L_N1_W]; a.beta = w[L_N1_B];
      PROF(e, PT_GEMM_OUTPROJ, st, launch_ln(e, t->tm_attb, t->tm_Wo[l], a, st));
    }
    {
      TcEpiArgs a = ep;
      a.N = F; a.K = D; a.bias = w[L_FF1_B]; a.out = t->ffb; a.ldc = F;
      PROF(e, PT_GEMM_FF1, st, (launch_tc<256, 4, EPI_GELU>(e, t->tm_xsb, t->tm_W1[l], a, F / 256, st)));
    }
    {
      TcEpiArgs a = ep;
      a.N = D; a.K = F; a.bias = w[L_FF2_B]; a.xs = e->xs; a.xsb = t->xsb; a.gamma = w[L_N2_W]; a.beta = w[L_N2_B];
      PROF(e, PT_GEMM_FF2, st, launch_ln(e, t->tm_ffb, t->tm_W2[l], a, st));
    }
    TRY(tc_debug_snap(e, l + 1, B, st));
  }
  return DSG_OK;
}

// OutputProcess (+ posterior when head_mode == 0)
static int tc_head(dsg_engine* e, int B, float* x, StepRef step, int head_mode, float* out, cudaStream_t st) {
  dsg_tc_state* t = e->tc;
  const dsg_model_desc& d = e->d;
  TcEpiArgs a;
  memset(&a, 0, sizeof a);
  a.M = B * e->S; a.N = d.njoints; a.K = d.latent_dim; a.S = e->S; a.T = d.n_poses; a.step = step;
  a.bias = e->w[W_OUT_B]; a.x = x; a.z = t->z; a.xb = t->xb; a.J = d.njoints; a.Jpad = t->Jpad; a.coef = e->coef;
  a.sampler = e->sampler; a.head_mode = head_mode; a.out = out;
  PROF(e, PT_GEMM_HEAD, st, (launch_tc<128, 4, EPI_HEAD>(e, t->tm_xsb, t->tm_Wout, a, t->Jpad / 128, st)));
  return DSG_OK;
}

int dsg_tc_denoise(dsg_engine* e, int B, const float* x, const int* tsel, StepRef step, float* out, cudaStream_t st) {
  TRY(tc_pack_x(e, B, x, st));
  TRY(tc_body(e, B, tsel, step, st));
  return tc_head(e, B, nullptr, step, 1, out, st);
}

// ---------------------------------------------------------------------------------------------------
// The hot loop.  Profiling / debug: plain launches on the caller's stream.  Otherwise: one captured step,
// replayed n_run times on the engine's stream (the legacy default stream cannot be captured).
// ---------------------------------------------------------------------------------------------------
static int tc_capture(dsg_engine* e, int B) {
  dsg_tc_state* t = e->tc;
  float* xd = t->xloop;
  if (t->exec) { cudaGraphExecDestroy(t->exec); t->exec = nullptr; }
  const StepRef step{e->d_loop, 0, 0, 0, 0, 0};
  const int64_t l0 = e->launches;
  cudaGraph_t graph = nullptr;
  CUDA_TRY(cudaStreamBeginCapture(t->main, cudaStreamCaptureModeThreadLocal));
  int rc = DSG_OK;
  do {
    if (cudaEventRecord(t->ev_fork, t->main) != cudaSuccess || cudaStreamWaitEvent(t->side, t->ev_fork, 0) != cudaSuccess) { rc = DSG_ERR_CUDA; break; }
    if ((rc = tc_noise(e, B, step, t->side))) break;
    if (cudaEventRecord(t->ev_join, t->side) != cudaSuccess) { rc = DSG_ERR_CUDA; break; }
    if ((rc = tc_body(e, B, nullptr, step, t->main))) break;
    if (cudaStreamWaitEvent(t->main, t->ev_join, 0) != cudaSuccess) { rc = DSG_ERR_CUDA; break; }
    if ((rc = tc_head(e, B, xd, step, 0, nullptr, t->main))) break;
    bump_step_kernel<<<1, 1, 0, t->main>>>(e->d_loop);
    e->launches++;
  } while (0);
  const cudaError_t ce = cudaStreamEndCapture(t->main, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); if (rc == DSG_ERR_CUDA) dsg_fail(rc, "graph capture failed: %s", cudaGetErrorString(cudaGetLastError())); return rc; }
  if (ce != cudaSuccess) return dsg_fail(DSG_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
  t->nodes_per_step = (int)(e->launches - l0);
  e->launches = l0;
  const cudaError_t ci = cudaGraphInstantiate(&t->exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ci != cudaSuccess) return dsg_fail(DSG_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ci));
  t->graph_B = B; t->graph_sampler = e->sampler;
  e->graph_valid = true;
  return DSG_OK;
}

int dsg_tc_run_steps(dsg_engine* e, int B, float* xd, int k0, int n_run, int first_index, uint64_t seed, int segment, cudaStream_t st) {
  dsg_tc_state* t = e->tc;
  // DSG_TC_MODE=kernels selects the multi-kernel graph path (default: the persistent clip kernel when the geometry allows)
  const bool force_kernels = getenv("DSG_TC_MODE") && !strcmp(getenv("DSG_TC_MODE"), "kernels");
  if (t->clip_ok && !force_kernels && !e->profiling) return clip_run(e, B, xd, k0, n_run, first_index, seed, segment, st);
  if (e->profiling || e->debug) {
    TRY(tc_pack_x(e, B, xd, st));
    for (int k = k0; k < k0 + n_run; ++k) {
      const StepRef step{nullptr, k, first_index, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), (uint32_t)segment};
      TRY(tc_body(e, B, nullptr, step, st));
      PROF(e, PT_NOISE, st, tc_noise(e, B, step, st));
      TRY(tc_head(e, B, xd, step, 0, nullptr, st));
    }
    return DSG_OK;
  }
  // hand over from the caller's stream to the engine's stream
  CUDA_TRY(cudaEventRecord(t->ev_in, st));
  CUDA_TRY(cudaStreamWaitEvent(t->main, t->ev_in, 0));
  // the graph works on an engine-owned copy of x, so one instantiation serves every segment / caller buffer
  const size_t xbytes = (size_t)B * e->d.njoints * e->d.n_poses * sizeof(float);
  CUDA_TRY(cudaMemcpyAsync(t->xloop, xd, xbytes, cudaMemcpyDeviceToDevice, t->main));
  TRY(dsg_upload_loop_params(e, k0, first_index, seed, segment, t->main));
  TRY(tc_pack_x(e, B, t->xloop, t->main));
  int k_start = 0;
  if (!e->graph_valid || !t->exec || t->graph_B != B || t->graph_sampler != e->sampler) {
    // first step outside the capture: every kernel is loaded / configured before the stream goes into capture mode
    const StepRef step{e->d_loop, 0, 0, 0, 0, 0};
    TRY(tc_noise(e, B, step, t->main));
    TRY(tc_body(e, B, nullptr, step, t->main));
    TRY(tc_head(e, B, t->xloop, step, 0, nullptr, t->main));
    bump_step_kernel<<<1, 1, 0, t->main>>>(e->d_loop);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    k_start = 1;
    if (n_run > 1) TRY(tc_capture(e, B));
  }
  for (int k = k_start; k < n_run; ++k) CUDA_TRY(cudaGraphLaunch(t->exec, t->main));
  e->launches += (int64_t)t->nodes_per_step * (n_run - k_start);
  CUDA_TRY(cudaMemcpyAsync(xd, t->xloop, xbytes, cudaMemcpyDeviceToDevice, t->main));
  CUDA_TRY(cudaEventRecord(t->ev_out, t->main));
  CUDA_TRY(cudaStreamWaitEvent(st, t->ev_out, 0));
  return DSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// Stand-alone check of the tcgen05 GEMM (tests/test_gpu_tc.py): C = bf16(A) * bf16(W)^T + bias, fp32 out.
// ---------------------------------------------------------------------------------------------------
extern "C" int dsg_selftest_gemm(int32_t device, int32_t bn, int32_t M, int32_t N, int32_t K, const float* A, const float* W,
                                 const float* bias, float* C) {
  if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0 || (K % 8)) return dsg_fail(DSG_ERR_BAD_SHAPE, "selftest_gemm: bad arguments");
  if (bn != 128 && bn != 256) return dsg_fail(DSG_ERR_BAD_SHAPE, "bn must be 128 or 256");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return dsg_fail(DSG_ERR_BAD_ARCH, "sm_%d%d: tcgen05 needs sm_100", prop.major, prop.minor);
  const int Mp = (M + BM - 1) / BM * BM, Np = (N + bn - 1) / bn * bn;
  float *dA = nullptr, *dW = nullptr, *dB = nullptr, *dC = nullptr;
  bf16 *bA = nullptr, *bW = nullptr;
  CUDA_TRY(cudaMalloc(&dA, (size_t)M * K * 4)); CUDA_TRY(cudaMalloc(&dW, (size_t)N * K * 4));
  CUDA_TRY(cudaMalloc(&dC, (size_t)M * N * 4)); CUDA_TRY(cudaMalloc(&dB, (size_t)N * 4));
  CUDA_TRY(cudaMalloc(&bA, (size_t)Mp * K * 2)); CUDA_TRY(cudaMalloc(&bW, (size_t)Np * K * 2));
  CUDA_TRY(cudaMemcpy(dA, A, (size_t)M * K * 4, cudaMemcpyDefault));
  CUDA_TRY(cudaMemcpy(dW, W, (size_t)N * K * 4, cudaMemcpyDefault));
  if (bias) CUDA_TRY(cudaMemcpy(dB, bias, (size_t)N * 4, cudaMemcpyDefault));
  TRY(pack_w(dA, bA, M, K, K, Mp, K));
  TRY(pack_w(dW, bW, N, K, K, Np, K));
  CUtensorMap ta, tb;
  TRY(make_tmap(&ta, bA, Mp, K, BM));
  TRY(make_tmap(&tb, bW, Np, K, bn));
  TcEpiArgs ep;
  memset(&ep, 0, sizeof ep);
  ep.M = M; ep.N = N; ep.K = K; ep.bias = bias ? dB : nullptr; ep.out = dC; ep.ldc = N;
  dsg_engine fake;
  if (bn == 128) TRY((launch_tc<128, 4, EPI_F32>(&fake, ta, tb, ep, Np / 128, 0)));
  else TRY((launch_tc<256, 4, EPI_F32>(&fake, ta, tb, ep, Np / 256, 0)));
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(C, dC, (size_t)M * N * 4, cudaMemcpyDefault));
  cudaFree(dA); cudaFree(dW); cudaFree(dB); cudaFree(dC); cudaFree(bA); cudaFree(bW);
  return DSG_OK;
}
