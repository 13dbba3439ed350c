// WavLM-Large conditioning forward on sm_100a: conv feature extractor as batched tcgen05
// GEMMs over overlapping-row TMA views of the channels-last input, grouped positional conv as 16 K = 8192 GEMMs per clip,
// 24 pre-norm transformer layers (tcgen05 GEMMs + mma.sync flash attention with the gated relative-position bias),
// final LayerNorm and the linear interpolation to n_poses frames (reference sample.py:44-48; WavLM/WavLM.py:323-375).
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "dsg_tc_host.cuh"
#include "dsg_wavlm_kernels.cuh"

namespace {
constexpr int E = 1024, FF = 4096, H = 16, NLAYER = 24, CC = 512, NCONV = 7;
const int CONV_K[NCONV] = {10, 3, 3, 3, 3, 2, 2};
const int CONV_S[NCONV] = {5, 2, 2, 2, 2, 2, 2};
// index into the weights array (diffusestylegesture_b200/wavlm_config.py:wavlm_state_dict_spec)
enum { WV_CONV0 = 0, WV_LN = 21, WV_LN_B, WV_PROJ_W, WV_PROJ_B, WV_PC_B, WV_PC_G, WV_PC_V, WV_RELB, WV_LAYER0 };
enum { WL_Q_W = 0, WL_Q_B, WL_K_W, WL_K_B, WL_V_W, WL_V_B, WL_O_W, WL_O_B, WL_GL_W, WL_GL_B, WL_GA, WL_LN1_W, WL_LN1_B, WL_FC1_W,
       WL_FC1_B, WL_FC2_W, WL_FC2_B, WL_LN2_W, WL_LN2_B, WL_PER_LAYER };
}  // namespace

struct dsg_wavlm {
  dsg_engine eng;                       // launch counter / device for the shared helpers
  int device = 0, MB = 0, N = 0, L[NCONV] = {0}, Lp = 0, Mpad = 0;
  std::vector<float*> w;                // fp32 copies of every tensor (biases, LayerNorm, small ones are used directly)
  float* wslab = nullptr;
  bf16* wconv[NCONV] = {nullptr};       // conv 1..6 packed [512][k*512]
  bf16 *wproj = nullptr, *wpc = nullptr;
  std::vector<bf16*> wqkv, wo, wfc1, wfc2;
  std::vector<float*> bqkv, wab;
  float *pcscale = nullptr, *posbias = nullptr;
  bf16 *act[2] = {nullptr, nullptr}, *hbf = nullptr, *qkv = nullptr, *att = nullptr, *ffb = nullptr, *xg = nullptr;
  float *convf = nullptr, *x = nullptr, *gate = nullptr, *lnout = nullptr, *wav = nullptr, *outbuf = nullptr;
  CUtensorMap tm_conv_a[NCONV], tm_conv_w[NCONV], tm_feat, tm_proj, tm_xg, tm_pc, tm_h, tm_att, tm_ff, tm_xout;
  std::vector<CUtensorMap> tm_qkv, tm_o, tm_fc1, tm_fc2;
  size_t smem_fa = 0;
  int xout_rows = 0;                    // rows the fp32 output map of the residual GEMMs was encoded for (= B * frames)
};

template <typename T>
static int walloc(T** p, size_t n) {
  CUDA_TRY(cudaMalloc((void**)p, n * sizeof(T)));
  CUDA_TRY(cudaMemset(*p, 0, n * sizeof(T)));
  return DSG_OK;
}

static std::vector<size_t> wavlm_sizes() {
  std::vector<size_t> s;
  int cin = 1;
  for (int i = 0; i < NCONV; ++i) { s.push_back((size_t)CC * cin * CONV_K[i]); s.push_back(CC); s.push_back(CC); cin = CC; }
  s.insert(s.end(), {(size_t)CC, (size_t)CC, (size_t)E * CC, (size_t)E, (size_t)E, (size_t)128, (size_t)E * 64 * 128, (size_t)320 * H});
  for (int l = 0; l < NLAYER; ++l) {
    const size_t per[WL_PER_LAYER] = {(size_t)E * E, E, (size_t)E * E, E, (size_t)E * E, E, (size_t)E * E, E, 8 * 64, 8, H, E, E,
                                      (size_t)FF * E, FF, (size_t)E * FF, E, E, E};
    s.insert(s.end(), per, per + WL_PER_LAYER);
  }
  s.push_back(E); s.push_back(E);
  return s;
}

extern "C" void dsg_wavlm_destroy(dsg_wavlm* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  std::vector<void*> ptrs = {m->wslab, m->wproj, m->wpc, m->pcscale, m->posbias, m->act[0], m->act[1], m->hbf, m->qkv, m->att, m->ffb,
                             m->xg, m->convf, m->x, m->gate, m->lnout, m->wav, m->outbuf};
  for (int i = 0; i < NCONV; ++i) ptrs.push_back(m->wconv[i]);
  for (auto* v : {&m->wqkv, &m->wo, &m->wfc1, &m->wfc2}) for (bf16* p : *v) ptrs.push_back(p);
  for (auto* v : {&m->bqkv, &m->wab}) for (float* p : *v) ptrs.push_back(p);
  for (void* p : ptrs) if (p) cudaFree(p);
  delete m;
}

#define WTRY(x) do { int rc__ = (x); if (rc__) { dsg_wavlm_destroy(m); return rc__; } } while (0)

extern "C" int dsg_wavlm_create(int32_t device, int32_t max_batch, int32_t n_samples, const float* const* weights, int32_t n_weights,
                                const float* pos_bias, dsg_wavlm** out) {
  if (!weights || !pos_bias || !out || max_batch <= 0 || n_samples < 400) return dsg_fail(DSG_ERR_BAD_SHAPE, "dsg_wavlm_create: bad arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return dsg_fail(DSG_ERR_BAD_ARCH, "no CUDA device: libdsg has no CPU fallback"); }
  if (device < 0 || device >= ndev) return dsg_fail(DSG_ERR_BAD_SHAPE, "device ordinal");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return dsg_fail(DSG_ERR_BAD_ARCH, "device is sm_%d%d; libdsg is built for sm_100a only", prop.major, prop.minor);
  CUDA_TRY(cudaSetDevice(device));
  const std::vector<size_t> sizes = wavlm_sizes();
  if ((int)sizes.size() != n_weights) return dsg_fail(DSG_ERR_BAD_SHAPE, "expected %d WavLM tensors, got %d", (int)sizes.size(), n_weights);
  dsg_wavlm* m = new dsg_wavlm();
  m->device = device; m->MB = max_batch; m->N = n_samples;
  m->eng.d.device = device; m->eng.num_sms = prop.multiProcessorCount;
  int n = n_samples;
  for (int i = 0; i < NCONV; ++i) { n = (n - CONV_K[i]) / CONV_S[i] + 1; m->L[i] = n; }
  const int Lf = m->L[NCONV - 1], MB = max_batch;
  if (Lf < 16 || Lf > 224) { dsg_wavlm_destroy(m); return dsg_fail(DSG_ERR_BAD_SHAPE, "%d frames: the attention kernel covers 16..224 frames per segment", Lf); }
  m->Lp = ((Lf + 128 + 7) / 8) * 8;
  m->Mpad = (MB * Lf + BM - 1) / BM * BM;
  // ---- fp32 copies of all tensors
  size_t total = 0;
  for (size_t s : sizes) total += (s + 3) & ~size_t(3);
  WTRY(walloc(&m->wslab, total));
  m->w.resize(sizes.size());
  size_t off = 0;
  for (size_t i = 0; i < sizes.size(); ++i) {
    m->w[i] = m->wslab + off;
    if (cudaMemcpy(m->w[i], weights[i], sizes[i] * sizeof(float), cudaMemcpyDefault) != cudaSuccess) {
      dsg_wavlm_destroy(m); return dsg_fail(DSG_ERR_CUDA, "copy of WavLM tensor %d failed", (int)i); }
    off += (sizes[i] + 3) & ~size_t(3);
  }
  // ---- packed bf16 operands
  for (int i = 1; i < NCONV; ++i) {
    const int K = CONV_K[i] * CC;
    WTRY(walloc(&m->wconv[i], (size_t)CC * K));
    wl::pack_conv_w_kernel<<<592, 256>>>(m->w[WV_CONV0 + 3 * i], m->wconv[i], CC, CC, CONV_K[i], nullptr);
    WTRY(make_tmap(&m->tm_conv_w[i], m->wconv[i], CC, K, 256));
  }
  WTRY(walloc(&m->wproj, (size_t)E * CC));
  WTRY(pack_w(m->w[WV_PROJ_W], m->wproj, E, CC, CC, E, CC));
  WTRY(make_tmap(&m->tm_proj, m->wproj, E, CC, 256));
  WTRY(walloc(&m->pcscale, (size_t)128));
  wl::weight_norm_scale_kernel<<<128, 256>>>(m->w[WV_PC_V], m->w[WV_PC_G], m->pcscale, E, 64, 128);
  WTRY(walloc(&m->wpc, (size_t)E * 64 * 128));
  wl::pack_conv_w_kernel<<<592, 256>>>(m->w[WV_PC_V], m->wpc, E, 64, 128, m->pcscale);
  WTRY(make_tmap(&m->tm_pc, m->wpc, E, 64 * 128, 64));
  m->wqkv.resize(NLAYER); m->wo.resize(NLAYER); m->wfc1.resize(NLAYER); m->wfc2.resize(NLAYER); m->bqkv.resize(NLAYER); m->wab.resize(NLAYER);
  m->tm_qkv.resize(NLAYER); m->tm_o.resize(NLAYER); m->tm_fc1.resize(NLAYER); m->tm_fc2.resize(NLAYER);
  std::vector<float> gl(8 * 64), gb(8), wab(130);
  for (int l = 0; l < NLAYER; ++l) {
    float* const* w = &m->w[WV_LAYER0 + WL_PER_LAYER * l];
    WTRY(walloc(&m->wqkv[l], (size_t)3 * E * E));
    WTRY(pack_w(w[WL_Q_W], m->wqkv[l], E, E, E, E, E));
    WTRY(pack_w(w[WL_K_W], m->wqkv[l] + (size_t)E * E, E, E, E, E, E));
    WTRY(pack_w(w[WL_V_W], m->wqkv[l] + (size_t)2 * E * E, E, E, E, E, E));
    WTRY(walloc(&m->bqkv[l], (size_t)3 * E));
    cudaMemcpy(m->bqkv[l], w[WL_Q_B], E * 4, cudaMemcpyDeviceToDevice);
    cudaMemcpy(m->bqkv[l] + E, w[WL_K_B], E * 4, cudaMemcpyDeviceToDevice);
    cudaMemcpy(m->bqkv[l] + 2 * E, w[WL_V_B], E * 4, cudaMemcpyDeviceToDevice);
    WTRY(walloc(&m->wo[l], (size_t)E * E));     WTRY(pack_w(w[WL_O_W], m->wo[l], E, E, E, E, E));
    WTRY(walloc(&m->wfc1[l], (size_t)FF * E));  WTRY(pack_w(w[WL_FC1_W], m->wfc1[l], FF, E, E, FF, E));
    WTRY(walloc(&m->wfc2[l], (size_t)E * FF));  WTRY(pack_w(w[WL_FC2_W], m->wfc2[l], E, FF, FF, E, FF));
    WTRY(make_tmap(&m->tm_qkv[l], m->wqkv[l], 3 * E, E, 256));
    WTRY(make_tmap(&m->tm_o[l], m->wo[l], E, E, 256));
    WTRY(make_tmap(&m->tm_fc1[l], m->wfc1[l], FF, E, 256));
    WTRY(make_tmap(&m->tm_fc2[l], m->wfc2[l], E, FF, 256));
    // the two 4-row sums of grep_linear (modules_WavLM.py:527-529: .view(..., 2, 4).sum(-1))
    cudaMemcpy(gl.data(), w[WL_GL_W], 8 * 64 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(gb.data(), w[WL_GL_B], 8 * 4, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 64; ++i) {
      wab[i] = gl[0 * 64 + i] + gl[1 * 64 + i] + gl[2 * 64 + i] + gl[3 * 64 + i];
      wab[64 + i] = gl[4 * 64 + i] + gl[5 * 64 + i] + gl[6 * 64 + i] + gl[7 * 64 + i];
    }
    wab[128] = gb[0] + gb[1] + gb[2] + gb[3]; wab[129] = gb[4] + gb[5] + gb[6] + gb[7];
    WTRY(walloc(&m->wab[l], (size_t)132));
    cudaMemcpy(m->wab[l], wab.data(), 130 * 4, cudaMemcpyHostToDevice);
  }
  WTRY(walloc(&m->posbias, (size_t)H * Lf * Lf));
  if (cudaMemcpy(m->posbias, pos_bias, (size_t)H * Lf * Lf * 4, cudaMemcpyDefault) != cudaSuccess) {
    dsg_wavlm_destroy(m); return dsg_fail(DSG_ERR_CUDA, "copy of pos_bias failed"); }
  // ---- activations
  const size_t a0 = (size_t)MB * m->L[0] * CC;
  WTRY(walloc(&m->act[0], a0 + 8192)); WTRY(walloc(&m->act[1], (size_t)MB * m->L[1] * CC + 8192));
  WTRY(walloc(&m->convf, (size_t)MB * m->L[1] * CC));
  WTRY(walloc(&m->wav, (size_t)MB * n_samples));
  WTRY(walloc(&m->hbf, (size_t)m->Mpad * E));
  WTRY(walloc(&m->qkv, (size_t)m->Mpad * 3 * E));
  WTRY(walloc(&m->att, (size_t)m->Mpad * E));
  WTRY(walloc(&m->ffb, (size_t)m->Mpad * FF));
  WTRY(walloc(&m->x, (size_t)m->Mpad * E));
  WTRY(walloc(&m->lnout, (size_t)m->Mpad * E));
  WTRY(walloc(&m->gate, (size_t)MB * H * Lf));
  WTRY(walloc(&m->xg, (size_t)MB * 16 * m->Lp * 64 + 8192));
  WTRY(walloc(&m->outbuf, (size_t)MB * Lf * E));
  // conv i reads act[(i-1)&1] as rows of k*512 contiguous elements starting every s*512 elements (channels-last im2col view)
  for (int i = 1; i < NCONV; ++i)
    WTRY(make_tmap3(&m->tm_conv_a[i], m->act[(i - 1) & 1], (uint64_t)CONV_K[i] * CC, m->L[i], MB, (uint64_t)CONV_S[i] * CC,
                    (uint64_t)m->L[i - 1] * CC, BM));
  WTRY(make_tmap(&m->tm_feat, m->hbf, m->Mpad, CC, BM));            // LayerNorm'ed conv features [M, 512] (uses the head of hbf)
  WTRY(make_tmap3(&m->tm_xg, m->xg, (uint64_t)128 * 64, Lf, (uint64_t)MB * 16, 64, (uint64_t)m->Lp * 64, BM));
  WTRY(make_tmap(&m->tm_h, m->hbf, m->Mpad, E, BM));
  WTRY(make_tmap(&m->tm_att, m->att, m->Mpad, E, BM));
  WTRY(make_tmap(&m->tm_ff, m->ffb, m->Mpad, FF, BM));
  m->smem_fa = (size_t)3 * 224 * wl::FA_LD * sizeof(bf16);
  CUDA_TRY(cudaFuncSetAttribute(wl::flash_attn_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_fa));
  if (cudaDeviceSynchronize() != cudaSuccess) {
    const int rc = dsg_fail(DSG_ERR_CUDA, "WavLM set-up kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
    dsg_wavlm_destroy(m); return rc; }
  *out = m;
  return DSG_OK;
}

template <typename TIn, typename TOut, int C, bool GELU>
static int launch_ln(dsg_wavlm* m, const TIn* in, TOut* out, const float* g, const float* b, long long rows, cudaStream_t st) {
  const long long blocks = (rows * 32 + 255) / 256;
  wl::ln_rows_kernel<TIn, TOut, C, GELU><<<(unsigned)blocks, 256, 0, st>>>(in, out, g, b, rows);
  m->eng.launches++;
  CUDA_TRY(cudaGetLastError());
  return DSG_OK;
}

static int wavlm_run(dsg_wavlm* m, int B, const float* wav_d, int n_poses, float* out_d, cudaStream_t st) {
  dsg_engine* e = &m->eng;
  const int Lf = m->L[NCONV - 1], M = B * Lf;
  TcEpiArgs z;
  memset(&z, 0, sizeof z);
  if (m->xout_rows != M) {               // out_proj / fc2 accumulate into x through TMA reduce-adds: clip at this batch's rows
    TRY(make_tmap_f32_out(&m->tm_xout, m->x, (uint64_t)M, E, E));
    m->xout_rows = M;
  }
  // ---- conv feature extractor
  wl::conv0_ln_gelu_kernel<<<m->eng.num_sms * 8, 256, 0, st>>>(wav_d, m->act[0], m->w[WV_CONV0], m->w[WV_CONV0 + 1], m->w[WV_CONV0 + 2], B, m->N, m->L[0]);
  e->launches++;
  CUDA_TRY(cudaGetLastError());
  for (int i = 1; i < NCONV; ++i) {
    TcEpiArgs a = z;
    a.M = m->L[i]; a.N = CC; a.K = CONV_K[i] * CC; a.rows_per_z = m->L[i]; a.z_div = 1; a.out = m->convf; a.ldc = CC;
    TRY((launch_tc<256, 4, EPI_F32>(e, m->tm_conv_a[i], m->tm_conv_w[i], a, CC / 256, st, B)));
    TRY((launch_ln<float, bf16, CC, true>(m, m->convf, m->act[i & 1], m->w[WV_CONV0 + 3 * i + 1], m->w[WV_CONV0 + 3 * i + 2], (long long)B * m->L[i], st)));
  }
  // ---- WavLM.layer_norm + post_extract_proj  (WavLM.py:341-348)
  TRY((launch_ln<bf16, bf16, CC, false>(m, m->act[(NCONV - 1) & 1], m->hbf, m->w[WV_LN], m->w[WV_LN_B], M, st)));
  {
    TcEpiArgs a = z;
    a.M = M; a.N = E; a.K = CC; a.bias = m->w[WV_PROJ_B]; a.out = m->x; a.ldc = E;
    TRY((launch_tc<256, 4, EPI_F32>(e, m->tm_feat, m->tm_proj, a, E / 256, st)));
  }
  // ---- x += GELU(pos_conv(x))   (WavLM.py:514-527, 577-579)
  {
    const long long tot = (long long)B * Lf * E;
    wl::pack_posconv_kernel<<<(unsigned)((tot + 255) / 256 < 4736 ? (tot + 255) / 256 : 4736), 256, 0, st>>>(m->x, m->xg, B, Lf, m->Lp);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    TcEpiArgs a = z;
    a.M = Lf; a.N = E; a.K = 128 * 64; a.rows_per_z = Lf; a.z_div = 16; a.bias = m->w[WV_PC_B]; a.out = m->x; a.ldc = E;
    TRY((launch_tc<64, 4, EPI_PCONV>(e, m->tm_xg, m->tm_pc, a, 1, st, B * 16)));
  }
  // ---- 24 pre-norm layers (WavLM.py:689-714)
  for (int l = 0; l < NLAYER; ++l) {
    float* const* w = &m->w[WV_LAYER0 + WL_PER_LAYER * l];
    TRY((launch_ln<float, bf16, E, false>(m, m->x, m->hbf, w[WL_LN1_W], w[WL_LN1_B], M, st)));
    wl::gate_kernel<<<(unsigned)(((long long)M * H + 255) / 256), 256, 0, st>>>(m->hbf, m->wab[l], w[WL_GA], m->gate, B, Lf, H);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    { TcEpiArgs a = z; a.M = M; a.N = 3 * E; a.K = E; a.bias = m->bqkv[l]; a.out = m->qkv; a.ldc = 3 * E;
      TRY((launch_tc_persistent<EPI_BF16>(e, m->tm_h, m->tm_qkv[l], nullptr, a, st))); }
    wl::flash_attn_bias_kernel<<<B * H * 2, 224, m->smem_fa, st>>>(m->qkv, m->att, m->gate, m->posbias, Lf, E, H);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
    { TcEpiArgs a = z; a.M = M; a.N = E; a.K = E; a.bias = w[WL_O_B]; a.out = m->x; a.ldc = E;
      TRY((launch_tc_persistent<EPI_RESID>(e, m->tm_att, m->tm_o[l], &m->tm_xout, a, st))); }
    TRY((launch_ln<float, bf16, E, false>(m, m->x, m->hbf, w[WL_LN2_W], w[WL_LN2_B], M, st)));
    { TcEpiArgs a = z; a.M = M; a.N = FF; a.K = E; a.bias = w[WL_FC1_B]; a.out = m->ffb; a.ldc = FF;
      TRY((launch_tc_persistent<EPI_GELU>(e, m->tm_h, m->tm_fc1[l], nullptr, a, st))); }
    { TcEpiArgs a = z; a.M = M; a.N = E; a.K = FF; a.bias = w[WL_FC2_B]; a.out = m->x; a.ldc = E;
      TRY((launch_tc_persistent<EPI_RESID>(e, m->tm_ff, m->tm_fc2[l], &m->tm_xout, a, st))); }
  }
  // ---- encoder.layer_norm (WavLM.py:567-568) and the interpolation to n_poses frames (sample.py:47)
  float* const* wl_end = &m->w[WV_LAYER0 + WL_PER_LAYER * NLAYER];
  if (n_poses > 0) {
    TRY((launch_ln<float, float, E, false>(m, m->x, m->lnout, wl_end[0], wl_end[1], M, st)));
    const long long tot = (long long)B * n_poses * E;
    wl::interp_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(m->lnout, out_d, B, Lf, n_poses, E);
    e->launches++;
    CUDA_TRY(cudaGetLastError());
  } else {
    TRY((launch_ln<float, float, E, false>(m, m->x, out_d, wl_end[0], wl_end[1], M, st)));
  }
  return DSG_OK;
}

static bool wl_is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

extern "C" int dsg_wavlm_forward(dsg_wavlm* m, int32_t batch, const float* wav, int32_t n_poses, float* out, void* stream) {
  if (!m || !wav || !out || batch <= 0 || n_poses < 0 || n_poses > m->L[NCONV - 1])
    return dsg_fail(DSG_ERR_BAD_SHAPE, "dsg_wavlm_forward: bad arguments (n_poses must be in 0..frames)");
  CUDA_TRY(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int Lf = m->L[NCONV - 1];
  const size_t per_out = (size_t)(n_poses > 0 ? n_poses : Lf) * E;
  const bool wav_dev = wl_is_device_ptr(wav), out_dev = wl_is_device_ptr(out);
  for (int b0 = 0; b0 < batch; b0 += m->MB) {                       // sub-batches of at most max_batch clips
    const int B = batch - b0 < m->MB ? batch - b0 : m->MB;
    const float* wd = wav + (size_t)b0 * m->N;
    if (!wav_dev) {
      CUDA_TRY(cudaMemcpyAsync(m->wav, wd, (size_t)B * m->N * sizeof(float), cudaMemcpyHostToDevice, st));
      wd = m->wav;
    }
    float* od = out_dev ? out + (size_t)b0 * per_out : m->outbuf;
    TRY(wavlm_run(m, B, wd, n_poses, od, st));
    if (!out_dev) CUDA_TRY(cudaMemcpyAsync(out + (size_t)b0 * per_out, od, (size_t)B * per_out * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (!out_dev || !wav_dev) CUDA_TRY(cudaStreamSynchronize(st));  // staging buffers are reused by the next sub-batch
  }
  return DSG_OK;
}

extern "C" int32_t dsg_wavlm_frames(const dsg_wavlm* m) { return m ? m->L[NCONV - 1] : -1; }
extern "C" int64_t dsg_wavlm_launch_count(const dsg_wavlm* m) { return m ? m->eng.launches : -1; }
