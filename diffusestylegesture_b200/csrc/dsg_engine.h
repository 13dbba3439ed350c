// Internal engine state shared by dsg_engine.cu (fp32 path, C ABI) and dsg_tc.cu (tcgen05 path).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dsg.h"
#include <cuda_bf16.h>
#include "dsg_common.cuh"

int dsg_fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t err__ = (expr);                                                                         \
    if (err__ != cudaSuccess) {                                                                         \
      cudaGetLastError();                                                                               \
      return dsg_fail(DSG_ERR_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(err__)); \
    }                                                                                                   \
  } while (0)
#define TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

// index of each tensor in the `weights` array (== diffusestylegesture_b200/config.py:state_dict_spec order)
enum {
  W_AUD_W = 0, W_AUD_B, W_POSE_W, W_POSE_B, W_IN2_W, W_IN2_B, W_T0_W, W_T0_B, W_T2_W, W_T2_B,
  W_STY_W, W_STY_B, W_TXT_W, W_TXT_B, W_OUT_W, W_OUT_B, W_LAYER0
};
enum { L_INPROJ_W = 0, L_INPROJ_B, L_OUTPROJ_W, L_OUTPROJ_B, L_FF1_W, L_FF1_B, L_FF2_W, L_FF2_B, L_N1_W, L_N1_B, L_N2_W, L_N2_B };

struct Stage { void* p = nullptr; size_t cap = 0; };

// kernel classes for the optional per-kernel event timing (dsg_profile)
enum ProfTag { PT_GEMM_IN = 0, PT_LOCAL_ATTN, PT_GEMM_QKV, PT_SELF_ATTN, PT_GEMM_OUTPROJ, PT_LAYERNORM, PT_GEMM_FF1,
               PT_GEMM_FF2, PT_GEMM_HEAD, PT_POSTERIOR, PT_NOISE, PT_OTHER, PT_COUNT };
struct ProfSpan { int tag; cudaEvent_t a, b; };
struct dsg_tc_state;

struct dsg_engine {
  dsg_model_desc d;
  int S = 0, num_sms = 0;
  // weights (fp32 slab) and derived tables
  float* wslab = nullptr;
  std::vector<float*> w;
  float *te = nullptr, *TW = nullptr, *Wxp = nullptr, *bxp = nullptr;
  float2* cs_local = nullptr;
  // schedule
  int sampler = 0, nsteps = 0;
  float4* coef = nullptr;
  int* tmap = nullptr;
  std::vector<float> qsample;
  // conditioning
  float *emb1 = nullptr, *cvec = nullptr, *enc = nullptr, *cond = nullptr;
  int cond_batch = 0;
  // workspace (fp32 path)
  float *h = nullptr, *xs = nullptr, *qkv = nullptr, *att = nullptr, *ff = nullptr, *tmp = nullptr, *x0 = nullptr;
  int* tsel = nullptr;
  long long* clip_ids = nullptr;     // keys of the x_T draw (one stream per clip)
  long long* noise_ids = nullptr;    // keys of the per-step draws: == clip_ids, or clip_ids[0] everywhere (const_noise)
  std::vector<long long> h_clip_ids;
  float* plms_buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // PLMS: 4 eps slots, mean_pred, second x0 (lazy)
  float *pe_d = nullptr, *l1 = nullptr;     // create-time scratch (freed by destroy on every path)
  LoopParams* d_loop = nullptr;      // device loop state for graph replay
  int* d_k = nullptr;
  size_t smem_self = 0, smem_local = 0;
  int local_threads = 128;
  // debug taps
  bool debug = false;
  float* dbg = nullptr;
  // host-pointer staging
  Stage stage[8];
  int64_t launches = 0;
  // per-kernel event timing (off by default; never on inside a timed bench region)
  bool profiling = false;
  std::vector<ProfSpan> prof_spans;
  size_t prof_used = 0;
  double prof_ms[PT_COUNT] = {0};
  int64_t prof_n[PT_COUNT] = {0};
  // tensor-core path
  dsg_tc_state* tc = nullptr;
  bool graph_valid = false;
};

// RAII-less span helpers: PROF(e, tag, st, launch-expr)
int dsg_prof_begin(dsg_engine* e, int tag, cudaStream_t st);
void dsg_prof_end(dsg_engine* e, cudaStream_t st);
#define PROF(e, tag, st, expr)                                   \
  do {                                                           \
    if ((e)->profiling) dsg_prof_begin((e), (tag), (st));        \
    int rc_p__ = (expr);                                         \
    if ((e)->profiling) dsg_prof_end((e), (st));                 \
    if (rc_p__) return rc_p__;                                   \
  } while (0)

struct GemmF32Args;
int launch_gemm_f32(dsg_engine* e, const GemmF32Args& g, bool a_m_contig, bool swap_mn, cudaStream_t st);
int launch_layernorm(dsg_engine* e, const float* in, float* out, const float* gamma, const float* beta, int rows, int D,
                     cudaStream_t st);
int launch_local_attention(dsg_engine* e, int B, const float* h, long long h_clip_stride, int h_row0, float* xs,
                           __nv_bfloat16* xsb, const int* tsel, StepRef step, cudaStream_t st);
int launch_self_attention(dsg_engine* e, int B, const float* qkv, float* out, cudaStream_t st);
int launch_self_attention_bf16(dsg_engine* e, int B, const __nv_bfloat16* qkv, __nv_bfloat16* out, cudaStream_t st);
int elementwise_grid(const dsg_engine* e, long long quads);
int dsg_upload_loop_params(dsg_engine* e, int k0, int first_index, uint64_t seed, int segment, cudaStream_t st);

// Entry points run on the engine's device and restore the caller's current device on return.
struct DeviceGuard {
  int prev = -1, cur = -1;
  explicit DeviceGuard(int dev) : cur(dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); }
  ~DeviceGuard() { if (prev >= 0 && prev != cur) cudaSetDevice(prev); }
};
int launch_posterior(dsg_engine* e, int B, float* x, const float* x0, StepRef step, int index_imm, int draw_imm,
                     uint64_t seed, int segment, cudaStream_t st);
int dsg_denoise_step(dsg_engine* e, int B, const float* x, const int* tsel, StepRef step, float* out, cudaStream_t st);
// loop iterations k0 .. k0 + n_run - 1 (iteration k: sampler index first_index - k, noise draw 1 + k)
int dsg_run_steps(dsg_engine* e, int B, float* xd, int k0, int n_run, int first_index, uint64_t seed, int segment, cudaStream_t st);

// tensor-core path (dsg_tc.cu)
int dsg_tc_create(dsg_engine* e);
void dsg_tc_destroy(dsg_engine* e);
const float* dsg_tc_h(dsg_engine* e);
const long long* dsg_tc_clip_prof(dsg_engine* e);
int dsg_tc_denoise(dsg_engine* e, int B, const float* x, const int* tsel, StepRef step, float* out, cudaStream_t st);
int dsg_tc_run_steps(dsg_engine* e, int B, float* xd, int k0, int n_run, int first_index, uint64_t seed, int segment, cudaStream_t st);
