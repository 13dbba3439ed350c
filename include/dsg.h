/* dsg.h — C ABI of libdsg.so, the B200 (sm_100a) sampling engine for the DiffuseStyleGesture hot path.
 *
 * The reference is 100 % Python and has NO FFI; its "plugin interface" for this path is the pair
 *   model(x, timesteps, y=...)                     reference main/model/mdm.py:166
 *   diffusion.p_sample_loop(model, shape, ...)     reference main/diffusion/gaussian_diffusion.py:608
 * reached from sample.py (reference main/mydiffusion_zeggs/sample.py:51-56, 253-264, 376).  Each entry
 * point below names the reference code it replaces.  The Python host mirror of that interface
 * (diffusestylegesture_b200/{mdm,gaussian_diffusion,respace,sample}.py) binds these symbols with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - Plain pointers and sizes only.  Every data pointer may be a HOST or a DEVICE pointer (unified
 *     addressing: the engine inspects it with cudaPointerGetAttributes).  Host buffers are staged through
 *     engine-owned device memory with cudaMemcpyAsync on `stream` (pinned host memory makes that truly
 *     asynchronous); device buffers are used in place.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is enqueued on
 *     it; no entry point synchronises the device unless it must return data into a HOST buffer.
 *   - Every function returns DSG_OK (0) or a negative dsg_status; dsg_last_error() gives the message for
 *     the calling thread.  There is no CPU fallback: without an sm_100 device dsg_engine_create fails
 *     with DSG_ERR_BAD_ARCH.
 *   - Layouts: x / seed are the reference's [B, njoints, 1, frames] fp32 (frames innermost);
 *     audio is [B, audio_frames, audio_dim]; style is [B, style_in].
 *   - An engine is bound to one device and is not thread-safe (one engine per GPU / per rank).  Entry points
 *     switch to the engine's device for the duration of the call and restore the caller's current device.
 */
#ifndef DSG_H_
#define DSG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsg_engine dsg_engine;

typedef enum dsg_status {
  DSG_OK = 0,
  DSG_ERR_BAD_SHAPE = -1,    /* a size/argument outside what the descriptor allows               */
  DSG_ERR_BAD_ARCH = -2,     /* no CUDA device of compute capability 10.x                         */
  DSG_ERR_UNSUPPORTED = -3,  /* a reference option the engine does not implement                  */
  DSG_ERR_CUDA = -4,         /* a CUDA runtime/driver call failed (message has the CUDA string)   */
  DSG_ERR_STATE = -5         /* call order: schedule / conditioning not set                       */
} dsg_status;

enum { DSG_PRECISION_FP32 = 0,   /* fp32 CUDA-core kernels: the validation path (matches the fp32 reference to ~1e-5) */
       DSG_PRECISION_BF16 = 1 }; /* bf16 operands on tcgen05 tensor cores, fp32 accumulate / residual / LN / posterior  */
enum { DSG_SAMPLER_DDPM = 0,     /* GaussianDiffusion.p_sample       (gaussian_diffusion.py:506-558)  */
       DSG_SAMPLER_DDIM = 1,     /* GaussianDiffusion.ddim_sample, eta = 0 (gaussian_diffusion.py:742-792) */
       DSG_SAMPLER_PLMS = 2 };   /* GaussianDiffusion.plms_sample, order 2..4 (gaussian_diffusion.py:1005-1103); DDIM coefficient rows */
enum { DSG_VARIANT_ATTN3 = 3,    /* cond_mode cross_local_attention3_style1 (main/model/mdm.py:194-233)  */
       DSG_VARIANT_ATTN4 = 4,    /* cond_mode cross_local_attention4_style1 (BEAT-TWH-main/model/mdm.py:187-224) */
       DSG_VARIANT_ATTN5 = 5 };  /* cond_mode cross_local_attention5_style1 ("++": + seed_last / embed_text_last, :226-264);
                                    its two extra tensors embed_text_last.{weight,bias} come LAST in the weight list */

/* Geometry of MDM(...) — the constructor arguments of reference main/model/mdm.py:11-151
 * (BEAT-TWH-main/model/mdm.py:11-118 for the "+" variant). */
typedef struct dsg_model_desc {
  int32_t variant;       /* DSG_VARIANT_*                                          */
  int32_t njoints;       /* J: njoints * nfeats (nfeats == 1)                     */
  int32_t n_poses;       /* T: frames per segment                                  */
  int32_t n_seed;        /* seed frames                                            */
  int32_t latent_dim;    /* D                                                      */
  int32_t ff_size;       /* F                                                      */
  int32_t num_layers;    /* L                                                      */
  int32_t num_heads;     /* global self-attention heads                            */
  int32_t local_heads;   /* MDM.num_head == 8                                      */
  int32_t local_window;  /* LocalAttention window_size                             */
  int32_t audio_dim;     /* width of y['audio']                                    */
  int32_t audio_latent;  /* A: WavEncoder output width                             */
  int32_t style_in;      /* width of y['style']                                    */
  int32_t style_latent;  /* embed_style output width                               */
  int32_t num_timesteps; /* rows of the timestep-embedding table (original diffusion steps, 1000) */
  int32_t max_batch;     /* workspace is sized for this many clips                 */
  int32_t precision;     /* DSG_PRECISION_*                                        */
  int32_t device;        /* CUDA device ordinal                                    */
} dsg_model_desc;

/* Replaces MDM.__init__ + load_model_wo_clip + model.to(device) (mdm.py:11-151, main/utils/model_util.py:8-12,
 * sample.py:369-374).  `weights[i]` points to the fp32 tensor i of the ordered parameter list
 * (diffusestylegesture_b200/config.py:state_dict_spec — the reference state_dict keys); `pe` is the
 * PositionalEncoding buffer rows [num_timesteps, D] (mdm.py:377-384).  The engine copies / repacks
 * everything; the caller keeps ownership of its buffers. */
int dsg_engine_create(const dsg_model_desc* desc, const float* const* weights, int32_t n_weights,
                      const float* pe, dsg_engine** out);
void dsg_engine_destroy(dsg_engine* e);

/* Replaces the float64->float32 table lookups of _extract_into_tensor (gaussian_diffusion.py:1607-1620) and
 * _WrappedModel's timestep map (respace.py:117-129).  coef is [nsteps][4] fp32, row i (the sampler's index):
 *   DDPM: { posterior_mean_coef1[i], posterior_mean_coef2[i], exp(0.5*posterior_log_variance_clipped[i]), 0 }
 *   DDIM: { sqrt_recip_alphas_cumprod[i], sqrt_recipm1_alphas_cumprod[i], sqrt(abar_prev[i]), sqrt(1-abar_prev[i]) }
 * qsample is [nsteps][2] = { sqrt_alphas_cumprod[i], sqrt_one_minus_alphas_cumprod[i] } (q_sample, :236-254).
 * timestep_map[i] is the original timestep fed to the denoiser.  Host pointers. */
int dsg_set_schedule(dsg_engine* e, int32_t sampler, int32_t nsteps, const float* coef, const float* qsample,
                     const int32_t* timestep_map);

/* Step-invariant part of MDM.forward, run once per segment instead of once per step
 * (mdm.py:180-183, 190 and the token/audio column blocks of input_process2, mdm.py:202-206). */
int dsg_set_conditioning(dsg_engine* e, int32_t batch, const float* style, const float* seed,
                         const float* audio, void* stream);

/* MDM.forward(x, timesteps, y) for the conditioning last set (mdm.py:166-358): out = predicted x_0.
 * `timesteps` are ORIGINAL timestep ids, one per clip (host pointer, int32). */
int dsg_denoise(dsg_engine* e, int32_t batch, const float* x, const int32_t* timesteps, float* out, void* stream);

/* One posterior / add-noise update on caller buffers (p_sample after the model call, gaussian_diffusion.py:
 * 264-271, 542-557; ddim_sample :768-791): x <- f(x0, x, z).  `index` = sampler index i, `draw` = noise draw
 * number.  Unit-test and profiling entry for the HBM-bound kernel. */
int dsg_posterior_step(dsg_engine* e, int32_t batch, float* x, const float* x0, int32_t index, uint64_t seed,
                       const int64_t* clip_ids, int32_t segment, int32_t draw, void* stream);

/* GaussianDiffusion.p_sample_loop / ddim_sample_loop (gaussian_diffusion.py:608-740, 889-1003) for the
 * conditioning last set.  x: [batch, J, 1, T] fp32, in/out.  If noise_given == 0 the engine draws x_T itself
 * (draw 0 of the counter-based stream: Philox4x32-10, key = seed, counter = (element/4, draw, clip, segment));
 * otherwise x holds the caller's `noise`.  init_image (nullable) and skip_timesteps follow :706-713.
 * clip_ids (host, nullable = 0..batch-1; each 0 <= id < 2^32) key the noise stream so results do not depend on sharding. */
int dsg_sample_loop(dsg_engine* e, int32_t batch, float* x, int32_t noise_given, uint64_t seed,
                    const int64_t* clip_ids, int32_t segment, int32_t skip_timesteps,
                    const float* init_image, void* stream);

/* dsg_sample_loop with the remaining options of the reference loops:
 *   flags & DSG_LOOP_CONST_NOISE  const_noise=True (gaussian_diffusion.py:544-545): every clip receives clip 0's step noise
 *                                 (x_T stays per clip: th.randn(*shape), :704);
 *   dump_iters / dump_out         dump_steps (gaussian_diffusion.py:647-669): a copy of the sample after loop iteration i
 *                                 (0 = the first, noisiest step) for each listed i, ascending; dump_out is
 *                                 [n_dump, batch, J, 1, T] fp32, host or device;
 *   plms_order                    order of plms_sample_loop (DSG_SAMPLER_PLMS schedule only; the reference default is 2,
 *                                 order 1 fails inside the reference on its first step and is rejected here).
 * opts == NULL is dsg_sample_loop. */
enum { DSG_LOOP_CONST_NOISE = 1 };
typedef struct dsg_loop_opts {
  int32_t flags;
  int32_t plms_order;
  int32_t n_dump;
  const int32_t* dump_iters;   /* host */
  float* dump_out;
} dsg_loop_opts;
int dsg_sample_loop_ex(dsg_engine* e, int32_t batch, float* x, int32_t noise_given, uint64_t seed,
                       const int64_t* clip_ids, int32_t segment, int32_t skip_timesteps,
                       const float* init_image, const dsg_loop_opts* opts, void* stream);

/* dsg_set_conditioning for DSG_VARIANT_ATTN5: seed_last [batch, J, 1, n_seed] is y['seed_last'] (BEAT-TWH-main/model/mdm.py:229);
 * audio covers n_poses - 2 n_seed frames.  seed_last == NULL is dsg_set_conditioning. */
int dsg_set_conditioning_ex(dsg_engine* e, int32_t batch, const float* style, const float* seed,
                            const float* audio, const float* seed_last, void* stream);

/* Batched form of the segment hand-off in inference() (sample.py:266-288): root-position shift and the
 * first-frame 1/2-1/2 blend (the reference's `len(last_poses)` quirk: n == 1 per clip).
 * prev_tail [batch, J, 1, n_seed], sample [batch, J, 1, T] (in/out). */
int dsg_stitch_segment(dsg_engine* e, int32_t batch, const float* prev_tail, float* sample, int32_t smoothing,
                       void* stream);

/* Introspection for tests / bench: number of kernels the engine has launched since creation. */
int64_t dsg_kernel_launch_count(const dsg_engine* e);
/* Copy an internal activation of the last dsg_denoise call to `dst` (host or device), for per-op parity
 * tests: name in {"tok","h_in","h_local","xs0","xs1",...,"xsL"}; returns element count or negative status. */
int64_t dsg_debug_read(dsg_engine* e, const char* name, int32_t batch, float* dst, int64_t capacity);

/* Optional per-kernel-class CUDA-event timing (bench.py's roofline leg; never enable it inside a timed region:
 * it brackets every launch with two event records).  enable: 1 = start (resets totals), 0 = stop.
 * dsg_profile_read returns launches and summed device ms for class `tag` (0 <= tag, name != NULL). */
int dsg_profile(dsg_engine* e, int32_t enable);
int dsg_profile_read(dsg_engine* e, int32_t tag, int64_t* count, double* total_ms);
const char* dsg_profile_tag_name(int32_t tag);

/* ---- WavLM-Large conditioning front-end (reference main/mydiffusion_zeggs/sample.py:30-48, WavLM/WavLM.py:323-375) ----
 * Replaces wavlm_init + wav2wavlm: `weights[i]` = fp32 tensor i of diffusestylegesture_b200/wavlm_config.py:
 * wavlm_state_dict_spec (the reference WavLM state_dict keys); `pos_bias` = compute_bias(L, L) of layer 0
 * ([heads, L, L] fp32, L = frames for n_samples; modules_WavLM.py:444-455), evaluated once by the host mirror.
 * dsg_wavlm_forward: wav [batch, n_samples] fp32 (host or device) -> out [batch, n_poses, 1024] (features linearly
 * interpolated to n_poses frames, align_corners = True), or [batch, L, 1024] when n_poses == 0 (extract_features()[0]).
 * Batches larger than max_batch are processed in sub-batches. */
typedef struct dsg_wavlm dsg_wavlm;
int dsg_wavlm_create(int32_t device, int32_t max_batch, int32_t n_samples, const float* const* weights, int32_t n_weights,
                     const float* pos_bias, dsg_wavlm** out);
int dsg_wavlm_forward(dsg_wavlm* m, int32_t batch, const float* wav, int32_t n_poses, float* out, void* stream);
int32_t dsg_wavlm_frames(const dsg_wavlm* m);
int64_t dsg_wavlm_launch_count(const dsg_wavlm* m);
void dsg_wavlm_destroy(dsg_wavlm* m);

/* Stand-alone check of the tcgen05 GEMM building block (tests): C[M,N] = bf16(A[M,K]) * bf16(W[N,K])^T + bias, fp32
 * accumulate and output.  bn = 128 or 256 (UMMA N); K a multiple of 8.  Host or device pointers. */
int dsg_selftest_gemm(int32_t device, int32_t bn, int32_t M, int32_t N, int32_t K, const float* A, const float* W,
                      const float* bias, float* C);

const char* dsg_last_error(void);
const char* dsg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DSG_H_ */
