// Shared device helpers: counter-based noise stream, small math.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DSG_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) — the shared noise stream of the engine and of oracle/dsg_oracle.py
// (philox4x32_10 / philox_normal there).  key = 64-bit seed; counter = (element/4, draw, clip, segment).
// ---------------------------------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

DSG_DEVINL Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1;
    c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

DSG_DEVINL float u01_from_u32(uint32_t r) {            // (0,1): top 24 bits, centred — exact in fp32
  return (float)(r >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f;
}

// 4 standard normals for elements 4q..4q+3 of a clip tensor: Box-Muller on the four Philox outputs,
//   z0 = r cos(2 pi u2), z1 = r sin(2 pi u2), r = sqrt(-2 ln u1)   (same for outputs 2,3).
// B200-first arithmetic: ln through MUFU.LG2 (__logf) and sin/cos through MUFU.SIN/COS on the angle shifted into
// (-pi, pi) where the approximation is tight (cos(2 pi u) = -cos(2 pi u - pi)).  ~18 instructions per normal
// instead of ~40 with libm-accurate logf/sincosf; |error| vs the exact definition (oracle/dsg_oracle.py:
// philox_normal) is <= ~1e-6 typically, up to ~4e-5 for the rare u1 -> 1 (tiny radius) draws.
DSG_DEVINL float2 box_muller(uint32_t a, uint32_t b) {
  const float rad = sqrtf(-2.0f * __logf(u01_from_u32(a)));
  float s, c;
  __sincosf(fmaf(6.283185307179586f, u01_from_u32(b), -3.14159265358979f), &s, &c);
  return make_float2(-rad * c, -rad * s);
}
DSG_DEVINL float4 philox_normal4(uint32_t q, uint32_t draw, uint32_t clip, uint32_t segment, uint32_t k0, uint32_t k1) {
  const Philox4 r = philox4x32_10(q, draw, clip, segment, k0, k1);
  const float2 p0 = box_muller(r.x, r.y), p1 = box_muller(r.z, r.w);
  return make_float4(p0.x, p0.y, p1.x, p1.y);
}

DSG_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
DSG_DEVINL float silu(float x) { return x / (1.0f + expf(-x)); }

DSG_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DSG_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Per-step scalars every step-dependent kernel needs.  Either immediates, or read from a small device struct
// (CUDA-graph replay: one captured step is replayed n times; a 1-thread kernel bumps `k` at the end of each).
struct LoopParams { int k; int first_index; uint32_t key0, key1, segment; int pad[3]; };
struct StepRef {
  const LoopParams* d;   // device loop state (nullable)
  int k_imm;             // loop iteration k (0 = noisiest step) when d == nullptr
  int first_imm;         // sampler index of k == 0  (nsteps - skip - 1)
  uint32_t key0_imm, key1_imm, seg_imm;
  DSG_DEVINL int k() const { return d ? d->k : k_imm; }
  DSG_DEVINL int index() const { return (d ? d->first_index : first_imm) - k(); }
  DSG_DEVINL uint32_t key0() const { return d ? d->key0 : key0_imm; }
  DSG_DEVINL uint32_t key1() const { return d ? d->key1 : key1_imm; }
  DSG_DEVINL uint32_t segment() const { return d ? d->segment : seg_imm; }
};
