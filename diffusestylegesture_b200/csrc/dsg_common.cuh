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

// 4 standard normals for elements 4q..4q+3 of a clip tensor.
DSG_DEVINL float4 philox_normal4(uint32_t q, uint32_t draw, uint32_t clip, uint32_t segment, uint32_t k0, uint32_t k1) {
  const Philox4 r = philox4x32_10(q, draw, clip, segment, k0, k1);
  const float two_pi = 6.283185307179586f;
  float4 o;
  {
    const float rad = sqrtf(-2.0f * logf(u01_from_u32(r.x)));
    float s, c; sincosf(two_pi * u01_from_u32(r.y), &s, &c);
    o.x = rad * c; o.y = rad * s;
  }
  {
    const float rad = sqrtf(-2.0f * logf(u01_from_u32(r.z)));
    float s, c; sincosf(two_pi * u01_from_u32(r.w), &s, &c);
    o.z = rad * c; o.w = rad * s;
  }
  return o;
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

// Per-step scalars every step-dependent kernel needs.  Either immediate (k_imm >= 0) or read from a device
// counter (CUDA-graph replay: one graph per step, a 1-thread kernel bumps the counter).
struct StepRef {
  const int* d_k;     // device loop-iteration counter (nullable)
  int k_imm;          // loop iteration k (0 = noisiest step) when d_k == nullptr
  int first_index;    // sampler index of k == 0  (nsteps - skip - 1)
  DSG_DEVINL int k() const { return d_k ? *d_k : k_imm; }
  DSG_DEVINL int index() const { return first_index - k(); }
};
