// tcgen05 GEMM for sm_100a:  D[TMEM, fp32] = A[smem, bf16, K-major] * W[smem, bf16, K-major]^T
//   - operands staged by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) through a STAGES-deep mbarrier ring,
//   - one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, cta_group::1), accumulator in tensor memory,
//   - four epilogue warps read the accumulator with tcgen05.ld (32 lanes x 32 columns per instruction) and
//     apply the fused epilogue of the Linear they stand for (see TcEpi below).
// Warp roles (256 threads): 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 3 = idle, 4..7 = epilogue
// (epilogue warp w owns TMEM lanes 32*(w%4) .. +31 = accumulator rows).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "dsg_common.cuh"

namespace tc {

constexpr int BM = 128;          // UMMA_M (cta_group::1)
constexpr int BK = 64;           // 64 bf16 = 128 B = one swizzle-128B row
constexpr int UMMA_K = 16;       // fixed for 16-bit inputs

enum TcEpi {
  EPI_F32 = 0,      // C fp32 [M, ldc] = acc + bias                                       (self-test / generic)
  EPI_IN = 1,       // h[b,s,:] = acc + cond[b,s-1,:] + TW[t,:]     InputProcess+input_process2 (mdm.py:196-206)
  EPI_BF16 = 2,     // bf16 [M, ldc] = acc + bias                    in_proj of self-attention
  EPI_GELU = 3,     // bf16 [M, ldc] = gelu_erf(acc + bias)          linear1 + activation
  EPI_LN = 4,       // xs = LayerNorm(acc + bias + xs) -> fp32 + bf16 copies   out_proj / linear2 + norm (post-norm layer)
  EPI_HEAD = 5,     // x0 = acc + bias; posterior update of x (fp32 [B,J,T]) + bf16 repack [B,S,Jpad]   OutputProcess + p_sample
  EPI_RESID = 6,    // fp32 [M, ldc] += acc + bias                 out_proj / fc2 of a pre-norm layer (WavLM)
  EPI_PCONV = 7     // fp32 [M, ldc] += gelu(acc + bias)           grouped positional conv (WavLM)
};

struct TcEpiArgs {
  int M, N, K;                 // logical GEMM sizes (N = valid output columns; tile columns beyond N are skipped)
  // batched (3-D A map) form: blockIdx.z = batch * z_div + group; rows m < rows_per_z of that batch; output row =
  // batch * rows_per_z + m; the group selects B rows / output columns [group * BN, +BN).  rows_per_z == 0: flat 2-D GEMM.
  int rows_per_z, z_div;
  const float* bias;           // [N]
  // EPI_F32 / EPI_BF16 / EPI_GELU
  void* out; int ldc;
  // EPI_IN
  const float* cond; const float* TW; const int* tsel; const int* tmap; StepRef step; int S, T;
  // EPI_LN
  float* xs; __nv_bfloat16* xsb; const float* gamma; const float* beta;
  // EPI_HEAD
  float* x; const float* z; __nv_bfloat16* xb; int J, Jpad; const float4* coef; int sampler; int head_mode;  // 0 = posterior, 1 = write x0 to `out`
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
DSG_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DSG_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
DSG_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// (Measured in round 2: giving try_wait a suspend-time hint of 20 us does not reduce the step time — 307.6 vs 301.7 us — although
// 20 % of the clip kernel's executed warp instructions are these polls; the plain form stays.)
DSG_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
DSG_DEVINL void mbar_arrive_cta(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DSG_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
DSG_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
DSG_DEVINL bool elect_one_lane() {            // one lane of a converged warp (the same lane on every call)
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
DSG_DEVINL void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
DSG_DEVINL void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
DSG_DEVINL void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DSG_DEVINL void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of the accumulator -> 32 registers per thread (thread = row).
// _issue variants do not wait: several loads can be in flight before one tmem_ld_wait().
DSG_DEVINL void tmem_ld32_issue(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
DSG_DEVINL void tmem_ld16_issue(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
DSG_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
DSG_DEVINL void tmem_ld32(uint32_t taddr, float* v) { tmem_ld32_issue(taddr, v); tmem_ld_wait(); }
DSG_DEVINL void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile stored by TMA with SWIZZLE_128B:
// rows of 128 B, 8-row (1024 B) swizzle atoms stacked along M/N.  Fields (cute::UMMA::SmemDescriptor):
// [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major: 1), [32,46) SBO>>4 = 1024>>4, [46,48) version = 1,
// [61,64) layout = 2 (SWIZZLE_128B).
DSG_DEVINL uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor, kind::f16: D = fp32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bits 17-22, M>>4 at bits 24-28 (cute::UMMA::InstrDescriptor).
DSG_DEVINL constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// same with A = B = fp16 (format code 0)
DSG_DEVINL constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// BN <= 256: one UMMA per k-step.  BN = 384 / 512 (a full LayerNorm row of the "+" denoisers, D = 384 / 512): the B tile is
// two TMA boxes of BN/2 rows and every k-step issues two UMMAs of N = BN/2 into adjacent TMEM column ranges.
template <int BN> struct TcTile {
  static constexpr int NSPLIT = BN > 256 ? 2 : 1;
  static constexpr int NSUB = BN / NSPLIT;                                   // UMMA N and TMA box rows of the B operand
  static constexpr int TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : (BN <= 256 ? 256 : 512)));
  static_assert(NSUB % 16 == 0 && NSUB <= 256 && BN <= 512, "UMMA N: multiple of 16, <= 256 (M = 128)");
};
template <int BN, int STAGES>
struct TcSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int RED_OFF = BAR_OFF + 128;                               // LayerNorm partials [2][128][2] fp32
  static constexpr int PRM_OFF = RED_OFF + 2048;                              // this tile's bias | gamma | beta (3 x BN fp32)
  static constexpr int TOTAL = PRM_OFF + 3 * BN * 4 + 1024;                   // + alignment slack
};

DSG_DEVINL float posterior_apply(int sampler, const float4 c, float x0, float xt, float z, bool nz) {
  if (sampler == 0) {
    float r = __fadd_rn(__fmul_rn(c.x, x0), __fmul_rn(c.y, xt));
    if (nz) r = __fadd_rn(r, __fmul_rn(c.z, z));
    return r;
  }
  const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(c.x, xt), x0), c.y);
  return __fadd_rn(__fmul_rn(x0, c.z), __fmul_rn(c.w, eps));
}

// erf to 1.5e-7 absolute (Abramowitz & Stegun 7.1.26) on MUFU.RCP / MUFU.EX2 — ~14 instructions instead of ~30 for
// erff; far below the bf16 rounding of the value it feeds (F.gelu in nn.TransformerEncoderLayer, mdm.py:79-86).
DSG_DEVINL float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = fmaf(-p * t, __expf(-ax * ax), 1.0f);
  return copysignf(r, x);
}
DSG_DEVINL float gelu_fast(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f)); }

DSG_DEVINL void store_bf16x32(__nv_bfloat16* o, const float* v) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[i], v[i + 1]), p1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]), p3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
    u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
    *reinterpret_cast<uint4*>(o + i) = u;
  }
}

// The epilogue runs on ALL 8 warps: warp w and warp w+4 share TMEM lane quarter w%4 (accumulator rows
// 32*(w%4)..+31) and split the BN columns into interleaved 32-column chunks (even chunks: warps 0-3, odd: 4-7).
// One epilogue warp per scheduler is dependent-issue bound; two per scheduler with 32 independent elements each
// keep the ALU pipes busy.
template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(256, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcEpiArgs ep) {
  using SM = TcSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* red = reinterpret_cast<float*>(smem + SM::RED_OFF);     // [2 halves][128 rows][2] LayerNorm partials
  // Per-column parameters of this tile are staged in shared memory once, under the barrier every CTA passes anyway: with the
  // shared-memory carve-out at its maximum the L1 is a few KB, and a `__ldg(bias + c)` per 32-column chunk of the epilogue was
  // one exposed L2 round trip per chunk and pass (ncu: the LayerNorm GEMM of the "+" path spent 25 of its 40 us there).
  float* prm = reinterpret_cast<float*>(smem + SM::PRM_OFF);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;     // warp-uniform for the compiler
  const int zdiv = ep.z_div > 0 ? ep.z_div : 1;
  const int zb = (int)blockIdx.z / zdiv, zg = (int)blockIdx.z - zb * zdiv;
  const int m0 = blockIdx.x * BM, n0 = (blockIdx.y + zg) * BN;
  const int num_kb = (ep.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TcTile<BN>::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    for (int i = threadIdx.x; i < BN; i += blockDim.x) {
      const int n = n0 + i;
      prm[i] = (EPI != EPI_IN && ep.bias != nullptr && n < ep.N) ? __ldg(ep.bias + n) : 0.f;
      if constexpr (EPI == EPI_LN) { prm[BN + i] = __ldg(ep.gamma + i); prm[2 * BN + i] = __ldg(ep.beta + i); }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int NSPLIT = TcTile<BN>::NSPLIT, NSUB = TcTile<BN>::NSUB;

  // Both roles run warp-converged with warp-uniform operands; only the TMA / tcgen05 instruction itself is issued by one
  // elected lane.  (Inside `if (lane == 0)` the compiler treats the operands as divergent and wraps every UTMALDG / UTCHMMA
  // in an R2UR / ELECT / BRA.U.ANY loop of ~20 dependent instructions: the issuing thread becomes the limit.)
  if (warp == 0) {
    // ===== TMA producer =====
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&empty_bar[s], ph ^ 1u);
      uint8_t* a_dst = smem + s * SM::STAGE_BYTES;
      if (elect_one_lane()) {
        mbar_expect_tx(&full_bar[s], SM::STAGE_BYTES);
        if (ep.rows_per_z > 0) tma_load_3d(a_dst, &tmA, &full_bar[s], kb * BK, m0, (int)blockIdx.z);
        else tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
#pragma unroll
        for (int i = 0; i < NSPLIT; ++i)
          tma_load_2d(a_dst + SM::A_BYTES + i * NSUB * 128, &tmB, &full_bar[s], kb * BK, n0 + i * NSUB);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc_bf16(BM, NSUB);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&full_bar[s], ph);
      tcgen05_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * SM::STAGE_BYTES);
      const uint32_t b_addr = a_addr + SM::A_BYTES;
      if (elect_one_lane()) {
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t adesc = make_sw128_desc(a_addr + k * UMMA_K * 2);
#pragma unroll
          for (int i = 0; i < NSPLIT; ++i)
            umma_bf16(tmem_u + i * NSUB, adesc, make_sw128_desc(b_addr + i * NSUB * 128 + k * UMMA_K * 2), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tcgen05_commit(&empty_bar[s]);            // frees the smem stage when these MMAs retire
      }
      __syncwarp();
    }
    if (elect_one_lane()) tcgen05_commit(tmem_full);                  // accumulator complete
  }
  __syncwarp();

  // ===== epilogue (all 8 warps) =====
  mbar_wait(tmem_full, 0);
  tcgen05_fence_after();
  const int wq = warp & 3, half = warp >> 2;
  const int rloc = wq * 32 + lane;
  const int row = ep.rows_per_z > 0 ? zb * ep.rows_per_z + m0 + rloc : m0 + rloc;
  const bool row_ok = ep.rows_per_z > 0 ? (m0 + rloc < ep.rows_per_z) : (row < ep.M);
  const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16);
  float v[32];
  // coalesced epilogues: a 32 x 33 fp32 tile per warp in the (now idle) operand stages; rows of this warp: row_w0 .. +rows_w-1
  float* tile = reinterpret_cast<float*>(smem) + warp * 1088;
  const int row_w0 = row - lane;
  const int rows_w = ep.rows_per_z > 0 ? min(32, max(0, ep.rows_per_z - (m0 + wq * 32))) : min(32, max(0, ep.M - (m0 + wq * 32)));

  if constexpr (EPI == EPI_F32 || EPI == EPI_BF16 || EPI == EPI_GELU || EPI == EPI_RESID || EPI == EPI_PCONV) {
    constexpr bool RMW = (EPI == EPI_RESID || EPI == EPI_PCONV);
    // read-modify-write epilogues: the old value is READ row-per-thread (8 independent 16-byte loads, next chunk one iteration
    // ahead), the new one is WRITTEN coalesced through the warp's tile
    float4 on[8];
    const bool vec_ok = RMW && (ep.ldc & 3) == 0;
    const float* orow = reinterpret_cast<const float*>(ep.out) + (long long)row * ep.ldc + n0;
    if constexpr (RMW) {
      if (vec_ok && row_ok && n0 + half * 32 + 32 <= ep.N) {
#pragma unroll
        for (int i = 0; i < 8; ++i) on[i] = *reinterpret_cast<const float4*>(orow + half * 32 + 4 * i);
      }
    }
#pragma unroll 1
    for (int c = half * 32; c < BN; c += 64) {
      float4 oc[8];
      if constexpr (RMW) {
#pragma unroll
        for (int i = 0; i < 8; ++i) oc[i] = on[i];
        if (vec_ok && row_ok && c + 64 < BN && n0 + c + 64 + 32 <= ep.N) {
#pragma unroll
          for (int i = 0; i < 8; ++i) on[i] = *reinterpret_cast<const float4*>(orow + c + 64 + 4 * i);
        }
      }
      tmem_ld32(taddr + c, v);
      const int n = n0 + c;
      if constexpr (RMW) {
        if (n >= ep.N) continue;                  // warp-uniform (the coalesced path below needs the whole warp)
      } else {
        if (!row_ok || n >= ep.N) continue;
      }
      if (n + 32 <= ep.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(prm + c + i);
          v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (n + i < ep.N) v[i] += prm[c + i];
      }
      if constexpr (EPI == EPI_GELU || EPI == EPI_PCONV) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
      }
      if constexpr (EPI == EPI_RESID || EPI == EPI_PCONV) {
        // out[row, n .. n+31] += v, coalesced: the warp's 32 x 32 block goes through its shared-memory tile and is applied row by
        // row (lane = column: one 128-byte read-modify-write per row instead of 32 scattered 16-byte ones per instruction)
        const bool full = vec_ok && n + 32 <= ep.N;                 // warp-uniform
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 t = oc[i >> 2];
            tile[lane * 33 + i] = v[i] + t.x; tile[lane * 33 + i + 1] = v[i + 1] + t.y;
            tile[lane * 33 + i + 2] = v[i + 2] + t.z; tile[lane * 33 + i + 3] = v[i + 3] + t.w;
          }
          __syncwarp();
          float* ob = reinterpret_cast<float*>(ep.out) + (long long)row_w0 * ep.ldc + n + lane;
#pragma unroll 8
          for (int r = 0; r < 32; ++r)
            if (r < rows_w) ob[(long long)r * ep.ldc] = tile[r * 33 + lane];
          __syncwarp();
        } else if (row_ok) {
          float* o = reinterpret_cast<float*>(ep.out) + (long long)row * ep.ldc + n;
#pragma unroll
          for (int i = 0; i < 32; ++i) if (n + i < ep.N) o[i] += v[i];
        }
      } else if constexpr (EPI == EPI_F32) {
        float* o = reinterpret_cast<float*>(ep.out) + (long long)row * ep.ldc + n;
        if (n + 32 <= ep.N && (ep.ldc & 3) == 0) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (n + i < ep.N) o[i] = v[i];
        }
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row * ep.ldc + n;
        if (n + 32 <= ep.N) {
          store_bf16x32(o, v);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (n + i < ep.N) o[i] = __float2bfloat16_rn(v[i]);
        }
      }
    }
  } else if constexpr (EPI == EPI_IN) {
    // row = b*S + s; s == 0 is the token slot (filled by the local-attention kernel): skipped.
    const int b = row / ep.S, s = row - b * ep.S;
    const bool ok = row_ok && s > 0;
    const float* cond = ep.cond + ((long long)b * ep.T + (s - 1)) * ep.N;
    const int trow = ok ? (ep.tsel ? ep.tsel[b] : ep.tmap[ep.step.index()]) : 0;
    const float* tw = ep.TW + (long long)trow * ep.N;
    float* o = reinterpret_cast<float*>(ep.out) + (long long)row * ep.N;
#pragma unroll 1
    for (int c = half * 32; c < BN; c += 64) {
      tmem_ld32(taddr + c, v);
      const int n = n0 + c;
      if (!ok || n >= ep.N) continue;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 cc = __ldg(reinterpret_cast<const float4*>(cond + n + i));
        const float4 tt = __ldg(reinterpret_cast<const float4*>(tw + n + i));
        *reinterpret_cast<float4*>(o + n + i) =
            make_float4(v[i] + cc.x + tt.x, v[i + 1] + cc.y + tt.y, v[i + 2] + cc.z + tt.z, v[i + 3] + cc.w + tt.w);
      }
    }
  } else if constexpr (EPI == EPI_LN) {
    // full rows live in this CTA (BN == N == D).  Pass 1 adds bias + residual, parks v back in TMEM and accumulates
    // this warp's share of the row statistics; the two column halves meet through shared memory; pass 2 normalises
    // and writes the fp32 residual stream + its bf16 copy.
    // Pass 2 WRITES coalesced: the warp's 32 rows x 32 columns go through its shared-memory tile and leave row by row with
    // lane = column (128- / 64-byte lines).  Row-per-thread stores are 32 scattered 16-byte transactions per instruction: the
    // LSU, one transaction per clock, made pass 2 17 of the 41 us of the "+" LayerNorm GEMM (clock64 instrumentation).
    float sum = 0.f, sq = 0.f;
    float* xw = ep.xs + (long long)row_w0 * ep.N;
    const float* xr = ep.xs + (long long)row * ep.N;
    // pass 1 READS row-per-thread (8 independent 16-byte loads per chunk, the next chunk requested one iteration ahead: loads
    // need memory-level parallelism, and a row-by-row coalesced loop serialises them — measured 41 -> 60 us)
    float4 rn[8];
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 8; ++i) rn[i] = *reinterpret_cast<const float4*>(xr + half * 32 + 4 * i);
    }
#pragma unroll 1
    for (int c = half * 32; c < BN; c += 64) {
      float4 rc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) rc[i] = rn[i];
      if (row_ok && c + 64 < BN) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rn[i] = *reinterpret_cast<const float4*>(xr + c + 64 + 4 * i);
      }
      tmem_ld32(taddr + c, v);
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 r4 = rc[i >> 2];
          const float4 b4 = *reinterpret_cast<const float4*>(prm + c + i);
          v[i] += r4.x + b4.x; v[i + 1] += r4.y + b4.y; v[i + 2] += r4.z + b4.z; v[i + 3] += r4.w + b4.w;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) { sum += v[i]; sq = fmaf(v[i], v[i], sq); }
      }
      tmem_st32(taddr + c, v);
    }
    red[(half * 128 + rloc) * 2] = sum;
    red[(half * 128 + rloc) * 2 + 1] = sq;
    __syncthreads();
    sum = red[rloc * 2] + red[(128 + rloc) * 2];
    sq = red[rloc * 2 + 1] + red[(128 + rloc) * 2 + 1];
    const float mean = sum / (float)BN;
    const float rstd = rsqrtf(fmaxf(sq / (float)BN - mean * mean, 0.f) + 1e-5f);
    __nv_bfloat16* xbw = ep.xsb + (long long)row_w0 * ep.N;
#pragma unroll 1
    for (int c = half * 32; c < BN; c += 64) {
      tmem_ld32(taddr + c, v);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(prm + BN + c + i);
        const float4 b4 = *reinterpret_cast<const float4*>(prm + 2 * BN + c + i);
        tile[lane * 33 + i] = (v[i] - mean) * rstd * g4.x + b4.x;
        tile[lane * 33 + i + 1] = (v[i + 1] - mean) * rstd * g4.y + b4.y;
        tile[lane * 33 + i + 2] = (v[i + 2] - mean) * rstd * g4.z + b4.z;
        tile[lane * 33 + i + 3] = (v[i + 3] - mean) * rstd * g4.w + b4.w;
      }
      __syncwarp();
#pragma unroll 8
      for (int r = 0; r < 32; ++r)
        if (r < rows_w) xw[(long long)r * ep.N + c + lane] = tile[r * 33 + lane];
      const int cc = 2 * (lane & 15);
#pragma unroll 8
      for (int it = 0; it < 16; ++it) {
        const int r = 2 * it + (lane >> 4);
        if (r < rows_w)
          *reinterpret_cast<__nv_bfloat162*>(xbw + (long long)r * ep.N + c + cc) = __floats2bfloat162_rn(tile[r * 33 + cc], tile[r * 33 + cc + 1]);
      }
      __syncwarp();
    }
  } else if constexpr (EPI == EPI_HEAD) {
    // row = b*S + s -> frame f = s-1 of clip b; column n = joint channel j.  For fixed j a warp's 32 rows are
    // 32 consecutive frames: x[b][j][f..f+31] is one coalesced 128-byte line.  All loads of a chunk are issued
    // before the first store (x is read-modify-written in place: the compiler cannot prove the stores do not alias).
    const int b = row / ep.S, s = row - b * ep.S;
    const bool ok = row_ok && s > 0;
    const int f = s - 1;
    const long long xoff = (long long)b * ep.J * ep.T + f;
    __nv_bfloat16* xbr = ep.xb + (long long)row * ep.Jpad;
    float4 cf = make_float4(0.f, 0.f, 0.f, 0.f);
    bool nz = false;
    if (ep.head_mode == 0) {
      const int index = ep.step.index();
      cf = ep.coef[index];
      nz = (index != 0) && (ep.sampler == 0);
    }
#pragma unroll 1
    for (int c = half * 32; c < BN; c += 64) {
      tmem_ld32(taddr + c, v);
      const int n = n0 + c;
      if (!ok || n >= ep.N) continue;
      const long long base = xoff + (long long)n * ep.T;
      if (ep.head_mode == 0) {
        float xt[32], zz[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) xt[i] = (n + i < ep.N) ? ep.x[base + (long long)i * ep.T] : 0.f;
        if (nz) {
#pragma unroll
          for (int i = 0; i < 32; ++i) zz[i] = (n + i < ep.N) ? __ldg(ep.z + base + (long long)i * ep.T) : 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) zz[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float x0 = v[i] + prm[c + i];
          v[i] = (n + i < ep.N) ? posterior_apply(ep.sampler, cf, x0, xt[i], zz[i], nz) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) if (n + i < ep.N) ep.x[base + (long long)i * ep.T] = v[i];
        store_bf16x32(xbr + n, v);           // Jpad is a multiple of 32: the padded columns receive zeros
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (n + i < ep.N) reinterpret_cast<float*>(ep.out)[base + (long long)i * ep.T] = v[i] + prm[c + i];
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TcTile<BN>::TMEM_COLS) : "memory");
  }
}

}  // namespace tc
