// The clip kernel: ONE persistent CTA per clip runs the whole sampling loop of a segment — every DDPM step, every
// layer — with the activations resident in shared / tensor memory and only the weights streaming in (TMA, L2-resident
// bf16, 13.4 MB per step).  Clips are independent, so there is no inter-CTA synchronisation at
// all: no kernel boundaries, no grid barriers, no activation traffic through L2.  ZEGGS geometry (D 256, F 1024,
// 4 x 64 global heads, 8 x 32 local heads, window 11, T 88, J 1141) is compiled in.
//
// Warp roles (512 threads, warp = 4 * sub + q4; q4 = TMEM lane quarter = scheduler):
//   q4 == 3 (rows 96..127 carry no token): 3 = TMA producer of the main weight ring (+ x_t k-blocks, x_t / z chunks)
//                                          7 = tcgen05.mma issuer of the weight GEMMs       11 = noise pre-draw (Philox)
//                                          15 = attention issuer (S = Q K^T, O = P V) and TMA producer of the linear2 ring
//       These roles run WARP-CONVERGED with warp-uniform operands; only the tcgen05 / TMA instruction itself is issued by an
//       elected lane (elect_one): inside `if (lane == 0)` the compiler serialises every UTCHMMA / UTMALDG through a
//       uniform-register waterfall loop (~110 cycles of issue per MMA).
//   q4 <  3: 12 worker warps (4 per scheduler): all epilogues, softmax, local attention; sub = column quarter of an epilogue.
// Tensor memory (512 columns) = 4 quarters of 128 fp32 columns (Q0..Q3), handed back and forth per op with ready/free
// mbarriers.  FFN: linear1(c) accumulates in Q2 / Q3, the GELU epilogue writes the fp16 hidden back over the first 64 columns of
// the same quarter, linear2(c) reads it from there as a TENSOR-MEMORY A operand and accumulates into Q0|Q1.  Shared memory:
//   XS  [128 x 256] bf16, UMMA K-major SWIZZLE_128B, stored row-group-major (xs_off)  — the residual stream AND the A
//                                        operand; rows 96..127 carry no token: that contiguous 16 KB is the 4th weight stage
//   BUF [128 x 256] bf16, same format — A ring of the input GEMM / rope'd h for local attention / attention output (A of
//                                        out_proj) / 3 stages of the linear2 weight ring (FFN) / x_t, z chunks of the pose head
//   W   3 x 16 KB weight stages ([128 or 64 rows] x 64 k) + the one inside XS: the main TMA + mbarrier ring
//   AT_Q/K/P/V: per-head operands of the tcgen05 attention (no-swizzle core-matrix layouts, 40 KB; during the FFN: 2 stages of
//   the linear2 ring); LayerNorm partials; per-layer parameters; barriers.
#pragma once
#include "dsg_tc_gemm.cuh"
#include "dsg_tc_kernels.cuh"

namespace clip {
using namespace tc;

constexpr int D = 256, F = 1024, NH = 4, HD = 64, LH = 8, LHD = 32, WIN = 11, T = 88, S = 89, J = 1141, JPAD = 1152, NL = 8;
constexpr int NS = 4;                       // weight ring stages: 3 at OFF_W + the token-less rows 96..127 of XS (16 KB)
constexpr int WSTAGE = 16384;               // [128 rows x 64 k] bf16
constexpr int KT = 16384;                   // one A k-tile [128 x 64] bf16
constexpr int OFF_XS = 0;
constexpr int OFF_BUF = 4 * KT;
constexpr int OFF_W = 8 * KT;
constexpr int AT_BYTES = 41472;             // attention operand staging (Q | K -> P, V); doubles as pose-head chunk slots
constexpr int OFF_Q = OFF_W + 3 * WSTAGE;
constexpr int OFF_W3 = OFF_XS + 12 * 4096;      // 4th weight stage (see xs_off: XS is stored row-group-major)
constexpr int OFF_RED = OFF_Q + AT_BYTES;
// Global attention on tcgen05: operands in the canonical no-swizzle K-major layout [k-chunk of 8][8-row group][8 rows][16 B]
// (descriptor: LBO = byte distance between k-chunks = groups * 128, SBO = byte distance between 8-row groups = 128).
constexpr int AT_Q = OFF_Q;                      // A of S = Q K^T: [128 x 64]  (8 chunks x 16 groups x 128 B = 16 KB)
constexpr int AT_K = OFF_Q + 16384;              // B of S:        [ 96 x 64]  (8 chunks x 12 groups x 128 B = 12 KB)
constexpr int AT_P = OFF_Q;                      // A of O = P V:  [128 x 96]  (12 chunks x 16 groups = 24 KB), over Q | K once S is done
constexpr int AT_V = OFF_Q + 28672;              // B of O = V^T:  [ 64 x 96]  (12 chunks x 8 groups = 12 KB)
constexpr int LBO_Q = 16 * 128, LBO_K = 12 * 128, LBO_P = 16 * 128, LBO_V = 8 * 128;
static_assert(AT_V + 12 * LBO_V <= OFF_RED, "attention operand staging must fit in the Q/K/V region");
constexpr int OFF_B1 = OFF_RED + 2048;           // per-layer parameters (7 KB): b1[1024] fp16 | fp32 bq[256] | bo'[256] | b2[256] | 512 B | LayerNorm partials 1.5 KB;
                                                 // during the pose head its first 4.5 KB hold the head bias instead
constexpr int OFF_LNP = OFF_BUF + 3 * KT + 12288; // rows 96..127 of BUF k-tile 3: never a token, and beyond the pose-head chunk slots (which start landing before the last LayerNorm ends): fp32 g1 | be1 | g2 | be2
// pose-head phase: x_t / z chunks of 32 joint channels ([32][88] fp32 = 11,264 B each) are bulk-copied into 4 slots carved out
// of BUF and the (idle) attention staging area
constexpr int HCH = 32, HBYTES = HCH * T * 4, NHS = 4, NCHUNK = JPAD / HCH;
constexpr int XA_ROWS_BYTES = 96 * 128;          // the token-carrying rows of one k-block of the x_t image
constexpr int XA_BYTES = (JPAD / 64) * KT;       // per clip: x_t as bf16 A-operand k-blocks [18][128 x 64], SWIZZLE_128B image
constexpr int OFF_BAR = OFF_B1 + 4096 + 3072;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
constexpr int ZLD = 264;                    // rope'd h staging row stride (bf16), lives in BUF: 104 rows x 528 B
constexpr int ZROWS = 104;
static_assert(ZROWS * ZLD * 2 <= 4 * KT, "Z staging must fit in BUF");
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(5 * HBYTES <= 4 * KT && 3 * HBYTES <= AT_BYTES, "pose-head slots");

// per-layer fp32 parameter block (biases, LayerNorm): offsets in floats
constexpr int P_BQKV = 0, P_BO = 768, P_G1 = 1024, P_BE1 = 1280, P_B1 = 1536, P_B2 = 2560, P_G2 = 2816, P_BE2 = 3072, P_SIZE = 3328;
// weight slab row map (K = 256 slab): layer l at l*2048: [qkv reordered per head 768 | out_proj 256 | linear1 1024]; pose head after the layers
constexpr int R_LAYER = 2048, R_QKV = 0, R_WO = 768, R_W1 = 1024, R_HEAD = NL * R_LAYER;

// barrier indices
enum { B_WFULL = 0, B_WEMPTY = 4, B_AFULL = 8, B_AEMPTY = 12, B_ACCR = 16, B_ACCF = 20, B_XSR = 24, B_BUFR = 25, B_BUFF = 27,
       B_ZR = 29, B_ZF = 30, B_HFULL = 31, B_HEMPTY = 35, B_HGO = 39, B_XAR = 40, B_QKR = 41, B_SR = 42, B_PR = 43, B_OR = 44,
       B_W2FULL = 45, B_W2EMPTY = 50, B_RX = 55, B_PFREE = 56, B_COUNT = 57 };
// FFN (round 2): the GELU'd hidden chunk never touches shared memory.  The epilogue writes it (fp16 pairs) back over the first
// 64 columns of its own linear1 accumulator quarter with tcgen05.st, and linear2 reads it from there as a TENSOR-MEMORY A
// operand (tcgen05.mma [d], [a_tmem], b_desc): no st.shared / proxy fence in the epilogue, no A re-reads by the tensor core, and
// BUF + the attention staging area are free during the FFN: they are the 5-stage ring of the linear2 weight tiles (80 KB),
// fed by the attention issuer's thread, so that a linear2 tile waiting for its GELU chunk never blocks the linear1 stream.
// K order inside a 128-unit chunk: TMEM column j holds units (j, 64 + j) -> K positions (2j, 2j + 1); the linear2 slab is
// packed in that order (pack_w2_perm_kernel), so every epilogue thread overwrites only columns it has read itself.
constexpr int NS2 = 5;
DSG_DEVINL int w2stage_off(int slot) { return slot < 2 ? OFF_Q + slot * WSTAGE : OFF_BUF + (slot - 2) * KT; }
static_assert(2 * WSTAGE <= AT_BYTES, "two linear2 stages live in the attention staging area");

struct ClipParams {
  float* x;                 // [B][J][T] fp32, in/out
  uint8_t* xa;              // [B][XA_BYTES]: bf16(x_t) as the input GEMM's A k-blocks (written by pack_xa_kernel, then by the head epilogue)
  float* z;                 // [B][J*T] fp32 noise scratch
  const float* cond;        // [B][T][D]
  const float* emb1;        // [B][D]
  const float* te;          // [n_t][D]
  const float* TW;          // [n_t][D]
  const float2* cs;         // [S][16]  rope (cos, sin) for local head dim 32
  const float* lparams;     // [NL][P_SIZE]
  const float* bout;        // [JPAD]
  const float4* coef;
  const int* tmap;
  const long long* clip_ids;
  const LoopParams* lp;
  int B, n_run, sampler;
  float* dbg; long long dbg_slot; int debug;     // taps: dbg[slot * dbg_slot + (clip*S + s)*D + col]
  long long* prof;          // optional [32] cycle counters written by CTA 0 (see PF_* below), nullable
};

DSG_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DSG_DEVINL void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
constexpr int NW = 12, NWT = NW * 32;          // worker warps / threads
DSG_DEVINL void workers_sync() { asm volatile("bar.sync 1, 384;" ::: "memory"); }
// the 4 worker warps that share a TMEM lane quarter (= one token row group): row statistics (LayerNorm sums, softmax max / sum)
// are exchanged only among them, so they need not wait for the other two quarters
DSG_DEVINL void quarter_sync(int q4) { asm volatile("bar.sync %0, 128;" ::"r"(2 + q4) : "memory"); }
// ---- CTA-pair mode (CL = 2, see the header of clip_kernel): distributed shared memory + remote mbarrier arrivals
DSG_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
DSG_DEVINL uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
DSG_DEVINL void st_remote16(uint32_t raddr, const uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
DSG_DEVINL void bulk_copy_to_peer(uint32_t rdst, const void* src, uint32_t bytes, uint32_t rbar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(rdst), "r"(smem_u32(src)), "r"(bytes), "r"(rbar) : "memory");
}
DSG_DEVINL void mbar_arrive_expect_tx_remote(uint32_t rbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(rbar), "r"(bytes) : "memory");
}
DSG_DEVINL void mbar_arrive_remote(uint32_t rbar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
}
DSG_DEVINL void mbar_arrive_release_cluster(uint64_t* bar) {      // local arrival that also publishes to the peer's observers
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DSG_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {      // acquire at cluster scope: the arrivals may be remote
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
DSG_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
DSG_DEVINL void tmem_ld8_issue(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
DSG_DEVINL void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
               "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand is read from tensor memory (lanes = rows, 16-bit elements packed two per
// 32-bit column, K-major), B through a shared-memory descriptor
DSG_DEVINL void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// One lane of a converged warp (the same lane every time on this hardware: CUTLASS relies on it for MMA + commit pairs).
// The issuer roles run warp-converged with warp-uniform operands and elect only the tcgen05 / TMA instruction itself: a role
// that lives inside `if (lane == 0)` is divergent code to the compiler, which then wraps every UTCHMMA in a per-lane
// R2UR / ELECT / BRA.U.ANY loop (~20 dependent instructions per MMA: the MMA thread, not the tensor pipe, was the limit).
DSG_DEVINL bool elect_one() { return elect_one_lane(); }
DSG_DEVINL uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
// 2^x on MUFU.EX2 directly (ex2.approx.ftz: 2 ulp; exp2f() adds a subnormal-range rescale = 4 more instructions per element)
DSG_DEVINL float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// packed fp32 pairs (sm_100: add / fma .f32x2 = FADD2 / FFMA2): two lanes per instruction for the LayerNorm statistics
DSG_DEVINL void add2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 a, b;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\tadd.rn.f32x2 a, a, b;\n\tmov.b64 {%0, %1}, a;\n\t}\n"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
DSG_DEVINL void mul2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 a, b;\n\tmov.b64 a, {%0, %1};\n\tmov.b64 b, {%2, %3};\n\tmul.rn.f32x2 a, a, b;\n\tmov.b64 {%0, %1}, a;\n\t}\n"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
DSG_DEVINL void fma2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmov.b64 c, {%0, %1};\n\tfma.rn.f32x2 c, a, b, c;\n\t"
      "mov.b64 {%0, %1}, c;\n\t}\n" : "+f"(c0), "+f"(c1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
DSG_DEVINL void tie4(float* v) { asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]) :: "memory"); }
DSG_DEVINL void ldsm_x2(uint32_t& r0, uint32_t& r1, const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
// byte offset of element (row r, column c) inside a [128 x 256] bf16 operand stored as 4 SWIZZLE_128B k-tiles
DSG_DEVINL uint32_t a_off(int r, int c) {
  return (uint32_t)((c >> 6) * KT + (r >> 3) * 1024 + (r & 7) * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
}
// XS is stored row-group-major: 8-row group g of k-tile kt at g * 4096 + kt * 1024 (UMMA descriptors: SBO = 4096), so that the
// token-less rows 96..127 of all four k-tiles form one contiguous 16 KB block — the 4th weight stage
DSG_DEVINL uint32_t xs_off(int r, int c) {
  return (uint32_t)((r >> 3) * 4096 + (c >> 6) * 1024 + (r & 7) * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + (c & 7) * 2);
}
DSG_DEVINL uint64_t make_sw128_desc_sbo(uint32_t smem_addr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
DSG_DEVINL uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {     // SWIZZLE_NONE, K-major
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
DSG_DEVINL int wstage_off(int slot) { return slot < 3 ? OFF_W + slot * WSTAGE : OFF_W3; }
DSG_DEVINL uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  return u;
}
// GELU(x) = 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))) on two fp16 lanes; saturates correctly when x^2 overflows
DSG_DEVINL uint32_t gelu_h2(const __half2 x) {
  const __half2 c1 = __floats2half2_rn(0.0356774081f, 0.0356774081f), c0 = __floats2half2_rn(0.7978845608f, 0.7978845608f);
  const __half2 u = __hmul2(__hfma2(__hmul2(x, x), c1, c0), x);
  uint32_t ui = *reinterpret_cast<const uint32_t*>(&u), ti;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(ti) : "r"(ui));
  const __half2 hx = __hmul2(x, __floats2half2_rn(0.5f, 0.5f));
  const __half2 g = __hfma2(hx, *reinterpret_cast<const __half2*>(&ti), hx);
  return *reinterpret_cast<const uint32_t*>(&g);
}
DSG_DEVINL void unpack8(const uint4 u, float* v) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(p[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// cycle counters (CTA 0): where each role spends its time
enum { PF_TOTAL = 0, PF_MMA_WAIT_W, PF_MMA_WAIT_OTHER, PF_PROD_WAIT_EMPTY, PF_W_STAGE, PF_W_IN_WAIT, PF_W_IN_EPI, PF_W_LOCAL,
       PF_W_QKV_WAIT, PF_W_ATT, PF_W_LN_WAIT, PF_W_LN, PF_W_GELU_WAIT, PF_W_GELU, PF_W_HEAD_WAIT, PF_W_HEAD, PF_W_ZWAIT,
       PF_W_EXTRACT, PF_W_SYNC1, PF_W_ATT_MMA, PF_W_ATT_MERGE,
       PF_MMA_FFN_TOTAL, PF_MMA_FFN_WAIT_W, PF_MMA_FFN_WAIT_O, PF_MMA_QKV_TOTAL, PF_MMA_QKV_WAIT_W, PF_MMA_QKV_WAIT_O, PF_COUNT };

// consumers of tcgen05.ld results must not be scheduled above tcgen05.wait::ld: pass the registers through an empty
// volatile asm placed after the wait
DSG_DEVINL void tie32(float* v) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
               "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]),
               "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]),
               "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]) :: "memory");
}

struct Phases {            // one phase bit per barrier, toggled on every completed wait
  uint64_t bits;
  DSG_DEVINL void wait(uint64_t* bars, int id) { mbar_wait(&bars[id], (uint32_t)(bits >> id) & 1u); bits ^= 1ull << id; }
  DSG_DEVINL void waitc(uint64_t* bars, int id) { mbar_wait_cluster(&bars[id], (uint32_t)(bits >> id) & 1u); bits ^= 1ull << id; }
};
DSG_DEVINL void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
DSG_DEVINL void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// byte offsets of head slot s: x_t chunk, z chunk
DSG_DEVINL int hslot_x(int s) { return s == 0 ? OFF_BUF : (s == 1 ? OFF_BUF + 2 * HBYTES : (s == 2 ? OFF_Q : OFF_BUF + 4 * HBYTES)); }
DSG_DEVINL int hslot_z(int s) { return s == 0 ? OFF_BUF + HBYTES : (s == 1 ? OFF_BUF + 3 * HBYTES : (s == 2 ? OFF_Q + HBYTES : OFF_Q + 2 * HBYTES)); }

// x fp32 [B][J][T] -> xa (the clip kernel's A k-blocks): row = frame + 1, column = joint channel, bf16, SWIZZLE_128B image.
// Rows 0 and 89..127 and the columns >= J stay zero (the buffer is allocated zeroed and never written there).
static __global__ void __launch_bounds__(256) pack_xa_kernel(const float* __restrict__ x, uint8_t* __restrict__ xa) {
  const int kb = blockIdx.x, clip = blockIdx.y;
  const float* xc = x + (long long)clip * J * T;
  uint8_t* at = xa + (long long)clip * XA_BYTES + (long long)kb * KT;
  for (int it = threadIdx.x; it < 32 * T; it += 256) {
    const int pr = it / T, f = it - pr * T;
    const int j = kb * 64 + 2 * pr;
    const float lo = (j < J) ? xc[(long long)j * T + f] : 0.f, hi = (j + 1 < J) ? xc[(long long)(j + 1) * T + f] : 0.f;
    const int rr = f + 1, cc = 2 * pr;
    *reinterpret_cast<uint32_t*>(at + (rr >> 3) * 1024 + (rr & 7) * 128 + (((cc >> 3) ^ (rr & 7)) << 4) + (cc & 7) * 2) = pack_bf16x2(lo, hi);
  }
}

// ---------------------------------------------------------------------------------------------------
// CL = 1: one CTA per clip.  CL = 2 (batches smaller than half the SMs): a CLUSTER of two CTAs per clip.  Each CTA keeps the
// whole residual stream (XS) but streams and multiplies only its share of the weights:
//   in_proj + global attention: 2 of the 4 heads per CTA; the head outputs go into BOTH CTAs' BUF (a bulk copy of the two
//       k-tiles into the peer's shared memory, counted on the peer's B_BUFR), so each CTA runs the full out_proj + LayerNorm 1;
//   FFN: 4 of the 8 hidden chunks per CTA = a K-split of linear2; the partial sums are exchanged (as bf16) by ROW OWNERSHIP (rank 0
//       owns TMEM lane quarters 0 and 2, rank 1 quarter 1): a non-owner warp writes its 64 columns into the owner's receive
//       buffer (BUF: the linear2 ring keeps to the attention staging area in this mode), the owner adds them, runs LayerNorm 2
//       and the bf16 rows go into both CTAs' XS (one 16 KB bulk copy per owned quarter, counted on the peer's B_XSR);
//   pose head: tiles 0..4 / 5..8; x_t and its bf16 image meet in global memory (B_XAR counts both CTAs' worker warps).
//   input GEMM + local attention: the 128 latent columns = 4 local heads of the rank; the halves of XS are exchanged as 12 bulk
//       copies of 2 KB (one per 8-row group), announced on the RECEIVER's B_XSR by one of its own arrivals (arrive.expect_tx);
//   out_proj and LayerNorm 1 are computed by both CTAs (8 % of the weight bytes).
// Write-after-read across the pair is covered by data dependencies except for BUF, for which the peer sends a token (B_PFREE)
// when its reads are over: once per step after the local attention (Z staging), once per layer after out_proj.
template <bool PROF, int CL>
__global__ void __launch_bounds__(512, 1)
clip_kernel(const __grid_constant__ CUtensorMap tm_in,     // Wxp   [256, 1152]   box 128 x 64
            const __grid_constant__ CUtensorMap tm_w128,   // K=256 slab          box 128 x 64
            const __grid_constant__ CUtensorMap tm_w64,    // K=256 slab          box  64 x 64
            const __grid_constant__ CUtensorMap tm_w2,     // linear2 slab [NL*256, 1024]  box 128 x 64
            const ClipParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
  const int warp = (int)uniform_u32(threadIdx.x >> 5), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&bars[B_WFULL + i], 1); mbar_init(&bars[B_WEMPTY + i], 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars[B_AFULL + i], 1); mbar_init(&bars[B_AEMPTY + i], 1);
      mbar_init(&bars[B_HFULL + i], 1); mbar_init(&bars[B_HEMPTY + i], NW);
      mbar_init(&bars[B_ACCR + i], 1);  mbar_init(&bars[B_ACCF + i], NW);
    }
    mbar_init(&bars[B_XSR], NW);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[B_BUFR + i], NW); mbar_init(&bars[B_BUFF + i], 1); }
    mbar_init(&bars[B_ZR], 1); mbar_init(&bars[B_ZF], NW);
    mbar_init(&bars[B_HGO], 1); if constexpr (CL == 1) mbar_init(&bars[B_XAR], NW);
    mbar_init(&bars[B_QKR], NW); mbar_init(&bars[B_SR], 1); mbar_init(&bars[B_PR], NW); mbar_init(&bars[B_OR], 1);
    for (int i = 0; i < NS2; ++i) { mbar_init(&bars[B_W2FULL + i], 1); mbar_init(&bars[B_W2EMPTY + i], 1); }
    if constexpr (CL > 1) {
      mbar_init(&bars[B_XAR], CL * NW);                               // (re-initialised: both CTAs' workers publish x_t)
      mbar_init(&bars[B_RX], cluster_ctarank() == 0 ? 8 : 4);         // the peer's non-owner warps of my row quarters
      mbar_init(&bars[B_PFREE], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_in) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w128) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w64) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w2) : "memory");
  }
  // zero the operand buffers once: rows that are never written (token slot of the A ring, rows >= S, Z pad rows) stay finite
  for (int i = threadIdx.x; i < (OFF_W) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < (OFF_RED - OFF_Q) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + OFF_Q)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 7) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();                // the peer's barriers are initialised before anything arrives on them
  tcgen05_fence_after();
  const uint32_t tmem = uniform_u32(*tmem_slot);           // (warp-uniform for the compiler: see elect_one)
  // CTA pair: rank, the peer's shared memory window, this CTA's share of the heads / hidden chunks / pose-head tiles
  const int rank = CL > 1 ? (int)uniform_u32(cluster_ctarank()) : 0;
  const uint32_t peer_smem = CL > 1 ? uniform_u32(map_to_cta(smem_u32(smem), (uint32_t)(rank ^ 1))) : 0u;
  const uint32_t peer_bars = peer_smem + OFF_BAR;
  constexpr int NHL = NH / CL, NCH = 8 / CL, NS2L = CL == 1 ? NS2 : 2;
  const int h_base = rank * NHL, c_base = rank * NCH;
  const int tile0 = CL == 1 ? 0 : (rank == 0 ? 0 : 5), tile1 = CL == 1 ? JPAD / 128 : (rank == 0 ? 5 : JPAD / 128);
  const int cid0 = (int)blockIdx.x / CL, ncl = (int)gridDim.x / CL;      // clip slot of this CTA (pair), clips in flight
  const int draw0 = (int)uniform_u32((uint32_t)P.lp->k);   // loop iteration of this launch's first step (dump_steps cuts the loop)
  const int first_index = (int)uniform_u32((uint32_t)P.lp->first_index) - draw0;
  const uint32_t key0 = P.lp->key0, key1 = P.lp->key1, segment = P.lp->segment;

  const int q4 = warp & 3, sub = warp >> 2;
  if (q4 == 3 && sub == 0) {
    // =================================================== TMA weight producer ===================================================
    {                                               // warp-converged, elected issue (see elect_one)
      Phases ph{(((1ull << NS) - 1) << B_WEMPTY) | (0xFull << B_AEMPTY) | (0xFull << B_HEMPTY)};     // "empty" barriers start free
      int slot = 0;
      long long t_wait = 0;
      const bool prof = PROF && P.prof != nullptr && blockIdx.x == 0;
      auto load = [&](const CUtensorMap* m, int row, int kcol, uint32_t bytes) {
        const long long c0 = prof ? clock64() : 0;
        ph.wait(bars, B_WEMPTY + slot);
        if (prof) t_wait += clock64() - c0;
        if (elect_one()) {
          mbar_expect_tx(&bars[B_WFULL + slot], bytes);
          tma_load_2d(smem + wstage_off(slot), m, &bars[B_WFULL + slot], kcol, row);
        }
        __syncwarp();
        slot = (slot + 1 == NS) ? 0 : slot + 1;
      };
      bool first = true;
      for (int clip = cid0; clip < P.B; clip += ncl)
        for (int k = 0; k < P.n_run; ++k) {
          const int index = first_index - k;
          const bool nz = (index != 0) && (P.sampler == 0);
          // x_t as bf16 A k-blocks: written by pack_xa_kernel before the launch, afterwards by the head epilogue of the previous
          // step (B_XAR: those stores are complete and fenced for the async proxy, and BUF is no longer read)
          if (!first) {
            if constexpr (CL > 1) { ph.waitc(bars, B_XAR); fence_async_all(); }     // half of the image was written by the peer
            else ph.wait(bars, B_XAR);
          }
          first = false;
          const uint8_t* xa = P.xa + (long long)clip * XA_BYTES;
          for (int kb = 0; kb < JPAD / 64; ++kb) {
            const int sl = kb & 3;
            ph.wait(bars, B_AEMPTY + sl);
            if (elect_one()) {
              // rows 0..95 of the k-block only (12 KB): rows 96..127 carry no token; what the stage holds there only reaches
              // accumulator rows nobody reads
              mbar_expect_tx(&bars[B_AFULL + sl], XA_ROWS_BYTES);
              bulk_load(smem + OFF_BUF + sl * KT, xa + (long long)kb * KT, XA_ROWS_BYTES, &bars[B_AFULL + sl]);
            }
            __syncwarp();
            // (CTA pair: each CTA computes the 128 latent columns = 4 local heads of its rank)
            for (int nh = (CL > 1 ? rank : 0); nh < (CL > 1 ? rank + 1 : 2); ++nh) load(&tm_in, nh * 128, kb * 64, WSTAGE);
          }
          for (int l = 0; l < NL; ++l) {
            const int rb = l * R_LAYER;
            for (int h = h_base; h < h_base + NHL; ++h) {   // per head: q|k rows as one 128-row box, v as a 64-row box
              for (int kb = 0; kb < 4; ++kb) load(&tm_w128, rb + R_QKV + h * 192, kb * 64, WSTAGE);
              for (int kb = 0; kb < 4; ++kb) load(&tm_w64, rb + R_QKV + h * 192 + 128, kb * 64, WSTAGE / 2);
            }
            for (int nh = 0; nh < 2; ++nh)
              for (int kb = 0; kb < 4; ++kb) load(&tm_w128, rb + R_WO + nh * 128, kb * 64, WSTAGE);
            auto ff1 = [&](int c) { for (int kb = 0; kb < 4; ++kb) load(&tm_w128, rb + R_W1 + c * 128, kb * 64, WSTAGE); };
            for (int c = c_base; c < c_base + NCH; ++c) ff1(c);   // (the linear2 tiles travel through their own ring: see the attention issuer)
          }
          // pose head: weights + the x_t / z chunks the posterior needs (BUF is free once the last linear2 has completed)
          for (int kb = 0; kb < 4; ++kb) load(&tm_w128, R_HEAD + tile0 * 128, kb * 64, WSTAGE);
          ph.wait(bars, B_HGO);
          if (nz) ph.wait(bars, B_ZR);
          const float* xc = P.x + (long long)clip * J * T;
          const float* zc = P.z + (long long)cid0 * J * T;                // noise scratch is per clip slot (see the noise warp)
          auto hload = [&](int c) {
            const int sl = c & 3, j0 = c * HCH;
            const uint32_t bytes = (uint32_t)((J - j0 < HCH ? J - j0 : HCH) * T * 4);
            ph.wait(bars, B_HEMPTY + sl);
            if (elect_one()) {
              mbar_expect_tx(&bars[B_HFULL + sl], nz ? 2 * bytes : bytes);
              bulk_load(smem + hslot_x(sl), xc + (long long)j0 * T, bytes, &bars[B_HFULL + sl]);
              if (nz) bulk_load(smem + hslot_z(sl), zc + (long long)j0 * T, bytes, &bars[B_HFULL + sl]);
            }
            __syncwarp();
          };
          for (int c = 4 * tile0; c < 4 * tile0 + NHS; ++c) hload(c);
          for (int t = tile0 + 1; t < tile1; ++t) {
            for (int kb = 0; kb < 4; ++kb) load(&tm_w128, R_HEAD + t * 128, kb * 64, WSTAGE);
            for (int c = 4 * t; c < 4 * t + 4; ++c) hload(c);
          }
        }
      if (prof && lane == 0) P.prof[PF_PROD_WAIT_EMPTY] = t_wait;
    }
  } else if (q4 == 3 && sub == 1) {
    // =================================================== MMA issuer ===================================================
    {                                               // the whole warp runs the role converged; one elected lane issues
      Phases ph{(0xFull << B_ACCF)};               // accumulators start free
      int slot = 0;
      const uint32_t xs_addr = smem_u32(smem + OFF_XS), buf_addr = smem_u32(smem + OFF_BUF), smem_addr0 = smem_u32(smem);
      constexpr uint32_t idesc128 = make_idesc_bf16(128, 128), idesc64 = make_idesc_bf16(128, 64), idesc128h = make_idesc_f16(128, 128);
      // one weight tile: 4 UMMAs (K = 64) of A k-tile `a_tile` against the current stage
      long long t_w = 0, t_o = 0, t_ffn = 0, t_ffn_w = 0, t_ffn_o = 0, t_qkv = 0, t_qkv_w = 0, t_qkv_o = 0;
      const bool prof = PROF && P.prof != nullptr && blockIdx.x == 0;
      const long long t_begin = clock64();
      auto tile = [&](uint32_t a_tile, uint32_t d_col, uint32_t idesc, bool acc_first, uint32_t a_sbo = 1024) {
        const long long c0 = prof ? clock64() : 0;
        ph.wait(bars, B_WFULL + slot);
        if (prof) t_w += clock64() - c0;
        tcgen05_fence_after();
        const uint32_t b_tile = smem_addr0 + wstage_off(slot);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tmem + d_col, make_sw128_desc_sbo(a_tile + kk * 32, a_sbo), make_sw128_desc(b_tile + kk * 32), idesc, (acc_first || kk > 0) ? 1u : 0u);
          tcgen05_commit(&bars[B_WEMPTY + slot]);
        }
        __syncwarp();
        slot = (slot + 1 == NS) ? 0 : slot + 1;
      };
      int slot2 = 0;
      // a linear2 tile: B from the linear2 ring, A = the fp16 hidden chunk in tensor memory (8 columns per K = 16 step)
      auto tile2 = [&](uint32_t a_tmem, uint32_t d_col, uint32_t idesc, bool acc_first) {
        const long long c0 = prof ? clock64() : 0;
        ph.wait(bars, B_W2FULL + slot2);
        if (prof) t_w += clock64() - c0;
        tcgen05_fence_after();
        const uint32_t b_tile = smem_addr0 + w2stage_off(slot2);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ts(tmem + d_col, tmem + a_tmem + kk * 8, make_sw128_desc(b_tile + kk * 32), idesc, (acc_first || kk > 0) ? 1u : 0u);
          tcgen05_commit(&bars[B_W2EMPTY + slot2]);
        }
        __syncwarp();
        slot2 = (slot2 + 1 == NS2L) ? 0 : slot2 + 1;
      };
      auto owait = [&](int id) { const long long c0 = prof ? clock64() : 0; ph.wait(bars, id); if (prof) t_o += clock64() - c0; };
      // a barrier the peer's workers arrive on too (acquire at cluster scope; the writers ran fence.proxy.async before arriving)
      // taps in CTA-pair mode: XS is complete (both CTAs' rows) exactly when this warp has acquired B_XSR
      auto pair_dump = [&](int slot, int clip) {
        if constexpr (CL > 1) {
          if (!(P.debug & 1) || rank != 0) return;
          float* o = P.dbg + (long long)slot * P.dbg_slot + (long long)clip * S * D;
          for (int i = lane; i < S * (D / 8); i += 32) {
            const int rr = i / (D / 8), c8 = i % (D / 8);
            float t8[8];
            unpack8(*reinterpret_cast<const uint4*>(smem + OFF_XS + xs_off(rr, c8 * 8)), t8);
            for (int j = 0; j < 8; ++j) o[(long long)rr * D + c8 * 8 + j] = t8[j];
          }
          __syncwarp();
        }
      };
      auto owait_pair = [&](int id) {
        if constexpr (CL > 1) {
          const long long c0 = prof ? clock64() : 0; ph.waitc(bars, id); if (prof) t_o += clock64() - c0;
        } else owait(id);
      };
      for (int clip = cid0; clip < P.B; clip += ncl)
        for (int k = 0; k < P.n_run; ++k) {
          // ---- input GEMM: A k-blocks staged by the workers into the BUF ring, D = Q0|Q1
          owait(B_ACCF + 0); owait(B_ACCF + 1);
          for (int kb = 0; kb < JPAD / 64; ++kb) {
            owait(B_AFULL + (kb & 3));
            tcgen05_fence_after();
            for (int nh = (CL > 1 ? rank : 0); nh < (CL > 1 ? rank + 1 : 2); ++nh) tile(buf_addr + (kb & 3) * KT, nh * 128, idesc128, kb > 0);
            if (elect_one()) { tcgen05_commit(&bars[B_AEMPTY + (kb & 3)]); }
          }
          if (elect_one()) { tcgen05_commit(&bars[B_ACCR + 0]); tcgen05_commit(&bars[B_ACCR + 1]); }
          for (int l = 0; l < NL; ++l) {
            // ---- in_proj, one head at a time into alternating TMEM halves: q | k | v = 3 x 64 columns
            owait_pair(B_XSR);
            pair_dump(l, clip);
            tcgen05_fence_after();
            const long long qkv_t0 = prof ? clock64() : 0, qkv_w0 = t_w, qkv_o0 = t_o;
            auto qkv = [&](int h) {
              const int hb = h & 1;
              owait(B_ACCF + 2 * hb); owait(B_ACCF + 2 * hb + 1);
              tcgen05_fence_after();
              for (int kb = 0; kb < 4; ++kb) tile(xs_addr + kb * 1024, hb * 256, idesc128, kb > 0, 4096);            // q | k
              for (int kb = 0; kb < 4; ++kb) tile(xs_addr + kb * 1024, hb * 256 + 128, idesc64, kb > 0, 4096);      // v
              if (elect_one()) { tcgen05_commit(&bars[B_ACCR + 2 * hb]); tcgen05_commit(&bars[B_ACCR + 2 * hb + 1]); }
            };
            // (S = Q K^T and O = P V of every head are issued by the attention issuer, warp 15; a head's TMEM half comes back
            //  after its O epilogue, which is what in_proj(h + 2) waits for)
            for (int h = 0; h < NHL; ++h) qkv(h);
            if (prof) { t_qkv += clock64() - qkv_t0; t_qkv_w += t_w - qkv_w0; t_qkv_o += t_o - qkv_o0; }
            // ---- out_proj: A = attention output in BUF, D = Q0|Q1
            owait_pair(B_BUFR + 0); owait_pair(B_BUFR + 1);
            owait(B_ACCF + 0); owait(B_ACCF + 1);
            tcgen05_fence_after();
            for (int nh = 0; nh < 2; ++nh)
              for (int kb = 0; kb < 4; ++kb) tile(buf_addr + kb * KT, nh * 128, idesc128, kb > 0);
            if (elect_one()) { tcgen05_commit(&bars[B_ACCR + 0]); tcgen05_commit(&bars[B_ACCR + 1]); }
            if (elect_one()) { tcgen05_commit(&bars[B_BUFF + 0]); tcgen05_commit(&bars[B_BUFF + 1]); }
            // ---- FFN: linear1 in 8 chunks of 128 hidden units (D = Q2 / Q3 alternating), linear2 accumulates into Q0|Q1
            owait(B_XSR);
            tcgen05_fence_after();
            // linear1(c) -> quarter 2 + (c & 1); GELU(c) leaves the hidden chunk in the same quarter; linear2(c) reads it from
            // there; linear1(c + 2) then overwrites the quarter (the tensor pipe executes one thread's MMAs in issue order)
            const long long ffn_t0 = prof ? clock64() : 0, ffn_w0 = t_w, ffn_o0 = t_o;
            auto ff1 = [&](int c, bool wait_free) {
              if (wait_free) { owait(B_ACCF + 2 + (c & 1)); tcgen05_fence_after(); }
              for (int kb = 0; kb < 4; ++kb) tile(xs_addr + kb * 1024, (2 + (c & 1)) * 128, idesc128, kb > 0, 4096);
              if (elect_one()) { tcgen05_commit(&bars[B_ACCR + 2 + (c & 1)]); }
            };
            ff1(0, true); ff1(1, true);
            for (int c = 0; c < NCH; ++c) {
              owait(B_BUFR + (c & 1));                   // hidden chunk c is in tensor memory
              if (c == 0) { owait(B_ACCF + 0); owait(B_ACCF + 1); }
              tcgen05_fence_after();
              for (int nh = 0; nh < 2; ++nh)
                for (int kb2 = 0; kb2 < 2; ++kb2)
                  tile2((uint32_t)((2 + (c & 1)) * 128 + kb2 * 32), nh * 128, idesc128h, c > 0 || kb2 > 0);
              if (c + 2 < NCH) ff1(c + 2, false);
            }
            if (elect_one()) { tcgen05_commit(&bars[B_ACCR + 0]); tcgen05_commit(&bars[B_ACCR + 1]); }
            if (prof) { t_ffn += clock64() - ffn_t0; t_ffn_w += t_w - ffn_w0; t_ffn_o += t_o - ffn_o0; }
            // every read of BUF by the tensor core is complete (CTA pair: BUF is the receive buffer of LayerNorm 2 until the
            // workers are through with it; they give the go themselves)
            if (CL == 1 && l == NL - 1 && elect_one()) tcgen05_commit(&bars[B_HGO]);
          }
          // ---- pose head: 9 tiles of 128 joint channels, D rotates over the 4 quarters
          owait_pair(B_XSR);
          pair_dump(NL, clip);
          tcgen05_fence_after();
          for (int t = 0; t < tile1 - tile0; ++t) {
            owait(B_ACCF + (t & 3));
            tcgen05_fence_after();
            for (int kb = 0; kb < 4; ++kb) tile(xs_addr + kb * 1024, (t & 3) * 128, idesc128, kb > 0, 4096);
            if (elect_one()) { tcgen05_commit(&bars[B_ACCR + (t & 3)]); }
          }
        }
      if (prof && lane == 0) {
        P.prof[PF_TOTAL] = clock64() - t_begin; P.prof[PF_MMA_WAIT_W] = t_w; P.prof[PF_MMA_WAIT_OTHER] = t_o;
        P.prof[PF_MMA_FFN_TOTAL] = t_ffn; P.prof[PF_MMA_FFN_WAIT_W] = t_ffn_w; P.prof[PF_MMA_FFN_WAIT_O] = t_ffn_o;
        P.prof[PF_MMA_QKV_TOTAL] = t_qkv; P.prof[PF_MMA_QKV_WAIT_W] = t_qkv_w; P.prof[PF_MMA_QKV_WAIT_O] = t_qkv_o;
      }
    }
  } else if (q4 == 3 && sub == 2) {
    // =================================================== noise pre-draw ===================================================
    Phases ph{1ull << B_ZF};
    for (int clip = cid0; clip < P.B; clip += ncl) {
      const uint32_t cid = (uint32_t)P.clip_ids[clip];
      // the noise scratch is indexed by clip slot (CTA or CTA pair), not by clip: 148 x 401 KB = 59 MB whatever the batch, small enough to be pinned in
      // L2 by the launch's access-policy window (dsg_tc.cu: clip_run), so the step noise never travels to HBM and back
      float* zc = P.z + (long long)cid0 * J * T;
      // (CTA pair: each CTA draws the channels of its own pose-head tiles)
      const int q_lo = tile0 * 128 * T / 4, q_hi = (tile1 * 128 < J ? tile1 * 128 : J) * T / 4;
      for (int k = 0; k < P.n_run; ++k) {
        const int index = first_index - k;
        if (index == 0 || P.sampler != 0) continue;
        ph.wait(bars, B_ZF);
        for (int q = q_lo + lane; q < q_hi; q += 32)
          *reinterpret_cast<float4*>(zc + 4 * q) = philox_normal4((uint32_t)q, (uint32_t)(1 + draw0 + k), cid, segment, key0, key1);
        fence_async_all();                               // z is read back through the async proxy (bulk copies)
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[B_ZR]);
      }
    }
  } else if (q4 == 3) {
    // =================================================== attention issuer ===================================================
    // Global attention of every head on the tensor core, from its own thread so that it never queues behind a weight tile:
    // S = Q K^T into columns [0, 96) of the head's TMEM half (the q | k accumulators are already extracted), then O = P V
    // into columns [96, 160).  Operands: canonical no-swizzle core-matrix layouts (AT_*); V is the MN-major B operand.
    {                                               // warp-converged, elected issue (see elect_one)
      Phases ph{((1ull << NS2L) - 1) << B_W2EMPTY};
      const uint32_t smem_addr0 = smem_u32(smem);
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 96), idesc_o = make_idesc_bf16(128, 64) | (1u << 16);
      int slot2 = 0;
      for (int clip = cid0; clip < P.B; clip += ncl)
        for (int k = 0; k < P.n_run; ++k)
          for (int l = 0; l < NL; ++l) {
            for (int h = 0; h < NHL; ++h) {
              const uint32_t d0 = tmem + (uint32_t)((h & 1) * 256);
              ph.wait(bars, B_QKR);
              tcgen05_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  umma_bf16(d0, make_nosw_desc(smem_addr0 + AT_Q + j * 2 * LBO_Q, LBO_Q, 128),
                            make_nosw_desc(smem_addr0 + AT_K + j * 2 * LBO_K, LBO_K, 128), idesc_s, j > 0 ? 1u : 0u);
                tcgen05_commit(&bars[B_SR]);
              }
              __syncwarp();
              ph.wait(bars, B_PR);
              tcgen05_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 6; ++j)
                  umma_bf16(d0 + 96, make_nosw_desc(smem_addr0 + AT_P + j * 2 * LBO_P, LBO_P, 128),
                            make_nosw_desc(smem_addr0 + AT_V + j * 2 * LBO_V, LBO_V, 128), idesc_o, j > 0 ? 1u : 0u);
                tcgen05_commit(&bars[B_OR]);
              }
              __syncwarp();
              ph.wait(bars, B_OR);           // (the next head's operands cannot arrive before the workers have seen this anyway)
            }
            // the attention operands are dead until the next layer, and BUF once out_proj has read it: they are the ring of this
            // layer's linear2 weight tiles
            bool buf_free = false;
            for (int c = c_base; c < c_base + NCH; ++c)
              for (int nh = 0; nh < 2; ++nh)
                for (int kb2 = 0; kb2 < 2; ++kb2) {
                  if (slot2 >= 2 && !buf_free) { ph.wait(bars, B_BUFF + 0); ph.wait(bars, B_BUFF + 1); buf_free = true; }
                  ph.wait(bars, B_W2EMPTY + slot2);
                  if (elect_one()) {
                    mbar_expect_tx(&bars[B_W2FULL + slot2], WSTAGE);
                    tma_load_2d(smem + w2stage_off(slot2), &tm_w2, &bars[B_W2FULL + slot2], c * 128 + kb2 * 64, l * 256 + nh * 128);
                  }
                  __syncwarp();
                  slot2 = (slot2 + 1 == NS2L) ? 0 : slot2 + 1;
                }
            if (!buf_free) { ph.wait(bars, B_BUFF + 0); ph.wait(bars, B_BUFF + 1); }      // keep the phase bits in step
          }
    }
  } else {
    // =================================================== workers ===================================================
    // warp = 4 * sub + q4: q4 = TMEM lane quarter (rows 32 q4 .. +31; quarter 3 carries no token and hosts the service
    // warps), sub = column quarter of every epilogue.  Three schedulers x four worker warps each.
    const int wl = sub * 3 + q4;                     // 0..11
    const int r = q4 * 32 + lane;                    // accumulator row == token slot s (row 0 = token, rows 1..88 = frames), 0..95
    const int wt = wl * 32 + lane;                   // 0..383
    const uint32_t tlane = tmem + ((uint32_t)(q4 * 32) << 16);
    Phases ph{(0x3ull << B_BUFF)};
    __nv_bfloat16* Zs = reinterpret_cast<__nv_bfloat16*>(smem + OFF_BUF);
    float* red_s = reinterpret_cast<float*>(smem + OFF_RED);            // [4][96] row sums
    float* red_q = reinterpret_cast<float*>(smem + OFF_B1 + 5632);      // [4][96] row sums of squares
    float* b1s = reinterpret_cast<float*>(smem + OFF_B1);
    uint8_t* XS = smem + OFF_XS;
    uint8_t* BUF = smem + OFF_BUF;
    const bool prof = PROF && P.prof != nullptr && blockIdx.x == 0 && wt == 0;
    long long pf[PROF ? PF_COUNT : 1];
#pragma unroll
    for (int i = 0; i < (PROF ? PF_COUNT : 1); ++i) pf[i] = 0;
    long long tmark = PROF ? clock64() : 0;
    auto lap = [&](int id) { if constexpr (PROF) { if (prof) { const long long c = clock64(); pf[id] += c - tmark; tmark = c; } } };

    auto release_acc = [&](int qa, int qb, bool xs_ready, int buf_ready) {
      tcgen05_fence_before();
      if (xs_ready || buf_ready >= 0) fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars[B_ACCF + qa]);
        if (qb >= 0) mbar_arrive(&bars[B_ACCF + qb]);
        if (xs_ready) mbar_arrive(&bars[B_XSR]);
        if (buf_ready >= 0) mbar_arrive(&bars[B_BUFR + buf_ready]);
      }
    };
    // residual + bias + LayerNorm on the accumulator in Q0|Q1, result -> XS (bf16).  Thread = (row, 64-column quarter): the
    // row's 64 values stay in registers across the one barrier that exchanges the partial sums.
    auto layernorm_epilogue = [&](const float* bias, const float* gamma, const float* beta, bool buf_token) {
      ph.wait(bars, B_ACCR + 0); ph.wait(bars, B_ACCR + 1);
      lap(PF_W_LN_WAIT);
      // CTA pair: out_proj has read BUF for the last time, the peer may send its linear2 partial sums into it
      if (CL > 1 && buf_token && wt == 0) mbar_arrive_remote(peer_bars + B_PFREE * 8);
      tcgen05_fence_after();
      float v[64];
      const int col0 = sub * 64;
      tmem_ld32_issue(tlane + col0, v); tmem_ld32_issue(tlane + col0 + 32, v + 32);
      tmem_ld_wait(); tie32(v); tie32(v + 32);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&bars[B_ACCF + 0]); mbar_arrive(&bars[B_ACCF + 1]); }     // the accumulator is in registers
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float rs[8];
        unpack8(*reinterpret_cast<const uint4*>(XS + xs_off(r, col0 + i * 8)), rs);
        const float4 b0 = *reinterpret_cast<const float4*>(bias + col0 + i * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + col0 + i * 8 + 4);
        add2(rs[0], rs[1], b0.x, b0.y); add2(rs[2], rs[3], b0.z, b0.w); add2(rs[4], rs[5], b1.x, b1.y); add2(rs[6], rs[7], b1.z, b1.w);
        add2(v[i * 8 + 0], v[i * 8 + 1], rs[0], rs[1]); add2(v[i * 8 + 2], v[i * 8 + 3], rs[2], rs[3]);
        add2(v[i * 8 + 4], v[i * 8 + 5], rs[4], rs[5]); add2(v[i * 8 + 6], v[i * 8 + 7], rs[6], rs[7]);
      }
      float sum1 = 0.f, sq1 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i += 2) { add2(sum, sum1, v[i], v[i + 1]); fma2(sq, sq1, v[i], v[i + 1], v[i], v[i + 1]); }
      sum += sum1; sq += sq1;
      red_s[sub * 96 + r] = sum; red_q[sub * 96 + r] = sq;
      quarter_sync(q4);
      sum = red_s[r] + red_s[96 + r] + red_s[192 + r] + red_s[288 + r];
      sq = red_q[r] + red_q[96 + r] + red_q[192 + r] + red_q[288 + r];
      const float mean = sum * (1.0f / D);
      const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + 1e-5f);
      const float nmr = -mean * rstd;                  // (v - mean) * rstd * g + b as two FMAs per element
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + col0 + i);
        const float4 b4 = *reinterpret_cast<const float4*>(beta + col0 + i);
        float t0 = nmr, t1 = nmr, t2 = nmr, t3 = nmr, o0 = b4.x, o1 = b4.y, o2 = b4.z, o3 = b4.w;
        fma2(t0, t1, v[i], v[i + 1], rstd, rstd); fma2(t2, t3, v[i + 2], v[i + 3], rstd, rstd);
        fma2(o0, o1, t0, t1, g4.x, g4.y); fma2(o2, o3, t2, t3, g4.z, g4.w);
        v[i] = o0; v[i + 1] = o1; v[i + 2] = o2; v[i + 3] = o3;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) *reinterpret_cast<uint4*>(XS + xs_off(r, col0 + i * 8)) = pack8(v + i * 8);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_XSR]);
      // `red` is rewritten by the next LayerNorm only after several more worker barriers
      lap(PF_W_LN);
    };
    // CTA pair, LayerNorm 2: the accumulator holds this CTA's K-half of linear2.  Rows are owned by TMEM lane quarter (rank 0:
    // quarters 0 and 2, rank 1: quarter 1): a non-owner warp ships its 64 columns to the owner's receive buffer (BUF, laid out
    // [column quarter][16-byte chunk][row] so that a warp's store is one contiguous 512 B), the owner adds them to its own half
    // and runs the LayerNorm; the bf16 rows go into both CTAs' XS.
    constexpr int RXROWS = 57;
    auto layernorm_pair = [&](const float* bias, const float* gamma, const float* beta) {
      ph.wait(bars, B_ACCR + 0); ph.wait(bars, B_ACCR + 1);
      lap(PF_W_LN_WAIT);
      tcgen05_fence_after();
      float v[64];
      const int col0 = sub * 64;
      tmem_ld32_issue(tlane + col0, v); tmem_ld32_issue(tlane + col0 + 32, v + 32);
      tmem_ld_wait(); tie32(v); tie32(v + 32);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&bars[B_ACCF + 0]); mbar_arrive(&bars[B_ACCF + 1]); }
      ph.waitc(bars, B_PFREE);                          // the peer's out_proj is done with its BUF (kept in step by every worker)
      const bool mine = (q4 == 1) == (rank == 1);
      const int rr = (q4 == 2 ? 32 : 0) + lane;         // row inside the owner's receive buffer
      // (the partial sums travel as bf16: distributed shared memory moves ~20 B per cycle, and an fp32 exchange of 89 x 256 values
      //  would cost more than the half of linear2 it saves; the rounding, 2^-9 of half the FFN output, is below that of the bf16
      //  residual stream it is added to)
      const uint32_t rx_off = (uint32_t)(OFF_BUF + (sub * 8 * RXROWS + rr) * 16);
      if (!mine) {
        if (r < S) {
#pragma unroll
          for (int i = 0; i < 8; ++i) st_remote16(peer_smem + rx_off + i * (RXROWS * 16), pack8(v + 8 * i));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(peer_bars + B_RX * 8);
        lap(PF_W_LN);
        return;
      }
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float rs[8];
        unpack8(*reinterpret_cast<const uint4*>(XS + xs_off(r, col0 + i * 8)), rs);
        const float4 b0 = *reinterpret_cast<const float4*>(bias + col0 + i * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + col0 + i * 8 + 4);
        add2(rs[0], rs[1], b0.x, b0.y); add2(rs[2], rs[3], b0.z, b0.w); add2(rs[4], rs[5], b1.x, b1.y); add2(rs[6], rs[7], b1.z, b1.w);
        add2(v[i * 8 + 0], v[i * 8 + 1], rs[0], rs[1]); add2(v[i * 8 + 2], v[i * 8 + 3], rs[2], rs[3]);
        add2(v[i * 8 + 4], v[i * 8 + 5], rs[4], rs[5]); add2(v[i * 8 + 6], v[i * 8 + 7], rs[6], rs[7]);
      }
      ph.waitc(bars, B_RX);                             // (residual and bias are in: the wait for the peer is the last thing)
      if (r < S) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float p8[8];
          unpack8(*reinterpret_cast<const uint4*>(smem + rx_off + i * (RXROWS * 16)), p8);
          add2(v[8 * i], v[8 * i + 1], p8[0], p8[1]); add2(v[8 * i + 2], v[8 * i + 3], p8[2], p8[3]);
          add2(v[8 * i + 4], v[8 * i + 5], p8[4], p8[5]); add2(v[8 * i + 6], v[8 * i + 7], p8[6], p8[7]);
        }
      }
      float sum1 = 0.f, sq1 = 0.f;
#pragma unroll
      for (int i = 0; i < 64; i += 2) { add2(sum, sum1, v[i], v[i + 1]); fma2(sq, sq1, v[i], v[i + 1], v[i], v[i + 1]); }
      sum += sum1; sq += sq1;
      red_s[sub * 96 + r] = sum; red_q[sub * 96 + r] = sq;
      quarter_sync(q4);
      sum = red_s[r] + red_s[96 + r] + red_s[192 + r] + red_s[288 + r];
      sq = red_q[r] + red_q[96 + r] + red_q[192 + r] + red_q[288 + r];
      const float mean = sum * (1.0f / D);
      const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + 1e-5f);
      const float nmr = -mean * rstd;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + col0 + i);
        const float4 b4 = *reinterpret_cast<const float4*>(beta + col0 + i);
        float t0 = nmr, t1 = nmr, t2 = nmr, t3 = nmr, o0 = b4.x, o1 = b4.y, o2 = b4.z, o3 = b4.w;
        fma2(t0, t1, v[i], v[i + 1], rstd, rstd); fma2(t2, t3, v[i + 2], v[i + 3], rstd, rstd);
        fma2(o0, o1, t0, t1, g4.x, g4.y); fma2(o2, o3, t2, t3, g4.z, g4.w);
        v[i] = o0; v[i + 1] = o1; v[i + 2] = o2; v[i + 3] = o3;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t off = (uint32_t)OFF_XS + xs_off(r, col0 + i * 8);
        const uint4 u = pack8(v + i * 8);
        *reinterpret_cast<uint4*>(smem + off) = u;
      }
      // the quarter's 32 rows are one contiguous 16 KB block of XS (row-group-major): one bulk copy per owned quarter into the peer's
      fence_async_smem();
      quarter_sync(q4);
      if (sub == 0 && lane == 0)
        bulk_copy_to_peer(peer_smem + OFF_XS + q4 * 4 * 4096, smem + OFF_XS + q4 * 4 * 4096, 4 * 4096, peer_bars + B_XSR * 8);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars[B_XSR]);
        if (sub == 0) mbar_arrive_expect_tx_remote(peer_bars + B_XSR * 8, 4 * 4096);      // announces the copy's bytes
        else mbar_arrive_remote(peer_bars + B_XSR * 8);
      }
      lap(PF_W_LN);
    };
    auto debug_dump = [&](int slot, int clip) {
      if (!(P.debug & 1) || CL > 1) return;                  // (CTA pair: the MMA warp writes the taps, see pair_dump)
      workers_sync();
      if (r < S) {
        float* o = P.dbg + (long long)slot * P.dbg_slot + ((long long)clip * S + r) * D + sub * 64;
        for (int c8 = 0; c8 < 8; ++c8) {
          float t8[8];
          unpack8(*reinterpret_cast<const uint4*>(XS + xs_off(r, sub * 64 + c8 * 8)), t8);
          for (int i = 0; i < 8; ++i) o[c8 * 8 + i] = t8[i];
        }
      }
    };

    for (int clip = cid0; clip < P.B; clip += ncl) {
      float* xc = P.x + (long long)clip * J * T;
      uint8_t* xac = P.xa + (long long)clip * XA_BYTES;
      const float* condc = P.cond + (long long)clip * T * D;
      for (int k = 0; k < P.n_run; ++k) {
        const int index = first_index - k;
        const int trow = P.tmap[index];
        const bool nz = (index != 0) && (P.sampler == 0);
        // (the A operand of the input GEMM is bulk-copied by the producer warp from the bf16 k-block image of x_t)
        ph.wait(bars, B_BUFF + 0); ph.wait(bars, B_BUFF + 1);
        lap(PF_W_STAGE);
        // ---------------- input epilogue: + cond + TW[t], rotary (position = frame), bf16 -> Z staging (in BUF)
        ph.wait(bars, B_ACCR + 0); ph.wait(bars, B_ACCR + 1);
        lap(PF_W_IN_WAIT);
        tcgen05_fence_after();
        // BUF held x_t / z chunks and A k-blocks since the last step: the pad rows of the Z staging (read as V with zero weight
        // by the last windows) must be finite
        for (int i = wt; i < (ZROWS - T) * 33; i += NWT)
          reinterpret_cast<uint4*>(Zs + (T + i / 33) * ZLD)[i % 33] = make_uint4(0u, 0u, 0u, 0u);
        {
          const int f = r - 1;
          const bool ok = r >= 1 && r <= T;
          float v[32];
#pragma unroll 1
          for (int c2 = 0; c2 < (CL > 1 ? 1 : 2); ++c2) {
            // == local head (sub*2 + c2) * 32; CTA pair: the rank's 128 columns, 32 per column quarter
            const int col0 = CL > 1 ? rank * 128 + sub * 32 : sub * 64 + c2 * 32;
            tmem_ld32(tlane + col0, v);
            if (!ok) continue;
            const float* cr = condc + (long long)f * D + col0;
            const float* tw = P.TW + (long long)trow * D + col0;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 c = __ldg(reinterpret_cast<const float4*>(cr + i));
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(tw + i));
              v[i] += c.x + t4.x; v[i + 1] += c.y + t4.y; v[i + 2] += c.z + t4.z; v[i + 3] += c.w + t4.w;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float2 cs = __ldg(P.cs + f * 16 + i);
              const float a = v[i], b = v[i + 16];
              v[i] = a * cs.x - b * cs.y;
              v[i + 16] = b * cs.x + a * cs.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(Zs + f * ZLD + col0 + i * 8) = pack8(v + i * 8);
          }
        }
        release_acc(0, 1, false, -1);
        workers_sync();
        lap(PF_W_IN_EPI);
        // ---------------- windowed causal local attention on mma.sync (q = k = v = Z), rotary (position = frame + 1) -> XS
        for (int item = wl; item < 64 / CL; item += NW) {
          const int w = CL > 1 ? item >> 2 : item >> 3, lh = CL > 1 ? rank * 4 + (item & 3) : item & 7;
          const int q0 = WIN * w, k0 = (w == 0) ? 0 : WIN * (w - 1), nk = q0 + WIN - k0;
          float sc[4][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) { sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f; }
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            uint32_t a[4], b0, b1, b2, b3, c0, c1;
            ldsm_x4(a[0], a[1], a[2], a[3], Zs + (q0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ZLD + lh * 32 + kk * 16 + (lane >> 4) * 8);
            ldsm_x4(b0, b1, b2, b3, Zs + (k0 + (lane & 7) + (lane >> 4) * 8) * ZLD + lh * 32 + kk * 16 + ((lane >> 3) & 1) * 8);
            ldsm_x2(c0, c1, Zs + (k0 + 16 + (lane & 7)) * ZLD + lh * 32 + kk * 16 + ((lane >> 3) & 1) * 8);
            mma_bf16_16816(sc[0], a, b0, b1);
            mma_bf16_16816(sc[1], a, b2, b3);
            mma_bf16_16816(sc[2], a, c0, c1);
          }
          const int g = lane >> 2, t2 = (lane & 3) * 2;
          float mx0 = -3.0e38f, mx1 = -3.0e38f;
#pragma unroll
          for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cc = nt * 8 + t2 + e;
              if (!(cc < nk && k0 + cc <= q0 + g)) sc[nt][e] = -3.0e38f;
              if (!(cc < nk && k0 + cc <= q0 + g + 8)) sc[nt][2 + e] = -3.0e38f;
              mx0 = fmaxf(mx0, sc[nt][e]); mx1 = fmaxf(mx1, sc[nt][2 + e]);
            }
          mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
          mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
          const float sl2 = 1.4426950408889634f * 0.17677669529663687f;     // log2(e) / sqrt(32)
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int nt = 0; nt < 3; ++nt) {
            sc[nt][0] = ex2_fast((sc[nt][0] - mx0) * sl2); sc[nt][1] = ex2_fast((sc[nt][1] - mx0) * sl2);
            sc[nt][2] = ex2_fast((sc[nt][2] - mx1) * sl2); sc[nt][3] = ex2_fast((sc[nt][3] - mx1) * sl2);
            s0 += sc[nt][0] + sc[nt][1]; s1 += sc[nt][2] + sc[nt][3];
          }
          s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
          float oc[4][4];
#pragma unroll
          for (int dt = 0; dt < 4; ++dt) { oc[dt][0] = oc[dt][1] = oc[dt][2] = oc[dt][3] = 0.f; }
#pragma unroll
          for (int kt = 0; kt < 2; ++kt) {
            uint32_t a[4];
            a[0] = pack_bf16x2(sc[2 * kt][0], sc[2 * kt][1]); a[1] = pack_bf16x2(sc[2 * kt][2], sc[2 * kt][3]);
            a[2] = pack_bf16x2(sc[2 * kt + 1][0], sc[2 * kt + 1][1]); a[3] = pack_bf16x2(sc[2 * kt + 1][2], sc[2 * kt + 1][3]);
#pragma unroll
            for (int dp = 0; dp < 2; ++dp) {
              uint32_t b0, b1, b2, b3;
              ldsm_x4_t(b0, b1, b2, b3, Zs + (k0 + kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ZLD + lh * 32 + dp * 16 + (lane >> 4) * 8);
              mma_bf16_16816(oc[2 * dp], a, b0, b1);
              mma_bf16_16816(oc[2 * dp + 1], a, b2, b3);
            }
          }
          const float i0 = 1.0f / s0, i1 = 1.0f / s1;
#pragma unroll
          for (int hr = 0; hr < 2; ++hr) {
            const int rr = g + 8 * hr;
            if (rr >= WIN) continue;
            const int pos = q0 + rr + 1;                 // XS row (token is row 0) == rotary position
            const float inv = hr ? i1 : i0;
#pragma unroll
            for (int dt = 0; dt < 2; ++dt) {
              float ol[2], oh[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float2 cs = __ldg(P.cs + pos * 16 + dt * 8 + t2 + e);
                const float a = oc[dt][2 * hr + e] * inv, b = oc[dt + 2][2 * hr + e] * inv;
                ol[e] = a * cs.x - b * cs.y;
                oh[e] = b * cs.x + a * cs.y;
              }
              *reinterpret_cast<uint32_t*>(XS + xs_off(pos, lh * 32 + dt * 8 + t2)) = pack_bf16x2(ol[0], ol[1]);
              *reinterpret_cast<uint32_t*>(XS + xs_off(pos, lh * 32 + 16 + dt * 8 + t2)) = pack_bf16x2(oh[0], oh[1]);
            }
          }
        }
        if (wt < D / CL) {   // token row: tok = emb_1 + emb_t (rotary at position 0 is the identity)
          const int tc = CL > 1 ? rank * 128 + wt : wt;
          const float tv = __ldg(P.emb1 + (long long)clip * D + tc) + __ldg(P.te + (long long)trow * D + tc);
          *reinterpret_cast<__nv_bfloat16*>(XS + xs_off(0, tc)) = __float2bfloat16_rn(tv);
        }
        fence_async_smem();
        workers_sync();                                  // Z staging (BUF) is dead from here on
        if constexpr (CL > 1) {
          // this CTA's 128 columns (k-tiles 2 rank, 2 rank + 1: 2 KB per 8-row group of the row-group-major XS) go into the peer's XS
          // as 12 bulk copies counted on the peer's B_XSR; the bytes coming the other way are announced on OUR barrier by warp 0
          if (wl == 0 && lane < 12) {
            const uint32_t off = (uint32_t)(OFF_XS + lane * 4096 + rank * 2048);
            bulk_copy_to_peer(peer_smem + off, smem + off, 2048, peer_bars + B_XSR * 8);
          }
          __syncwarp();
          if (lane == 0) { if (wl == 0) mbar_expect_tx(&bars[B_XSR], 12 * 2048); else mbar_arrive(&bars[B_XSR]); }
        } else if (lane == 0) mbar_arrive(&bars[B_XSR]);
        if (CL > 1 && wt == 0) mbar_arrive_remote(peer_bars + B_PFREE * 8);      // ... and the peer may write head outputs into it
        lap(PF_W_LOCAL);
        debug_dump(0, clip);

        // ---------------- transformer layers
        for (int l = 0; l < NL; ++l) {
          const float* lp = P.lparams + (long long)l * P_SIZE;
          float* lnp = reinterpret_cast<float*>(smem + OFF_LNP);
          workers_sync();                                // every warp is done with the previous layer's LayerNorm parameters
          // With 225 KB of shared memory the L1 is a few KB: every parameter read would be an L2 round trip, so the layer's
          // parameters are staged in shared memory once; the first barrier of the attention phase orders them before use.
          for (int i = wt; i < 448; i += NWT) {
            if (i < 256) {                               // linear1 bias as fp16; LayerNorm gains / biases (g1|be1 and g2|be2 are contiguous)
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(lp + P_B1) + i);
              reinterpret_cast<__half2*>(b1s)[2 * i] = __floats2half2_rn(t4.x, t4.y);
              reinterpret_cast<__half2*>(b1s)[2 * i + 1] = __floats2half2_rn(t4.z, t4.w);
              reinterpret_cast<float4*>(lnp)[i] = __ldg(reinterpret_cast<const float4*>(lp + ((i >> 7) ? P_G2 : P_G1)) + (i & 127));
            } else if (i < 320) {                        // q bias (per head, [h][64]); the k bias drops out of the softmax,
              const int j = i - 256, h = j >> 4, i4 = j & 15;   // the v bias is folded into bo' = bo + Wo bv at set-up
              reinterpret_cast<float4*>(b1s + 512)[j] = __ldg(reinterpret_cast<const float4*>(lp + P_BQKV + h * 192) + i4);
            } else if (i < 384) {
              reinterpret_cast<float4*>(b1s + 768)[i - 320] = __ldg(reinterpret_cast<const float4*>(lp + P_BO) + (i - 320));
            } else {
              reinterpret_cast<float4*>(b1s + 1024)[i - 384] = __ldg(reinterpret_cast<const float4*>(lp + P_B2) + (i - 384));
            }
          }
          workers_sync();
          for (int hl = 0; hl < NHL; ++hl) {
            const int hb = hl & 1, h = h_base + hl;
            ph.wait(bars, B_ACCR + 2 * hb); ph.wait(bars, B_ACCR + 2 * hb + 1);
            lap(PF_W_QKV_WAIT);
            tcgen05_fence_after();
            const uint32_t thalf = tlane + (uint32_t)(hb * 256);
            // ---- q | k | v of this head -> tcgen05 operands.  Thread = (token row r, 16-column slice `sub` of each of q, k, v).
            {
              float v48[48];
              tmem_ld16_issue(thalf + sub * 16, v48);
              tmem_ld16_issue(thalf + 64 + sub * 16, v48 + 16);
              tmem_ld16_issue(thalf + 128 + sub * 16, v48 + 32);
              tmem_ld_wait(); tie32(v48); tie4(v48 + 32); tie4(v48 + 36); tie4(v48 + 40); tie4(v48 + 44);
              uint8_t* qd = smem + AT_Q + (sub * 2) * LBO_Q + (r >> 3) * 128 + (r & 7) * 16;
              uint8_t* kd = smem + AT_K + (sub * 2) * LBO_K + (r >> 3) * 128 + (r & 7) * 16;
              uint8_t* vd = smem + AT_V + (r >> 3) * LBO_V + (sub * 2) * 128 + (r & 7) * 16;      // MN-major: [key group][d group][key][8 d]
              if (r < S) {
                const float* bq = b1s + 512 + h * 64 + sub * 16;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(bq + i);
                  add2(v48[i], v48[i + 1], b4.x, b4.y); add2(v48[i + 2], v48[i + 3], b4.z, b4.w);
                }
                *reinterpret_cast<uint4*>(qd) = pack8(v48);          *reinterpret_cast<uint4*>(qd + LBO_Q) = pack8(v48 + 8);
                *reinterpret_cast<uint4*>(kd) = pack8(v48 + 16);     *reinterpret_cast<uint4*>(kd + LBO_K) = pack8(v48 + 24);
                *reinterpret_cast<uint4*>(vd) = pack8(v48 + 32);     *reinterpret_cast<uint4*>(vd + 128) = pack8(v48 + 40);
              } else {
                const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(qd) = z4; *reinterpret_cast<uint4*>(qd + LBO_Q) = z4;
                *reinterpret_cast<uint4*>(kd) = z4; *reinterpret_cast<uint4*>(kd + LBO_K) = z4;
                *reinterpret_cast<uint4*>(vd) = z4; *reinterpret_cast<uint4*>(vd + 128) = z4;
              }
            }
            if constexpr (CL == 1) {
              if (l > 0 && hb == 0) ph.wait(bars, B_BUFF + (h >> 1));  // linear2 of the previous layer has consumed this BUF half
            } else {
              if (l > 0 && hb == 0) { ph.wait(bars, B_BUFF + 0); ph.wait(bars, B_BUFF + 1); }
              if (l == 0 && hl == 0) ph.waitc(bars, B_PFREE);          // the peer's local attention is done with its Z staging
            }
            tcgen05_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_QKR]);
            lap(PF_W_EXTRACT);
            // ---- softmax over the 96 key columns of S (thread = row r, keys 24 sub .. 24 sub + 23); P (bf16) -> A operand of P V
            ph.wait(bars, B_SR);
            lap(PF_W_SYNC1);
            tcgen05_fence_after();
            {
              float sc[24];
              tmem_ld16_issue(thalf + sub * 24, sc);
              tmem_ld8_issue(thalf + sub * 24 + 16, sc + 16);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 24; i += 4) tie4(sc + i);
              float mx = -3.0e38f;
#pragma unroll
              for (int i = 0; i < 24; ++i) {
                if (sub * 24 + i >= S) sc[i] = -3.0e38f;
                mx = fmaxf(mx, sc[i]);
              }
              red_s[sub * 96 + r] = mx;
              quarter_sync(q4);
              mx = fmaxf(fmaxf(red_s[r], red_s[96 + r]), fmaxf(red_s[192 + r], red_s[288 + r]));
              const float sl2 = 1.4426950408889634f * 0.125f;
              float sum = 0.f;
              const float mxs = mx * sl2;
#pragma unroll
              for (int i = 0; i < 24; i += 2) {
                float t0 = -mxs, t1 = -mxs;
                fma2(t0, t1, sc[i], sc[i + 1], sl2, sl2);
                sc[i] = ex2_fast(t0); sc[i + 1] = ex2_fast(t1);
              }
              float sum1 = 0.f;
#pragma unroll
              for (int i = 0; i < 24; i += 2) add2(sum, sum1, sc[i], sc[i + 1]);
              sum += sum1;
              red_q[sub * 96 + r] = sum;
              uint8_t* pd = smem + AT_P + (sub * 3) * LBO_P + (r >> 3) * 128 + (r & 7) * 16;
              *reinterpret_cast<uint4*>(pd) = pack8(sc);
              *reinterpret_cast<uint4*>(pd + LBO_P) = pack8(sc + 8);
              *reinterpret_cast<uint4*>(pd + 2 * LBO_P) = pack8(sc + 16);
            }
            tcgen05_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[B_PR]);
            lap(PF_W_ATT_MMA);
            // ---- O / sum -> BUF (A operand of out_proj), 16 head-dim columns per thread
            ph.wait(bars, B_OR);
            tcgen05_fence_after();
            {
              float o16[16];
              tmem_ld16_issue(thalf + 96 + sub * 16, o16);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; i += 4) tie4(o16 + i);
              const float inv = 1.0f / (red_q[r] + red_q[96 + r] + red_q[192 + r] + red_q[288 + r]);
#pragma unroll
              for (int i = 0; i < 16; i += 2) mul2(o16[i], o16[i + 1], inv, inv);
              const uint32_t o_off = (uint32_t)OFF_BUF + a_off(r, h * HD + sub * 16), o_off8 = (uint32_t)OFF_BUF + a_off(r, h * HD + sub * 16 + 8);
              const uint4 u0 = pack8(o16), u1 = pack8(o16 + 8);
              *reinterpret_cast<uint4*>(smem + o_off) = u0;
              *reinterpret_cast<uint4*>(smem + o_off8) = u1;
            }
            tcgen05_fence_before();
            fence_async_smem();
            if constexpr (CL > 1) {
              // out_proj runs on both CTAs: once both heads of this CTA are in BUF (k-tiles h - 1, h: 12 KB of token rows each) and
              // every worker's stores are fenced for the async proxy, ONE thread copies them into the peer's BUF.  (Per-thread
              // st.shared::cluster of the same data: 16 B per 32 B sector on the SM-to-SM path and a GPU-scope membar per warp,
              // 11 us per step slower.)
              if (hb == 1) {
                workers_sync();
                if (wt == 0) {
                  const uint32_t rb = peer_bars + (B_BUFR + (h >> 1)) * 8;
                  bulk_copy_to_peer(peer_smem + OFF_BUF + (h - 1) * KT, smem + OFF_BUF + (h - 1) * KT, XA_ROWS_BYTES, rb);
                  bulk_copy_to_peer(peer_smem + OFF_BUF + h * KT, smem + OFF_BUF + h * KT, XA_ROWS_BYTES, rb);
                }
              }
            }
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&bars[B_ACCF + 2 * hb]); mbar_arrive(&bars[B_ACCF + 2 * hb + 1]);      // the head's TMEM half is free
              if (hb == 1) {                                                                       // both heads of this BUF half are written
                mbar_arrive(&bars[B_BUFR + (h >> 1)]);
                if constexpr (CL > 1) {                                                            // (warp 0 announces the copies' bytes)
                  if (wl == 0) mbar_arrive_expect_tx_remote(peer_bars + (B_BUFR + (h >> 1)) * 8, 2 * XA_ROWS_BYTES);
                  else mbar_arrive_remote(peer_bars + (B_BUFR + (h >> 1)) * 8);
                }
              }
            }
            lap(PF_W_ATT_MERGE);
          }
          layernorm_epilogue(b1s + 768, lnp, lnp + 256, true);
          // ---- FFN: GELU epilogue per 128-unit chunk; the fp16 hidden goes back into the chunk's own accumulator quarter
          // (columns 0..63: column j = units (j, 64 + j)) as the tensor-memory A operand of linear2
          for (int c = 0; c < NCH; ++c) {
            const int qd = 2 + (c & 1);
            ph.wait(bars, B_ACCR + qd);
            lap(PF_W_GELU_WAIT);
            tcgen05_fence_after();
            {
              float va[32];
              const int cc0 = sub * 16;                  // this thread's units: cc0 .. cc0 + 15 and 64 + cc0 .. 64 + cc0 + 15
              tmem_ld16_issue(tlane + qd * 128 + cc0, va);
              tmem_ld16_issue(tlane + qd * 128 + 64 + cc0, va + 16);
              tmem_ld_wait(); tie32(va);
              // GELU in packed fp16 (tanh form on MUFU.TANH, 4 instructions per element): the hidden is stored as fp16, which
              // keeps 3 more mantissa bits than the bf16 it replaces; |half-tanh GELU - exact| rms 5e-4 vs 2e-3 for bf16(exact)
              const __half* b1h = reinterpret_cast<const __half*>(b1s) + (c_base + c) * 128 + cc0;
              uint32_t ha[16];
#pragma unroll
              for (int i = 0; i < 16; ++i)
                ha[i] = gelu_h2(__hadd2(__floats2half2_rn(va[i], va[16 + i]), __halves2half2(b1h[i], b1h[64 + i])));
              tmem_st16(tlane + qd * 128 + cc0, ha);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&bars[B_BUFR + (c & 1)]);
              if (c >= NCH - 2) mbar_arrive(&bars[B_ACCF + qd]);      // the quarter's next user (in_proj of the next layer / the pose head) waits for this
            }
            lap(PF_W_GELU);
          }
          if constexpr (CL > 1) layernorm_pair(b1s + 1024, lnp + 512, lnp + 768);
          else layernorm_epilogue(b1s + 1024, lnp + 512, lnp + 768, false);
          debug_dump(l + 1, clip);
        }

        // ---------------- pose head + posterior: x <- f(x0, x, z) in place (fp32, global) and as next step's bf16 A k-blocks.
        // Chunk c = 32 joint channels (TMEM quarter (c >> 2) & 3, columns (c & 3) * 32); the four column quarters take 8 each.
        // the head bias takes the place of the layer parameters once EVERY warp has read the linear2 bias of the last LayerNorm
        // (CTA pair: the non-owner warps leave that LayerNorm long before the owners have their partial sums)
        workers_sync();
        for (int i = wt; i < JPAD; i += NWT) b1s[i] = __ldg(P.bout + i);
        workers_sync();
        // CTA pair: every worker is through with the LayerNorm 2 receive buffer (BUF): the x_t / z chunks may land there
        if (CL > 1 && wt == 0) mbar_arrive(&bars[B_HGO]);
        lap(PF_W_ZWAIT);
        {
          const float4 cf = P.coef[index];
          const int f = r - 1;
          const bool ok = r >= 1 && r <= T;
          for (int tt = 0; tt < tile1 - tile0; ++tt) {
            const int t = tile0 + tt;
            // the whole 128-channel tile leaves tensor memory at once: one load round trip per tile instead of four, and the
            // accumulator quarter goes back to the MMA thread before the posterior work instead of after it
            float v32[4][8];                             // this thread's 8 channels of each of the tile's four chunks
            ph.wait(bars, B_ACCR + (tt & 3));
            tcgen05_fence_after();
#pragma unroll
            for (int q = 0; q < 4; ++q) tmem_ld8_issue(tlane + (tt & 3) * 128 + q * 32 + sub * 8, v32[q]);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) { tie4(v32[q]); tie4(v32[q] + 4); }
            release_acc(tt & 3, -1, false, -1);
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              const int c = 4 * t + sl;
              ph.wait(bars, B_HFULL + sl);
              lap(PF_W_HEAD_WAIT);
              const int j0 = c * HCH + sub * 8;
              float* v8 = v32[sl];
              if (ok && j0 < J) {
                const float* xs = reinterpret_cast<const float*>(smem + hslot_x(sl)) + sub * 8 * T + f;
                const float* zs = reinterpret_cast<const float*>(smem + hslot_z(sl)) + sub * 8 * T + f;
                const float* bo = b1s + j0;
                float* xg = xc + (long long)j0 * T + f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const bool in = j0 + i < J;
                  const float xt = in ? xs[i * T] : 0.f, zz = (nz && in) ? zs[i * T] : 0.f;
                  const float x0 = v8[i] + bo[i];
                  v8[i] = in ? posterior_apply(P.sampler, cf, x0, xt, zz, nz) : 0.f;
                  if (in) xg[(long long)i * T] = v8[i];
                }
                uint8_t* row = xac + (long long)(j0 >> 6) * KT + (r >> 3) * 1024 + (r & 7) * 128;
                *reinterpret_cast<uint4*>(row + ((((j0 & 63) >> 3) ^ (r & 7)) << 4)) = pack8(v8);
              }
              __syncwarp();
              if (lane == 0) mbar_arrive(&bars[B_HEMPTY + sl]);
              lap(PF_W_HEAD);
            }
          }
        }
        fence_async_all();                               // x and its k-block image are read back by bulk copies (async proxy)
        __syncwarp();
        if (lane == 0) {
          if constexpr (CL > 1) { mbar_arrive_release_cluster(&bars[B_XAR]); mbar_arrive_remote(peer_bars + B_XAR * 8); }
          else mbar_arrive(&bars[B_XAR]);
          if (nz) mbar_arrive(&bars[B_ZF]);
        }
      }
    }
    if constexpr (PROF) { if (prof) for (int i = PF_W_STAGE; i <= PF_W_ATT_MERGE; ++i) P.prof[i] = pf[i]; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CL > 1) cluster_sync_all();                // no CTA leaves while its peer can still store into / arrive on its shared memory
  if (warp == 7) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace clip
