// tcgen05 implicit-GEMM 1-D convolution for sm_100a: the contraction kernel of the tensor-core
// path (every conv of the generator except the 16->1 tail runs through it).
//
// Formulation ("time on M"):   D[t, co] = sum_j sum_ci  A[t + off_j, ci] * W_j[co, ci]
//   M = 128 output time steps per MMA (MSUB sub-tiles per CTA tile), N = Cout tile, K = Cin per tap.
//   A  activations, 16-bit, panel-blocked [B][C/PW][L][PW]
//   W  weights, 16-bit, pre-packed + pre-swizzled on the host, [tap][panel][N][PW]
//   D  fp32 accumulators in TMEM (MSUB * N columns, double buffered when they fit twice)
//
// Layout: 16-bit activations are "panel blocked" in HBM, [B][C/PW][L][PW] with PW = min(C, 64)
// channels, so one time step of a panel is one row of 128 / 64 / 32 bytes -- exactly the row of
// the UMMA K-major SWIZZLE_128B / 64B / 32B canonical layouts.  TMA (cp.async.bulk.tensor, same
// swizzle mode) stages the input tile with its +-(k-1)d/2 halo as [panel][rows][PW]; rows are
// contiguous, so the operand of filter tap j is the SAME staged tile with the descriptor start
// address advanced by off_j rows: dilated taps cost no data movement and no im2col.  (The swizzle
// XOR is a function of the absolute shared-memory address bits, so a start address in the middle
// of an 8-row swizzle atom addresses the rows the TMA wrote.)  A first version used the
// no-swizzle "interleaved" layout; ncu showed the tensor core reading it at ~16 B/cycle (3x off
// the MMA rate for N = 256 and ~200 cycles per MMA for narrow N), see profiles/README.md.
// TMA zero-fills rows outside [0, L): that is the per-layer zero padding of the reference convs
// (nn.py:98-166).  Weights are pre-swizzled on the host and stream through an mbarrier ring
// with 1-D bulk copies (or stay resident when the whole filter bank is small).  The polyphase
// transposed convs (archi.py:47-59) are the same kernel: phase phi is a conv with taps at rows
// (off_phi - m) whose output row is u*q + phi.
//
// Persistent CTAs: grid.x CTAs walk the (item, time-tile) list with stride grid.x, so barrier
// setup, TMEM allocation and (resident) weights are paid once and the three pipelines overlap
// across tiles:
//   warps 0..7  epilogue: TMEM -> registers -> bias / residual / MRF sum / leaky-ReLU -> global,
//               two warps per TMEM lane group, each taking half of the N columns
//   warp 8      A producer: TMA tile loads into a 1- or 2-deep ring            (a_full / a_empty)
//   warp 9      W producer: bulk copies of weight stages                       (w_full / w_empty)
//   warp 10     MMA issuer (one elected lane) + TMEM owner                     (acc_full / acc_empty)
//   (the issuer is the last warp: the scheduler prefers the highest eligible warp id)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace sa {
namespace tc {

constexpr int kThreads = 352;           // 11 warps
constexpr int kEpiWarps = 8;
constexpr int kMaxPhases = 8;
constexpr int kMaxStages = 8;           // weight ring depth (runtime n_wstages <= this)

// epilogue flags
enum : uint32_t {
  EPI_RES = 1u << 0,        // v += res32
  EPI_OUT32 = 1u << 1,      // out32 = v
  EPI_OUT16 = 1u << 2,      // out16 = 16-bit(lrelu(v, slope_out))
  EPI_SUM_SET = 1u << 3,    // sum32 = v                       (MRF, first ResBlock)
  EPI_SUM_ADD = 1u << 4,    // sum32 += v                      (MRF, middle ResBlocks)
  EPI_SUM_FIN = 1u << 5,    // v = (sum32 + v) / n_blocks      (MRF, last ResBlock; archi.py:86)
  EPI_BF16 = 1u << 6        // 16-bit type is bf16 (else fp16)
};

struct ConvParams {
  CUtensorMap tmap;         // activations [B][Cin/PW][L_in][PW], box {PW, box_rows, 1, 1}, swizzle = row bytes
  CUtensorMap wmap;         // conv_pair_tc.cuh only: packed weights as [rows][64] 16-bit, box {64, rows of one stage}
  const void* w;            // packed weights: phase p / n-tile t at w + (p * n_tiles + t) * w_tile_bytes
  const float* bias;        // [Cout_total]
  const float* res32;       // fp32 blocked [B][Cout_total/8][L_out][8]
  float* out32;
  float* sum32;
  void* out16;              // 16-bit panel-blocked [B][Cout/out_pw][L_out][out_pw]
  int* error_flag;          // raised when a barrier wait times out
  long long* timing;        // optional [16] cycle counters (diagnostics): MMA warp total / wait A / wait W / wait acc-empty;
                            // epilogue warp 0 total / wait acc-full; A producer total / wait a-empty
  int cin;                  // multiple of 16
  int cout_total;           // channels of the output tensor
  int m_rows;               // valid output rows per item on the M axis (L for conv, L_in for convT)
  int l_out;                // rows of the output tensor (m_rows * out_stride)
  int out_stride;           // 1 (conv) or u (transposed conv)
  int n_phases;             // 1 (conv) or u
  int n_tiles;              // Cout_total / N
  int tap_step;             // +dilation (conv) or -1 (transposed conv)
  int tap_base[kMaxPhases]; // row offset of tap 0 of each phase
  int n_taps[kMaxPhases];
  int row_lo[kMaxPhases];   // min tap offset of the phase (first staged row = m0 + row_lo)
  int rows_alloc;           // staged rows per panel (nseg * box_rows), multiple of 8
  int box_rows, nseg;
  int out_pw;               // panel width of the 16-bit output tensor
  int k16_per_stage;        // K=16 steps per weight stage
  int n_wstages;            // weight ring depth
  int w_resident;           // 1: all weight stages stay in smem (loaded once per CTA)
  int n_abuf;               // A-tile ring depth (1 or 2)
  int cluster;              // 1: launched as 2-CTA clusters; weight stages are multicast (each CTA loads half)
  int m_tiles;              // time tiles per item
  int total_tiles;          // m_tiles * B
  uint32_t w_tile_bytes;
  uint32_t flags;
  float slope_out;
  float n_blocks;           // divisor of EPI_SUM_FIN (true division, as xs / num_kernels)
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error, never as a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* error_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {          // ~2 s at 2 GHz
      if (error_flag) atomicExch(error_flag, 1);
      return false;
    }
  }
  return true;
}
// Same, for the producer and epilogue warps: back off between polls so that waiting warps do not take
// issue slots from the MMA warp (the scheduler picks the highest warp id among the eligible warps; the
// MMA issuer is therefore also the LAST warp of the CTA).
__device__ __forceinline__ bool mbar_wait_relaxed(uint32_t bar, uint32_t parity, int* error_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(40);
    if (clock64() - t0 > 4000000000LL) {
      if (error_flag) atomicExch(error_flag, 1);
      return false;
    }
  }
  return true;
}
// One lane of a converged warp.  The role loops below run warp-uniformly (all 32 lanes walk the
// tiles and poll the barriers) and only the TMA / MMA / commit instruction itself is issued by
// the elected lane: descriptors and addresses then stay in uniform registers.  (Running the
// whole loop under `if (lane == 0)` made ptxas wrap every UTCHMMA in an ELECT/BRA uniformization
// loop with R2UR moves and the kernel became MMA-issue bound; see profiles/README.md.)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
// Multicast variants for 2-CTA clusters: the data lands at the same CTA-relative offset in every CTA of
// `mask`, and so does the mbarrier complete_tx / arrive.
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers, 32 lanes x 16 columns of 32-bit (thread i gets 16 columns of lane base+i).
// Asynchronous until tmem_ld_wait().
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, swizzled canonical layout, version 1 (sm_100), as two
// 32-bit halves (only the start field, low 14 bits of `lo`, changes between MMAs):
//   lo: [0,14) start >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major; 1)
//   hi: [0,14) SBO >> 4 (8 rows = 8 * row bytes) | [14,16) version = 1 | [17,20) base offset
//       | [29,32) layout type: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t row_bytes) {
  const uint32_t layout = row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u;
  return (((8u * row_bytes) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}
// The base-offset field stays 0: the hardware applies the swizzle XOR to the absolute shared-memory
// address bits, so a start address in the middle of an 8-row atom is fine (verified on B200: setting
// base_offset = (start >> 7) & 7 breaks the result, leaving it 0 is bit-exact).
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
// byte offset inside a swizzled buffer whose base is aligned to the swizzle pattern (8 rows)
__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t row_bytes) {
  const uint32_t mask = row_bytes == 128 ? 7u : row_bytes == 64 ? 3u : 1u;
  return off ^ (((off >> 7) & mask) << 4);
}

// Instruction descriptor, kind::f16: fp32 accumulate, A/B fp16 (0) or bf16 (1), both K-major,
// M = 128, N = n.   bits [4,6) c_format=1 | [7,10) a_format | [10,13) b_format | [17,23) N>>3 | [24,29) M>>4
__device__ __forceinline__ uint32_t make_idesc(int n, bool bf16) {
  const uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ float lrelu_f(float v, float s) { return v >= 0.f ? v : v * s; }

__device__ __forceinline__ uint4 pack8(const float (&v)[8], bool bf16) {
  uint4 o;
  if (bf16) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  } else {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  }
  return o;
}

// 8 fp32 -> 8 x 16-bit with leaky-ReLU applied in packed 16-bit arithmetic: h = round(v); h = max(h, slope*h)
// (slope < 1).  One convert + two packed ops per PAIR of values instead of ~4 fp32 ops per value: the
// epilogues of the narrow layers are instruction-issue bound (profiles/README.md).  `keep` = false zeroes
// the result (rows outside the utterance).  Both the per-layer and the fused kernels use this routine, so
// their outputs stay bit-identical.
__device__ __forceinline__ uint4 pack8_lrelu(const float (&v)[8], float slope, bool keep, bool bf16) {
  uint4 o;
  uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
  if (bf16) {
    const __nv_bfloat162 s2 = __float2bfloat162_rn(slope);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      h = __hmax2(h, __hmul2(h, s2));
      ow[i] = keep ? *reinterpret_cast<uint32_t*>(&h) : 0u;
    }
  } else {
    const __half2 s2 = __float2half2_rn(slope);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
      h = __hmax2(h, __hmul2(h, s2));
      ow[i] = keep ? *reinterpret_cast<uint32_t*>(&h) : 0u;
    }
  }
  return o;
}

__device__ __forceinline__ float4 ldg_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void stg_f4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

// ---------------------------------------------------------------------------------------
// The kernel.  grid = (persistent CTAs, n_phases * n_tiles).  Dynamic smem:
//   [A ring: n_abuf * (cin/pw) * rows_alloc * 2pw][W ring: n_wstages * k16_per_stage * N * 32]
//   [bias N*4][barriers][tmem holder]      (base rounded up to 1024 B: swizzle pattern anchor)
// ---------------------------------------------------------------------------------------
template <int N, int MSUB, int PW>
__global__ void __launch_bounds__(kThreads, (N <= 64) ? 2 : 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int phase = blockIdx.y / p.n_tiles, ntile = blockIdx.y % p.n_tiles;

  constexpr int kAccCols = N * MSUB;                              // columns of one accumulator buffer
  constexpr int kNumAcc = (2 * kAccCols <= 512) ? 2 : 1;
  constexpr uint32_t kTmemCols = (kNumAcc * kAccCols <= 32) ? 32 : (kNumAcc * kAccCols <= 64) ? 64
                                 : (kNumAcc * kAccCols <= 128) ? 128 : (kNumAcc * kAccCols <= 256) ? 256 : 512;

  constexpr uint32_t row_bytes = (uint32_t)PW * 2u;                // 128 / 64 / 32
  constexpr int SPP = PW / 16;                                     // K=16 steps per panel row
  const int panels = p.cin / PW;
  const uint32_t panel_bytes = (uint32_t)p.rows_alloc * row_bytes; // multiple of the 8-row swizzle pattern
  const uint32_t a_bytes = (uint32_t)panels * panel_bytes;
  const uint32_t stage_bytes = (uint32_t)p.k16_per_stage * N * 32u;
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + (size_t)p.n_abuf * a_bytes;
  float* bias_s = reinterpret_cast<float*>(w_smem + (size_t)p.n_wstages * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + N);
  // barrier slots: a_full[2] a_empty[2] acc_full[2] acc_empty[2] w_full[8] w_empty[8]
  auto bar_a_full = [&](int i) { return smem_u32(&bars[0 + i]); };
  auto bar_a_empty = [&](int i) { return smem_u32(&bars[2 + i]); };
  auto bar_acc_full = [&](int i) { return smem_u32(&bars[4 + i]); };
  auto bar_acc_empty = [&](int i) { return smem_u32(&bars[6 + i]); };
  auto bar_w_full = [&](int s) { return smem_u32(&bars[8 + s]); };
  auto bar_w_empty = [&](int s) { return smem_u32(&bars[8 + kMaxStages + s]); };
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 8 + 2 * kMaxStages);

  const int n_taps = p.n_taps[phase];
  const int k16_per_tap = p.cin >> 4;
  const int n_k16 = n_taps * k16_per_tap;
  const int n_iters = (n_k16 + p.k16_per_stage - 1) / p.k16_per_stage;   // weight stages per tile
  const uint8_t* w_tile = static_cast<const uint8_t*>(p.w) + (size_t)(phase * p.n_tiles + ntile) * p.w_tile_bytes;

  constexpr int kWarpA = kEpiWarps, kWarpW = kEpiWarps + 1, kWarpMma = kEpiWarps + 2;
  if (warp == kWarpA && lane == 0) {
    prefetch_tmap(&p.tmap);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_a_full(i), 1);
      mbar_init(bar_a_empty(i), 1);
      mbar_init(bar_acc_full(i), 1);
      mbar_init(bar_acc_empty(i), kEpiWarps);
    }
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(bar_w_full(s), 1); mbar_init(bar_w_empty(s), p.cluster ? 2 : 1); }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  for (int i = threadIdx.x; i < N; i += kThreads) bias_s[i] = p.bias[ntile * N + i];
  tc_fence_before();
  __syncthreads();
  if (p.cluster) cluster_sync_all();              // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // In cluster mode both CTAs of a pair must run the same number of tiles (they share the weight ring's pace):
  // every CTA runs n_rounds rounds; a round past the end recomputes the last tile without storing it.
  const int n_rounds = (p.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int round) { return min((int)blockIdx.x + round * (int)gridDim.x, p.total_tiles - 1); };
  auto is_dummy = [&](int round) { return (int)blockIdx.x + round * (int)gridDim.x >= p.total_tiles; };
  const int my_rounds = p.cluster ? n_rounds : (((int)blockIdx.x < p.total_tiles) ? (p.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0);

  if (warp == kWarpA) {
    // ===== A producer: one TMA tile per work item =====
    {
      const bool leader = elect_one();
      int it = 0;
      for (; it < my_rounds; ++it) {
        const int tile = tile_of(it);
        const int buf = (p.n_abuf == 2) ? (it & 1) : 0, use = (p.n_abuf == 2) ? (it >> 1) : it;
        if (use > 0 && !mbar_wait_relaxed(bar_a_empty(buf), (use - 1) & 1, p.error_flag)) break;
        const int b = tile / p.m_tiles, m0 = (tile - b * p.m_tiles) * (128 * MSUB);
        const int row0 = m0 + p.row_lo[phase];
        const uint32_t dst = smem_u32(a_smem) + (uint32_t)buf * a_bytes;
        if (leader) {
          mbar_arrive_expect_tx(bar_a_full(buf), a_bytes);
          for (int c = 0; c < panels; ++c)
            for (int s = 0; s < p.nseg; ++s)
              tma_load_4d(dst + (uint32_t)c * panel_bytes + (uint32_t)(s * p.box_rows) * row_bytes, &p.tmap,
                          bar_a_full(buf), 0, row0 + s * p.box_rows, c, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == kWarpW) {
    // ===== W producer: weight stages through the ring (once, when resident) =====
    {
      const bool leader = elect_one();
      int slot = 0;
      uint32_t par = 1;                                           // parity of the previous use of `slot`
      bool wrapped = false;
      const uint32_t crank = p.cluster ? cluster_ctarank() : 0u;
      for (int round = 0; round < my_rounds; ++round) {
        if (p.w_resident && round != 0) break;
        bool ok = true;
        for (int i = 0; i < n_iters && ok; ++i) {
          if (wrapped) ok = mbar_wait_relaxed(bar_w_empty(slot), par, p.error_flag);
          if (!ok) break;
          const int k16 = min(p.k16_per_stage, n_k16 - i * p.k16_per_stage);
          const uint32_t bytes = (uint32_t)k16 * N * 32u;
          if (leader) {
            mbar_arrive_expect_tx(bar_w_full(slot), bytes);
            if (p.cluster) {                                      // this CTA fetches its half and multicasts it to the pair
              const uint32_t half = bytes >> 1;
              bulk_load_mc(smem_u32(w_smem) + (uint32_t)slot * stage_bytes + crank * half,
                           w_tile + (size_t)i * stage_bytes + crank * half, half, bar_w_full(slot), (uint16_t)0x3);
            } else {
              bulk_load(smem_u32(w_smem) + (uint32_t)slot * stage_bytes, w_tile + (size_t)i * stage_bytes, bytes,
                        bar_w_full(slot));
            }
          }
          __syncwarp();
          if (++slot == p.n_wstages) { slot = 0; par ^= 1u; wrapped = true; }
        }
        if (!ok) break;
      }
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    // The loop body is kept as lean as the microbenchmark in tools/mma_bench2.cu (which reaches
    // the hardware rate: 128 / 64 / 48 / 40 / 39 cycles per MMA for N = 256 .. 16): all
    // bookkeeping happens per (tap, panel) weight block, the SPP x MSUB MMAs of a block are
    // fully unrolled and differ only by constant adds on the descriptor start field.
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc(N, (p.flags & EPI_BF16) != 0);
      const int row_lo = p.row_lo[phase];
      constexpr uint32_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | ((row_bytes == 128 ? 2u : row_bytes == 64 ? 4u : 6u) << 29);
      constexpr uint32_t row16 = row_bytes >> 4;                  // one row, in 16-byte units
      constexpr uint32_t b_block16 = ((uint32_t)N * row_bytes) >> 4;
      const uint32_t a_panel16 = panel_bytes >> 4;
      const uint32_t b_lo0 = desc_lo(smem_u32(w_smem));
      const uint32_t tap_step16 = (uint32_t)p.tap_step * row16;
      const int blocks_per_stage = p.k16_per_stage / SPP;
      const int n_blocks_total = n_taps * panels;
      int it = 0, wslot = 0;
      uint32_t wpar = 0;
      bool ok = true, resident_ready = false;
      const bool timing = p.timing != nullptr;
      long long t_a = 0, t_w = 0, t_acc = 0, t_begin = timing ? clock64() : 0;
      for (; it < my_rounds && ok; ++it) {
        const int buf = (p.n_abuf == 2) ? (it & 1) : 0, use = (p.n_abuf == 2) ? (it >> 1) : it;
        const int acc = (kNumAcc == 2) ? (it & 1) : 0, acc_use = (kNumAcc == 2) ? (it >> 1) : it;
        long long tq = timing ? clock64() : 0;
        if (acc_use > 0) ok = mbar_wait(bar_acc_empty(acc), (acc_use - 1) & 1, p.error_flag);
        if (timing) { const long long t1 = clock64(); t_acc += t1 - tq; tq = t1; }
        if (ok) ok = mbar_wait(bar_a_full(buf), use & 1, p.error_flag);
        if (timing) t_a += clock64() - tq;
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)(acc * kAccCols);
        uint32_t a_tap = desc_lo(smem_u32(a_smem) + (uint32_t)buf * a_bytes) + (uint32_t)(p.tap_base[phase] - row_lo) * row16;
        uint32_t a_blk = a_tap, accum = 0;
        int panel = 0, blk = 0;
        for (int i = 0; i < n_iters; ++i) {
          int slot = i;
          const long long tw0 = timing ? clock64() : 0;
          if (!p.w_resident) {
            slot = wslot;
            ok = mbar_wait(bar_w_full(slot), wpar, p.error_flag);
            if (++wslot == p.n_wstages) { wslot = 0; wpar ^= 1u; }
          } else if (!resident_ready) {
            ok = mbar_wait(bar_w_full(slot), 0, p.error_flag);
          }
          if (timing) t_w += clock64() - tw0;
          if (!ok) break;
          tc_fence_after();
          uint32_t b_blk = b_lo0 + (uint32_t)slot * (stage_bytes >> 4);
          const int nb = min(blocks_per_stage, n_blocks_total - blk);
          for (int bi = 0; bi < nb; ++bi, ++blk) {
#pragma unroll
            for (int kk = 0; kk < SPP; ++kk) {
              const uint64_t bdesc = desc64(b_blk + 2u * kk, hi);
#pragma unroll
              for (int ms = 0; ms < MSUB; ++ms) {
                const uint64_t adesc = desc64(a_blk + 2u * kk + (uint32_t)ms * 128u * row16, hi);
                if (leader) umma_f16(d_base + (uint32_t)ms * N, adesc, bdesc, idesc, accum);
              }
              accum = 1;
            }
            b_blk += b_block16;
            if (++panel == panels) { panel = 0; a_tap += tap_step16; a_blk = a_tap; } else { a_blk += a_panel16; }
          }
          if (!p.w_resident && leader) {                          // slot free once these MMAs have read it
            if (p.cluster) umma_commit_mc(bar_w_empty(slot), (uint16_t)0x3);   // ... in both CTAs' rings
            else umma_commit(bar_w_empty(slot));
          }
          __syncwarp();
        }
        if (!ok) break;
        resident_ready = true;
        if (leader) {
          umma_commit(bar_a_empty(buf));                          // A tile consumed
          umma_commit(bar_acc_full(acc));                         // accumulators complete
        }
        __syncwarp();
      }
      if (timing && lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 0), (unsigned long long)(clock64() - t_begin));
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 1), (unsigned long long)t_a);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 2), (unsigned long long)t_w);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 3), (unsigned long long)t_acc);
      }
    }
  } else {
    // ===== epilogue: warps 0..7; TMEM lane group = warp % 4; column half = warp / 4 =====
    const int lg = warp & 3;
    const int half = warp >> 2;
    constexpr int kColsPerWarp = (N >= 32) ? N / 2 : N;            // N = 16: only half 0 has columns
    const bool has_cols = (N >= 32) || half == 0;
    const int col0 = (N >= 32) ? half * kColsPerWarp : 0;
    const bool bf16 = (p.flags & EPI_BF16) != 0;
    const int cchunks_total = p.cout_total >> 3;
    const int opc = p.out_pw >> 3;                                 // 8-channel chunks per output panel row
    const int opanels = p.cout_total / p.out_pw;
    const uint32_t flags = p.flags;
    int it = 0;
    const bool timing = p.timing != nullptr && warp == 0;
    long long t_full = 0, t_begin = timing ? clock64() : 0;
    for (; it < my_rounds; ++it) {
      const int tile = tile_of(it);
      const bool dummy = is_dummy(it);
      const int acc = (kNumAcc == 2) ? (it & 1) : 0, acc_use = (kNumAcc == 2) ? (it >> 1) : it;
      const long long tf0 = timing ? clock64() : 0;
      if (!mbar_wait_relaxed(bar_acc_full(acc), acc_use & 1, p.error_flag)) break;
      if (timing) t_full += clock64() - tf0;
      tc_fence_after();
      const int b = tile / p.m_tiles, m0 = (tile - b * p.m_tiles) * (128 * MSUB);
      if (has_cols) {
#pragma unroll
        for (int ms = 0; ms < MSUB; ++ms) {
          const int t = m0 + ms * 128 + lg * 32 + lane;            // output row on the M axis
          const bool valid = t < p.m_rows && !dummy;
          const size_t orow = (size_t)t * p.out_stride + phase;
          const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(acc * kAccCols + ms * N + col0);
          // The residual of up to 64 columns (4 groups) is requested before the first TMEM round trip, so a
          // tile costs kColsPerWarp/64 dependent HBM latencies per sub-tile instead of kColsPerWarp/16.
          constexpr int kGroups = kColsPerWarp / 16;
          constexpr int kBatch = kGroups < 4 ? kGroups : 4;
#pragma unroll 1
          for (int g0 = 0; g0 < kGroups; g0 += kBatch) {
            float4 qr[kBatch][4];
            if (valid && (flags & EPI_RES)) {
#pragma unroll
              for (int gi = 0; gi < kBatch; ++gi) {
                const int c8 = (col0 >> 3) + (g0 + gi) * 2;
                const size_t i0 = (((size_t)b * cchunks_total + (size_t)ntile * (N / 8) + c8) * (size_t)p.l_out + orow) * 8;
                const size_t i1 = i0 + (size_t)p.l_out * 8;
                qr[gi][0] = ldg_f4(p.res32 + i0); qr[gi][1] = ldg_f4(p.res32 + i0 + 4);
                qr[gi][2] = ldg_f4(p.res32 + i1); qr[gi][3] = ldg_f4(p.res32 + i1 + 4);
              }
            }
#pragma unroll
            for (int gi = 0; gi < kBatch; ++gi) {
              const int g = g0 + gi;
              uint32_t r[16];
              __syncwarp();                                        // tcgen05.ld is .sync.aligned
              tmem_ld16(t_addr + (uint32_t)(g * 16), r);
              const int c8 = (col0 >> 3) + g * 2;                  // first channel chunk inside this n-tile
              const size_t idx0 = (((size_t)b * cchunks_total + (size_t)ntile * (N / 8) + c8) * (size_t)p.l_out + orow) * 8;
              const size_t idx1 = idx0 + (size_t)p.l_out * 8;
              float4 qs[4];
              if (valid && (flags & (EPI_SUM_ADD | EPI_SUM_FIN))) {
                qs[0] = ldg_f4(p.sum32 + idx0); qs[1] = ldg_f4(p.sum32 + idx0 + 4);
                qs[2] = ldg_f4(p.sum32 + idx1); qs[3] = ldg_f4(p.sum32 + idx1 + 4);
              }
              tmem_ld_wait();
              if (valid) {
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) + bias_s[col0 + g * 16 + e];
                if (flags & EPI_RES) {
#pragma unroll
                  for (int h = 0; h < 4; ++h) { v[4 * h] += qr[gi][h].x; v[4 * h + 1] += qr[gi][h].y; v[4 * h + 2] += qr[gi][h].z; v[4 * h + 3] += qr[gi][h].w; }
                }
                if (flags & (EPI_SUM_ADD | EPI_SUM_FIN)) {
#pragma unroll
                  for (int h = 0; h < 4; ++h) {
                    v[4 * h] += qs[h].x; v[4 * h + 1] += qs[h].y; v[4 * h + 2] += qs[h].z; v[4 * h + 3] += qs[h].w;
                  }
                }
                if (flags & EPI_SUM_FIN) {
#pragma unroll
                  for (int e = 0; e < 16; ++e) v[e] = v[e] / p.n_blocks;            // xs / num_kernels
                }
                if (flags & (EPI_SUM_SET | EPI_SUM_ADD)) {
                  stg_f4(p.sum32 + idx0, v[0], v[1], v[2], v[3]); stg_f4(p.sum32 + idx0 + 4, v[4], v[5], v[6], v[7]);
                  stg_f4(p.sum32 + idx1, v[8], v[9], v[10], v[11]); stg_f4(p.sum32 + idx1 + 4, v[12], v[13], v[14], v[15]);
                }
                if (flags & EPI_OUT32) {
                  stg_f4(p.out32 + idx0, v[0], v[1], v[2], v[3]); stg_f4(p.out32 + idx0 + 4, v[4], v[5], v[6], v[7]);
                  stg_f4(p.out32 + idx1, v[8], v[9], v[10], v[11]); stg_f4(p.out32 + idx1 + 4, v[12], v[13], v[14], v[15]);
                }
                if (flags & EPI_OUT16) {
                  float lo[8], hi[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) { lo[e] = v[e]; hi[e] = v[8 + e]; }
                  const int cg = ntile * (N / 8) + c8;               // chunk index in the output tensor (even)
                  const size_t o16 = ((((size_t)b * opanels + cg / opc) * (size_t)p.l_out + orow) * opc + cg % opc) * 16;
                  *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out16) + o16) = pack8_lrelu(lo, p.slope_out, true, bf16);
                  *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out16) + o16 + 16) = pack8_lrelu(hi, p.slope_out, true, bf16);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty(acc));              // this warp is done with the accumulator buffer
    }
    if (timing && lane == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 4), (unsigned long long)(clock64() - t_begin));
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 5), (unsigned long long)t_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster) cluster_sync_all();              // no CTA leaves while its peer may still multicast into it
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tc
}  // namespace sa
