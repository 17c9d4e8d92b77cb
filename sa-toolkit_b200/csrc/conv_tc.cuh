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

// Ragged batches: when the caller gives the true length of every item (frames_per_item), the kernels enumerate only
// the tiles that can reach an item's kept samples: rows [0, (frames + margin) * rows_per_frame) of every layer, margin >
// receptive field of the whole generator (20 frames) so the kept region is bit-identical to the padded run.  Live
// tiles are compacted with a prefix sum over the items (built by warp 0 of every CTA, n_items <= kMaxMapItems), so
// the persistent CTAs stay balanced.
constexpr int kMaxMapItems = 256;
struct TileMapParams {
  const int* frames;        // device [n_items]; nullptr: every item owns all its tiles (padded semantics)
  int n_items;
  int rows_per_frame;       // rows of this kernel's tile axis per input frame
  int margin_frames;
};
// Called by all threads before a __syncthreads(); pre[i] = first live tile of item i, pre[n_items] = live tiles.
__device__ __forceinline__ void tilemap_build(int* pre, const TileMapParams& m, int rows_full, int tile_rows) {
  if (m.frames == nullptr || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  int running = 0;
  if (lane == 0) pre[0] = 0;
  for (int base = 0; base < m.n_items; base += 32) {
    const int b = base + lane;
    int v = 0;
    if (b < m.n_items) {
      const int rows = min(rows_full, (m.frames[b] + m.margin_frames) * m.rows_per_frame);
      v = (rows + tile_rows - 1) / tile_rows;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += u;
    }
    if (b < m.n_items) pre[b + 1] = running + v;
    running += __shfl_sync(0xffffffffu, v, 31);
  }
}
__device__ __forceinline__ int tilemap_total(const int* pre, const TileMapParams& m, int total_full) {
  return m.frames ? pre[m.n_items] : total_full;
}
__device__ __forceinline__ void tilemap_locate(const int* pre, const TileMapParams& m, int tiles_full, int tile, int& b, int& mt) {
  if (m.frames == nullptr) { b = tile / tiles_full; mt = tile - b * tiles_full; return; }
  int lo = 0, hi = m.n_items;                                // largest b with pre[b] <= tile
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (pre[mid] <= tile) lo = mid; else hi = mid;
  }
  b = lo; mt = tile - pre[lo];
}

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
  int m_tiles;              // time tiles per item
  int total_tiles;          // m_tiles * B
  TileMapParams map;        // ragged batches: live-tile enumeration (map.frames == nullptr: all tiles)
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
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes (or the hint, in ns, runs
// out), so a waiting warp neither polls shared memory nor takes issue slots, and wakes on the completion itself.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity, uint32_t ticks = 0u) {
  uint32_t ok;
  if (ticks)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ticks)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  return ok != 0;
}
// suspend-time hint (ns) of the relaxed waits, set once per process by tc_init (SATOOLS_B200_MBAR_NS; 0 = poll with
// nanosleep(40) between plain try_waits, the round-1 behaviour)
__device__ uint32_t g_mbar_suspend_ns = 20000;
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
  const uint32_t hint = g_mbar_suspend_ns;
  while (!mbar_try_wait(bar, parity, hint)) {
    if (hint == 0) __nanosleep(40);
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
// 256-bit global accesses (sm_100: LDG/STG.256): one 8-channel fp32 row (32 bytes, 32-byte aligned) per instruction.  The
// rows of a warp are consecutive in the blocked fp32 tensors, so a warp instruction covers 1 KB = 8 full lines, where two
// 128-bit instructions each touched the same 8 lines with half sectors.
__device__ __forceinline__ void ldg_f8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg_f8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void stg_u8(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// TMEM -> registers, 32 lanes x 32 columns in one instruction.  Every tcgen05.ld + wait::ld round trip costs the epilogue
// ~230 cycles whatever its width (B300_MICROARCH.md: 12 + MEMBAR ~113 + 2 BAR ~34 + ~35), so a 64-column row is read as
// two x32 loads -- the second in flight while the first half is converted -- instead of four x16 loads.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }


// 4 x 4 transpose of 16-byte items inside every aligned group of four lanes: on entry lane j (= lane & 3) holds
// a[k] = item (row j, piece k); on exit a[k] = item (row k, piece j).  Two butterfly rounds, 16 SHFL.
__device__ __forceinline__ uint4 shfl_xor_u4(uint4 x, int m) {
  x.x = __shfl_xor_sync(0xffffffffu, x.x, m); x.y = __shfl_xor_sync(0xffffffffu, x.y, m);
  x.z = __shfl_xor_sync(0xffffffffu, x.z, m); x.w = __shfl_xor_sync(0xffffffffu, x.w, m);
  return x;
}
__device__ __forceinline__ void quad_transpose(uint4 (&a)[4], int j) {
  const bool b0 = (j & 1) != 0, b1 = (j & 2) != 0;
  uint4 r;
  r = shfl_xor_u4(b0 ? a[0] : a[1], 1); if (b0) a[0] = r; else a[1] = r;
  r = shfl_xor_u4(b0 ? a[2] : a[3], 1); if (b0) a[2] = r; else a[3] = r;
  r = shfl_xor_u4(b1 ? a[0] : a[2], 2); if (b1) a[0] = r; else a[2] = r;
  r = shfl_xor_u4(b1 ? a[1] : a[3], 2); if (b1) a[1] = r; else a[3] = r;
}

// ---------------------------------------------------------------------------------------
// Epilogue of one accumulator tile, shared by conv_tc_kernel and conv_pair_kernel.  The calling warp owns TMEM lanes
// [32 lg, 32 lg + 32) = 32 output rows of every 128-row sub-tile, and COLS columns starting at col0 of the n-tile:
// U = MSUB * COLS / 16 units of 16 columns.  Per unit: TMEM -> registers, + bias, + residual (fp32 stream),
// multi-receptive-field sum, then the fp32 and / or 16-bit (leaky-ReLU'd) output streams.
//   * All addresses are one 64-bit base per tile plus small strides (an earlier version recomputed the blocked-tensor
//     indices per unit: ~200 issued instructions per unit, the epilogue was issue bound -- profiles/README.md).
//   * The residual of unit u+1 is requested while unit u is processed; unit 0's before the accumulator is waited for.
//   * 16-bit stores: thread-per-row stores put every lane of a warp store into a different 128-byte line.  When a
//     thread ends up with 64 contiguous output bytes (two units), they are transposed inside lane quads instead, so
//     one instruction writes 64 contiguous bytes of 8 rows.
// wait_acc() is called after the first residual request and must return false on a barrier timeout.
// ---------------------------------------------------------------------------------------
template <int N, int MSUB, int COLS, typename WaitAcc>
__device__ __forceinline__ bool epi_tile(const ConvParams& p, const float* bias_s, uint32_t tmem_acc, int b, int m0, int phase,
                                         int ntile, int lg, int lane, int col0, bool dummy, WaitAcc&& wait_acc, int j_lo = 0,
                                         int j_hi = 0x7fffffff) {
  // j_lo / j_hi: only tile rows [j_lo, j_hi) are stored (resblock_pair_tc.cuh: the edge rows of a fused tile are not valid)
  constexpr int kGroups = COLS / 16;
  constexpr int U = MSUB * kGroups;
  constexpr bool kQuadStores = (kGroups % 2 == 0);
  const uint32_t flags = p.flags;
  const bool bf16 = (flags & EPI_BF16) != 0;
  const bool has_res = (flags & EPI_RES) != 0;
  const bool has_sum_in = (flags & (EPI_SUM_ADD | EPI_SUM_FIN)) != 0;
  const int m_rows = p.m_rows, out_stride = p.out_stride;
  const size_t l_out = (size_t)p.l_out;
  const size_t plane = l_out * 8;                            // floats between consecutive 8-channel chunks (fp32 tensors)
  const int chunk0 = ntile * (N / 8) + (col0 >> 3);          // first 8-channel chunk of this warp in the output tensor
  const int cchunks_total = p.cout_total >> 3;
  const int t0 = m0 + lg * 32 + lane;
  // fp32 element offset of (item b, chunk0, output row of sub-tile 0); sub-tile ms adds 128 * out_stride rows
  const size_t base32 = (((size_t)b * cchunks_total + chunk0) * l_out + (size_t)t0 * out_stride + phase) * 8;
  const size_t ms_step32 = (size_t)128 * out_stride * 8;
  // 16-bit output [B][Cout / out_pw][L][out_pw]: rows of row16 bytes, opc chunks of 16 bytes per row
  const int opc_shift = p.out_pw >= 64 ? 3 : p.out_pw >= 32 ? 2 : 1;
  const size_t row16 = (size_t)p.out_pw * 2;
  const size_t panel16 = l_out * row16;
  const size_t item16 = (size_t)(cchunks_total >> opc_shift) * panel16;
  uint8_t* const out16 = static_cast<uint8_t*>(p.out16);
  // narrow tiles (one unit per sub-tile) run two CTAs per SM on a tight register budget: no load-ahead there
  constexpr int QB = kQuadStores ? 2 : 1;
  float4 qr[QB][4];
  uint4 pk[kQuadStores ? 4 : 1];
  auto fetch_res = [&](int u, float4 (&q)[4]) {
    const int ms = u / kGroups, g = u % kGroups;
    const int j = ms * 128 + lg * 32 + lane;
    if (!has_res || dummy || t0 + ms * 128 >= m_rows || j < j_lo || j >= j_hi) return;
    const float* a0 = p.res32 + base32 + ms * ms_step32 + (size_t)(2 * g) * plane;
    const float* a1 = a0 + plane;
    ldg_f8(a0, q[0], q[1]);
    ldg_f8(a1, q[2], q[3]);
  };
  fetch_res(0, qr[0]);                                       // does not depend on the accumulator
  if (!wait_acc()) return false;
  tc_fence_after();
  if (dummy) return true;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int ms = u / kGroups, g = u % kGroups;
    const int t = t0 + ms * 128;
    const int jrow = ms * 128 + lg * 32 + lane;
    const bool valid = t < m_rows && jrow >= j_lo && jrow < j_hi;
    uint32_t r[16];
    __syncwarp();                                            // tcgen05.ld is .sync.aligned
    tmem_ld16(tmem_acc + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ms * N + col0 + g * 16), r);
    const size_t i0 = base32 + ms * ms_step32 + (size_t)(2 * g) * plane, i1 = i0 + plane;
    float4 qs[4];
    if (QB == 1 && u > 0) fetch_res(u, qr[0]);
    if (valid && has_sum_in) {
      ldg_f8(p.sum32 + i0, qs[0], qs[1]);
      ldg_f8(p.sum32 + i1, qs[2], qs[3]);
    }
    tmem_ld_wait();
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]) + bias_s[col0 + g * 16 + e];
    if (QB == 2 && u + 1 < U) fetch_res(u + 1, qr[(u + 1) % QB]);
    if (valid) {
      if (has_res) {
#pragma unroll
        for (int h = 0; h < 4; ++h) { v[4 * h] += qr[u % QB][h].x; v[4 * h + 1] += qr[u % QB][h].y; v[4 * h + 2] += qr[u % QB][h].z; v[4 * h + 3] += qr[u % QB][h].w; }
      }
      if (has_sum_in) {
#pragma unroll
        for (int h = 0; h < 4; ++h) { v[4 * h] += qs[h].x; v[4 * h + 1] += qs[h].y; v[4 * h + 2] += qs[h].z; v[4 * h + 3] += qs[h].w; }
      }
      if (flags & EPI_SUM_FIN) {
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = v[e] / p.n_blocks;              // xs / num_kernels
      }
      if (flags & (EPI_SUM_SET | EPI_SUM_ADD)) {
        stg_f8(p.sum32 + i0, v);
        stg_f8(p.sum32 + i1, v + 8);
      }
      if (flags & EPI_OUT32) {
        stg_f8(p.out32 + i0, v);
        stg_f8(p.out32 + i1, v + 8);
      }
    }
    if (flags & EPI_OUT16) {
      float lo[8], hi8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { lo[e] = v[e]; hi8[e] = v[8 + e]; }
      if constexpr (kQuadStores) {
        pk[(u & 1) * 2] = pack8_lrelu(lo, p.slope_out, true, bf16);
        pk[(u & 1) * 2 + 1] = pack8_lrelu(hi8, p.slope_out, true, bf16);
        if (u & 1) {                                         // N >= 64 here, so rows are 128 bytes (out_pw = 64)
          const int j4 = lane & 3;
          quad_transpose(pk, j4);
          const int cg = chunk0 + (g - 1) * 2 + j4;          // this lane's 8-channel chunk of the output row
          const int tq = t - j4;                             // row of the quad's first lane
          uint8_t* o = out16 + (size_t)b * item16 + (size_t)(cg >> 3) * panel16 + ((size_t)tq * out_stride + phase) * 128 +
                       (size_t)(cg & 7) * 16;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (tq + k < m_rows && jrow - j4 + k >= j_lo && jrow - j4 + k < j_hi)
              *reinterpret_cast<uint4*>(o + (size_t)k * out_stride * 128) = pk[k];
        }
      } else if (valid) {
        const int cg = chunk0 + g * 2;                       // even chunk; cg and cg + 1 share a panel row
        uint8_t* o = out16 + (size_t)b * item16 + (size_t)(cg >> opc_shift) * panel16 +
                     ((size_t)t * out_stride + phase) * row16 + (size_t)(cg & ((1 << opc_shift) - 1)) * 16;
        stg_u8(o, pack8_lrelu(lo, p.slope_out, true, bf16), pack8_lrelu(hi8, p.slope_out, true, bf16));
      }
    }
  }
  return true;
}

// ---------------------------------------------------------------------------------------
// The kernel.  grid = (persistent CTAs, n_phases * n_tiles).  Dynamic smem:
//   [A ring: n_abuf * (cin/pw) * rows_alloc * 2pw][W ring: n_wstages * k16_per_stage * N * 32]
//   [bias N*4][barriers][tmem holder]      (base rounded up to 1024 B: swizzle pattern anchor)
// ---------------------------------------------------------------------------------------
template <int N, int MSUB, int PW>
__global__ void __launch_bounds__(kThreads, (N < 64) ? 2 : 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops (barrier addresses,
  // descriptors, counters) in uniform registers instead of converting them per use (R2UR), as CUTLASS' canonical_warp_idx_sync
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
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

  __shared__ int tile_pre[kMaxMapItems + 1];
  tilemap_build(tile_pre, p.map, p.m_rows, 128 * MSUB);          // visible after the __syncthreads() below
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
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(bar_w_full(s), 1); mbar_init(bar_w_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  for (int i = threadIdx.x; i < N; i += kThreads) bias_s[i] = p.bias[ntile * N + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // persistent CTAs: tile = blockIdx.x + round * gridDim.x
  auto tile_of = [&](int round) { return (int)blockIdx.x + round * (int)gridDim.x; };
  const int n_live = tilemap_total(tile_pre, p.map, p.total_tiles);
  const int my_rounds = ((int)blockIdx.x < n_live) ? (n_live - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == kWarpA) {
    // ===== A producer: one TMA tile per work item =====
    {
      const bool leader = elect_one();
      int it = 0;
      for (; it < my_rounds; ++it) {
        const int tile = tile_of(it);
        const int buf = (p.n_abuf == 2) ? (it & 1) : 0, use = (p.n_abuf == 2) ? (it >> 1) : it;
        if (use > 0 && !mbar_wait_relaxed(bar_a_empty(buf), (use - 1) & 1, p.error_flag)) break;
        int b, mt;
        tilemap_locate(tile_pre, p.map, p.m_tiles, tile, b, mt);
        const int m0 = mt * (128 * MSUB);
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
            bulk_load(smem_u32(w_smem) + (uint32_t)slot * stage_bytes, w_tile + (size_t)i * stage_bytes, bytes,
                      bar_w_full(slot));
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
#ifdef SA_DIAG
      const bool timing = p.timing != nullptr;
#else
      constexpr bool timing = false;   // in-kernel cycle counters: -DSA_DIAG builds only
#endif
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
          // (weights are TMA data: the wait above is the acquire; no tcgen05 fence needed here)
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
          if (!p.w_resident && leader) umma_commit(bar_w_empty(slot));   // slot free once these MMAs have read it
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
    // ===== epilogue: warps 0..7; TMEM lane group = warp % 4; column half = warp / 4 (epi_tile above) =====
    const int lg = warp & 3;
    const int half = warp >> 2;
    constexpr int kColsPerWarp = (N >= 32) ? N / 2 : N;            // N = 16: only half 0 has columns
    const bool has_cols = (N >= 32) || half == 0;
    const int col0 = (N >= 32) ? half * kColsPerWarp : 0;
#ifdef SA_DIAG
    const bool timing = p.timing != nullptr && warp == 0;
#else
    constexpr bool timing = false;   // in-kernel cycle counters: -DSA_DIAG builds only
#endif
    long long t_full = 0, t_begin = timing ? clock64() : 0;
    for (int it = 0; it < my_rounds; ++it) {
      const int tile = tile_of(it);
      const int acc = (kNumAcc == 2) ? (it & 1) : 0, acc_use = (kNumAcc == 2) ? (it >> 1) : it;
      int b, mt;
      tilemap_locate(tile_pre, p.map, p.m_tiles, tile, b, mt);
      const int m0 = mt * (128 * MSUB);
      auto wait_acc = [&]() {
        const long long tf0 = timing ? clock64() : 0;
        const bool ok = mbar_wait_relaxed(bar_acc_full(acc), acc_use & 1, p.error_flag);
        if (timing) t_full += clock64() - tf0;
        return ok;
      };
      if (!epi_tile<N, MSUB, kColsPerWarp>(p, bias_s, tmem_base + (uint32_t)(acc * kAccCols), b, m0, phase, ntile, lg, lane,
                                           col0, !has_cols, wait_acc))
        break;
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
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tc
}  // namespace sa
