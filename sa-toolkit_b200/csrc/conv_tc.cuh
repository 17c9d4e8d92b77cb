// tcgen05 implicit-GEMM 1-D convolution for sm_100a: the contraction kernel of the tensor-core
// path (every conv of the generator except the 16->1 tail runs through it).
//
// Formulation ("time on M"):   D[t, co] = sum_j sum_ci  A[t + off_j, ci] * W_j[co, ci]
//   M = 128 output time steps per MMA (MSUB sub-tiles per CTA), N = Cout tile, K = Cin per tap.
//   A  activations, 16-bit, channel-blocked  [B][C/8][L][8]   (one time step of 8 channels = 16 B)
//   W  weights, 16-bit, pre-packed on the host in MMA order [k16 step][2][N][8]
//   D  fp32 accumulators in TMEM (MSUB * N columns)
//
// Why this layout: in the no-swizzle K-major canonical layout of the UMMA shared-memory
// descriptor a core matrix is 8 rows x 16 bytes with rows 16 B apart, 8-row groups SBO apart
// and the two K-halves LBO apart.  Staging the input tile as [C/8][rows][8] makes SBO = 128 B
// (rows are contiguous), so the operand of filter tap j is the SAME staged tile with the
// descriptor start address advanced by off_j * 16 B: dilated taps cost no data movement and
// no im2col; the +-(k-1)d/2 halo lives in shared memory.  TMA (cp.async.bulk.tensor) loads the
// tile and zero-fills rows outside [0, L), which is exactly the per-layer zero padding of
// the reference convs (nn.py:98-166).  Weights stream through an mbarrier ring with 1-D bulk
// copies.  The polyphase transposed convs (archi.py:47-59) are the same kernel: phase phi is
// a conv with taps at rows (off_phi - m) whose output row is u*q + phi.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias / residual / MRF / leaky-ReLU -> global).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace sa {
namespace tc {

constexpr int kThreads = 192;
constexpr int kMaxPhases = 8;
constexpr int kStages = 4;              // weight ring depth

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
  CUtensorMap tmap;         // activations [B][Cin/8][L_in][8], box {8, box_rows, box_chunks, 1}
  const void* w;            // packed weights: phase p / n-tile t at w + (p * n_tiles + t) * w_tile_bytes
  const float* bias;        // [Cout_total]
  const float* res32;       // fp32 blocked [B][Cout_total/8][L_out][8]
  float* out32;
  float* sum32;
  void* out16;              // 16-bit blocked
  int* error_flag;          // set when a barrier wait times out
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
  int rows_alloc;           // staged rows per chunk (nseg * box_rows)
  int box_rows, nseg, box_chunks;
  int k16_per_stage;        // K=16 steps per weight-ring stage
  uint32_t w_tile_bytes;
  uint32_t flags;
  float slope_out;
  float inv_blocks;         // 1 / n_resblocks is NOT used (true division below); kept = n_blocks
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
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {          // ~2 s at 2 GHz
      if (error_flag) atomicExch(error_flag, 1);
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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

// 32 lanes x 8 columns of 32-bit: thread i of the warp gets columns [c, c+8) of TMEM lane base+i.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}

// UMMA shared-memory descriptor, K-major, no swizzle (layout_type 0), version 1 (sm_100).
//   bits [0,14) start >> 4 | [16,30) LBO >> 4 (between the two 16-byte K halves)
//   | [32,46) SBO >> 4 (between 8-row groups) | [46,48) version = 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
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

// ---------------------------------------------------------------------------------------
// The kernel.  grid = (m_tiles, n_phases * n_tiles, B).  Dynamic smem:
//   [A tile: (cin/8) * rows_alloc * 16][W ring: kStages * k16_per_stage * N * 32][bias N*4][barriers]
// ---------------------------------------------------------------------------------------
template <int N, int MSUB>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int phase = blockIdx.y / p.n_tiles, ntile = blockIdx.y % p.n_tiles;
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * (128 * MSUB);

  const int chunks = p.cin >> 3;
  const uint32_t chunk_stride = (uint32_t)p.rows_alloc * 16u;
  const uint32_t a_bytes = (uint32_t)chunks * chunk_stride;
  const uint32_t stage_bytes = (uint32_t)p.k16_per_stage * N * 32u;
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + a_bytes;
  float* bias_s = reinterpret_cast<float*>(w_smem + kStages * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + N);       // 8-byte aligned: all sizes are multiples of 16
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 + 2 * kStages);
  const uint32_t bar_a_full = smem_u32(&bars[0]);
  const uint32_t bar_acc_full = smem_u32(&bars[1]);
  auto bar_w_full = [&](int s) { return smem_u32(&bars[2 + s]); };
  auto bar_w_empty = [&](int s) { return smem_u32(&bars[2 + kStages + s]); };

  constexpr uint32_t kTmemCols = (N * MSUB <= 32) ? 32 : (N * MSUB <= 64) ? 64 : (N * MSUB <= 128) ? 128
                                 : (N * MSUB <= 256) ? 256 : 512;
  const int n_taps = p.n_taps[phase];
  const int k16_per_tap = p.cin >> 4;
  const int n_k16 = n_taps * k16_per_tap;
  const int n_iters = (n_k16 + p.k16_per_stage - 1) / p.k16_per_stage;
  const uint8_t* w_tile = static_cast<const uint8_t*>(p.w) + (size_t)(phase * p.n_tiles + ntile) * p.w_tile_bytes;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap);
    mbar_init(bar_a_full, 1);
    mbar_init(bar_acc_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(bar_w_full(s), 1); mbar_init(bar_w_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  for (int i = threadIdx.x; i < N; i += kThreads) bias_s[i] = p.bias[ntile * N + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_a_full, a_bytes);
      const int row0 = m0 + p.row_lo[phase];
      for (int c = 0; c < chunks; c += p.box_chunks)
        for (int s = 0; s < p.nseg; ++s)
          tma_load_4d(smem_u32(a_smem) + (uint32_t)c * chunk_stride + (uint32_t)(s * p.box_rows) * 16u, &p.tmap,
                      bar_a_full, 0, row0 + s * p.box_rows, c, b);
      for (int it = 0; it < n_iters; ++it) {
        const int slot = it % kStages;
        if (it >= kStages && !mbar_wait(bar_w_empty(slot), ((it / kStages) - 1) & 1, p.error_flag)) break;
        const int k16 = min(p.k16_per_stage, n_k16 - it * p.k16_per_stage);
        const uint32_t bytes = (uint32_t)k16 * N * 32u;
        mbar_arrive_expect_tx(bar_w_full(slot), bytes);
        bulk_load(smem_u32(w_smem) + slot * stage_bytes, w_tile + (size_t)it * stage_bytes, bytes, bar_w_full(slot));
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(N, (p.flags & EPI_BF16) != 0);
      const int row_lo = p.row_lo[phase];
      bool ok = mbar_wait(bar_a_full, 0, p.error_flag);
      tc_fence_after();
      int step = 0;
      for (int it = 0; it < n_iters && ok; ++it) {
        const int slot = it % kStages;
        ok = mbar_wait(bar_w_full(slot), (it / kStages) & 1, p.error_flag);
        if (!ok) break;
        tc_fence_after();
        const int k16 = min(p.k16_per_stage, n_k16 - it * p.k16_per_stage);
        for (int kk = 0; kk < k16; ++kk, ++step) {
          const int tap = step / k16_per_tap, cb = step - tap * k16_per_tap;
          const int off = p.tap_base[phase] + tap * p.tap_step - row_lo;          // staged row of output row m0
          const uint32_t a_addr = smem_u32(a_smem) + (uint32_t)(2 * cb) * chunk_stride + (uint32_t)off * 16u;
          const uint64_t bdesc = make_smem_desc(smem_u32(w_smem) + slot * stage_bytes + (uint32_t)kk * N * 32u,
                                                (uint32_t)N * 16u, 128u);
#pragma unroll
          for (int ms = 0; ms < MSUB; ++ms) {
            const uint64_t adesc = make_smem_desc(a_addr + (uint32_t)ms * 128u * 16u, chunk_stride, 128u);
            umma_f16(tmem_base + (uint32_t)ms * N, adesc, bdesc, idesc, step > 0 ? 1u : 0u);
          }
        }
        umma_commit(bar_w_empty(slot));           // frees the ring slot once these MMAs have read it
      }
      umma_commit(bar_acc_full);                  // accumulators complete
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane group = warp % 4 =====
    const int lg = warp & 3;
    const bool ok = mbar_wait(bar_acc_full, 0, p.error_flag);
    tc_fence_after();
    const bool bf16 = (p.flags & EPI_BF16) != 0;
    const int cchunks_total = p.cout_total >> 3;
    if (ok) {
#pragma unroll
      for (int ms = 0; ms < MSUB; ++ms) {
        const int t = m0 + ms * 128 + lg * 32 + lane;            // output row on the M axis
        const bool valid = t < p.m_rows;
        const long long orow = (long long)t * p.out_stride + phase;
#pragma unroll 1
        for (int c8 = 0; c8 < N / 8; ++c8) {
          float v[8];
          __syncwarp();                                          // tcgen05.ld is .sync.aligned
          tmem_ld8(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(ms * N + c8 * 8), v);   // warp-collective
          if (valid) {
          const size_t idx = (((size_t)b * cchunks_total + (size_t)ntile * (N / 8) + c8) * (size_t)p.l_out + (size_t)orow) * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] += bias_s[c8 * 8 + e];
          if (p.flags & EPI_RES) {
            const float4 r0 = *reinterpret_cast<const float4*>(p.res32 + idx);
            const float4 r1 = *reinterpret_cast<const float4*>(p.res32 + idx + 4);
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
          }
          if (p.flags & (EPI_SUM_ADD | EPI_SUM_FIN)) {
            const float4 s0 = *reinterpret_cast<const float4*>(p.sum32 + idx);
            const float4 s1 = *reinterpret_cast<const float4*>(p.sum32 + idx + 4);
            v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
            v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
          }
          if (p.flags & EPI_SUM_FIN) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = v[e] / p.inv_blocks;     // xs / num_kernels (true division)
          }
          if (p.flags & (EPI_SUM_SET | EPI_SUM_ADD)) {
            *reinterpret_cast<float4*>(p.sum32 + idx) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(p.sum32 + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (p.flags & EPI_OUT32) {
            *reinterpret_cast<float4*>(p.out32 + idx) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(p.out32 + idx + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          if (p.flags & EPI_OUT16) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = lrelu_f(v[e], p.slope_out);
            *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out16) + idx * 2) = pack8(v, bf16);
          }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tc
}  // namespace sa
