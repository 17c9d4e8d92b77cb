// Whole multi-receptive-field stage (three ResBlock1 with k = K0, K1, K2) fused in one kernel, for the
// narrowest stages (C = 32, 16):   h_out = (RB_K0(x) + RB_K1(x) + RB_K2(x)) / 3      (archi.py:82-86)
//
// Why: in-kernel cycle accounting of the single-ResBlock kernel (chain_tc.cuh) showed its MMA warp idle
// 35-50% of the time, waiting on the one dependency chain conv -> epilogue -> conv, and every ResBlock
// re-reads x and round-trips the running sum through HBM.  The three ResBlocks of a stage are independent
// chains over the same input tile, so one CTA runs them interleaved: while the epilogue warps finish
// conv c of chain j, the tensor pipe works on chains j+1, j+2; x is read once, the three residual
// streams live in registers and the mean is formed there (no sum buffer).
//
// Layout and building blocks are those of chain_tc.cuh: staged activations [rows][C] in the UMMA K-major
// swizzled layout (two buffers per chain), time on M, taps as row-shifted descriptors, accumulators in TMEM
// (one per chain and sub-tile: the next conv of a chain only starts after that chain's epilogue has drained
// it), one epilogue warp quad per sub-tile (one thread = one row, for all three chains).  All chains use
// the largest halo (that of K2) so that they produce the same valid rows.  Synchronisation is per (chain,
// conv), not per sub-tile: with three chains in flight the MMA warp rarely has to wait, and every barrier
// check costs it ~90 cycles even when it is already complete.
#pragma once
#include "chain_tc.cuh"

namespace sa {
namespace tc {

struct Chain3Params {
  const float* x32;         // stage input h, fp32 blocked [B][C/8][L][8]
  float* out32;             // stage output (EPI_OUT32)
  void* out16;              // lrelu(stage output), 16-bit [B][1][L][C] (EPI_OUT16)
  const void* w[3];         // per chain: n_convs convs, each [tap][C rows][C] pre-swizzled
  const float* bias[3];     // per chain [n_convs][C]
  int* error_flag;
  int L;
  int n_convs;              // 2 * n_dilations (same for the three chains)
  int dil[3][kChainMaxConvs];
  int pad[3][kChainMaxConvs];
  int halo;                 // halo of the widest chain
  int tiles_per_item, total_tiles;
  TileMapParams map;        // ragged batches: live-tile enumeration (conv_tc.cuh); tile axis = valid rows per tile
  uint32_t flags;           // EPI_OUT32 / EPI_OUT16 / EPI_BF16
  float slope_out;
};

template <int C, int MS, int K0, int K1, int K2>
__global__ void __launch_bounds__(chain_threads(MS, 4), 1) stage_chain3_kernel(const __grid_constant__ Chain3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int N = C;
  constexpr int R = MS * 128;
  constexpr int ROWS = R + 2 * kChainPad;
  constexpr uint32_t RB = 2u * C;
  constexpr uint32_t kBufBytes = ROWS * RB;
  constexpr int kCPT = C / 8;
  constexpr int K16 = C / 16;
  constexpr uint32_t kTapBytes = (uint32_t)N * RB;
  constexpr int KS[3] = {K0, K1, K2};
  constexpr uint32_t kWOff[3] = {0u, (uint32_t)K0 * kTapBytes, (uint32_t)(K0 + K1) * kTapBytes};
  constexpr uint32_t kWBytes = (uint32_t)(K0 + K1 + K2) * kTapBytes;   // one resident conv per chain
  constexpr int kThreads_ = chain_threads(MS, 4);
  constexpr uint32_t kTmemNeed = 3u * MS * N;
  constexpr uint32_t kTmemCols = kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
  static_assert(kTmemNeed <= 512, "accumulators do not fit in TMEM");

  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops (barrier addresses,
  // descriptors, counters) in uniform registers instead of converting them per use (R2UR), as CUTLASS' canonical_warp_idx_sync
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int kWarpW = 4 * MS, kWarpMma = 4 * MS + 1;          // epilogue warps first, MMA issuer last (see chain_tc.cuh)
  // smem: buf[chain][A|T], weights (one conv per chain), bias, barriers
  auto buf = [&](int j, int t) { return smem + (uint32_t)(j * 2 + t) * kBufBytes; };
  uint8_t* w_smem = smem + 6 * kBufBytes;
  float* bias_s = reinterpret_cast<float*>(w_smem + kWBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 3 * kChainMaxConvs * C);
  // barriers: ready[3][2][8] acc_full[3][8] w_full[3] w_empty[3]
  auto bar_ready = [&](int j, int t, int s) { return smem_u32(&bars[(j * 2 + t) * 8 + s]); };
  auto bar_acc_full = [&](int j, int s) { return smem_u32(&bars[48 + j * 8 + s]); };
  auto bar_w_full = [&](int j) { return smem_u32(&bars[72 + j]); };
  auto bar_w_empty = [&](int j) { return smem_u32(&bars[75 + j]); };
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 78);

  const int valid_rows = R - 2 * p.halo;
  __shared__ int tile_pre[kMaxMapItems + 1];
  tilemap_build(tile_pre, p.map, p.L, valid_rows);                // visible after the __syncthreads() below
  const bool bf16 = (p.flags & EPI_BF16) != 0;
  const int n_pairs = p.n_convs / 2;

  if (warp == kWarpW && lane == 0) {
    for (int j = 0; j < 3; ++j) {
      for (int s = 0; s < 8; ++s) {
        mbar_init(bar_ready(j, 0, s), 4 * MS); mbar_init(bar_ready(j, 1, s), 4 * MS);   // only s = 0 is used
        mbar_init(bar_acc_full(j, s), 1);
      }
      mbar_init(bar_w_full(j), 1); mbar_init(bar_w_empty(j), 1);
    }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  for (int i = threadIdx.x; i < 3 * p.n_convs * C; i += kThreads_) {
    const int j = i / (p.n_convs * C), rem = i - j * p.n_convs * C;
    bias_s[j * kChainMaxConvs * C + rem] = p.bias[j][rem];
  }
  for (uint32_t i = threadIdx.x; i < 6 * kBufBytes / 16; i += kThreads_)
    *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);    // PAD slack rows stay zero
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int n_live = tilemap_total(tile_pre, p.map, p.total_tiles);

  if (warp == kWarpW) {
    // ===== weight producer: conv c of chain j into chain j's slot =====
    const bool leader = elect_one();
    uint32_t n = 0;                                              // how often each slot has been filled
    bool ok = true;
    for (int tile = blockIdx.x; tile < n_live && ok; tile += gridDim.x)
      for (int c = 0; c < p.n_convs && ok; ++c, ++n)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (n > 0) ok = ok && mbar_wait_relaxed(bar_w_empty(j), (n - 1) & 1u, p.error_flag);
          if (!ok) break;
          const uint32_t bytes = (uint32_t)KS[j] * kTapBytes;
          if (leader) {
            mbar_arrive_expect_tx(bar_w_full(j), bytes);
            bulk_load(smem_u32(w_smem) + kWOff[j], static_cast<const uint8_t*>(p.w[j]) + (size_t)c * bytes, bytes, bar_w_full(j));
          }
          __syncwarp();
        }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer: conv by conv, the three chains in turn =====
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(N, bf16);
    constexpr uint32_t hi = ((8u * RB) >> 4) | (1u << 14) | ((RB == 128 ? 2u : RB == 64 ? 4u : 6u) << 29);
    constexpr uint32_t row16 = RB >> 4;
    constexpr uint32_t tap16 = kTapBytes >> 4;
    uint32_t it = 0, n = 0;
    bool ok = true;
    auto conv = [&](auto k_c, auto j_c, int c) {
      constexpr int K = decltype(k_c)::value;
      constexpr int j = decltype(j_c)::value;
      const uint32_t in_lo0 = desc_lo(smem_u32(buf(j, c & 1))) + (uint32_t)(kChainPad - p.pad[j][c]) * row16;
      const uint32_t rdy_parity = (it * (uint32_t)n_pairs + (uint32_t)(c / 2)) & 1u;
      const uint32_t dil16 = (uint32_t)p.dil[j][c] * row16;
      const uint32_t b_lo0 = desc_lo(smem_u32(w_smem) + kWOff[j]);
      ok = ok && mbar_wait(bar_w_full(j), n & 1u, p.error_flag);
      ok = ok && mbar_wait(bar_ready(j, c & 1, 0), rdy_parity, p.error_flag);      // every sub-tile of this chain staged
      if (!ok) return;
      tc_fence_after();
#pragma unroll
      for (int s = 0; s < MS; ++s) {
        const uint32_t d_tmem = tmem_base + (uint32_t)((j * MS + s) * N);
        uint32_t a_tap = in_lo0 + (uint32_t)(s * 128) * row16;
        uint32_t b_tap = b_lo0;
#pragma unroll
        for (int tap = 0; tap < K; ++tap) {
#pragma unroll
          for (int kk = 0; kk < K16; ++kk)
            if (leader) umma_f16(d_tmem, desc64(a_tap + 2u * kk, hi), desc64(b_tap + 2u * kk, hi), idesc, (tap | kk) ? 1u : 0u);
          a_tap += dil16;
          b_tap += tap16;
        }
      }
      if (leader) {
        umma_commit(bar_acc_full(j, 0));                         // all sub-tiles of this chain's conv
        umma_commit(bar_w_empty(j));                             // and its weights are consumed
      }
      __syncwarp();
    };
    for (int tile = blockIdx.x; tile < n_live && ok; tile += gridDim.x, ++it)
      for (int c = 0; c < p.n_convs && ok; ++c, ++n) {
        conv(std::integral_constant<int, K0>{}, std::integral_constant<int, 0>{}, c);
        if (ok) conv(std::integral_constant<int, K1>{}, std::integral_constant<int, 1>{}, c);
        if (ok) conv(std::integral_constant<int, K2>{}, std::integral_constant<int, 2>{}, c);
      }
  } else {
    // ===== epilogue: warp quad (warp - 2) / 4 owns sub-tile s; this thread owns row r for all three chains =====
    const int lg = warp & 3;
    const int s = warp >> 2;
    const int r = s * 128 + lg * 32 + lane;
    constexpr int cchunks = C / 8;
    float xr[3][C];                                              // the three residual streams of this row
    const uint32_t row_off = (uint32_t)(kChainPad + r) * RB;
    uint32_t soff[kCPT];
#pragma unroll
    for (int q = 0; q < kCPT; ++q) soff[q] = swz(row_off + (uint32_t)q * 16u, RB);
    uint32_t it = 0, n = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < n_live && ok; tile += gridDim.x, ++it) {
      int b, mt;
      tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile, b, mt);
      const int t = mt * valid_rows - p.halo + r;
      const bool inside = t >= 0 && t < p.L;
      const bool keep = inside && r >= p.halo && r < R - p.halo;
      // ---- load x once; it seeds the three residual streams and the three staged lrelu(x) tiles ----
#pragma unroll
      for (int q = 0; q < kCPT; ++q) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c4 = a;
        if (inside) {
          const float* src = p.x32 + (((size_t)b * cchunks + q) * (size_t)p.L + t) * 8;
          ldg_f8(src, a, c4);
        }
        float v[8] = {a.x, a.y, a.z, a.w, c4.x, c4.y, c4.z, c4.w};
        const uint4 pk = pack8_lrelu(v, 0.1f, true, bf16);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int e = 0; e < 8; ++e) xr[j][q * 8 + e] = v[e];
          *reinterpret_cast<uint4*>(buf(j, 0) + soff[q]) = pk;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(bar_ready(0, 0, 0)); mbar_arrive(bar_ready(1, 0, 0)); mbar_arrive(bar_ready(2, 0, 0)); }
      // The x load above sits at the head of every tile's dependency chain: pull the next tile's rows into L2 now.
      if (tile + (int)gridDim.x < n_live) {
        int bn, mtn;
        tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile + (int)gridDim.x, bn, mtn);
        const int tn = mtn * valid_rows - p.halo + r;
        if (tn >= 0 && tn < p.L) {
#pragma unroll
          for (int q = 0; q < kCPT; ++q)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x32 + (((size_t)bn * cchunks + q) * (size_t)p.L + tn) * 8));
        }
      }
      // ---- one conv of one chain ----
      auto step = [&](auto second_c, auto last_c, auto j_c, int c) {
        constexpr bool second = decltype(second_c)::value;
        constexpr bool last = decltype(last_c)::value;
        constexpr int j = decltype(j_c)::value;
        const float* bias_c = bias_s + (j * kChainMaxConvs + c) * C;
        uint8_t* out_buf = buf(j, second ? 0 : 1);
        ok = ok && mbar_wait_relaxed(bar_acc_full(j, 0), n & 1u, p.error_flag);
        if (!ok) return;
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((j * MS + s) * N);
#pragma unroll
        for (int g = 0; g < kCPT / 2; ++g) {
          uint32_t rr[16];
          __syncwarp();
          tmem_ld16(t_addr + (uint32_t)(g * 16), rr);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int q = g * 2 + h;
            float v[8];
            const float4 b0 = *reinterpret_cast<const float4*>(bias_c + q * 8), b1 = *reinterpret_cast<const float4*>(bias_c + q * 8 + 4);
            v[0] = __uint_as_float(rr[h * 8 + 0]) + b0.x; v[1] = __uint_as_float(rr[h * 8 + 1]) + b0.y;
            v[2] = __uint_as_float(rr[h * 8 + 2]) + b0.z; v[3] = __uint_as_float(rr[h * 8 + 3]) + b0.w;
            v[4] = __uint_as_float(rr[h * 8 + 4]) + b1.x; v[5] = __uint_as_float(rr[h * 8 + 5]) + b1.y;
            v[6] = __uint_as_float(rr[h * 8 + 6]) + b1.z; v[7] = __uint_as_float(rr[h * 8 + 7]) + b1.w;
            if constexpr (second) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { v[e] += xr[j][q * 8 + e]; xr[j][q * 8 + e] = v[e]; }
            }
            if constexpr (!last) *reinterpret_cast<uint4*>(out_buf + soff[q]) = pack8_lrelu(v, 0.1f, inside, bf16);
          }
        }
        if constexpr (!last) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_ready(j, second ? 0 : 1, 0));
        } else {
          tc_fence_before();                                     // TMEM reads done before the next tile's MMAs
        }
      };
      using F = std::false_type;
      using T = std::true_type;
      using J0 = std::integral_constant<int, 0>;
      using J1 = std::integral_constant<int, 1>;
      using J2 = std::integral_constant<int, 2>;
#pragma unroll 1
      for (int m = 0; m + 1 < n_pairs && ok; ++m) {
        step(F{}, F{}, J0{}, 2 * m); step(F{}, F{}, J1{}, 2 * m); step(F{}, F{}, J2{}, 2 * m); ++n;
        step(T{}, F{}, J0{}, 2 * m + 1); step(T{}, F{}, J1{}, 2 * m + 1); step(T{}, F{}, J2{}, 2 * m + 1); ++n;
      }
      const int cl = 2 * (n_pairs - 1);
      step(F{}, F{}, J0{}, cl); step(F{}, F{}, J1{}, cl); step(F{}, F{}, J2{}, cl); ++n;
      step(T{}, T{}, J0{}, cl + 1); step(T{}, T{}, J1{}, cl + 1); step(T{}, T{}, J2{}, cl + 1); ++n;
      // ---- multi-receptive-field mean: xs = 0; xs += r0; xs += r1; xs += r2; x = xs / 3 (archi.py:82-86) ----
      if (keep && ok) {
        uint4 pk16[kCPT];
#pragma unroll
        for (int q = 0; q < kCPT; ++q) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = ((xr[0][q * 8 + e] + xr[1][q * 8 + e]) + xr[2][q * 8 + e]) / 3.0f;
          const size_t idx = (((size_t)b * cchunks + q) * (size_t)p.L + t) * 8;
          if (p.flags & EPI_OUT32) {
            stg_f8(p.out32 + idx, v);
          }
          if (p.flags & EPI_OUT16) pk16[q] = pack8_lrelu(v, p.slope_out, true, bf16);
        }
        if (p.flags & EPI_OUT16) {                                 // the row's C 16-bit values are contiguous: 32-byte stores
          uint8_t* o = static_cast<uint8_t*>(p.out16) + ((size_t)b * (size_t)p.L + t) * cchunks * 16;
#pragma unroll
          for (int q = 0; q < kCPT; q += 2) stg_u8(o + q * 16, pk16[q], pk16[q + 1]);
        }
      }
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
