// Fused (conv1 -> conv2) pair of a ResBlock1 for C = 128 on CTA pairs (cta_group::2): nn.py:169-174, one launch per
//     x_{m+1} = x_m + conv2(lrelu(conv1(lrelu(x_m))))            (conv1: k taps, dilation d; conv2: k taps, dilation 1)
//
// Why.  Stage 1 (C = 128) is memory bound: one launch per conv moves 16 B per element and pair (conv1: read a 2 + write t 2;
// conv2: read t 2 + read x 4 + write x' 4 + write a' 2), 17.6 GB per forward at ~4.1 TB/s.  Here lrelu(conv1 + b1) never
// leaves the SM: it goes from conv1's accumulator through the epilogue warps into a shared-memory tile that is conv2's A
// operand.  12 B per element and pair, half the launches.
//
// Geometry.  A CTA owns a tile of 128 rows (one MMA sub-tile); the two CTAs of a pair are independent tiles that share the
// weights (each holds half of the B rows, conv_pair_tc.cuh).  conv1 computes the 128 rows [m0, m0 + 128) of t; conv2 then
// has valid outputs for the inner V = 128 - (k - 1) rows (its edge rows read the zero slack rows of the t tile and are not
// stored), so tiles step by V: (k - 1) / 128 of both convs is recomputed (1.6 / 4.7 / 7.8 % for k = 3 / 7 / 11).
//
// Pipeline.  (A first version with 256-row tiles had no spare TMEM: conv1 -> t tile -> conv2 -> epilogue ran as a chain and
// the launch was slower than the two per-conv launches, 5.06 vs 4.54 ms for the stage.)  With 128-row tiles everything but
// the t tile is double buffered -- TMEM: conv1's accumulator 2 x 128 columns, conv2's 2 x 128; shared memory: two A tiles --
// and the MMA warp issues   c1(0) | c1(1) c2(0) | c1(2) c2(1) | ...   so that the tensor pipe runs conv1 of the next tile
// while the epilogue warps turn this tile's conv1 accumulator into the t tile (pass 1), and both convs of the next tile
// while they run conv2's epilogue (pass 2: residual, outputs: epi_tile of conv_tc.cuh).  Pass 1 and pass 2 have their own
// warps (see kPfP1Warps).  The t tile is double buffered as well (with one t tile pass1(i + 1) had to wait for conv2(i), and
// the chain pass 1 -> conv2 -> pass 1 set the tile period of the k = 3 launches: 4.8 us against 2.2 us of MMAs).
// Shared memory: 2 A tiles (2 panels x (128 + 2 d (k-1)/2 rows)), 2 t tiles (2 panels x (128 + 16 rows)), weight ring
// shared by the two convs (stages in MMA issue order): 169 / 183 / 195 KB + 16 KB per stage for k = 3 / 7 / 11.
#pragma once
#include "conv_pair_tc.cuh"

namespace sa {
namespace tc {

struct PairFuseParams {
  ConvParams c2;             // conv2's side and everything shared: tmap = conv1's INPUT activations, wmap = conv2's packed weights,
                             // bias = conv2's, res32 / out32 / sum32 / out16 / flags / slope_out / n_blocks of conv2's epilogue,
                             // m_rows = l_out = L, cout_total = 128, out_stride 1, rows_alloc / box_rows / nseg of the A tile,
                             // k16_per_stage, n_wstages, w_tile_bytes, total_tiles, m_tiles (tiles of V rows), map
  CUtensorMap wmap1;         // conv1's packed weights (same tiling as conv2's)
  const float* bias1;
  int k, dil1;               // taps of both convs; dilation of conv1
  int V;                     // valid rows per tile
};

constexpr int kPfTPad = 8;                                 // slack rows on both sides of the t tile (>= (k - 1) / 2)
constexpr int kPfRows = 128;                               // rows of a tile
// Warps: 0-3 pass 1 (one per TMEM lane group, all 128 columns), 4-11 pass 2 (lane group x column half), then the A producer,
// the W producer and the MMA issuer.  Pass 1 has its own warps because its hand-over needs fence.proxy.async, which
// compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: in a warp that also runs pass 2 the fence waits for that warp's outstanding
// output stores (measured: the stores alone cost +0.11 ms per k = 3 launch).  15 warps leave 128 registers per thread: the
// residual loads of pass 2 stay in registers (with 19 warps / 96 registers ptxas spilled a loaded value right behind its
// load, which exposes the whole load latency).
constexpr int kPfP1Warps = 4, kPfP2Warps = 8;
constexpr int kPfThreads = (kPfP1Warps + kPfP2Warps + 3) * 32;
constexpr int kPfTRows = kPfRows + 2 * kPfTPad;            // 144 rows of 128 bytes per panel = 18 KB (a multiple of 1024)

// Hand-over of the t tile to the MMA: generic-proxy stores -> fence.proxy.async -> arrive on the leader's barrier with the
// default (.release.cta) semantics, as conv_pair_tc.cuh does for its accumulators.  (A .release.cluster arrive / .acquire.cluster
// wait compile to MEMBAR + ERRBAR and CCTL.IVALL: the arriving warp then waits until all its earlier global stores -- the
// previous tile's outputs -- are performed; ncu showed 8 % of all stall samples there.)

__global__ void __launch_bounds__(kPfThreads, 1) resblock_pair_kernel(const __grid_constant__ PairFuseParams q) {
  const ConvParams& p = q.c2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t crank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const bool is_leader = crank == 0;

  constexpr int N = 128, NH = N / 2;
  constexpr uint32_t kTmemCols = 512;                      // conv1: columns [0, 256), conv2: [256, 512), two buffers each
  constexpr uint32_t row_bytes = 128;
  constexpr int SPP = 4, panels = 2;
  const uint32_t panel_bytes_a = (uint32_t)p.rows_alloc * row_bytes;
  const uint32_t a_bytes = (uint32_t)panels * panel_bytes_a;
  constexpr uint32_t panel_bytes_t = (uint32_t)kPfTRows * row_bytes;
  constexpr uint32_t t_bytes = (uint32_t)panels * panel_bytes_t;
  const int blocks_per_stage = p.k16_per_stage / SPP;
  const uint32_t stage_rows = (uint32_t)blocks_per_stage * NH;
  const uint32_t stage_bytes = stage_rows * row_bytes;
  uint8_t* a_smem = smem;
  uint8_t* t_smem = smem + 2u * a_bytes;
  uint8_t* w_smem = t_smem + 2u * t_bytes;                 // two t tiles
  float* bias1_s = reinterpret_cast<float*>(w_smem + (size_t)p.n_wstages * stage_bytes);
  float* bias2_s = bias1_s + N;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias2_s + N);
  auto bar_a_full = [&](int i) { return smem_u32(&bars[0 + i]); };
  auto bar_a_empty = [&](int i) { return smem_u32(&bars[2 + i]); };
  auto bar_acc1_full = [&](int i) { return smem_u32(&bars[4 + i]); };
  auto bar_acc2_full = [&](int i) { return smem_u32(&bars[6 + i]); };
  auto bar_acc2_empty = [&](int i) { return smem_u32(&bars[8 + i]); };
  auto bar_t_ready = [&](int i) { return smem_u32(&bars[10 + i]); };
  auto bar_w_full = [&](int s) { return smem_u32(&bars[12 + s]); };
  auto bar_w_empty = [&](int s) { return smem_u32(&bars[12 + kMaxStages + s]); };
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 12 + 2 * kMaxStages);

  __shared__ int tile_pre[kMaxMapItems + 1];
  tilemap_build(tile_pre, p.map, p.m_rows, q.V);
  const int h2 = (q.k - 1) / 2, r1 = q.dil1 * h2;
  const int n_blocks_total = q.k * panels;
  const int n_iters = (n_blocks_total + blocks_per_stage - 1) / blocks_per_stage;    // weight stages per conv
  const int w_row0 = (int)(crank * (p.w_tile_bytes / row_bytes));                     // this CTA's half tile

  constexpr int kWarpA = kPfP1Warps + kPfP2Warps, kWarpW = kWarpA + 1, kWarpMma = kWarpA + 2;
  if (warp == kWarpA && lane == 0) {
    prefetch_tmap(&p.tmap);
    prefetch_tmap(&p.wmap);
    prefetch_tmap(&q.wmap1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_a_full(i), 1);                   // leader: its producer's arrive.expect_tx (bytes of both CTAs)
      mbar_init(bar_a_empty(i), 1);                  // multicast commit
      mbar_init(bar_acc1_full(i), 1);                // multicast commit
      mbar_init(bar_acc2_full(i), 1);                // multicast commit
      mbar_init(bar_acc2_empty(i), 2 * kPfP2Warps);      // leader: pass-2 warps of both CTAs
      mbar_init(bar_t_ready(i), 2 * kPfP1Warps);         // leader: the pass-1 warps of both CTAs have written their t tiles
    }
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(bar_w_full(s), 1); mbar_init(bar_w_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc_pair(smem_u32(tmem_holder), kTmemCols);
  for (int i = threadIdx.x; i < N; i += kPfThreads) { bias1_s[i] = q.bias1[i]; bias2_s[i] = p.bias[i]; }
  // the slack rows of the t tile stay zero: conv2's edge rows (not stored) read them
  for (uint32_t i = threadIdx.x; i < 2u * t_bytes / 16; i += kPfThreads) *reinterpret_cast<uint4*>(t_smem + i * 16) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int n_live = tilemap_total(tile_pre, p.map, p.total_tiles);
  const int n_rounds = (n_live + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int round) { return min((int)blockIdx.x + round * (int)gridDim.x, n_live - 1); };
  auto is_dummy = [&](int round) { return (int)blockIdx.x + round * (int)gridDim.x >= n_live; };

  if (warp == kWarpA) {
    // ===== A producer (both CTAs): rows [m0 - r1, m0 + 128 + r1) of lrelu(x), completion on the leader's barrier =====
    const bool leader_lane = elect_one();
    for (int it = 0; it < n_rounds; ++it) {
      const int buf = it & 1, use = it >> 1;
      if (use > 0 && !mbar_wait_relaxed(bar_a_empty(buf), (uint32_t)(use - 1) & 1u, p.error_flag)) break;
      int b, mt;
      tilemap_locate(tile_pre, p.map, p.m_tiles, tile_of(it), b, mt);
      const int row0 = mt * q.V - h2 - r1;
      const uint32_t dst = smem_u32(a_smem) + (uint32_t)buf * a_bytes;
      const uint32_t full0 = mapa_u32(bar_a_full(buf), 0);
      if (leader_lane) {
        if (is_leader) mbar_arrive_expect_tx(bar_a_full(buf), 2u * a_bytes);
        for (int c = 0; c < panels; ++c)
          for (int s = 0; s < p.nseg; ++s)
            tma_load_4d_pair(dst + (uint32_t)c * panel_bytes_a + (uint32_t)(s * p.box_rows) * row_bytes, &p.tmap, full0, 0,
                             row0 + s * p.box_rows, c, b);
      }
      // (An L2 prefetch of the fp32 streams of pass 2 from here -- cp.async.bulk.prefetch.L2, two tiles ahead -- was tried and
      // dropped: the lines were gone again before pass 2 read them; ncu showed the residual fetched from DRAM twice, 1.19 GB
      // of reads per k = 3 launch instead of 0.76 GB.)
      __syncwarp();
    }
  } else if (warp == kWarpW) {
    // ===== W producer (both CTAs): the stages of the convs in MMA issue order: c1(0) | c1(1) c2(0) | c1(2) c2(1) | ... =====
    const bool leader_lane = elect_one();
    int slot = 0;
    uint32_t par = 1;
    bool wrapped = false, ok = true;
    auto load_conv = [&](const CUtensorMap* wm) {
      for (int i = 0; i < n_iters && ok; ++i) {
        if (wrapped) ok = mbar_wait_relaxed(bar_w_empty(slot), par, p.error_flag);
        if (!ok) break;
        const uint32_t full0 = mapa_u32(bar_w_full(slot), 0);
        if (leader_lane) {
          if (is_leader) mbar_arrive_expect_tx(bar_w_full(slot), 2u * stage_bytes);
          tma_load_2d_pair(smem_u32(w_smem) + (uint32_t)slot * stage_bytes, wm, full0, 0, w_row0 + i * (int)stage_rows);
        }
        __syncwarp();
        if (++slot == p.n_wstages) { slot = 0; par ^= 1u; wrapped = true; }
      }
    };
    load_conv(&q.wmap1);
    for (int round = 0; round < n_rounds && ok; ++round) {
      if (round + 1 < n_rounds) load_conv(&q.wmap1);
      if (ok) load_conv(&p.wmap);
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer: the leader CTA only =====
    if (is_leader) {
      const bool leader_lane = elect_one();
      const uint32_t idesc = make_idesc_pair(N, (p.flags & EPI_BF16) != 0);
      constexpr uint32_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t row16 = row_bytes >> 4;
      constexpr uint32_t b_block16 = ((uint32_t)NH * row_bytes) >> 4;
      const uint32_t b_lo0 = desc_lo(smem_u32(w_smem));
      int wslot = 0;
      uint32_t wpar = 0;
      bool ok = true;
      // one conv: n_blocks_total (tap, panel) blocks of 4 K-steps
      auto run_conv = [&](uint32_t a_tap, uint32_t a_panel16, uint32_t tap_step16, uint32_t d_tmem) {
        uint32_t a_blk = a_tap, accum = 0;
        int panel = 0, blk = 0;
        for (int i = 0; i < n_iters && ok; ++i) {
          const int slot = wslot;
          ok = mbar_wait(bar_w_full(slot), wpar, p.error_flag);
          if (++wslot == p.n_wstages) { wslot = 0; wpar ^= 1u; }
          if (!ok) break;
          uint32_t b_blk = b_lo0 + (uint32_t)slot * (stage_bytes >> 4);
          const int nb = min(blocks_per_stage, n_blocks_total - blk);
          for (int bi = 0; bi < nb; ++bi, ++blk) {
#pragma unroll
            for (int kk = 0; kk < SPP; ++kk) {
              if (leader_lane) umma_f16_pair(d_tmem, desc64(a_blk + 2u * kk, hi), desc64(b_blk + 2u * kk, hi), idesc, accum);
              accum = 1;
            }
            b_blk += b_block16;
            if (++panel == panels) { panel = 0; a_tap += tap_step16; a_blk = a_tap; } else { a_blk += a_panel16; }
          }
          if (leader_lane) umma_commit_pair(bar_w_empty(slot));
          __syncwarp();
        }
      };
      auto conv1 = [&](int it) {                    // tile `it`: A buffer and accumulator it & 1
        const int buf = it & 1;
        ok = mbar_wait(bar_a_full(buf), (uint32_t)(it >> 1) & 1u, p.error_flag);
        if (!ok) return;
        tc_fence_after();
        // (its accumulator was drained by pass 1 of tile it - 2: t_ready of that tile was waited for before conv2(it - 2))
        run_conv(desc_lo(smem_u32(a_smem) + (uint32_t)buf * a_bytes), panel_bytes_a >> 4, (uint32_t)q.dil1 * row16,
                 tmem_base + (uint32_t)(buf * N));
        if (!ok) return;
        if (leader_lane) {
          umma_commit_pair(bar_a_empty(buf));
          umma_commit_pair(bar_acc1_full(buf));
        }
        __syncwarp();
      };
      conv1(0);
      for (int it = 0; it < n_rounds && ok; ++it) {
        if (it + 1 < n_rounds) conv1(it + 1);
        if (!ok) break;
        // conv2: the t tiles of both CTAs are written; its accumulator buffer was drained by pass 2 of tile it - 2
        ok = mbar_wait(bar_t_ready(it & 1), (uint32_t)(it >> 1) & 1u, p.error_flag);
        if (ok && it >= 2) ok = mbar_wait(bar_acc2_empty(it & 1), (uint32_t)((it >> 1) - 1) & 1u, p.error_flag);
        if (!ok) break;
        tc_fence_after();
        run_conv(desc_lo(smem_u32(t_smem) + (uint32_t)(it & 1) * t_bytes) + (uint32_t)(kPfTPad - h2) * row16, panel_bytes_t >> 4, row16,
                 tmem_base + 256u + (uint32_t)((it & 1) * N));
        if (!ok) break;
        if (leader_lane) umma_commit_pair(bar_acc2_full(it & 1));
        __syncwarp();
      }
    }
  } else if (warp < kPfP1Warps) {
    // ===== pass 1 (both CTAs): conv1's accumulator + bias -> lrelu -> the t tile (zero outside the utterance: conv2's zero
    // padding).  Warp = TMEM lane group = 32 rows, all 128 columns. =====
    const int lg = warp;
    const bool bf16 = (p.flags & EPI_BF16) != 0;
    for (int it = 0; it < n_rounds; ++it) {
      int b, mt;
      tilemap_locate(tile_pre, p.map, p.m_tiles, tile_of(it), b, mt);
      const int j = lg * 32 + lane;                              // row of the t tile
      const int t = mt * q.V - h2 + j;
      const bool inside = !is_dummy(it) && t >= 0 && t < p.m_rows;
      bool ok = mbar_wait_relaxed(bar_acc1_full(it & 1), (uint32_t)(it >> 1) & 1u, p.error_flag);
      // this t tile is free once conv2 of tile it - 2 has read it
      if (ok && it >= 2) ok = mbar_wait_relaxed(bar_acc2_full(it & 1), (uint32_t)((it >> 1) - 1) & 1u, p.error_flag);
      if (!ok) break;
      tc_fence_after();
      const uint32_t row_off = (uint32_t)(kPfTPad + j) * row_bytes;
#pragma unroll
      for (int pn = 0; pn < panels; ++pn) {
        uint32_t rr[2][32];
        __syncwarp();
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((it & 1) * N + pn * 64), rr[0]);
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((it & 1) * N + pn * 64 + 32), rr[1]);
        tmem_ld_wait();
        uint8_t* const dst = t_smem + (uint32_t)(it & 1) * t_bytes + (uint32_t)pn * panel_bytes_t;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(rr[c8 >> 2][(c8 & 3) * 8 + e]) + bias1_s[pn * 64 + c8 * 8 + e];
          const uint32_t lin = row_off + (uint32_t)c8 * 16u;
          *reinterpret_cast<uint4*>(dst + (lin ^ (((lin >> 7) & 7u) << 4))) = pack8_lrelu(v, 0.1f, inside, bf16);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(bar_t_ready(it & 1), 0));
    }
  } else {
    // ===== pass 2 (both CTAs): conv2's accumulator: + bias, + residual, multi-receptive-field sum, outputs (epi_tile of
    // conv_tc.cuh).  Warp = TMEM lane group x half of the 128 columns. =====
    const int lg = warp & 3;
    const int col0 = ((warp - kPfP1Warps) >> 2) * (N / 2);
    for (int it = 0; it < n_rounds; ++it) {
      int b, mt;
      tilemap_locate(tile_pre, p.map, p.m_tiles, tile_of(it), b, mt);
      auto wait_acc = [&]() { return mbar_wait_relaxed(bar_acc2_full(it & 1), (uint32_t)(it >> 1) & 1u, p.error_flag); };
      if (!epi_tile<N, 1, N / 2>(p, bias2_s, tmem_base + 256u + (uint32_t)((it & 1) * N), b, mt * q.V - h2, 0, 0, lg, lane, col0,
                                 is_dummy(it), wait_acc, h2, kPfRows - h2))
        break;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(bar_acc2_empty(it & 1), 0));
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace tc
}  // namespace sa
