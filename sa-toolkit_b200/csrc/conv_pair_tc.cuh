// tcgen05 implicit-GEMM convolution on CTA PAIRS (cta_group::2) for the wide layers (Cout tile N = 256 / 128).
//
// Why: the single-CTA kernel (conv_tc.cuh) reads, per MMA, 128 rows of A and all N rows of B from shared
// memory and streams every weight stage from L2 into every SM.  For N >= 128 both are at their limits
// (128 B/cycle of shared-memory bandwidth; ~42 B/cycle/SM from the L2 slices -- profiles/README.md), so the
// tensor pipe waits.  With cta_group::2 two CTAs of a cluster run ONE M = 256 MMA: each CTA holds its own
// 128 rows of A and only HALF of the B rows (N/2), the hardware exchanges the halves.  Per SM that is half the
// weight bytes from L2, half the weight bytes written to and read from shared memory, same MMA rate.
//
// Same formulation, data layout, tap addressing and epilogue as conv_tc.cuh (read that header first).
// Differences:
//   * grid.x CTAs form clusters (2,1,1); the pair walks the tile list together: CTA rank r of pair c takes tile
//     2c + r + round * grid.x.  Both run the same number of rounds (a round past the end recomputes the last tile
//     and does not store).
//   * weights are packed in N/2-row tiles ([phase][2 * ntile + r][tap][panel][N/2][64]); CTA r streams tile r of its
//     n-tile with 2-D TMA loads (the packed buffer seen as rows of 128 bytes, no TMA swizzle: the host already
//     swizzled it).
//   * every TMA load (A and W, both CTAs) completes on the LEADER's (rank 0) a_full / w_full barriers
//     (.cta_group::2 loads may signal the peer CTA's barrier); the leader's elected MMA lane issues
//     tcgen05.mma.cta_group::2 and releases smem slots / publishes accumulators in BOTH CTAs with multicast commits.
//   * each CTA's 16 epilogue warps drain their own TMEM (their 128 rows x N columns) and arrive on the leader's
//     acc_empty barrier (remote mbarrier.arrive for the peer).
#pragma once
#include "conv_tc.cuh"

namespace sa {
namespace tc {

__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
}
// Arrive on a barrier of another CTA of the cluster.  Default (.release.cta) semantics on purpose: a
// .release.cluster arrive makes the warp wait (MEMBAR) until all its earlier global stores are performed, which
// serialised the epilogue behind its own output stores (ncu: membar stall, profiles/README.md).  The hand-back only
// has to order the TMEM reads, and tcgen05.fence::before_thread_sync does that.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA loads of a CTA pair: data into this CTA's shared memory, completion bytes onto `bar` (a shared::cluster
// address, here always the leader's barrier).
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {     // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)0x3) : "memory");
}
// M = 256 (two CTAs x 128 rows), N = n.
__device__ __forceinline__ uint32_t make_idesc_pair(int n, bool bf16) {
  const uint32_t fmt = bf16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24);
}

// 16 epilogue warps (four per TMEM lane group, each a quarter of the N columns): the epilogue of the wide layers is
// what the MMA waits for (residual reads + three output streams), and twice the warps keep twice the loads in flight.
constexpr int kPairEpiWarps = 16;
constexpr int kPairThreads = (kPairEpiWarps + 3) * 32;   // + A producer, W producer, MMA issuer

// Dynamic smem (identical layout in both CTAs -- the MMA addresses both through one descriptor):
//   [A ring: n_abuf * (cin/64) * rows_alloc * 128][W ring: n_wstages * stage rows * 128][bias N*4][barriers][tmem holder]
template <int N, int MSUB>
__global__ void __launch_bounds__(kPairThreads, 1) conv_pair_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops (barrier addresses,
  // descriptors, counters) in uniform registers instead of converting them per use (R2UR), as CUTLASS' canonical_warp_idx_sync
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int phase = blockIdx.y / p.n_tiles, ntile = blockIdx.y % p.n_tiles;
  const uint32_t crank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);   // warp-uniform for the compiler, like the warp index
  const bool is_leader = crank == 0;

  constexpr int NH = N / 2;                                       // B rows held by one CTA
  constexpr int kAccCols = N * MSUB;
  static_assert(2 * kAccCols <= 512, "the pair kernel double-buffers its accumulators");
  constexpr uint32_t kTmemCols = (2 * kAccCols <= 256) ? 256 : 512;
  constexpr uint32_t row_bytes = 128;                             // PW = 64
  constexpr int SPP = 4;                                          // K=16 steps per panel row
  const int panels = p.cin >> 6;
  const uint32_t panel_bytes = (uint32_t)p.rows_alloc * row_bytes;
  const uint32_t a_bytes = (uint32_t)panels * panel_bytes;
  const int blocks_per_stage = p.k16_per_stage / SPP;             // [NH][64] weight blocks per stage
  const uint32_t stage_rows = (uint32_t)blocks_per_stage * NH;
  const uint32_t stage_bytes = stage_rows * row_bytes;
  uint8_t* a_smem = smem;
  uint8_t* w_smem = smem + (size_t)p.n_abuf * a_bytes;
  float* bias_s = reinterpret_cast<float*>(w_smem + (size_t)p.n_wstages * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + N);
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
  const int n_blocks_total = n_taps * panels;
  const int n_iters = (n_blocks_total + blocks_per_stage - 1) / blocks_per_stage;   // weight stages per tile
  // first 128-byte row of this CTA's half tile inside the packed weight buffer
  const int w_row0 = (int)(((size_t)(phase * p.n_tiles + ntile) * 2 + crank) * (p.w_tile_bytes / row_bytes));

  constexpr int kWarpA = kPairEpiWarps, kWarpW = kPairEpiWarps + 1, kWarpMma = kPairEpiWarps + 2;
  if (warp == kWarpA && lane == 0) {
    prefetch_tmap(&p.tmap);
    prefetch_tmap(&p.wmap);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_a_full(i), 1);                 // leader: its producer's arrive.expect_tx (bytes of both CTAs)
      mbar_init(bar_a_empty(i), 1);                // multicast commit
      mbar_init(bar_acc_full(i), 1);               // multicast commit
      mbar_init(bar_acc_empty(i), 2 * kPairEpiWarps);  // leader: epilogue warps of both CTAs
    }
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(bar_w_full(s), 1); mbar_init(bar_w_empty(s), 1); }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc_pair(smem_u32(tmem_holder), kTmemCols);
  for (int i = threadIdx.x; i < N; i += kPairThreads) bias_s[i] = p.bias[ntile * N + i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // both CTAs' barriers and TMEM exist before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const int n_live = tilemap_total(tile_pre, p.map, p.total_tiles);
  const int n_rounds = (n_live + (int)gridDim.x - 1) / (int)gridDim.x;
  auto tile_of = [&](int round) { return min((int)blockIdx.x + round * (int)gridDim.x, n_live - 1); };
  auto is_dummy = [&](int round) { return (int)blockIdx.x + round * (int)gridDim.x >= n_live; };

  if (warp == kWarpA) {
    // ===== A producer (both CTAs): this CTA's rows, completion on the leader's barrier =====
    const bool leader_lane = elect_one();
    for (int it = 0; it < n_rounds; ++it) {
      const int tile = tile_of(it);
      const int buf = (p.n_abuf == 2) ? (it & 1) : 0, use = (p.n_abuf == 2) ? (it >> 1) : it;
      if (use > 0 && !mbar_wait_relaxed(bar_a_empty(buf), (use - 1) & 1, p.error_flag)) break;
      int b, mt;
      tilemap_locate(tile_pre, p.map, p.m_tiles, tile, b, mt);
      const int m0 = mt * (128 * MSUB);
      const int row0 = m0 + p.row_lo[phase];
      const uint32_t dst = smem_u32(a_smem) + (uint32_t)buf * a_bytes;
      const uint32_t full0 = mapa_u32(bar_a_full(buf), 0);
      if (leader_lane) {
        if (is_leader) mbar_arrive_expect_tx(bar_a_full(buf), 2u * a_bytes);
        for (int c = 0; c < panels; ++c)
          for (int s = 0; s < p.nseg; ++s)
            tma_load_4d_pair(dst + (uint32_t)c * panel_bytes + (uint32_t)(s * p.box_rows) * row_bytes, &p.tmap, full0, 0,
                             row0 + s * p.box_rows, c, b);
      }
      __syncwarp();
    }
  } else if (warp == kWarpW) {
    // ===== W producer (both CTAs): this CTA's half of every weight stage =====
    const bool leader_lane = elect_one();
    int slot = 0;
    uint32_t par = 1;
    bool wrapped = false, ok = true;
    for (int round = 0; round < n_rounds && ok; ++round) {
      for (int i = 0; i < n_iters && ok; ++i) {
        if (wrapped) ok = mbar_wait_relaxed(bar_w_empty(slot), par, p.error_flag);
        if (!ok) break;
        const uint32_t full0 = mapa_u32(bar_w_full(slot), 0);
        if (leader_lane) {
          // a stage is always a whole box (the last one of a tile reads past its tile; those blocks are not used)
          if (is_leader) mbar_arrive_expect_tx(bar_w_full(slot), 2u * stage_bytes);
          tma_load_2d_pair(smem_u32(w_smem) + (uint32_t)slot * stage_bytes, &p.wmap, full0, 0,
                           w_row0 + i * (int)stage_rows);
        }
        __syncwarp();
        if (++slot == p.n_wstages) { slot = 0; par ^= 1u; wrapped = true; }
      }
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer: the leader CTA only =====
    if (is_leader) {
      const bool leader_lane = elect_one();
      const uint32_t idesc = make_idesc_pair(N, (p.flags & EPI_BF16) != 0);
      const int row_lo = p.row_lo[phase];
      constexpr uint32_t hi = ((8u * row_bytes) >> 4) | (1u << 14) | (2u << 29);
      constexpr uint32_t row16 = row_bytes >> 4;
      constexpr uint32_t b_block16 = ((uint32_t)NH * row_bytes) >> 4;
      const uint32_t a_panel16 = panel_bytes >> 4;
      const uint32_t b_lo0 = desc_lo(smem_u32(w_smem));
      const uint32_t tap_step16 = (uint32_t)p.tap_step * row16;
      int wslot = 0;
      uint32_t wpar = 0;
      bool ok = true;
#ifdef SA_DIAG
      const bool timing = p.timing != nullptr;
#else
      constexpr bool timing = false;   // in-kernel cycle counters: -DSA_DIAG builds only
#endif
      long long t_a = 0, t_w = 0, t_acc = 0, t_begin = timing ? clock64() : 0;
      for (int it = 0; it < n_rounds && ok; ++it) {
        const int buf = (p.n_abuf == 2) ? (it & 1) : 0, use = (p.n_abuf == 2) ? (it >> 1) : it;
        const int acc = it & 1, acc_use = it >> 1;
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
          const long long tw0 = timing ? clock64() : 0;
          const int slot = wslot;
          ok = mbar_wait(bar_w_full(slot), wpar, p.error_flag);
          if (++wslot == p.n_wstages) { wslot = 0; wpar ^= 1u; }
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
                if (leader_lane) umma_f16_pair(d_base + (uint32_t)ms * N, adesc, bdesc, idesc, accum);
              }
              accum = 1;
            }
            b_blk += b_block16;
            if (++panel == panels) { panel = 0; a_tap += tap_step16; a_blk = a_tap; } else { a_blk += a_panel16; }
          }
          if (leader_lane) umma_commit_pair(bar_w_empty(slot));    // slot free in both CTAs once these MMAs have read it
          __syncwarp();
        }
        if (!ok) break;
        if (leader_lane) {
          umma_commit_pair(bar_a_empty(buf));                      // A tiles of both CTAs consumed
          umma_commit_pair(bar_acc_full(acc));                     // accumulators of both CTAs complete
        }
        __syncwarp();
      }
      if (timing && lane == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 0), 2ull * (unsigned long long)(clock64() - t_begin));
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 1), 2ull * (unsigned long long)t_a);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 2), 2ull * (unsigned long long)t_w);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 3), 2ull * (unsigned long long)t_acc);
      }
    }
  } else {
    // ===== epilogue (both CTAs): own TMEM lanes, a quarter of the N columns per warp (epi_tile, conv_tc.cuh) =====
    const int lg = warp & 3;
    const int col0 = (warp >> 2) * (N / 4);
#ifdef SA_DIAG
    const bool timing = p.timing != nullptr && warp == 0;
#else
    constexpr bool timing = false;   // in-kernel cycle counters: -DSA_DIAG builds only
#endif
    long long t_full = 0, t_begin = timing ? clock64() : 0;
    for (int it = 0; it < n_rounds; ++it) {
      const int acc = it & 1, acc_use = it >> 1;
      const int tile = tile_of(it);
      int b, mt;
      tilemap_locate(tile_pre, p.map, p.m_tiles, tile, b, mt);
      const int m0 = mt * (128 * MSUB);
      auto wait_acc = [&]() {
        const long long tf0 = timing ? clock64() : 0;
        const bool ok = mbar_wait_relaxed(bar_acc_full(acc), acc_use & 1, p.error_flag);
        if (timing) t_full += clock64() - tf0;
        return ok;
      };
      if (!epi_tile<N, MSUB, N / 4>(p, bias_s, tmem_base + (uint32_t)(acc * kAccCols), b, m0, phase, ntile, lg, lane, col0,
                                    is_dummy(it), wait_acc))
        break;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(bar_acc_empty(acc), 0));   // the leader owns the accumulator hand-back
    }
    if (timing && lane == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 4), (unsigned long long)(clock64() - t_begin));
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 5), (unsigned long long)t_full);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // the peer may still read this CTA's B half / signal its barriers
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

}  // namespace tc
}  // namespace sa
