// Grouped (block-Toeplitz) fused ResBlock kernel, first epilogue mapping: one thread owns one whole 128-byte row of ONE
// stream's tile (8 epilogue warps per stream), the fused transposed conv hands over through a res_free barrier, the running
// multi-receptive-field sum uses the channel-blocked [C/8][L][8] layout.  It issues ~20 % fewer epilogue instructions than
// the half-row mapping of chain_group_tc.cuh and was the faster one for C = 32 (864-2400 cycles of MMAs per conv and stream
// cover the epilogue latency) until the issue loops went uniform; now it only wins the k = 3 block (0.85 vs 0.93 ms) and
// loses the stage (4.4-4.5 vs 4.25 ms), so it is an option (SATOOLS_B200_GROUP_V1 bit mask), not the default.
// Data flow, shared-memory layout, weight packing and parameters are those of chain_group_tc.cuh.
#pragma once
#include "chain_group_tc.cuh"

namespace sa {
namespace tc {
namespace v1 {

template <int C, bool BF16, int NS, int MS>
__global__ void __launch_bounds__(kGrpThreads, 1) group_chain_kernel(const __grid_constant__ GroupParams p) {
  static_assert(C == 16 || C == 32, "grouped formulation: C = 16 (G = 4) or C = 32 (G = 2)");
  static_assert(NS * MS * 4 == kGrpEpiWarps && NS <= kGrpMaxStreams, "NS * MS sub-tiles of 2 x 64 TMEM columns fill the 512 columns");
  static_assert(NS == kGrpMmaWarps, "one MMA-issuing warp per stream");
  constexpr int kGrpStreams = NS, kGrpMS = MS, kGrpRows = MS * 128;
  constexpr uint32_t kGrpBufBytes = grp_buf_bytes(MS);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int G = 64 / C;                                      // positions per 128-byte row
  constexpr int R = kGrpRows * G;                                // positions per stream tile
  constexpr uint32_t PB = 2u * C;                                // bytes per position
  constexpr uint32_t kPadBytes = kGrpPadRows * 128;
  constexpr int kGroupsPerPos = C / 16;                          // 16-column TMEM groups per position
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops (barrier addresses,
  // descriptors, counters) in uniform registers instead of converting them per use (R2UR), as CUTLASS' canonical_warp_idx_sync
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;      // (the shuffle form of the other kernels makes ptxas spill more here: 684 vs 128 bytes of spill loads)
  constexpr int kWarpW = kGrpEpiWarps, kWarpMma = kGrpEpiWarps + 1;
  // smem: buf[stream][A|T], weight ring, ones tile (128 rows x 32 B, SWIZZLE_32B), barriers
  auto buf = [&](int st, int t) { return smem + (uint32_t)(st * 2 + t) * kGrpBufBytes; };
  uint8_t* w_smem = smem + 2 * NS * kGrpBufBytes;
  uint8_t* ones_smem = w_smem + (size_t)p.n_wstages * kGrpStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones_smem + kGrpOnesBytes);
  // barriers: ready[stream][A|T] (8), acc_full[stream] (4), w_full[12], w_empty[12]
  auto bar_ready = [&](int st, int t) { return smem_u32(&bars[st * 2 + t]); };
  auto bar_acc_full = [&](int st) { return smem_u32(&bars[2 * kGrpMaxStreams + st]); };
  auto bar_w_full = [&](int i) { return smem_u32(&bars[3 * kGrpMaxStreams + i]); };
  auto bar_w_empty = [&](int i) { return smem_u32(&bars[3 * kGrpMaxStreams + kGrpMaxStages + i]); };
  // fuse_up: up_full[stream] (input tile landed), up_empty[stream] (buffer A free for the next tile), res_free[stream]
  // (the final epilogue has read the residual columns)
  auto bar_up_full = [&](int st) { return smem_u32(&bars[3 * kGrpMaxStreams + 2 * kGrpMaxStages + st]); };
  auto bar_up_empty = [&](int st) { return smem_u32(&bars[4 * kGrpMaxStreams + 2 * kGrpMaxStages + st]); };
  auto bar_res_free = [&](int st) { return smem_u32(&bars[5 * kGrpMaxStreams + 2 * kGrpMaxStages + st]); };
  auto bar_turn = [&](int w) { return smem_u32(&bars[6 * kGrpMaxStreams + 2 * kGrpMaxStages + w]); };   // issue turn of MMA warp w
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 7 * kGrpMaxStreams + 2 * kGrpMaxStages);

  const int valid = R - 2 * p.halo;
  __shared__ int tile_pre[kMaxMapItems + 1];
  tilemap_build(tile_pre, p.map, p.L, valid);                    // visible after the __syncthreads() below
  constexpr bool bf16 = BF16;                                    // both 16-bit flavours of the epilogue would double its code
  const int n_pairs = p.n_convs / 2;

  if (warp == kWarpW && lane == 0) {
    for (int st = 0; st < kGrpStreams; ++st) {
      mbar_init(bar_ready(st, 0), kGrpMS * 4);                   // the warps of the stream
      mbar_init(bar_ready(st, 1), kGrpMS * 4);
      mbar_init(bar_acc_full(st), 1);
      mbar_init(bar_up_full(st), 1);
      mbar_init(bar_up_empty(st), 1);
      mbar_init(bar_res_free(st), kGrpMS * 4);
    }
    if (p.fuse_up) prefetch_tmap(&p.up_map);
    for (int w = 0; w < kGrpMmaWarps; ++w) mbar_init(bar_turn(w), 1);
    mbar_arrive(bar_turn(0));                                    // stream 0 issues first
    for (int i = 0; i < kGrpMaxStages; ++i) { mbar_init(bar_w_full(i), 1); mbar_init(bar_w_empty(i), kGrpMmaWarps); }   // w_empty: every MMA warp has read the stage
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(smem_u32(tmem_holder), 512);
  // zero the four staged tiles once: the leading slack rows are never written, everything else only ever holds finite
  // 16-bit activations (a Toeplitz block multiplies positions outside its taps by an exact 0)
  for (uint32_t i = threadIdx.x; i < 2 * NS * kGrpBufBytes / 16; i += kGrpThreads)
    *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
  // the ones tile: A operand of the bias slice, [128 rows][16] with 1.0 in columns 0 and 1 (K-major, SWIZZLE_32B:
  // the 16-byte chunk at linear offset o lives at o ^ (((o >> 7) & 1) << 4); chunk 0 of every 32-byte row is even)
  for (uint32_t i = threadIdx.x; i < kGrpOnesBytes / 16; i += kGrpThreads) {
    const uint32_t o = i * 16u;
    const uint32_t one2 = bf16 ? 0x00003F80u | 0x3F800000u : 0x00003C00u | 0x3C000000u;   // {1.0, 1.0}
    *reinterpret_cast<uint4*>(ones_smem + (o ^ (((o >> 7) & 1u) << 4))) = make_uint4((i & 1u) ? 0u : one2, 0, 0, 0);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int n_live = tilemap_total(tile_pre, p.map, p.total_tiles);
  const int my_tiles = ((int)blockIdx.x < n_live) ? (n_live - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_iters = (my_tiles + NS - 1) / NS;                  // NS tiles (streams) per iteration

  if (warp == kWarpW) {
    // ===== weight producer: the stages of conv c, once per iteration (both streams read them) =====
    const bool leader = elect_one();
    int slot = 0;
    uint32_t par = 1;                                            // parity of the previous use of `slot`
    bool wrapped = false, ok = true;
    uint32_t n = 0;                                              // (iteration, chain) counter
    for (int it = 0; it < n_iters && ok; ++it)
     for (int j = 0; j < p.n_chains && ok; ++j, ++n) {
      if (p.fuse_up) {
        // the input tile of every stream: rows [t0 / G - 8, t0 / G + 264) of the item (out-of-range rows arrive as zeros:
        // the zero padding of the transposed conv), into buffer A once conv1 of the previous chain's last pair has read it
        for (int st = 0; st < NS && ok; ++st) {
          if (n > 0) ok = mbar_wait_relaxed(bar_up_empty(st), (n - 1) & 1u, p.error_flag);
          if (!ok) break;
          const int tile = (int)blockIdx.x + (NS * it + st) * (int)gridDim.x;
          if (tile < n_live) {
            int b, mt;
            tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile, b, mt);
            const int row0 = (mt * valid - p.halo) / G - kGrpPadRows;
            if (leader) {
              mbar_arrive_expect_tx(bar_up_full(st), kGrpBufBytes);
              const uint32_t dst = smem_u32(buf(st, 0));
              tma_load_3d(dst, &p.up_map, bar_up_full(st), 0, row0, b);
              tma_load_3d(dst + kGrpBufBytes / 2, &p.up_map, bar_up_full(st), 0, row0 + (int)(kGrpBufBytes / 256), b);
            }
          } else if (leader) {
            mbar_arrive(bar_up_full(st));                          // a stream without a tile: nothing to load
          }
          __syncwarp();
        }
      }
      for (int c = p.fuse_up ? -1 : 0; c < p.n_convs && ok; ++c) {
        const int n_stg = c < 0 ? p.up_stages : p.stages_per_conv[j];
        const uint8_t* src = c < 0 ? static_cast<const uint8_t*>(p.up_w)
                                   : static_cast<const uint8_t*>(p.w[j]) + (size_t)c * p.stages_per_conv[j] * kGrpStageBytes;
        for (int i = 0; i < n_stg; ++i) {
          if (wrapped) ok = mbar_wait_relaxed(bar_w_empty(slot), par, p.error_flag);
          if (!ok) break;
          if (leader) {
            mbar_arrive_expect_tx(bar_w_full(slot), kGrpStageBytes);
            bulk_load(smem_u32(w_smem) + (uint32_t)slot * kGrpStageBytes, src + (size_t)i * kGrpStageBytes, kGrpStageBytes,
                      bar_w_full(slot));
          }
          __syncwarp();
          if (++slot == p.n_wstages) { slot = 0; par ^= 1u; wrapped = true; }
        }
      }
     }
  } else if (warp >= kWarpMma) {
    // ===== MMA issuers, one warp per stream (warp-uniform loop, one elected lane issues) =====
    // Whatever the issuing thread does between its tcgen05.mma instructions is NOT hidden behind the queued MMAs
    // (tools/mma_bench5.cu: an already-complete try_wait + tcgen05.fence costs ~90 cycles, a commit ~50, a __syncwarp and
    // the loop bookkeeping ~60; N = 64 MMAs then run at 62-85 cycles instead of 48), but the tensor pipe takes MMAs from a
    // second warp meanwhile (tools/mma_bench6.cu: two issuing warps with that overhead: 48.0 cycles per MMA in aggregate).
    // The streams are independent dependency chains, so each gets its own issuer; both read the shared weight ring (every
    // stage is waited for and released by both: w_empty counts two arrivals).
    const int my_st = warp - kWarpMma;
    // The first version walked the slices with a runtime count and per-MMA predicates: 16 issued instructions per MMA
    // and ~100 cycles per MMA on this single warp (ncu source view, profiles/r2_group_v1_*).  The slice count is a
    // compile-time constant per filter length now: one conv of one stream is straight-line code whose descriptors
    // differ by constant adds.
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(64, bf16);
    constexpr uint32_t hiA = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);     // SWIZZLE_128B, 8-row groups 1024 B apart
    constexpr uint32_t hiB = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);      // SWIZZLE_32B, 8-row groups 256 B apart
    const uint32_t b_lo0 = desc_lo(smem_u32(w_smem));
    const uint64_t ones_desc = desc64(desc_lo(smem_u32(ones_smem)), hiB);
    const int n_wst = p.n_wstages;
    int slot = 0;
    uint32_t par = 0;
    bool ok = true;
    uint32_t turn = 0;                                           // issue turns this warp has taken
#ifdef SA_DIAG
    const bool timing = p.timing != nullptr;
#else
    constexpr bool timing = false;                                             // in-kernel cycle counters: -DSA_DIAG builds only
#endif
    long long t_ready = 0, t_w = 0, t_begin = timing ? clock64() : 0;
    // one conv (NSL Toeplitz slices + the bias slice) of both streams
    // kind 0: conv1 (reads A, fresh accumulator), 1: conv2 (reads T, accumulates onto the residual), 2: the fused
    // transposed conv (reads the TMA-staged input tile in A, writes the residual columns)
    auto conv = [&](auto nsl_c, auto kind_c, uint32_t ph0, int c, uint32_t n) {
      constexpr int NSL = decltype(nsl_c)::value;
      constexpr int KIND = decltype(kind_c)::value;
      constexpr bool CONV2 = KIND == 1;                                        // accumulates onto the residual
      constexpr int NTOT = NSL + 1;
      constexpr int NSTG = (NTOT + kGrpSlicesPerStage - 1) / kGrpSlicesPerStage;
      // bytes before the row's first position that the first slice starts at: (k - 1) / 2 positions, or one INPUT position
      // (= 2 output positions' worth of bytes: twice the channels) for the transposed conv
      constexpr uint32_t lead = KIND == 2 ? 2u * PB : (uint32_t)((NSL / kGroupsPerPos - G) / 2) * PB;
      constexpr int t_in = KIND == 1 ? 1 : 0;                                  // conv1 / up read A, conv2 reads T
      constexpr int d_idx = KIND == 0 ? 0 : 1;                                 // conv1 -> accumulator, conv2 / up -> residual
      const uint32_t rdy_parity = (ph0 + (uint32_t)(c >> 1)) & 1u;            // ph0: pairs completed before this chain
      const int st = my_st;
      if (!ok) return;
      // 1. everything this conv depends on, outside the issue turn: activations of this stream, all weight stages
      const long long tr0 = timing ? clock64() : 0;
      if (KIND == 2) {
        ok = mbar_wait(bar_up_full(st), n & 1u, p.error_flag);                 // input tile landed
        if (ok && n > 0) ok = mbar_wait(bar_res_free(st), (n - 1) & 1u, p.error_flag);   // previous residual consumed
      } else {
        ok = mbar_wait(bar_ready(st, t_in), rdy_parity, p.error_flag);
      }
      if (timing) t_ready += clock64() - tr0;
      if (!ok) return;
      const int slot0 = slot;
      {
        const long long tw0 = timing ? clock64() : 0;
#pragma unroll
        for (int i = 0; i < NSTG; ++i) {
          ok = ok && mbar_wait(bar_w_full(slot), par, p.error_flag);
          if (++slot == n_wst) { slot = 0; par ^= 1u; }
        }
        if (timing) t_w += clock64() - tw0;
      }
      if (!ok) return;
      // 2. the issue turn: the two MMA warps alternate conv by conv (stream 0, stream 1, stream 0, ...), so that one stream's
      // MMAs run under the other stream's epilogue as with a single issuer -- left to themselves the two issuers fall into
      // lockstep (both streams in their MMA phase, then both in their epilogue phase with the tensor pipe idle: measured
      // slower than one issuer) -- while the waits above and the commits below overlap the other warp's MMAs.
      ok = mbar_wait(bar_turn(my_st), turn & 1u, p.error_flag);
      ++turn;
      if (!ok) return;
      tc_fence_after();
      const uint32_t a_lo0 = desc_lo(smem_u32(buf(st, t_in)) + kPadBytes - lead);
      const uint32_t d_tmem = tmem_base + (uint32_t)((st * 2 + d_idx) * kGrpMS * 64);
      // 3. NTOT x MS MMAs back to back
      {
        int sl = slot0;
#pragma unroll
        for (int i = 0; i < NSTG; ++i) {
          const uint32_t b_lo = b_lo0 + (uint32_t)sl * (kGrpStageBytes >> 4);
#pragma unroll
          for (int s = 0; s < kGrpMS; ++s) {
#pragma unroll
            for (int qq = 0; qq < kGrpSlicesPerStage; ++qq) {
              const int q = i * kGrpSlicesPerStage + qq;
              if (q < NTOT) {
                const uint64_t adesc = (q < NSL) ? desc64(a_lo0 + (uint32_t)(s * 128 * 8 + 2 * q), hiA) : ones_desc;
                if (leader)
                  umma_f16(d_tmem + (uint32_t)(s * 64), adesc, desc64(b_lo + (uint32_t)qq * (kGrpSliceBytes >> 4), hiB), idesc,
                           (CONV2 || q > 0) ? 1u : 0u);
              }
            }
          }
          if (++sl == n_wst) sl = 0;
        }
      }
      // 4. hand the turn over, then the completion tracking (commits follow this thread's MMAs whenever they are issued)
      if (leader) {
        mbar_arrive(bar_turn(my_st ^ 1));
        int sl = slot0;
#pragma unroll
        for (int i = 0; i < NSTG; ++i) {
          umma_commit(bar_w_empty(sl));                              // this stream has read the stage
          if (++sl == n_wst) sl = 0;
        }
        umma_commit(bar_acc_full(st));
        if (KIND == 0 && c == p.n_convs - 2 && p.fuse_up) umma_commit(bar_up_empty(st));   // buffer A read for the last time
      }
      __syncwarp();
    };
    using K0 = std::integral_constant<int, 0>;
    using K1 = std::integral_constant<int, 1>;
    using K2 = std::integral_constant<int, 2>;
    auto chain = [&](auto nsl_c, uint32_t ph0, uint32_t n) {
      for (int c = 0; c < p.n_convs && ok; c += 2) {
        conv(nsl_c, K0{}, ph0, c, n);
        if (ok) conv(nsl_c, K1{}, ph0, c + 1, n);
      }
    };
    constexpr int CPP = kGroupsPerPos;
    constexpr int NSL_UP = (G / 2 + 2) * 2 * CPP;                 // (own + 2 halo) input positions x (2 C / 16) channel blocks
    uint32_t ph0 = 0, n = 0;
    for (int it = 0; it < n_iters && ok; ++it)
      for (int j = 0; j < p.n_chains && ok; ++j, ph0 += (uint32_t)n_pairs, ++n) {
        if (p.fuse_up) conv(std::integral_constant<int, NSL_UP>{}, K2{}, 0u, 0, n);
        if (!ok) break;
        const int nsl = p.n_slices[j];
        if (nsl == (G + 2) * CPP) chain(std::integral_constant<int, (G + 2) * CPP>{}, ph0, n);           // k = 3
        else if (nsl == (G + 6) * CPP) chain(std::integral_constant<int, (G + 6) * CPP>{}, ph0, n);      // k = 7
        else if (nsl == (G + 10) * CPP) chain(std::integral_constant<int, (G + 10) * CPP>{}, ph0, n);    // k = 11
        else { if (p.error_flag) atomicExch(p.error_flag, 1); ok = false; }                             // not instantiated (the host checks)
      }
    if (timing && lane == 0 && my_st == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 0), (unsigned long long)(clock64() - t_begin));
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 1), (unsigned long long)t_ready);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 2), (unsigned long long)t_w);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 3), (unsigned long long)(clock64() - t_begin - t_ready - t_w));
    }
  } else {
    // ===== epilogue warps: one thread = one 128-byte row (G positions x C channels) of its stream's tile =====
    // Instruction diet (ncu source view of the first versions: ~300 issued instructions per conv and thread at ~5.6
    // cycles each = the 1.7k-cycle epilogue latency that the MMA warp waited for): no bias add (it is in the MMA), one
    // branch per row for the zero padding instead of per-store selects, tile-invariant swizzled store offsets, TMEM loads
    // one group ahead of the conversion, diagnostics compiled out unless SA_DIAG.
    const int st = warp / (4 * MS);
    const int s = (warp >> 2) % MS;
    const int lg = warp & 3;
    const int r = s * 128 + lg * 32 + lane;                      // row within the stream tile
    constexpr int cchunks = C / 8;
    uint8_t* const bufA = buf(st, 0);
    uint8_t* const bufT = buf(st, 1);
    const uint32_t t_acc1 = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(((st * 2 + 0) * kGrpMS + s) * 64);
    const uint32_t t_res = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(((st * 2 + 1) * kGrpMS + s) * 64);
    auto swz128 = [](uint32_t lin) { return lin ^ (((lin >> 7) & 7u) << 4); };
    // tile-invariant d-major maps of this row's G positions, for every pair with a dilated conv1 (16 bits each):
    //   physP(m, g) = swizzled byte offset of time (G r + g) in the d_m-major input tile of pair m's conv1
    //   tauI(m, g)  = time whose conv1 output this row holds at position g in pair m (>= R: none)
    uint32_t pmap[kGrpMaxPairs][G];
#pragma unroll
    for (int m = 0; m < kGrpMaxPairs; ++m) {
      const int d = (m < n_pairs) ? p.dil[m] : 1;
      const int Q = (R + d - 1) / d;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int tau = G * r + g;
        const int pp = (d == 1) ? tau : (tau % d) * Q + tau / d;
        const int ti = (d == 1) ? tau : d * (tau % Q) + tau / Q;
        pmap[m][g] = swz128(kPadBytes + (uint32_t)pp * PB) | ((uint32_t)ti << 16);
      }
    }
    auto physP = [&](int m, int g) { return pmap[m][g] & 0xFFFFu; };
    auto tauI = [&](int m, int g) { return (int)(pmap[m][g] >> 16); };
    const uint32_t lin_row = kPadBytes + (uint32_t)r * 128u;     // this row in natural order
    const uint32_t xrow = (uint32_t)(r & 7) << 4;                // its swizzle XOR (kPadBytes is a multiple of 1024)
    // 16-column group gi of this row, natural order: chunks 2 gi and 2 gi + 1 of the 128-byte row
    auto st_nat = [&](uint8_t* b, int gi, const uint4& lo, const uint4& hi8) {
      *reinterpret_cast<uint4*>(b + lin_row + (((uint32_t)(2 * gi) * 16u) ^ xrow)) = lo;
      *reinterpret_cast<uint4*>(b + lin_row + (((uint32_t)(2 * gi + 1) * 16u) ^ xrow)) = hi8;
    };
    // 16-column group gi at a position whose chunk 0 lives at the swizzled offset phys0: its chunk cc is at phys0 ^ 16 cc
    auto st_at = [&](uint8_t* b, uint32_t phys0, int gi, const uint4& lo, const uint4& hi8) {
      const uint32_t o = phys0 ^ ((uint32_t)(gi % kGroupsPerPos) * 32u);
      *reinterpret_cast<uint4*>(b + o) = lo;
      *reinterpret_cast<uint4*>(b + (o ^ 16u)) = hi8;
    };
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    auto pack16 = [&](const auto& rr, int o, uint4& lo, uint4& hi8, float slope) {       // 16 columns from offset o of rr
      float a[8], c8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { a[e] = __uint_as_float(rr[o + e]); c8[e] = __uint_as_float(rr[o + 8 + e]); }
      lo = pack8_lrelu(a, slope, true, bf16);
      hi8 = pack8_lrelu(c8, slope, true, bf16);
    };
    // TMEM -> registers in two 32-column loads, the second in flight while the first is used: f(gi, regs, offset) for the
    // 16-column groups gi = 0..3.  (tcgen05.ld / wait are .sync.aligned: the whole warp must execute them converged, so
    // every caller branches on warp-uniform conditions only; f may diverge inside.)
    auto for_groups = [&](uint32_t taddr, auto&& f) {
      uint32_t ra[32], rb[32];
      __syncwarp();
      tmem_ld32(taddr, ra);
      tmem_ld_wait();
      tmem_ld32(taddr + 32u, rb);
      f(0, ra, 0);
      f(1, ra, 16);
      __syncwarp();
      tmem_ld_wait();
      f(2, rb, 0);
      f(3, rb, 16);
      __syncwarp();
    };
    bool ok = true;
    uint32_t aph = 0;                                            // completed phases of this stream's acc_full barrier
#ifdef SA_DIAG
    const bool timing = p.timing != nullptr && warp == 0;
    long long t_x = 0, t_acc = 0, t_begin = timing ? clock64() : 0;
#define GRP_T0(v) const long long v = timing ? clock64() : 0
#define GRP_ADD(acc, v) if (timing) acc += clock64() - v
#else
#define GRP_T0(v)
#define GRP_ADD(acc, v)
#endif
    for (int it = 0; it < n_iters && ok; ++it) {
      const int tile = (int)blockIdx.x + (NS * it + st) * (int)gridDim.x;
      const bool live = tile < n_live;                            // the last iteration may have streams without a tile
      int b = 0, mt = 0;
      if (live) tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile, b, mt);
      const int t0 = mt * valid - p.halo;                         // time of the tile's first position
      const int t_row = t0 + G * r;
      const bool inside = live && t_row >= 0 && t_row < p.L;      // L, halo, valid are multiples of G: whole rows
      const bool keep = inside && G * r >= p.halo && G * r < R - p.halo;
      const bool interior = live && t0 >= 0 && t0 + R <= p.L;     // no position of the tile is outside the utterance
      const bool w_all = __all_sync(0xffffffffu, inside), w_any = __any_sync(0xffffffffu, inside);   // warp-uniform
      // The next tile of this stream: have its rows in L2 by the time they are needed.
      {
        const int tile_n = tile + NS * (int)gridDim.x;
        if (tile_n < n_live) {
          int bn, mtn;
          tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile_n, bn, mtn);
          const int tn = mtn * valid - p.halo + G * r;
          if (tn >= 0 && tn < p.L) {
#pragma unroll
            for (int q = 0; q < cchunks; ++q)
#pragma unroll
              for (int g = 0; g < G; g += 4)                       // 32 B per position: one 128-byte line holds 4
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x32 + (((size_t)bn * cchunks + q) * (size_t)p.L + (size_t)(tn + g)) * 8));
          }
        }
      }
#pragma unroll 1
      for (int j = 0; j < p.n_chains && ok; ++j) {
      const uint32_t flags = p.flags[j];
      GRP_T0(tx0);
      // ---- x: residual stream -> tensor memory, lrelu(x) -> input tile of pair 0's conv1 ----
      if (p.fuse_up) {
        // x = the stage's transposed conv, computed by the MMA warp from the staged input tile into the residual
        // columns: read it back once to stage lrelu(x); rows outside the utterance are the zero padding of conv1
        ok = mbar_wait_relaxed(bar_acc_full(st), aph & 1u, p.error_flag);
        ++aph;
        tc_fence_after();
        if (ok) {
          const int d0 = p.dil[0];
          if (w_any) {
            for_groups(t_res, [&](int gi, const uint32_t (&rr)[32], int o) {
              uint4 lo, hi8;
              pack16(rr, o, lo, hi8, 0.1f);
              if (!inside) { lo = zero4; hi8 = zero4; }
              if (d0 == 1) st_nat(bufA, gi, lo, hi8); else st_at(bufA, physP(0, gi / kGroupsPerPos), gi, lo, hi8);
            });
          } else {
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
              if (d0 == 1) st_nat(bufA, gi, zero4, zero4); else st_at(bufA, physP(0, gi / kGroupsPerPos), gi, zero4, zero4);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_ready(st, 0));
        }
      } else
      {
        const int d0 = p.dil[0];
        float4 xq[4][4];                                          // all loads in flight before the first use
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          const int g = gi / kGroupsPerPos;
          const int ch0 = (gi % kGroupsPerPos) * 16;
          xq[gi][0] = xq[gi][1] = xq[gi][2] = xq[gi][3] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (inside) {
            const float* src = p.x32 + (((size_t)b * cchunks + (ch0 >> 3)) * (size_t)p.L + (size_t)(t_row + g)) * 8;
            ldg_f8(src, xq[gi][0], xq[gi][1]);
            ldg_f8(src + (size_t)p.L * 8, xq[gi][2], xq[gi][3]);
          }
        }
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          const int g = gi / kGroupsPerPos;
          const float4 q0 = xq[gi][0], q1 = xq[gi][1], q2 = xq[gi][2], q3 = xq[gi][3];
          const uint32_t rr[16] = {__float_as_uint(q0.x), __float_as_uint(q0.y), __float_as_uint(q0.z), __float_as_uint(q0.w),
                                   __float_as_uint(q1.x), __float_as_uint(q1.y), __float_as_uint(q1.z), __float_as_uint(q1.w),
                                   __float_as_uint(q2.x), __float_as_uint(q2.y), __float_as_uint(q2.z), __float_as_uint(q2.w),
                                   __float_as_uint(q3.x), __float_as_uint(q3.y), __float_as_uint(q3.z), __float_as_uint(q3.w)};
          __syncwarp();
          tmem_st16(t_res + (uint32_t)(gi * 16), rr);
          uint4 lo, hi8;
          pack16(rr, 0, lo, hi8, 0.1f);                           // rows outside the utterance were loaded as zeros
          if (d0 == 1) st_nat(bufA, gi, lo, hi8); else st_at(bufA, physP(0, g), gi, lo, hi8);
        }
        tmem_st_wait();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready(st, 0));
      }
      GRP_ADD(t_x, tx0);
      // a running sum left in HBM by an earlier launch: have it in L2 when the final epilogue needs it
      if (p.n_chains == 1 && keep && (flags & (EPI_SUM_ADD | EPI_SUM_FIN))) {
#pragma unroll
        for (int q = 0; q < cchunks; ++q)
#pragma unroll
          for (int g = 0; g < G; g += 4)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sum32 + (((size_t)b * cchunks + q) * (size_t)p.L + (size_t)(t_row + g)) * 8));
      }
      // ---- the (conv1, conv2) pairs ----
#pragma unroll
      for (int m = 0; m < kGrpMaxPairs; ++m) {
        if (m < n_pairs && ok) {
          const int d = p.dil[m];
          // conv1: TMEM -> lrelu -> conv2's input tile in natural order
          {
            GRP_T0(ta0);
            ok = mbar_wait_relaxed(bar_acc_full(st), aph & 1u, p.error_flag);
            ++aph;
            tc_fence_after();
            GRP_ADD(t_acc, ta0);
          }
          if (ok) {
            if (d == 1) {
              if (w_all) {
                for_groups(t_acc1, [&](int gi, const uint32_t (&rr)[32], int o) {
                  uint4 lo, hi8;
                  pack16(rr, o, lo, hi8, 0.1f);
                  st_nat(bufT, gi, lo, hi8);
                });
              } else if (!w_any) {
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) st_nat(bufT, gi, zero4, zero4);
              } else {                                            // the utterance ends inside this warp's rows
                for_groups(t_acc1, [&](int gi, const uint32_t (&rr)[32], int o) {
                  uint4 lo, hi8;
                  pack16(rr, o, lo, hi8, 0.1f);
                  st_nat(bufT, gi, inside ? lo : zero4, inside ? hi8 : zero4);
                });
              }
            } else if (interior) {
              for_groups(t_acc1, [&](int gi, const uint32_t (&rr)[32], int o) {
                uint4 lo, hi8;
                pack16(rr, o, lo, hi8, 0.1f);
                const int tau = tauI(m, gi / kGroupsPerPos);
                if (tau < R) st_at(bufT, swz128(kPadBytes + (uint32_t)tau * PB), gi, lo, hi8);
              });
            } else {
              for_groups(t_acc1, [&](int gi, const uint32_t (&rr)[32], int o) {
                uint4 lo, hi8;
                pack16(rr, o, lo, hi8, 0.1f);
                const int tau = tauI(m, gi / kGroupsPerPos);
                const int tt = t0 + tau;
                const bool ins = live && tt >= 0 && tt < p.L;
                if (tau < R) st_at(bufT, swz128(kPadBytes + (uint32_t)tau * PB), gi, ins ? lo : zero4, ins ? hi8 : zero4);
              });
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready(st, 1));
          }
          // conv2: the accumulator IS the residual stream x_{m+1}
          {
            GRP_T0(ta0);
            ok = ok && mbar_wait_relaxed(bar_acc_full(st), aph & 1u, p.error_flag);
            ++aph;
            tc_fence_after();
            GRP_ADD(t_acc, ta0);
          }
          if (ok) {
            if (m + 1 < n_pairs) {
              const int mn = m + 1 < kGrpMaxPairs ? m + 1 : 0;
              const int dn = p.dil[mn];
              if (w_all) {
                for_groups(t_res, [&](int gi, const uint32_t (&rr)[32], int o) {
                  uint4 lo, hi8;
                  pack16(rr, o, lo, hi8, 0.1f);
                  if (dn == 1) st_nat(bufA, gi, lo, hi8); else st_at(bufA, physP(mn, gi / kGroupsPerPos), gi, lo, hi8);
                });
              } else if (!w_any) {
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                  if (dn == 1) st_nat(bufA, gi, zero4, zero4); else st_at(bufA, physP(mn, gi / kGroupsPerPos), gi, zero4, zero4);
                }
              } else {
                for_groups(t_res, [&](int gi, const uint32_t (&rr)[32], int o) {
                  uint4 lo, hi8;
                  pack16(rr, o, lo, hi8, 0.1f);
                  if (!inside) { lo = zero4; hi8 = zero4; }
                  if (dn == 1) st_nat(bufA, gi, lo, hi8); else st_at(bufA, physP(mn, gi / kGroupsPerPos), gi, lo, hi8);
                });
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_ready(st, 0));
            } else {
              // final epilogue: multi-receptive-field combine (archi.py:82-86) + output streams
              if (flags & (EPI_SUM_ADD | EPI_SUM_FIN)) {          // the running sum is in L2 (prefetched, or just written)
                for_groups(t_res, [&](int gi, const uint32_t (&rr)[32], int o) {
                  if (!keep) return;
                  const int g = gi / kGroupsPerPos;
                  const int ch0 = (gi % kGroupsPerPos) * 16;
                  const size_t i0 = (((size_t)b * cchunks + (ch0 >> 3)) * (size_t)p.L + (size_t)(t_row + g)) * 8;
                  const size_t i1 = i0 + (size_t)p.L * 8;
                  float4 s0, s1, s2, s3;
                  ldg_f8(p.sum32 + i0, s0, s1);
                  ldg_f8(p.sum32 + i1, s2, s3);
                  float v[16] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
#pragma unroll
                  for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(rr[o + e]) + v[e];
                  if (flags & EPI_SUM_FIN) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = v[e] / p.n_blocks;
                  }
                  if (flags & EPI_SUM_ADD) { stg_f8(p.sum32 + i0, v); stg_f8(p.sum32 + i1, v + 8); }
                  if (flags & EPI_OUT32) { stg_f8(p.out32 + i0, v); stg_f8(p.out32 + i1, v + 8); }
                  if (flags & EPI_OUT16) {
                    float lo[8], hi8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { lo[e] = v[e]; hi8[e] = v[8 + e]; }
                    uint8_t* o = static_cast<uint8_t*>(p.out16) + (((size_t)b * (size_t)p.L + (size_t)(t_row + g)) * C + ch0) * 2;
                    stg_u8(o, pack8_lrelu(lo, p.slope_out, true, bf16), pack8_lrelu(hi8, p.slope_out, true, bf16));
                  }
                });
              } else {
                for_groups(t_res, [&](int gi, const uint32_t (&rr)[32], int o) {
                  if (!keep) return;
                  const int g = gi / kGroupsPerPos;
                  const int ch0 = (gi % kGroupsPerPos) * 16;
                  const size_t i0 = (((size_t)b * cchunks + (ch0 >> 3)) * (size_t)p.L + (size_t)(t_row + g)) * 8;
                  const size_t i1 = i0 + (size_t)p.L * 8;
                  float v[16];
#pragma unroll
                  for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(rr[o + e]);
                  if (flags & EPI_SUM_SET) { stg_f8(p.sum32 + i0, v); stg_f8(p.sum32 + i1, v + 8); }
                  if (flags & EPI_OUT32) { stg_f8(p.out32 + i0, v); stg_f8(p.out32 + i1, v + 8); }
                  if (flags & EPI_OUT16) {
                    uint4 lo, hi8;
                    pack16(rr, o, lo, hi8, p.slope_out);
                    uint8_t* o = static_cast<uint8_t*>(p.out16) + (((size_t)b * (size_t)p.L + (size_t)(t_row + g)) * C + ch0) * 2;
                    stg_u8(o, lo, hi8);
                  }
                });
              }
              tc_fence_before();                                 // TMEM reads done before the next tile overwrites the residual
              if (p.fuse_up) {                                   // ... which the MMA warp does itself when the upsampler is fused
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_res_free(st));
              }
            }
          }
        }
      }
      }   // chains
    }
#ifdef SA_DIAG
    if (timing && lane == 0) {
      const long long tot = clock64() - t_begin;
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 4), (unsigned long long)tot);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 5), (unsigned long long)t_x);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 6), (unsigned long long)t_acc);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 7), (unsigned long long)(tot - t_x - t_acc));
    }
#endif
#undef GRP_T0
#undef GRP_ADD
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace v1
}  // namespace tc
}  // namespace sa
