// Fused ResBlock1 for the narrowest stages (C = 32, 16) in the "grouped" (block-Toeplitz) formulation.
//
// Why.  With time on M and Cout on N (conv_tc.cuh, chain_tc.cuh) a C-channel layer issues MMAs with N = C.  The tensor
// pipe re-reads the 128 x 16 A tile (4 KB) from shared memory for every MMA, at 128 B/cycle: an MMA costs
// max(N/2, (4 KB + N * 32 B) / 128 B) cycles = 48 / 40 / 39 for N = 64 / 32 / 16 (tools/mma_bench2.cu), so C = 32 and
// C = 16 run at 16/40 and 8/39 of the tensor rate whatever else the kernel does (caps 0.40 / 0.21).
//
// What.  Put G = 64 / C consecutive time steps ("positions") on one 128-byte shared-memory row.  An MMA row is then a
// group of G positions, its N = 64 columns are (g', co), and one K = 16 "slice" -- 16 input channels of ONE input
// position at offset c from the row's first position -- contributes to every output position of the row through the
// block-Toeplitz weight block
//     B_q[g' * C + co][ci] = W[co][ci][j = c - g']   (0 outside the k taps)
// so that D[m][g' * C + co] = sum_j sum_ci W[co][ci][j] * X[G m + g' + j - pad][ci].  A k-tap conv of 128 * G positions
// costs (G + k - 1) * C/16 MMAs of N = 64 (48 cycles each) instead of G * k * C/16 MMAs of N = C (39 - 40 cycles each):
// 2.3x fewer tensor-pipe cycles for C = 16, 1.45x for C = 32.  The A operand of slice q is the staged tile with the
// descriptor start advanced by 32 * q bytes (the swizzle XOR acts on absolute address bits: a start in the middle of a
// 128-byte row is as exact as a start in the middle of an 8-row atom, conv_tc.cuh).
//
// Dilation.  A dilated conv (d = 3, 5) is d interleaved dilation-1 convs.  Its input tile is therefore staged in
// "d-major" order, position(tau) = (tau mod d) * Q + tau div d with Q = ceil(R / d): inside a residue class consecutive
// positions are d time steps apart and the same Toeplitz blocks apply.  Rows near the class boundaries are garbage --
// they are exactly the rows within (k-1)/2 * d of the tile ends, i.e. halo rows.  The permutation costs nothing: it is
// the shared-memory address of the epilogue stores (conv2 writes the next conv1's input in d-major order, the dilated
// conv1's epilogue writes conv2's input back in natural order); per-thread positions are tile invariant.
//
// Residual stream in tensor memory.  conv2 accumulates ON TOP of the residual: the fp32 stream x lives in the TMEM
// columns that conv2 uses as accumulator (written once per tile with tcgen05.st, MMA with accumulate = 1), so
// x_{m+1} = x_m + conv2(...) needs no registers and no add.  Sums differ from the per-layer path in the last fp32 bit only.
//
// Bias in the MMA.  The epilogue of these layers is instruction-issue bound (ncu: 6 instructions per element and phase
// in the first version, more issue cycles than MMA cycles), so the bias add moves into the tensor pipe: every conv gets
// one more slice whose A operand is a constant tile of ones (columns 0 and 1) and whose B block holds the bias split in
// two 16-bit halves (hi + lo: exact to 2^-22 for fp16, 2^-16 for bf16).  conv2's bias then also lands in the residual.
//
// Two tiles per CTA ("streams").  With full (conv-granular) dependencies a single chain would leave the tensor pipe
// idle during every epilogue, so each CTA works on two independent tiles of the same ResBlock: the MMA warp issues
// conv c of stream A, then of stream B, while the epilogue warps of the other stream drain.  Both streams read the same
// weight stages from a shared-memory ring (a stage is released after stream B has used it).
//
// Warp roles (576 threads): 16 epilogue warps (stream = w / 8, sub-tile = (w / 4) % 2, TMEM lane group = w % 4; one
// thread owns one 128-byte row = G positions x C channels), weight producer, MMA issuer (last, see chain_tc.cuh).
// TMEM: per stream and sub-tile 64 columns conv1 accumulator + 64 columns residual / conv2 accumulator = 512 columns.
#pragma once
#include "chain_tc.cuh"

namespace sa {
namespace tc {

// NS streams (independent tiles in flight per CTA) of MS sub-tiles (128 rows of 128 bytes) each; NS * MS = 4 fills the
// 512 TMEM columns.  (2, 2): 256-row tiles, less recomputed halo; (4, 1): four dependency chains for the MMA warp to
// rotate over -- with two, the C = 16 kernel's MMA warp waited 57 % of the time for epilogues (in-kernel counters).
constexpr int kGrpMaxStreams = 4;
constexpr int kGrpPadRows = 8;                             // slack rows on both sides (one 1024-byte swizzle atom)
constexpr int kGrpEpiWarps = 16;                           // NS * MS * 4
constexpr int kGrpMmaWarps = 2;                            // one MMA-issuing warp per stream
constexpr int kGrpThreads = 32 * (kGrpEpiWarps + 1 + kGrpMmaWarps);   // 608
constexpr int kGrpMaxStages = 12;                          // weight ring depth (runtime n_wstages <= this)
constexpr uint32_t kGrpSliceBytes = 64 * 32;               // one Toeplitz block: [64 rows (g', co)][16 ci] 16-bit
constexpr int kGrpSlicesPerStage = 4;
constexpr uint32_t kGrpStageBytes = kGrpSlicesPerStage * kGrpSliceBytes;   // 8 KB
constexpr int kGrpMaxPairs = 3;
constexpr int kGrpMaxChains = 3;
__host__ __device__ constexpr uint32_t grp_buf_bytes(int ms) { return (uint32_t)(ms * 128 + 2 * kGrpPadRows) * 128u; }   // multiple of 1024
__host__ __device__ constexpr int grp_tile_positions(int c, int ms) { return ms * 128 * (64 / c); }
constexpr uint32_t kGrpOnesBytes = 128 * 32;                               // A operand of the bias slice

struct GroupParams {
  const float* x32;         // block input, fp32 blocked [B][C/8][L][8]
  float* sum32;             // MRF running sum, fp32 blocked
  float* out32;             // EPI_OUT32
  void* out16;              // EPI_OUT16: lrelu(output), 16-bit [B][1][L][C]
  // One launch runs n_chains ResBlocks over every tile, one after the other (the three multi-receptive-field chains of a
  // stage read the same input tile: x is re-read from L2 and the running sum stays in L2 between them, so a stage moves
  // ~2.5 GB through DRAM instead of ~7.3 GB as three launches); n_chains = 1: a single ResBlock (or a part of one).
  // fuse_up: the stage's transposed conv (k = 4, stride 2; archi.py:80-81) runs inside this kernel, as one more
  // block-Toeplitz conv from a TMA-staged tile of the previous stage's 16-bit output straight into the residual columns of
  // tensor memory: the fp32 stage input never exists in HBM (no upsampler launch, no 4 B/element write + re-reads).
  CUtensorMap up_map;                // previous stage's lrelu'd 16-bit output as [B][rows][64], SWIZZLE_128B, box {64, 136}
  const void* up_w;                  // up_stages stages: the up_slices Toeplitz blocks of the transposed conv + its bias block
  int fuse_up, up_stages;
  int n_chains;
  const void* w[kGrpMaxChains];      // per chain: n_convs convs, each stages_per_conv stages of 8 KB (4 slices per stage: the
                                     // n_slices Toeplitz blocks, then the bias block, zero padded)
  int n_slices[kGrpMaxChains];       // (G + k - 1) * C / 16 Toeplitz slices (+ 1 bias slice)
  int stages_per_conv[kGrpMaxChains];// ceil((n_slices + 1) / 4)
  uint32_t flags[kGrpMaxChains];     // EPI_* of each chain's final epilogue (SUM_SET / SUM_ADD / SUM_FIN, OUT32, OUT16)
  int* error_flag;
  long long* timing;        // optional [16] cycle counters: MMA warp total / wait ready / wait weights
  int L;                    // positions per item (multiple of G)
  int n_convs;              // 2 * n_pairs
  int dil[kGrpMaxPairs];    // dilation of conv1 of each pair
  int halo;                 // recomputed positions per side (sum of all conv reaches), multiple of G
  int n_wstages;            // ring depth, >= stages_per_conv + 1
  int tiles_per_item, total_tiles;
  TileMapParams map;        // ragged batches (conv_tc.cuh); tile axis = valid positions per tile
  float slope_out;
  float n_blocks;
};

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

template <int C, bool BF16, int NS, int MS>
__global__ void __launch_bounds__(kGrpThreads, 1) group_chain_kernel(const __grid_constant__ GroupParams p) {
  static_assert(C == 16 || C == 32, "grouped formulation: C = 16 (G = 4) or C = 32 (G = 2)");
  static_assert(NS * MS * 4 == kGrpEpiWarps && NS <= kGrpMaxStreams, "NS * MS sub-tiles of 2 x 64 TMEM columns fill the 512 columns");
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
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
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
  // fuse_up: up_full[stream] (input tile landed), up_empty[stream] (buffer A free for the next tile), up_done[stream] (the
  // transposed conv's accumulator is complete: a barrier of its own, because the MMA warp issues it right behind the previous
  // chain's last conv2 -- two completions of acc_full ahead of the epilogue would alias its parity)
  auto bar_up_full = [&](int st) { return smem_u32(&bars[3 * kGrpMaxStreams + 2 * kGrpMaxStages + st]); };
  auto bar_up_empty = [&](int st) { return smem_u32(&bars[4 * kGrpMaxStreams + 2 * kGrpMaxStages + st]); };
  auto bar_up_done = [&](int st) { return smem_u32(&bars[5 * kGrpMaxStreams + 2 * kGrpMaxStages + st]); };
  auto bar_turn = [&](int w) { return smem_u32(&bars[6 * kGrpMaxStreams + 2 * kGrpMaxStages + w]); };   // issue turn of MMA warp w
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 7 * kGrpMaxStreams + 2 * kGrpMaxStages);

  const int valid = R - 2 * p.halo;
  __shared__ int tile_pre[kMaxMapItems + 1];
  tilemap_build(tile_pre, p.map, p.L, valid);                    // visible after the __syncthreads() below
  constexpr bool bf16 = BF16;                                    // both 16-bit flavours of the epilogue would double its code
  const int n_pairs = p.n_convs / 2;

  if (warp == kWarpW && lane == 0) {
    for (int st = 0; st < kGrpStreams; ++st) {
      mbar_init(bar_ready(st, 0), kGrpEpiWarps);                 // every epilogue warp serves every stream
      mbar_init(bar_ready(st, 1), kGrpEpiWarps);
      mbar_init(bar_acc_full(st), 1);
      mbar_init(bar_up_full(st), 1);
      mbar_init(bar_up_empty(st), 1);
      mbar_init(bar_up_done(st), 1);
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
    long long t_kind[4] = {0, 0, 0, 0};                                        // wait by kind: conv1 ready, conv2 ready, up_full, res_free
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
        if (timing) t_kind[2] += clock64() - tr0;
      } else {
        ok = mbar_wait(bar_ready(st, t_in), rdy_parity, p.error_flag);
        if (timing) t_kind[KIND] += clock64() - tr0;
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
      // fuse_up: the column halves swap roles from chain to chain (see the epilogue): the transposed conv of chain n writes
      // the half that was conv1's accumulator in chain n - 1 (drained long ago), never the residual still being read
      const uint32_t d_tmem = tmem_base + (uint32_t)((st * 2 + (d_idx ^ (int)(p.fuse_up ? (n & 1u) : 0u))) * kGrpMS * 64);
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
        umma_commit(KIND == 2 ? bar_up_done(st) : bar_acc_full(st));
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
      for (int i = 0; i < 4; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 10 + i), (unsigned long long)t_kind[i]);
    }
  } else {
    // ===== epilogue warps: one thread = HALF a 128-byte row (32 accumulator columns) of BOTH streams' tiles =====
    // Every epilogue step (TMEM -> leaky-ReLU -> 16-bit -> shared memory -> fence -> arrive) sits on the dependency chain
    // conv c -> conv c + 1 of its stream, and in-kernel counters put it at ~2000 cycles against 700-1400 cycles of MMAs per
    // conv and stream: the MMA warp waited for it.  With one thread per row, half of the epilogue warps idled while the
    // other stream's were the critical path; now all 16 warps drain each stream's accumulator (lane group lg = warp % 4 as
    // the hardware requires, sub-tile s, column half h) and alternate between the streams in the MMA warp's issue order:
    // a step touches 32 columns per thread -- one tcgen05.ld.x32, four 16-byte stores -- and twice as many warps hide the
    // latencies.  Register pressure halves with it (no spills at 96 registers).
    static_assert(NS == 2 && MS == 2, "epilogue mapping: 16 warps = 4 lane groups x 2 sub-tiles x 2 column halves, two streams");
    const int h = warp >> 3;                                     // column half: columns [32 h, 32 h + 32) = row bytes [64 h, 64 h + 64)
    const int s = (warp >> 2) & 1;
    const int lg = warp & 3;
    const int r = s * 128 + lg * 32 + lane;                      // row within a stream tile
    constexpr int cchunks = C / 8;
    constexpr int PP = G / 2;                                    // positions per half row
    // TMEM address of this thread's 32 columns in column set idx (0 / 1) of stream st
    auto t_cols = [&](int st, int idx) {
      return tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(((st * 2 + idx) * kGrpMS + s) * 64 + 32 * h);
    };
    uint32_t n_chain = 0;                                        // chains this thread has started (both streams)
    const float inv_n = 1.0f / p.n_blocks;
    auto swz128 = [](uint32_t lin) { return lin ^ (((lin >> 7) & 7u) << 4); };
    // tile-invariant d-major maps of this thread's PP positions g = PP h + i, for every pair with a dilated conv1:
    //   physP(m, i) = swizzled byte offset of time (G r + g) in the d_m-major input tile of pair m's conv1
    //   tauI(m, i)  = time whose conv1 output this row holds at position g in pair m (>= R: none)
    uint32_t pmap[kGrpMaxPairs][PP];
#pragma unroll
    for (int m = 0; m < kGrpMaxPairs; ++m) {
      const int d = (m < n_pairs) ? p.dil[m] : 1;
      const int Q = (R + d - 1) / d;
#pragma unroll
      for (int i = 0; i < PP; ++i) {
        const int tau = G * r + PP * h + i;
        const int pp = (d == 1) ? tau : (tau % d) * Q + tau / d;
        const int ti = (d == 1) ? tau : d * (tau % Q) + tau / Q;
        pmap[m][i] = swz128(kPadBytes + (uint32_t)pp * PB) | ((uint32_t)ti << 16);
      }
    }
    auto physP = [&](int m, int i) { return pmap[m][i] & 0xFFFFu; };
    auto tauI = [&](int m, int i) { return (int)(pmap[m][i] >> 16); };
    const uint32_t lin_row = kPadBytes + (uint32_t)r * 128u;     // this row in natural order
    const uint32_t xrow = (uint32_t)(r & 7) << 4;                // its swizzle XOR (kPadBytes is a multiple of 1024)
    // Local 16-column group u (0 / 1) of this thread = group gi = 2 h + u of the row = chunks 2 gi and 2 gi + 1.
    auto st_nat = [&](uint8_t* bp, int u, const uint4& lo, const uint4& hi8) {
      *reinterpret_cast<uint4*>(bp + lin_row + (((uint32_t)(4 * h + 2 * u) * 16u) ^ xrow)) = lo;
      *reinterpret_cast<uint4*>(bp + lin_row + (((uint32_t)(4 * h + 2 * u + 1) * 16u) ^ xrow)) = hi8;
    };
    // local position index of group u, and its 16-channel block inside the position
    auto pos_of = [&](int u) { return (kGroupsPerPos == 1) ? u : 0; };
    auto blk_of = [&](int u) { return (kGroupsPerPos == 1) ? 0 : u; };
    // group u at a position whose chunk 0 lives at the swizzled offset phys0: its chunk cc is at phys0 ^ 16 cc
    auto st_at = [&](uint8_t* bp, uint32_t phys0, int u, const uint4& lo, const uint4& hi8) {
      const uint32_t o = phys0 ^ ((uint32_t)blk_of(u) * 32u);
      *reinterpret_cast<uint4*>(bp + o) = lo;
      *reinterpret_cast<uint4*>(bp + (o ^ 16u)) = hi8;
    };
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    auto pack16 = [&](const auto& rr, int o, uint4& lo, uint4& hi8, float slope) {       // 16 columns from offset o of rr
      float a[8], c8[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { a[e] = __uint_as_float(rr[o + e]); c8[e] = __uint_as_float(rr[o + 8 + e]); }
      lo = pack8_lrelu(a, slope, true, bf16);
      hi8 = pack8_lrelu(c8, slope, true, bf16);
    };
    // TMEM -> registers, this thread's 32 columns in one load: f(u, regs, offset) for its two 16-column groups.
    // (tcgen05.ld / wait are .sync.aligned: the whole warp must execute them converged, so every caller branches on
    // warp-uniform conditions only; f may diverge inside.)
    auto for_groups = [&](uint32_t taddr, auto&& f) {
      uint32_t rr[32];
      __syncwarp();
      tmem_ld32(taddr, rr);
      tmem_ld_wait();
      f(0, rr, 0);
      f(1, rr, 16);
      __syncwarp();
    };
    bool ok = true;
    uint32_t aph = 0;                                            // completed phases of the acc_full barriers (same for both streams)
#ifdef SA_DIAG
    const bool timing = p.timing != nullptr && warp == 0;
    long long t_x = 0, t_acc = 0, t_fin = 0, t_begin = timing ? clock64() : 0;
#define GRP_T0(v) const long long v = timing ? clock64() : 0
#define GRP_ADD(acc, v) if (timing) acc += clock64() - v
#else
#define GRP_T0(v)
#define GRP_ADD(acc, v)
#endif
    // per-stream tile state, packed: flag bits
    constexpr uint32_t F_LIVE = 1u, F_INSIDE = 2u, F_KEEP = 4u, F_INTERIOR = 8u, F_WALL = 16u, F_WANY = 32u;
    for (int it = 0; it < n_iters && ok; ++it) {
      int tb[NS], tt0[NS];                                        // item and time of the first position of each stream's tile
      uint32_t tf[NS];
#pragma unroll
      for (int st = 0; st < NS; ++st) {
        const int tile = (int)blockIdx.x + (NS * it + st) * (int)gridDim.x;
        const bool live = tile < n_live;                          // the last iteration may have streams without a tile
        int b = 0, mt = 0;
        if (live) tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile, b, mt);
        const int t0 = mt * valid - p.halo;
        const int t_row = t0 + G * r;
        const bool inside = live && t_row >= 0 && t_row < p.L;    // L, halo, valid are multiples of G: whole rows
        const bool keep = inside && G * r >= p.halo && G * r < R - p.halo;
        const bool interior = live && t0 >= 0 && t0 + R <= p.L;   // no position of the tile is outside the utterance
        const bool w_all = __all_sync(0xffffffffu, inside), w_any = __any_sync(0xffffffffu, inside);   // warp-uniform
        tb[st] = b; tt0[st] = t0;
        tf[st] = (live ? F_LIVE : 0u) | (inside ? F_INSIDE : 0u) | (keep ? F_KEEP : 0u) | (interior ? F_INTERIOR : 0u) |
                 (w_all ? F_WALL : 0u) | (w_any ? F_WANY : 0u);
        // The next tile of this stream: have its rows in L2 by the time they are needed (fp32 stage input only).
        if (!p.fuse_up) {
          const int tile_n = tile + NS * (int)gridDim.x;
          if (tile_n < n_live) {
            int bn, mtn;
            tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile_n, bn, mtn);
            const int tn = mtn * valid - p.halo + G * r + PP * h;
            if (tn >= 0 && tn < p.L) {
#pragma unroll
              for (int q = 0; q < cchunks; ++q)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x32 + (((size_t)bn * cchunks + q) * (size_t)p.L + (size_t)tn) * 8));
            }
          }
        }
      }
#pragma unroll 1
      for (int j = 0; j < p.n_chains && ok; ++j, ++n_chain) {
      const uint32_t flags = p.flags[j];
      // fuse_up: the two 64-column sets of a sub-tile swap roles from chain to chain -- the next tile's transposed conv
      // lands in the columns that held conv1's accumulator while the final epilogue still reads the old residual, so the
      // MMA warp never waits for the global-memory part of a tile's last epilogue.
      const int i_acc = (p.fuse_up && (n_chain & 1u)) ? 1 : 0, i_res = i_acc ^ 1;
      const bool add = (flags & (EPI_SUM_ADD | EPI_SUM_FIN)) != 0;
      // The running multi-receptive-field sum lives in a layout private to the grouped kernels: 32-row blocks, inside a
      // block [piece 0..7][row][8 floats], so that lane l of a warp store writes 32 bytes right behind lane l - 1 (a
      // thread-per-row store into the channel-blocked [C/8][L][8] layout put the 32 lanes into 32 different lines:
      // 32 LSU wavefronts per instruction).  This thread owns pieces 4 h .. 4 h + 3 of its row.
      auto sum_ptr = [&](int st) {
        const int grow = tb[st] * (p.L / G) + (tt0[st] + G * r) / G;   // row in the whole batch (used under `keep` only)
        return p.sum32 + (size_t)(grow >> 5) * 2048 + (size_t)(grow & 31) * 8 + (size_t)(4 * h) * 256;
      };
      // ---- x: residual stream -> tensor memory, lrelu(x) -> input tile of pair 0's conv1 ----
#pragma unroll
      for (int st = 0; st < NS; ++st) {
        if (!ok) break;
        GRP_T0(tx0);
        uint8_t* const bufA = buf(st, 0);
        const bool inside = (tf[st] & F_INSIDE) != 0;
        const int d0 = p.dil[0];
        if (p.fuse_up) {
          // x = the stage's transposed conv, computed by the MMA warp from the staged input tile into the residual
          // columns: read it back once to stage lrelu(x); rows outside the utterance are the zero padding of conv1
          ok = mbar_wait_relaxed(bar_up_done(st), n_chain & 1u, p.error_flag);
          tc_fence_after();
          if (!ok) break;
          if (tf[st] & F_WANY) {
            for_groups(t_cols(st, i_res), [&](int u, const uint32_t (&rr)[32], int o) {
              uint4 lo, hi8;
              pack16(rr, o, lo, hi8, 0.1f);
              if (!inside) { lo = zero4; hi8 = zero4; }
              if (d0 == 1) st_nat(bufA, u, lo, hi8); else st_at(bufA, physP(0, pos_of(u)), u, lo, hi8);
            });
          } else {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (d0 == 1) st_nat(bufA, u, zero4, zero4); else st_at(bufA, physP(0, pos_of(u)), u, zero4, zero4);
            }
          }
        } else {
          const int t_pos = tt0[st] + G * r + PP * h;             // first position of this half row
          float4 xq[2][4];                                        // all loads in flight before the first use
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            xq[u][0] = xq[u][1] = xq[u][2] = xq[u][3] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inside) {
              const float* src = p.x32 + (((size_t)tb[st] * cchunks + 2 * blk_of(u)) * (size_t)p.L + (size_t)(t_pos + pos_of(u))) * 8;
              ldg_f8(src, xq[u][0], xq[u][1]);
              ldg_f8(src + (size_t)p.L * 8, xq[u][2], xq[u][3]);
            }
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 q0 = xq[u][0], q1 = xq[u][1], q2 = xq[u][2], q3 = xq[u][3];
            const uint32_t rr[16] = {__float_as_uint(q0.x), __float_as_uint(q0.y), __float_as_uint(q0.z), __float_as_uint(q0.w),
                                     __float_as_uint(q1.x), __float_as_uint(q1.y), __float_as_uint(q1.z), __float_as_uint(q1.w),
                                     __float_as_uint(q2.x), __float_as_uint(q2.y), __float_as_uint(q2.z), __float_as_uint(q2.w),
                                     __float_as_uint(q3.x), __float_as_uint(q3.y), __float_as_uint(q3.z), __float_as_uint(q3.w)};
            __syncwarp();
            tmem_st16(t_cols(st, i_res) + (uint32_t)(u * 16), rr);
            uint4 lo, hi8;
            pack16(rr, 0, lo, hi8, 0.1f);                         // rows outside the utterance were loaded as zeros
            if (d0 == 1) st_nat(bufA, u, lo, hi8); else st_at(bufA, physP(0, pos_of(u)), u, lo, hi8);
          }
          tmem_st_wait();
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready(st, 0));
        GRP_ADD(t_x, tx0);
        // a running sum left in HBM by an earlier launch: have it in L2 when the final epilogue needs it
        if (p.n_chains == 1 && (tf[st] & F_KEEP) && add) {
          const float* sp = sum_ptr(st);
#pragma unroll
          for (int i = 0; i < 4; ++i) asm volatile("prefetch.global.L2 [%0];" ::"l"(sp + i * 256));
        }
      }
      // ---- the (conv1, conv2) pairs ----
#pragma unroll
      for (int m = 0; m < kGrpMaxPairs; ++m) {
        if (m < n_pairs && ok) {
          const int d = p.dil[m];
          // conv1: TMEM -> lrelu -> conv2's input tile in natural order
#pragma unroll
          for (int st = 0; st < NS; ++st) {
            if (!ok) break;
            uint8_t* const bufT = buf(st, 1);
            const uint32_t f = tf[st];
            const bool inside = (f & F_INSIDE) != 0;
            {
              GRP_T0(ta0);
              ok = mbar_wait_relaxed(bar_acc_full(st), aph & 1u, p.error_flag);
              tc_fence_after();
              GRP_ADD(t_acc, ta0);
            }
            if (!ok) break;
            const uint32_t t_acc1 = t_cols(st, i_acc);
            if (d == 1) {
              if (f & F_WALL) {
                for_groups(t_acc1, [&](int u, const uint32_t (&rr)[32], int o) {
                  uint4 lo, hi8;
                  pack16(rr, o, lo, hi8, 0.1f);
                  st_nat(bufT, u, lo, hi8);
                });
              } else if (!(f & F_WANY)) {
#pragma unroll
                for (int u = 0; u < 2; ++u) st_nat(bufT, u, zero4, zero4);
              } else {                                            // the utterance ends inside this warp's rows
                for_groups(t_acc1, [&](int u, const uint32_t (&rr)[32], int o) {
                  uint4 lo, hi8;
                  pack16(rr, o, lo, hi8, 0.1f);
                  st_nat(bufT, u, inside ? lo : zero4, inside ? hi8 : zero4);
                });
              }
            } else if (f & F_INTERIOR) {
              for_groups(t_acc1, [&](int u, const uint32_t (&rr)[32], int o) {
                uint4 lo, hi8;
                pack16(rr, o, lo, hi8, 0.1f);
                const int tau = tauI(m, pos_of(u));
                if (tau < R) st_at(bufT, swz128(kPadBytes + (uint32_t)tau * PB), u, lo, hi8);
              });
            } else {
              const bool live = (f & F_LIVE) != 0;
              const int t0 = tt0[st];
              for_groups(t_acc1, [&](int u, const uint32_t (&rr)[32], int o) {
                uint4 lo, hi8;
                pack16(rr, o, lo, hi8, 0.1f);
                const int tau = tauI(m, pos_of(u));
                const int tt = t0 + tau;
                const bool ins = live && tt >= 0 && tt < p.L;
                if (tau < R) st_at(bufT, swz128(kPadBytes + (uint32_t)tau * PB), u, ins ? lo : zero4, ins ? hi8 : zero4);
              });
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready(st, 1));
          }
          ++aph;
          // conv2: the accumulator IS the residual stream x_{m+1}
#pragma unroll
          for (int st = 0; st < NS; ++st) {
            if (!ok) break;
            uint8_t* const bufA = buf(st, 0);
            const uint32_t f = tf[st];
            const bool inside = (f & F_INSIDE) != 0, keep = (f & F_KEEP) != 0;
            const uint32_t t_res = t_cols(st, i_res);
            if (m + 1 < n_pairs) {
              {
                GRP_T0(ta0);
                ok = mbar_wait_relaxed(bar_acc_full(st), aph & 1u, p.error_flag);
                tc_fence_after();
                GRP_ADD(t_acc, ta0);
              }
              if (!ok) break;
              const int mn = m + 1 < kGrpMaxPairs ? m + 1 : 0;
              const int dn = p.dil[mn];
              if (f & F_WALL) {
                for_groups(t_res, [&](int u, const uint32_t (&rr)[32], int o) {
                  uint4 lo, hi8;
                  pack16(rr, o, lo, hi8, 0.1f);
                  if (dn == 1) st_nat(bufA, u, lo, hi8); else st_at(bufA, physP(mn, pos_of(u)), u, lo, hi8);
                });
              } else if (!(f & F_WANY)) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  if (dn == 1) st_nat(bufA, u, zero4, zero4); else st_at(bufA, physP(mn, pos_of(u)), u, zero4, zero4);
                }
              } else {
                for_groups(t_res, [&](int u, const uint32_t (&rr)[32], int o) {
                  uint4 lo, hi8;
                  pack16(rr, o, lo, hi8, 0.1f);
                  if (!inside) { lo = zero4; hi8 = zero4; }
                  if (dn == 1) st_nat(bufA, u, lo, hi8); else st_at(bufA, physP(mn, pos_of(u)), u, lo, hi8);
                });
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_ready(st, 0));
            } else {
              // final epilogue: multi-receptive-field combine (archi.py:82-86) + output streams.  The 32 floats of the
              // running sum are requested before the wait for conv2.  The 16-bit output row ([B][L][C]: 128 contiguous
              // bytes per row, 64 per thread) is transposed inside lane quads, so one store instruction writes 64
              // contiguous bytes of 8 rows instead of 16 bytes into each of 32 lines.
              float* const sp = sum_ptr(st);
              float4 sq[8];
              if (keep && add) {
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) ldg_f8(sp + pc * 256, sq[2 * pc], sq[2 * pc + 1]);
              }
              __syncwarp();
              {
                GRP_T0(ta0);
                ok = mbar_wait_relaxed(bar_acc_full(st), aph & 1u, p.error_flag);
                tc_fence_after();
                GRP_ADD(t_acc, ta0);
              }
              if (!ok) break;
              GRP_T0(tf0);
              const bool o16 = (flags & EPI_OUT16) != 0;
              const int b = tb[st], t0 = tt0[st];
              uint4 pk[4];
              for_groups(t_res, [&](int u, const uint32_t (&rr)[32], int o) {
                float lo[8], hi8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { lo[e] = __uint_as_float(rr[o + e]); hi8[e] = __uint_as_float(rr[o + 8 + e]); }
                if (add) {
                  const float4 s0 = sq[4 * u + 0], s1 = sq[4 * u + 1], s2 = sq[4 * u + 2], s3 = sq[4 * u + 3];
                  const float sl[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                  const float sh[8] = {s2.x, s2.y, s2.z, s2.w, s3.x, s3.y, s3.z, s3.w};
#pragma unroll
                  for (int e = 0; e < 8; ++e) { lo[e] = lo[e] + sl[e]; hi8[e] = hi8[e] + sh[e]; }
                  if (flags & EPI_SUM_FIN) {
                    // x / n as q = x * (1/n) plus one residual correction (q + (x - n q) * (1/n)): the correctly rounded
                    // quotient for normal operands without the division's slow-path calls
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                      const float ql = lo[e] * inv_n, qh = hi8[e] * inv_n;
                      lo[e] = fmaf(fmaf(-p.n_blocks, ql, lo[e]), inv_n, ql);
                      hi8[e] = fmaf(fmaf(-p.n_blocks, qh, hi8[e]), inv_n, qh);
                    }
                  }
                }
                if (keep) {
                  if (flags & (EPI_SUM_SET | EPI_SUM_ADD)) { stg_f8(sp + (2 * u) * 256, lo); stg_f8(sp + (2 * u + 1) * 256, hi8); }
                  if (flags & EPI_OUT32) {                       // debug taps / fp32 hand-off: the channel-blocked layout
                    const size_t i0 = (((size_t)b * cchunks + 2 * blk_of(u)) * (size_t)p.L + (size_t)(t0 + G * r + PP * h + pos_of(u))) * 8;
                    stg_f8(p.out32 + i0, lo);
                    stg_f8(p.out32 + i0 + (size_t)p.L * 8, hi8);
                  }
                }
                if (o16) { pk[2 * u] = pack8_lrelu(lo, p.slope_out, true, bf16); pk[2 * u + 1] = pack8_lrelu(hi8, p.slope_out, true, bf16); }
              });
              if (o16) {                                         // warp-uniform
                const int j4 = lane & 3;
                quad_transpose(pk, j4);                          // now pk[k] = 16-byte piece j4 of the half row of the quad's lane k
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int rk = r - j4 + k;
                  const int tk = t0 + G * rk;
                  if ((f & F_LIVE) && tk >= 0 && tk < p.L && G * rk >= p.halo && G * rk < R - p.halo)
                    *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out16) + ((size_t)b * (size_t)p.L + (size_t)tk) * (C * 2) + 64 * h + 16 * j4) = pk[k];
                }
              }
              tc_fence_before();                                 // TMEM reads done before the columns are written again
              GRP_ADD(t_fin, tf0);
            }
          }
          ++aph;
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
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 8), (unsigned long long)t_fin);
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

}  // namespace tc
}  // namespace sa
