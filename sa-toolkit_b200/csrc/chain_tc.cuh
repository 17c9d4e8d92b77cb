// Fused ResBlock1 kernel for the narrow stages (C = 64, 32, 16) on tcgen05.
//
// One CTA keeps a time tile of MS*128 rows (halo included) in shared memory and runs all six
// convs of a ResBlock1 (nn.py:168-175) on it without touching HBM in between:
//     x -> [ lrelu -> conv1(k, d) -> lrelu -> conv2(k, 1) -> + x ] for d in (1, 3, 5)
// and folds the multi-receptive-field combine (archi.py:82-86) into the last epilogue.
//
//   bufA  [PAD + MS*128 + PAD rows][C] 16-bit, UMMA K-major swizzled (row = 2C bytes: SWIZZLE_128B /
//         64B / 32B for C = 64 / 32 / 16)         lrelu(x)            operand of conv1
//   bufT  same shape                              lrelu(conv1 + b)    operand of conv2
//   x     fp32 residual stream, in REGISTERS of the epilogue threads (one thread owns one row)
//   D     fp32 accumulators in TMEM, two buffers per sub-tile (conv parity)
//
// Same implicit-GEMM formulation as conv_tc.cuh: time on M, Cout on N, the staged tile is the
// A operand and tap j is the tile with the descriptor start advanced by (j*d - pad) rows.
// Rows computed from data beyond the tile are garbage by construction; the halo
// H = sum of all conv reaches (12 / 36 / 60 rows for k = 3 / 7 / 11) is discarded at the end,
// so a tile yields MS*128 - 2H valid rows.  Rows outside the utterance [0, L) are forced to
// zero in every staged activation: that is the per-layer zero padding of the reference.
//
// Schedule inside a CTA (sub-tile major): MMA(c, s) needs the activations of sub-tiles s-1..s+1
// written by the epilogue of conv c-1, so the tensor pipe works on sub-tile s+1.. while the
// eight epilogue warps drain sub-tile s.  Weights of one conv sit in a ring (slot = stage) and
// are released on the last sub-tile, which lets the next conv's weights stream in behind.
//
// Warp roles (32*WPS*MS + 64 threads): first WPS (4 or 8) epilogue warps per sub-tile, then the weight
// producer, and LAST the MMA issuer + TMEM owner (the scheduler prefers the highest eligible warp id, and
// the waiting epilogue warps back off with nanosleep, so the issuer is never starved of issue slots);
// epilogue warps: WPS per sub-tile (TMEM lane group = warp % 4; with 8 warps the second
// quad takes the upper half of the channels): ONE THREAD OWNS ONE ROW (or half of it) of the tile
// for the whole ResBlock, so the residual stream is a fixed array in its registers, there is no
// loop over sub-tiles (small code, every sub-tile epilogue runs concurrently) and each sub-tile
// talks to the MMA warp through its own acc_full / ready barriers.  (Earlier mappings -- all 8
// warps on one sub-tile, or two groups on alternating sub-tiles -- were 2-4x slower: ~250
// instructions per 256 elements and 32 KB of unrolled code, profiles/README.md.)
#pragma once
#include "conv_tc.cuh"
#include <type_traits>

namespace sa {
namespace tc {

__host__ __device__ constexpr int chain_threads(int ms, int wps) { return 64 + 32 * wps * ms; }
constexpr int kChainPad = 32;            // slack rows on both sides of the staged tile (>= max tap reach 25)
constexpr int kChainMaxConvs = 8;
constexpr int kChainMaxSlots = 16;
constexpr int kChainTapsPerStage64 = 2;

struct ChainParams {
  const float* x32;         // stage input h, fp32 blocked [B][C/8][L][8]
  float* sum32;             // MRF running sum, fp32 blocked
  float* out32;             // stage output (EPI_OUT32)
  void* out16;              // lrelu(stage output), 16-bit blocked (EPI_OUT16)
  const void* w;            // weights of the n_convs convs, conv-major, each [tap][C rows][C] pre-swizzled
  const float* bias;        // [n_convs][C]
  int* error_flag;
  int L;                    // rows per item
  int n_convs;              // 2 * n_dilations
  int ktaps;
  int dil[kChainMaxConvs];
  int pad[kChainMaxConvs];
  int halo;                 // H
  int tiles_per_item, total_tiles;
  TileMapParams map;        // ragged batches: live-tile enumeration (conv_tc.cuh); tile axis = valid rows per tile
  int k16_per_stage, stages_per_conv, n_slots;
  long long* timing;        // optional [16] cycle counters (diagnostics): MMA warp: total, wait ready, wait weights,
                            // issue; epilogue warp 2: total, x load + staging, wait accumulator, work
  uint32_t flags;           // EPI_* of the final epilogue (SUM_SET / SUM_ADD / SUM_FIN / OUT32 / OUT16 / BF16)
  float slope_out;
  float n_blocks;
};

// K = filter taps, WPS = epilogue warps per sub-tile.  Weight ring: C = 64 streams one tap ([64 rows][64],
// 8 KB) per stage through p.n_slots >= K slots (a ring deeper than one conv lets the first taps of the
// next conv land before the current conv's last sub-tile is done); C <= 32 holds a whole conv per stage
// in 2 slots.
// RT ("residual in tensor memory", round 2): conv2 accumulates on top of the residual stream, which lives in the TMEM buffer
// conv2 uses (the epilogue adds the running sum of the conv2 biases when it reads it), so the epilogue threads hold no
// residual registers and do no residual add, and the tile boundary is pipelined: the next tile's x is loaded and staged
// (shared memory + the TMEM buffer that held conv1's accumulator: the two buffers swap roles from tile to tile) while the
// last conv2 of the running tile executes, and the last epilogue hands its TMEM buffer back (res_free) as soon as the row is
// in registers.  The MMA warp then walks from one tile into the next without waiting for global memory.  fp32 sums are
// associated differently from the per-layer path (last-bit differences), so RT = false stays the bit-identical reference
// form (SATOOLS_B200_GROUP=0) and the form of the C <= 32 instantiations.
template <int C, int MS, int K, int WPS, bool RT = false>
__global__ void __launch_bounds__(chain_threads(MS, WPS), (C <= 32 && MS <= 3) ? 2 : 1)
resblock_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int N = C;
  constexpr int R = MS * 128;
  constexpr int ROWS = R + 2 * kChainPad;
  constexpr uint32_t RB = 2u * C;                                // row bytes: 128 / 64 / 32
  constexpr uint32_t kBufBytes = ROWS * RB;
  // TM (tap-major, C = 64): for tap { for sub-tile { MMAs } }: a tap's weights serve every sub-tile and are
  // released at once, so a short ring streams them (measured with the sub-tile-major order: 14% of the MMA
  // warp's time waiting for weights, 22% for activations); one ready / acc_full barrier per conv.
  constexpr bool TM = false;     // measured slower on B200 (stage 2: 7.7 -> 9.1 ms: the epilogue no longer overlaps the MMAs)
  constexpr int kThreads_ = chain_threads(MS, WPS);
  constexpr int kCPT = C / 8 / (WPS / 4);                        // 8-channel chunks per epilogue thread
  static_assert(WPS == 4 || WPS == 8, "4 or 8 epilogue warps per sub-tile");
  static_assert(kCPT >= 2 && kCPT % 2 == 0, "a thread handles pairs of channel chunks");
  constexpr uint32_t kTmemNeed = 2u * MS * N;
  constexpr uint32_t kTmemCols = kTmemNeed <= 32 ? 32 : kTmemNeed <= 64 ? 64 : kTmemNeed <= 128 ? 128 : kTmemNeed <= 256 ? 256 : 512;
  static_assert(kTmemNeed <= 512, "accumulators do not fit in TMEM");

  // Weight stages: C = 64 streams kChainTapsPerStage64 taps (16 KB) per stage -- every stage costs the issuing thread a
  // try_wait on sub-tile 0 and a commit on the last sub-tile (~90 + ~50 cycles that are not hidden behind the queued MMAs,
  // tools/mma_bench5.cu), so fewer, larger stages; C <= 32 holds a whole conv per stage.
  constexpr int TPS = (C == 64) ? kChainTapsPerStage64 : K;      // taps per stage
  constexpr int SPC = (K + TPS - 1) / TPS;                       // weight stages per conv
  const int NSLOTS = (C == 64) ? p.n_slots : 2;
  constexpr int K16 = C / 16;                                    // K=16 steps per tap
  constexpr uint32_t kTapBytes = (uint32_t)N * RB;               // one tap = one [N rows][C] weight block
  constexpr uint32_t stage_bytes = (uint32_t)TPS * kTapBytes;    // slot pitch (the last stage of a conv may hold fewer taps)
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the role loops (barrier addresses,
  // descriptors, counters) in uniform registers instead of converting them per use (R2UR), as CUTLASS' canonical_warp_idx_sync
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int kWarpW = WPS * MS, kWarpMma = WPS * MS + 1;      // epilogue warps are 0 .. WPS*MS-1
  uint8_t* bufA = smem;
  uint8_t* bufT = smem + kBufBytes;
  uint8_t* w_smem = smem + 2 * kBufBytes;
  float* bias_s = reinterpret_cast<float*>(w_smem + (size_t)NSLOTS * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + kChainMaxConvs * C);
  // barrier slots: ready[2][8] acc_full[2][8] w_full[16] w_empty[16]
  auto bar_ready = [&](int buf, int s) { return smem_u32(&bars[buf * 8 + s]); };
  auto bar_acc_full = [&](int par, int s) { return smem_u32(&bars[16 + par * 8 + s]); };
  auto bar_w_full = [&](int i) { return smem_u32(&bars[32 + i]); };
  auto bar_w_empty = [&](int i) { return smem_u32(&bars[32 + kChainMaxSlots + i]); };
  auto bar_res_free = [&](int s) { return smem_u32(&bars[32 + 2 * kChainMaxSlots + s]); };   // RT: the last epilogue has read sub-tile s
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 40 + 2 * kChainMaxSlots);

  const int valid_rows = R - 2 * p.halo;
  __shared__ int tile_pre[kMaxMapItems + 1];
  tilemap_build(tile_pre, p.map, p.L, valid_rows);                // visible after the __syncthreads() below
  const bool bf16 = (p.flags & EPI_BF16) != 0;

  if (warp == kWarpW && lane == 0) {
    for (int s = 0; s < 8; ++s) {
      mbar_init(bar_ready(0, s), TM ? WPS * MS : WPS);           // the warps owning the sub-tile (TM: the whole tile)
      mbar_init(bar_ready(1, s), TM ? WPS * MS : WPS);
      mbar_init(bar_acc_full(0, s), 1); mbar_init(bar_acc_full(1, s), 1);
      mbar_init(bar_res_free(s), WPS);
    }
    for (int i = 0; i < kChainMaxSlots; ++i) { mbar_init(bar_w_full(i), 1); mbar_init(bar_w_empty(i), 1); }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(smem_u32(tmem_holder), kTmemCols);
  if constexpr (RT) {
    // conv1: its own bias; conv2 of pair m: b2_0 + .. + b2_m (the TMEM residual accumulates the conv2 outputs without them)
    for (int i = threadIdx.x; i < C; i += kThreads_) {
      float run = 0.f;
      for (int c = 0; c < p.n_convs; ++c) {
        const float bv = p.bias[c * C + i];
        if (c & 1) { run += bv; bias_s[c * C + i] = run; } else { bias_s[c * C + i] = bv; }
      }
    }
  } else {
    for (int i = threadIdx.x; i < p.n_convs * C; i += kThreads_) bias_s[i] = p.bias[i];
  }
  // zero both staged tiles once: the PAD slack rows are never written afterwards
  for (uint32_t i = threadIdx.x; i < 2 * kBufBytes / 16; i += kThreads_)
    *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int n_live = tilemap_total(tile_pre, p.map, p.total_tiles);

  if (warp == kWarpW) {
    // ===== weight producer =====
    {
      const bool leader = elect_one();
      int slot = 0;
      uint32_t par = 1;                                          // parity of the previous use of `slot`
      bool wrapped = false, ok = true;
      for (int tile = blockIdx.x; tile < n_live && ok; tile += gridDim.x)
        for (int c = 0; c < p.n_convs && ok; ++c) {
          const uint8_t* src = static_cast<const uint8_t*>(p.w) + (size_t)c * K * kTapBytes;
#pragma unroll 1
          for (int i = 0; i < SPC; ++i) {
            if (wrapped) ok = mbar_wait_relaxed(bar_w_empty(slot), par, p.error_flag);
            if (!ok) break;
            const uint32_t bytes = (uint32_t)min(TPS, K - i * TPS) * kTapBytes;
            if (leader) {
              mbar_arrive_expect_tx(bar_w_full(slot), bytes);
              bulk_load(smem_u32(w_smem) + (uint32_t)slot * stage_bytes, src + (size_t)i * stage_bytes, bytes, bar_w_full(slot));
            }
            __syncwarp();
            if (++slot == NSLOTS) { slot = 0; par ^= 1u; wrapped = true; }
          }
        }
    }
  } else if (warp == kWarpMma) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    // Straight-line code per (conv, sub-tile): K * K16 MMAs whose descriptors differ by constant
    // adds; the only waits are the activation-ready barrier and, on sub-tile 0, the weight stages.
    {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc(N, bf16);
      constexpr uint32_t hi = ((8u * RB) >> 4) | (1u << 14) | ((RB == 128 ? 2u : RB == 64 ? 4u : 6u) << 29);
      constexpr uint32_t row16 = RB >> 4;
      constexpr uint32_t tap16 = kTapBytes >> 4;
      const uint32_t b_lo0 = desc_lo(smem_u32(w_smem));
      uint32_t it = 0;
      int slot0 = 0;                                             // ring position of the running conv's first stage
      uint32_t par0 = 0;
      bool ok = true;
#ifdef SA_DIAG
      const bool timing = p.timing != nullptr;
#else
      constexpr bool timing = false;   // in-kernel cycle counters: -DSA_DIAG builds only
#endif
      long long t_ready = 0, t_w = 0, t_begin = timing ? clock64() : 0;
      long long t_issue = 0, n_issue = 0;      // SA_DIAG: cycles inside the MMA issue block of sub-tiles > 0 (no waits inside)
      for (int tile = blockIdx.x; tile < n_live && ok; tile += gridDim.x, ++it) {
        for (int c = 0; c < p.n_convs && ok; ++c) {
          const uint32_t in_lo0 = desc_lo(smem_u32((c & 1) ? bufT : bufA)) + (uint32_t)(kChainPad - p.pad[c]) * row16;
          const uint32_t rdy_parity = (it * (uint32_t)(p.n_convs / 2) + (uint32_t)(c / 2)) & 1u;
          const uint32_t dil16 = (uint32_t)p.dil[c] * row16;
          int slot_end = slot0;
          uint32_t par_end = par0;
          if (TM) {
            const long long tr0 = timing ? clock64() : 0;
            ok = mbar_wait(bar_ready(c & 1, 0), rdy_parity, p.error_flag);     // every sub-tile staged
            if (timing) t_ready += clock64() - tr0;
            if (!ok) break;
            tc_fence_after();
            const uint32_t d_tmem0 = tmem_base + (uint32_t)((c & 1) * MS * N);
            uint32_t a_tap = in_lo0;
            int slot = slot0;
            uint32_t par = par0;
#pragma unroll
            for (int tap = 0; tap < K; ++tap) {
              const long long tw0 = timing ? clock64() : 0;
              ok = ok && mbar_wait(bar_w_full(slot), par, p.error_flag);
              tc_fence_after();
              if (timing) t_w += clock64() - tw0;
              const uint32_t b_tap = b_lo0 + (uint32_t)slot * (stage_bytes >> 4);
#pragma unroll
              for (int s = 0; s < MS; ++s)
#pragma unroll
                for (int kk = 0; kk < K16; ++kk)
                  if (leader) umma_f16(d_tmem0 + (uint32_t)(s * N), desc64(a_tap + (uint32_t)(s * 128) * row16 + 2u * kk, hi),
                                       desc64(b_tap + 2u * kk, hi), idesc, (tap | kk) ? 1u : 0u);
              if (leader) umma_commit(bar_w_empty(slot));        // this tap's weights are consumed
              __syncwarp();
              a_tap += dil16;
              if (++slot == NSLOTS) { slot = 0; par ^= 1u; }
            }
            if (leader) umma_commit(bar_acc_full(c & 1, 0));
            __syncwarp();
            slot0 = slot; par0 = par;
            continue;
          }
#pragma unroll
          for (int s = 0; s < MS; ++s) {
            // inputs of sub-tiles s-1..s+1 must be staged; s-1 and s were confirmed in earlier iterations
            // (an already-complete try_wait still costs ~90 cycles on this single issuing warp)
            const long long tr0 = timing ? clock64() : 0;
#ifdef SA_DIAG
            const bool freerun = (p.flags & (1u << 30)) != 0;    // diagnostics: do not wait for the epilogue (results invalid)
#else
            constexpr bool freerun = false;
#endif
            if (s == 0 && ok && !freerun) ok = mbar_wait(bar_ready(c & 1, 0), rdy_parity, p.error_flag);
            if (s + 1 < MS && ok && !freerun) ok = mbar_wait(bar_ready(c & 1, s + 1), rdy_parity, p.error_flag);
            if (timing) t_ready += clock64() - tr0;
            if constexpr (RT) {                                  // the buffer this block overwrites held the previous tile's residual
              if (c == 0 && it > 0 && ok) ok = mbar_wait(bar_res_free(s), (it - 1) & 1u, p.error_flag);
            }
            if (ok) {
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + (uint32_t)((((c & 1) ^ (RT ? (int)(it & 1u) : 0)) * MS + s) * N);
              uint32_t a_tap = in_lo0 + (uint32_t)(s * 128) * row16;
              int slot = slot0;
              uint32_t par = par0;
              uint32_t b_tap = b_lo0 + (uint32_t)slot * (stage_bytes >> 4);
#ifdef SA_DIAG
              const long long ti0 = (timing && s > 0) ? clock64() : 0;
#endif
#pragma unroll
              for (int tap = 0; tap < K; ++tap) {
                const bool stage_begin = (tap % TPS) == 0, stage_end = ((tap + 1) % TPS) == 0 || tap == K - 1;
                if (s == 0 && stage_begin) {                     // later sub-tiles reuse the landed weights
                  const long long tw0 = timing ? clock64() : 0;
                  ok = ok && mbar_wait(bar_w_full(slot), par, p.error_flag);   // TMA data: the wait is the acquire, no tcgen05 fence
                  if (timing) t_w += clock64() - tw0;
                }
#pragma unroll
                for (int kk = 0; kk < K16; ++kk)
                  if (leader) umma_f16(d_tmem, desc64(a_tap + 2u * kk, hi), desc64(b_tap + 2u * kk, hi), idesc, ((RT && (c & 1)) || (tap | kk)) ? 1u : 0u);
                a_tap += dil16;
                if (s == MS - 1 && stage_end) {                  // last sub-tile: the slot may be refilled
                  if (leader) umma_commit(bar_w_empty(slot));
                }
                if (stage_end) {                                 // next weight stage
                  if (++slot == NSLOTS) { slot = 0; par ^= 1u; b_tap = b_lo0; } else { b_tap = b_lo0 + (uint32_t)slot * (stage_bytes >> 4); }
                } else {
                  b_tap += tap16;                                // taps are consecutive inside a stage
                }
              }
              slot_end = slot; par_end = par;
#ifdef SA_DIAG
              if (timing && s > 0) { t_issue += clock64() - ti0; ++n_issue; }
#endif
              if (leader) umma_commit(bar_acc_full(c & 1, s));
              __syncwarp();
            }
          }
          slot0 = slot_end; par0 = par_end;
        }
      }
      if (timing && lane == 0) {
        const long long tot = clock64() - t_begin;
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 0), (unsigned long long)tot);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 1), (unsigned long long)t_ready);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 2), (unsigned long long)t_w);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 3), (unsigned long long)(tot - t_ready - t_w));
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 10), (unsigned long long)t_issue);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 11), (unsigned long long)n_issue);
      }
    }
  } else {
    // ===== epilogue: warps (warp - 2) / WPS own sub-tile s; this thread owns row r (channel chunks ch0..) =====
    const int lg = warp & 3;
    const int s = warp / WPS;
    const int ch0 = (WPS == 8) ? ((warp % WPS) >> 2) * kCPT : 0;         // first 8-channel chunk of this thread
    const int r = s * 128 + lg * 32 + lane;
    constexpr int cchunks = C / 8;
    float xr[kCPT * 8];                                          // fp32 residual stream of this thread's channels
    const uint32_t row_off = (uint32_t)(kChainPad + r) * RB;
    uint32_t soff[kCPT];                                         // swizzled staging offsets of this thread's chunks (tile invariant)
#pragma unroll
    for (int q = 0; q < kCPT; ++q) soff[q] = swz(row_off + (uint32_t)(ch0 + q) * 16u, RB);
    uint32_t it = 0;
    bool ok = true;
#ifdef SA_DIAG
    const bool timing = p.timing != nullptr && warp == 0;
#else
    constexpr bool timing = false;   // in-kernel cycle counters: -DSA_DIAG builds only
#endif
    long long t_p0 = 0, t_acc = 0, t_ld = 0, t_fence = 0, t_begin = timing ? clock64() : 0;
#ifdef SA_DIAG
    if (p.flags & (1u << 30)) ok = false;                        // diagnostics: free-running MMA warp, no epilogue at all
#endif
    if constexpr (RT) {
      static_assert(kCPT == 4, "RT: 32 accumulator columns per epilogue thread (C = 64, 8 warps per sub-tile)");
      const uint32_t tcol = ((uint32_t)(lg * 32) << 16) + (uint32_t)(s * N + ch0 * 8);
      auto t_buf = [&](uint32_t bufi) { return tmem_base + tcol + bufi * (uint32_t)(MS * N); };
      const int n_pairs = p.n_convs / 2;
      auto locate = [&](int tile, int& b, int& t) {
        int mt;
        tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile, b, mt);
        t = mt * valid_rows - p.halo + r;
      };
      // x of a tile: global -> registers ...
      auto load_x = [&](int b, int t, float4 (&xa)[kCPT], float4 (&xb)[kCPT]) {
#pragma unroll
        for (int q = 0; q < kCPT; ++q) {
          xa[q] = make_float4(0.f, 0.f, 0.f, 0.f); xb[q] = xa[q];
          if (t >= 0 && t < p.L) ldg_f8(p.x32 + (((size_t)b * cchunks + ch0 + q) * (size_t)p.L + t) * 8, xa[q], xb[q]);
        }
      };
      // ... -> the residual buffer of tile number `itn` of this CTA (TMEM) and lrelu(x) -> bufA; then the inputs of pair 0 are ready
      auto put_x = [&](uint32_t itn, const float4 (&xa)[kCPT], const float4 (&xb)[kCPT]) {
        const uint32_t res = t_buf(1u ^ (itn & 1u));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const float4 a0 = xa[2 * g], a1 = xb[2 * g], a2 = xa[2 * g + 1], a3 = xb[2 * g + 1];
          const uint32_t rr[16] = {__float_as_uint(a0.x), __float_as_uint(a0.y), __float_as_uint(a0.z), __float_as_uint(a0.w),
                                   __float_as_uint(a1.x), __float_as_uint(a1.y), __float_as_uint(a1.z), __float_as_uint(a1.w),
                                   __float_as_uint(a2.x), __float_as_uint(a2.y), __float_as_uint(a2.z), __float_as_uint(a2.w),
                                   __float_as_uint(a3.x), __float_as_uint(a3.y), __float_as_uint(a3.z), __float_as_uint(a3.w)};
          __syncwarp();
          tmem_st16(res + (uint32_t)(g * 16), rr);
        }
#pragma unroll
        for (int q = 0; q < kCPT; ++q) {
          const float v[8] = {xa[q].x, xa[q].y, xa[q].z, xa[q].w, xb[q].x, xb[q].y, xb[q].z, xb[q].w};
          *reinterpret_cast<uint4*>(bufA + soff[q]) = pack8_lrelu(v, 0.1f, true, bf16);    // rows outside the utterance were loaded as zeros
        }
        tmem_st_wait();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready(0, s));
      };
      int b = 0, t = 0;
      if ((int)blockIdx.x < n_live) {                            // prologue: the first tile's x
        locate((int)blockIdx.x, b, t);
        float4 xa[kCPT], xb[kCPT];
        load_x(b, t, xa, xb);
        put_x(0u, xa, xb);
      }
      for (int tile = blockIdx.x; tile < n_live && ok; tile += gridDim.x, ++it) {
        const bool inside = t >= 0 && t < p.L;
        const bool keep = inside && r >= p.halo && r < R - p.halo;
        const uint32_t acc_b = it & 1u, res_b = acc_b ^ 1u;        // TMEM buffers of this tile: conv1 accumulator, residual / conv2
        const int tile_n = tile + (int)gridDim.x;
        const bool has_next = tile_n < n_live;
        int bn = 0, tn = 0;
        if (has_next) {
          locate(tile_n, bn, tn);
          if (tn >= 0 && tn < p.L) {                             // the next tile's rows: into L2 now, into registers under the last conv2
#pragma unroll
            for (int q = 0; q < kCPT; ++q)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x32 + (((size_t)bn * cchunks + ch0 + q) * (size_t)p.L + tn) * 8));
          }
        }
        if (keep && (p.flags & (EPI_SUM_ADD | EPI_SUM_FIN))) {
#pragma unroll
          for (int q = 0; q < kCPT; ++q)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sum32 + (((size_t)b * cchunks + ch0 + q) * (size_t)p.L + t) * 8));
        }
#pragma unroll 1
        for (int m = 0; m < n_pairs && ok; ++m) {
          const uint32_t par = (it * (uint32_t)n_pairs + (uint32_t)m) & 1u;
          const bool last = (m == n_pairs - 1);
          // ---- conv1: accumulator + bias -> lrelu -> conv2's input ----
          {
            ok = mbar_wait_relaxed(bar_acc_full(0, s), par, p.error_flag);
            if (!ok) break;
            tc_fence_after();
            uint32_t rr[32];
            __syncwarp();
            tmem_ld32(t_buf(acc_b), rr);
            tmem_ld_wait();
            const float* bias_c = bias_s + (2 * m) * C + ch0 * 8;
#pragma unroll
            for (int q = 0; q < kCPT; ++q) {
              const float4 b0 = *reinterpret_cast<const float4*>(bias_c + q * 8), b1 = *reinterpret_cast<const float4*>(bias_c + q * 8 + 4);
              const float v[8] = {__uint_as_float(rr[q * 8 + 0]) + b0.x, __uint_as_float(rr[q * 8 + 1]) + b0.y,
                                  __uint_as_float(rr[q * 8 + 2]) + b0.z, __uint_as_float(rr[q * 8 + 3]) + b0.w,
                                  __uint_as_float(rr[q * 8 + 4]) + b1.x, __uint_as_float(rr[q * 8 + 5]) + b1.y,
                                  __uint_as_float(rr[q * 8 + 6]) + b1.z, __uint_as_float(rr[q * 8 + 7]) + b1.w};
              *reinterpret_cast<uint4*>(bufT + soff[q]) = pack8_lrelu(v, 0.1f, inside, bf16);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready(1, s));
          }
          // ---- under the last conv2: the next tile's x ----
          if (last && has_next) {
            float4 xa[kCPT], xb[kCPT];
            load_x(bn, tn, xa, xb);
            // bufA rows of this sub-tile are also read by the last conv1's MMAs of sub-tile s + 1 (those of s - 1 were issued,
            // hence complete, before those of s)
            if (s + 1 < MS) ok = mbar_wait_relaxed(bar_acc_full(0, s + 1), par, p.error_flag);
            if (!ok) break;
            put_x(it + 1u, xa, xb);
          }
          // ---- conv2: the accumulator is the residual stream; + the running conv2 bias ----
          {
            ok = mbar_wait_relaxed(bar_acc_full(1, s), par, p.error_flag);
            if (!ok) break;
            tc_fence_after();
            uint32_t rr[32];
            __syncwarp();
            tmem_ld32(t_buf(res_b), rr);
            tmem_ld_wait();
            if (last) {                                          // the row is in registers: the buffer may be overwritten
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_res_free(s));
            }
            const float* bias_c = bias_s + (2 * m + 1) * C + ch0 * 8;
#pragma unroll
            for (int q = 0; q < kCPT; ++q) {
              const float4 b0 = *reinterpret_cast<const float4*>(bias_c + q * 8), b1 = *reinterpret_cast<const float4*>(bias_c + q * 8 + 4);
              float v[8] = {__uint_as_float(rr[q * 8 + 0]) + b0.x, __uint_as_float(rr[q * 8 + 1]) + b0.y,
                            __uint_as_float(rr[q * 8 + 2]) + b0.z, __uint_as_float(rr[q * 8 + 3]) + b0.w,
                            __uint_as_float(rr[q * 8 + 4]) + b1.x, __uint_as_float(rr[q * 8 + 5]) + b1.y,
                            __uint_as_float(rr[q * 8 + 6]) + b1.z, __uint_as_float(rr[q * 8 + 7]) + b1.w};
              if (!last) {
                *reinterpret_cast<uint4*>(bufA + soff[q]) = pack8_lrelu(v, 0.1f, inside, bf16);
              } else if (keep) {
                // final epilogue: multi-receptive-field combine + stores (v = x_final)
                const size_t idx = (((size_t)b * cchunks + ch0 + q) * (size_t)p.L + t) * 8;
                if (p.flags & (EPI_SUM_ADD | EPI_SUM_FIN)) {
                  float4 s0, s1;
                  ldg_f8(p.sum32 + idx, s0, s1);
                  v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
                  v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
                }
                if (p.flags & EPI_SUM_FIN) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) v[e] = v[e] / p.n_blocks;
                }
                if (p.flags & (EPI_SUM_SET | EPI_SUM_ADD)) stg_f8(p.sum32 + idx, v);
                if (p.flags & EPI_OUT32) stg_f8(p.out32 + idx, v);
                if (p.flags & EPI_OUT16) {
                  const size_t o16 = (((size_t)b * (size_t)p.L + t) * cchunks + ch0 + q) * 16;   // [B][1][L][C]
                  *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out16) + o16) = pack8_lrelu(v, p.slope_out, true, bf16);
                }
              }
            }
            if (!last) {
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_ready(0, s));
            }
          }
        }
        b = bn; t = tn;
      }
    } else
    for (int tile = blockIdx.x; tile < n_live && ok; tile += gridDim.x, ++it) {
      const long long tp0 = timing ? clock64() : 0;
      int b, mt;
      tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile, b, mt);
      const int t = mt * valid_rows - p.halo + r;                // global row of this thread
      const bool inside = t >= 0 && t < p.L;
      const bool keep = inside && r >= p.halo && r < R - p.halo; // rows this tile is responsible for
      // ---- load x, keep it in registers, stage lrelu(x) ----
#pragma unroll
      for (int q = 0; q < kCPT; ++q) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c4 = a;
        if (inside) {
          const float* src = p.x32 + (((size_t)b * cchunks + ch0 + q) * (size_t)p.L + t) * 8;
          ldg_f8(src, a, c4);
        }
        xr[q * 8 + 0] = a.x; xr[q * 8 + 1] = a.y; xr[q * 8 + 2] = a.z; xr[q * 8 + 3] = a.w;
        xr[q * 8 + 4] = c4.x; xr[q * 8 + 5] = c4.y; xr[q * 8 + 6] = c4.z; xr[q * 8 + 7] = c4.w;
      }
#pragma unroll
      for (int q = 0; q < kCPT; ++q) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = xr[q * 8 + e];
        *reinterpret_cast<uint4*>(bufA + soff[q]) = pack8_lrelu(v, 0.1f, true, bf16);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready(0, TM ? 0 : s));
      // the multi-receptive-field partial sum is only needed by this tile's last epilogue: have it in L2 by then
      if (keep && (p.flags & (EPI_SUM_ADD | EPI_SUM_FIN))) {
#pragma unroll
        for (int q = 0; q < kCPT; ++q)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sum32 + (((size_t)b * cchunks + ch0 + q) * (size_t)p.L + t) * 8));
      }
      // The x load above sits at the head of every tile's dependency chain: pull the next tile's rows into L2 now.
      if (tile + (int)gridDim.x < n_live) {
        int bn, mtn;
        tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile + (int)gridDim.x, bn, mtn);
        const int tn = mtn * valid_rows - p.halo + r;
        if (tn >= 0 && tn < p.L) {
#pragma unroll
          for (int q = 0; q < kCPT; ++q)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x32 + (((size_t)bn * cchunks + ch0 + q) * (size_t)p.L + tn) * 8));
        }
      }
      if (timing) t_p0 += clock64() - tp0;
      // ---- the convs: (conv1, conv2) pairs; one generic step with compile-time flags so the residual add, the
      // final stores and the staging stores are straight-line code without predicated moves ----
      auto step = [&](auto second_c, auto last_c, int c) {
        constexpr bool second = decltype(second_c)::value;       // conv2 of a pair: x += ..
        constexpr bool last = decltype(last_c)::value;           // last conv of the block: final epilogue
        const uint32_t acc_parity = (it * (uint32_t)(p.n_convs / 2) + (uint32_t)(c / 2)) & 1u;
        const float* bias_c = bias_s + c * C + ch0 * 8;
        uint8_t* out_buf = second ? bufA : bufT;
        const long long ta0 = timing ? clock64() : 0;
        ok = ok && mbar_wait_relaxed(bar_acc_full(c & 1, TM ? 0 : s), acc_parity, p.error_flag);
        if (!ok) return;
        tc_fence_after();
        if (timing) t_acc += clock64() - ta0;
        const uint32_t t_addr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(((c & 1) * MS + s) * N + ch0 * 8);
#pragma unroll
        for (int g = 0; g < kCPT / 2; ++g) {                     // 16 columns = 2 channel chunks per TMEM round trip
          uint32_t rr[16];
          const long long tl0 = timing ? clock64() : 0;
          __syncwarp();
          tmem_ld16(t_addr + (uint32_t)(g * 16), rr);
          tmem_ld_wait();
          if (timing) t_ld += clock64() - tl0;
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
              for (int e = 0; e < 8; ++e) { v[e] += xr[q * 8 + e]; xr[q * 8 + e] = v[e]; }
            }
            if constexpr (!last) {
              *reinterpret_cast<uint4*>(out_buf + soff[q]) = pack8_lrelu(v, 0.1f, inside, bf16);
            } else if (keep) {
              // final epilogue: multi-receptive-field combine + stores (v = x_final)
              const size_t idx = (((size_t)b * cchunks + ch0 + q) * (size_t)p.L + t) * 8;
              if (p.flags & (EPI_SUM_ADD | EPI_SUM_FIN)) {
                float4 s0, s1;
                ldg_f8(p.sum32 + idx, s0, s1);
                v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w;
                v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
              }
              if (p.flags & EPI_SUM_FIN) {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = v[e] / p.n_blocks;
              }
              if (p.flags & (EPI_SUM_SET | EPI_SUM_ADD)) {
                stg_f8(p.sum32 + idx, v);
              }
              if (p.flags & EPI_OUT32) {
                stg_f8(p.out32 + idx, v);
              }
              if (p.flags & EPI_OUT16) {
                const size_t o16 = (((size_t)b * (size_t)p.L + t) * cchunks + ch0 + q) * 16;   // [B][1][L][C]
                *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out16) + o16) = pack8_lrelu(v, p.slope_out, true, bf16);
              }
            }
          }
        }
        if constexpr (!last) {
          const long long tf0 = timing ? clock64() : 0;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_ready(second ? 0 : 1, TM ? 0 : s));
          if (timing) t_fence += clock64() - tf0;
        }
      };
      const int n_pairs = p.n_convs / 2;
#pragma unroll 1
      for (int m = 0; m + 1 < n_pairs && ok; ++m) {
        step(std::false_type{}, std::false_type{}, 2 * m);
        step(std::true_type{}, std::false_type{}, 2 * m + 1);
      }
      step(std::false_type{}, std::false_type{}, 2 * (n_pairs - 1));
      step(std::true_type{}, std::true_type{}, 2 * (n_pairs - 1) + 1);
    }
    if (timing && lane == 0) {
      const long long tot = clock64() - t_begin;
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 4), (unsigned long long)tot);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 5), (unsigned long long)t_p0);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 6), (unsigned long long)t_acc);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 7), (unsigned long long)(tot - t_p0 - t_acc));
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 8), (unsigned long long)t_ld);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 9), (unsigned long long)t_fence);
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
