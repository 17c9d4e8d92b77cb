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
// x_{m+1} = x_m + conv2(...) needs no registers and no add; the conv2 biases are added when the epilogue reads
// (cumulative: x_{m+1} = D + b2_0 + .. + b2_m).  Sums differ from the per-layer path in the last fp32 bit only.
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

constexpr int kGrpMS = 2;                                  // sub-tiles (128 rows of 128 bytes) per stream tile
constexpr int kGrpStreams = 2;
constexpr int kGrpRows = kGrpMS * 128;                     // rows of one stream tile
constexpr int kGrpPadRows = 8;                             // slack rows on both sides (one 1024-byte swizzle atom)
constexpr int kGrpEpiWarps = kGrpStreams * kGrpMS * 4;     // 16
constexpr int kGrpThreads = 32 * (kGrpEpiWarps + 2);       // 576
constexpr int kGrpMaxStages = 12;                          // weight ring depth (runtime n_wstages <= this)
constexpr uint32_t kGrpSliceBytes = 64 * 32;               // one Toeplitz block: [64 rows (g', co)][16 ci] 16-bit
constexpr int kGrpSlicesPerStage = 4;
constexpr uint32_t kGrpStageBytes = kGrpSlicesPerStage * kGrpSliceBytes;   // 8 KB
constexpr int kGrpMaxPairs = 3;
constexpr uint32_t kGrpBufBytes = (kGrpRows + 2 * kGrpPadRows) * 128;      // 34816 = 34 KB, multiple of 1024

struct GroupParams {
  const float* x32;         // block input, fp32 blocked [B][C/8][L][8]
  float* sum32;             // MRF running sum, fp32 blocked
  float* out32;             // EPI_OUT32
  void* out16;              // EPI_OUT16: lrelu(output), 16-bit [B][1][L][C]
  const void* w;            // n_convs convs, each stages_per_conv stages of 8 KB (4 Toeplitz slices, zero padded)
  const float* bias;        // [n_convs][C]
  int* error_flag;
  long long* timing;        // optional [16] cycle counters: MMA warp total / wait ready / wait weights
  int L;                    // positions per item (multiple of G)
  int n_convs;              // 2 * n_pairs
  int dil[kGrpMaxPairs];    // dilation of conv1 of each pair
  int halo;                 // recomputed positions per side (sum of all conv reaches), multiple of G
  int n_slices;             // (G + k - 1) * C / 16
  int stages_per_conv;      // ceil(n_slices / 4)
  int n_wstages;            // ring depth, >= stages_per_conv + 1
  int tiles_per_item, total_tiles;
  TileMapParams map;        // ragged batches (conv_tc.cuh); tile axis = valid positions per tile
  uint32_t flags;           // EPI_* of the final epilogue
  float slope_out;
  float n_blocks;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int C, bool BF16>
__global__ void __launch_bounds__(kGrpThreads, 1) group_chain_kernel(const __grid_constant__ GroupParams p) {
  static_assert(C == 16 || C == 32, "grouped formulation: C = 16 (G = 4) or C = 32 (G = 2)");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int G = 64 / C;                                      // positions per 128-byte row
  constexpr int R = kGrpRows * G;                                // positions per stream tile
  constexpr uint32_t PB = 2u * C;                                // bytes per position
  constexpr uint32_t kPadBytes = kGrpPadRows * 128;
  constexpr int kGroupsPerPos = C / 16;                          // 16-column TMEM groups per position
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpW = kGrpEpiWarps, kWarpMma = kGrpEpiWarps + 1;
  // smem: buf[stream][A|T], weight ring, bias [n_convs][C] + cumulative conv2 bias [n_pairs][C], barriers
  auto buf = [&](int st, int t) { return smem + (uint32_t)(st * 2 + t) * kGrpBufBytes; };
  uint8_t* w_smem = smem + 4 * kGrpBufBytes;
  float* bias_s = reinterpret_cast<float*>(w_smem + (size_t)p.n_wstages * kGrpStageBytes);
  float* cbias_s = bias_s + 2 * kGrpMaxPairs * C;
  uint64_t* bars = reinterpret_cast<uint64_t*>(cbias_s + kGrpMaxPairs * C);
  // barriers: ready[stream][A|T] (4), acc_full[stream] (2), w_full[12], w_empty[12]
  auto bar_ready = [&](int st, int t) { return smem_u32(&bars[st * 2 + t]); };
  auto bar_acc_full = [&](int st) { return smem_u32(&bars[4 + st]); };
  auto bar_w_full = [&](int i) { return smem_u32(&bars[6 + i]); };
  auto bar_w_empty = [&](int i) { return smem_u32(&bars[6 + kGrpMaxStages + i]); };
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 6 + 2 * kGrpMaxStages);

  const int valid = R - 2 * p.halo;
  __shared__ int tile_pre[kMaxMapItems + 1];
  tilemap_build(tile_pre, p.map, p.L, valid);                    // visible after the __syncthreads() below
  constexpr bool bf16 = BF16;                                    // both 16-bit flavours of the epilogue would double its code
  const int n_pairs = p.n_convs / 2;

  if (warp == kWarpW && lane == 0) {
    for (int st = 0; st < kGrpStreams; ++st) {
      mbar_init(bar_ready(st, 0), kGrpMS * 4);                   // the eight warps of the stream
      mbar_init(bar_ready(st, 1), kGrpMS * 4);
      mbar_init(bar_acc_full(st), 1);
    }
    for (int i = 0; i < kGrpMaxStages; ++i) { mbar_init(bar_w_full(i), 1); mbar_init(bar_w_empty(i), 1); }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc(smem_u32(tmem_holder), 512);
  for (int i = threadIdx.x; i < p.n_convs * C; i += kGrpThreads) bias_s[i] = p.bias[i];
  for (int i = threadIdx.x; i < n_pairs * C; i += kGrpThreads) {  // b2_0 + .. + b2_m per channel
    const int m = i / C, ch = i - m * C;
    float s = 0.f;
    for (int mm = 0; mm <= m; ++mm) s += p.bias[(2 * mm + 1) * C + ch];
    cbias_s[i] = s;
  }
  // zero the four staged tiles once: the leading slack rows are never written, everything else only ever holds finite
  // 16-bit activations (a Toeplitz block multiplies positions outside its taps by an exact 0)
  for (uint32_t i = threadIdx.x; i < 4 * kGrpBufBytes / 16; i += kGrpThreads)
    *reinterpret_cast<uint4*>(smem + i * 16) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  const int n_live = tilemap_total(tile_pre, p.map, p.total_tiles);
  const int my_tiles = ((int)blockIdx.x < n_live) ? (n_live - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int n_iters = (my_tiles + 1) / 2;                        // two tiles (streams) per iteration

  if (warp == kWarpW) {
    // ===== weight producer: the stages of conv c, once per iteration (both streams read them) =====
    const bool leader = elect_one();
    int slot = 0;
    uint32_t par = 1;                                            // parity of the previous use of `slot`
    bool wrapped = false, ok = true;
    for (int it = 0; it < n_iters && ok; ++it)
      for (int c = 0; c < p.n_convs && ok; ++c) {
        const uint8_t* src = static_cast<const uint8_t*>(p.w) + (size_t)c * p.stages_per_conv * kGrpStageBytes;
        for (int i = 0; i < p.stages_per_conv; ++i) {
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
  } else if (warp == kWarpMma) {
    // ===== MMA issuer (warp-uniform loop, one elected lane issues) =====
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(64, bf16);
    constexpr uint32_t hiA = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);     // SWIZZLE_128B, 8-row groups 1024 B apart
    constexpr uint32_t hiB = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);      // SWIZZLE_32B, 8-row groups 256 B apart
    const uint32_t b_lo0 = desc_lo(smem_u32(w_smem));
    const int pad_pos = (p.n_slices / kGroupsPerPos - G) / 2;                  // (k - 1) / 2 positions
    const int n_sl = p.n_slices, n_stg = p.stages_per_conv;
    int slot0 = 0;
    uint32_t par0 = 0;
    bool ok = true;
    const bool timing = p.timing != nullptr;
    long long t_ready = 0, t_w = 0, t_begin = timing ? clock64() : 0;
    for (int it = 0; it < n_iters && ok; ++it) {
      for (int c = 0; c < p.n_convs && ok; ++c) {
        const int t_in = c & 1;                                                // conv1 reads A, conv2 reads T
        const uint32_t rdy_parity = ((uint32_t)it * (uint32_t)n_pairs + (uint32_t)(c >> 1)) & 1u;
        int slot = slot0;
        uint32_t par = par0;
#pragma unroll 1
        for (int st = 0; st < kGrpStreams && ok; ++st) {
          const long long tr0 = timing ? clock64() : 0;
          ok = mbar_wait(bar_ready(st, t_in), rdy_parity, p.error_flag);
          if (timing) t_ready += clock64() - tr0;
          if (!ok) break;
          tc_fence_after();
          // first slice of the first row: kPadBytes - pad_pos positions
          const uint32_t a_lo0 = desc_lo(smem_u32(buf(st, t_in)) + kPadBytes - (uint32_t)pad_pos * PB);
          const uint32_t d_tmem = tmem_base + (uint32_t)((st * 2 + t_in) * kGrpMS * 64);
          slot = slot0; par = par0;
          int q0 = 0;
#pragma unroll 1
          for (int i = 0; i < n_stg; ++i, q0 += kGrpSlicesPerStage) {
            if (st == 0) {
              const long long tw0 = timing ? clock64() : 0;
              ok = mbar_wait(bar_w_full(slot), par, p.error_flag);
              if (timing) t_w += clock64() - tw0;
              if (!ok) break;
              tc_fence_after();
            }
            const uint32_t b_lo = b_lo0 + (uint32_t)slot * (kGrpStageBytes >> 4);
            const uint32_t a_lo = a_lo0 + 2u * (uint32_t)q0;                   // 32 bytes per slice
#pragma unroll
            for (int s = 0; s < kGrpMS; ++s) {
#pragma unroll
              for (int qq = 0; qq < kGrpSlicesPerStage; ++qq) {
                if (q0 + qq < n_sl) {
                  const uint32_t accum = (t_in == 1 || q0 + qq > 0) ? 1u : 0u;  // conv2 accumulates onto the residual
                  if (leader)
                    umma_f16(d_tmem + (uint32_t)(s * 64), desc64(a_lo + (uint32_t)(s * 128 * 8) + 2u * qq, hiA),
                             desc64(b_lo + (uint32_t)qq * (kGrpSliceBytes >> 4), hiB), idesc, accum);
                }
              }
            }
            if (st == kGrpStreams - 1 && leader) umma_commit(bar_w_empty(slot));   // both streams have read the stage
            __syncwarp();
            if (++slot == p.n_wstages) { slot = 0; par ^= 1u; }
          }
          if (!ok) break;
          if (leader) umma_commit(bar_acc_full(st));
          __syncwarp();
        }
        slot0 = slot; par0 = par;
      }
    }
    if (timing && lane == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 0), (unsigned long long)(clock64() - t_begin));
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 1), (unsigned long long)t_ready);
      atomicAdd(reinterpret_cast<unsigned long long*>(p.timing + 2), (unsigned long long)t_w);
    }
  } else {
    // ===== epilogue warps: one thread = one 128-byte row (G positions x C channels) of its stream's tile =====
    const int st = warp >> 3;
    const int s = (warp >> 2) & (kGrpMS - 1);
    const int lg = warp & 3;
    const int r = s * 128 + lg * 32 + lane;                      // row within the stream tile
    constexpr int cchunks = C / 8;
    uint8_t* const bufA = buf(st, 0);
    uint8_t* const bufT = buf(st, 1);
    const uint32_t t_acc1 = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(((st * 2 + 0) * kGrpMS + s) * 64);
    const uint32_t t_res = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(((st * 2 + 1) * kGrpMS + s) * 64);
    // tile-invariant d-major maps of this row's G positions, for every pair with a dilated conv1:
    //   posP[m][g] = position of time (G r + g) in the d_m-major input tile of pair m's conv1
    //   tauI(m, g) = time whose conv1 output this row holds at column group g in pair m (>= R: none)
    uint32_t pmap[kGrpMaxPairs][G];                               // posP | tauI << 16 (both < 2^16)
#pragma unroll
    for (int m = 0; m < kGrpMaxPairs; ++m) {
      const int d = (m < n_pairs) ? p.dil[m] : 1;
      const int Q = (R + d - 1) / d;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int tau = G * r + g;
        const int pp = (d == 1) ? tau : (tau % d) * Q + tau / d;
        const int ti = (d == 1) ? tau : d * (tau % Q) + tau / Q;
        pmap[m][g] = (uint32_t)pp | ((uint32_t)ti << 16);
      }
    }
    auto posP = [&](int m, int g) { return (int)(pmap[m][g] & 0xFFFFu); };
    auto tauI = [&](int m, int g) { return (int)(pmap[m][g] >> 16); };
    // swizzled byte offset (within a staged tile) of the 16-byte chunk at byte `ch` of position `pos`, and of its neighbour
    auto store_pos = [&](uint8_t* b, int pos, uint32_t ch, const uint4& lo, const uint4& hi8) {
      const uint32_t lin = kPadBytes + (uint32_t)pos * PB + ch;
      const uint32_t x = ((lin >> 7) & 7u) << 4;
      *reinterpret_cast<uint4*>(b + (lin ^ x)) = lo;
      *reinterpret_cast<uint4*>(b + ((lin + 16u) ^ x)) = hi8;
    };
    bool ok = true;
    for (int it = 0; it < n_iters && ok; ++it) {
      const int tile = (int)blockIdx.x + (2 * it + st) * (int)gridDim.x;
      const bool live = tile < n_live;                            // the last iteration may have one stream without a tile
      int b = 0, mt = 0;
      if (live) tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile, b, mt);
      const int t0 = mt * valid - p.halo;                         // time of the tile's first position
      const int t_row = t0 + G * r;
      const bool inside = live && t_row >= 0 && t_row < p.L;      // L, halo, valid are multiples of G: whole rows
      const bool keep = inside && G * r >= p.halo && G * r < R - p.halo;
      // ---- x: residual stream -> tensor memory, lrelu(x) -> input tile of pair 0's conv1 ----
      {
        const int d0 = p.dil[0];
#pragma unroll
        for (int gi = 0; gi < 4; ++gi) {
          const int g = gi / kGroupsPerPos;
          const int ch0 = (gi % kGroupsPerPos) * 16;
          float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0, q2 = q0, q3 = q0;
          if (inside) {
            const float* src = p.x32 + (((size_t)b * cchunks + (ch0 >> 3)) * (size_t)p.L + (size_t)(t_row + g)) * 8;
            ldg_f8(src, q0, q1);
            ldg_f8(src + (size_t)p.L * 8, q2, q3);
          }
          const float v[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
          uint32_t rr[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) rr[e] = __float_as_uint(v[e]);
          __syncwarp();
          tmem_st16(t_res + (uint32_t)(gi * 16), rr);
          float lo[8], hi8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { lo[e] = v[e]; hi8[e] = v[8 + e]; }
          store_pos(bufA, d0 == 1 ? G * r + g : posP(0, g), (uint32_t)ch0 * 2u, pack8_lrelu(lo, 0.1f, true, bf16),
                    pack8_lrelu(hi8, 0.1f, true, bf16));
        }
        tmem_st_wait();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_ready(st, 0));
      }
      // the next tile of this stream and this tile's running sum: have them in L2 when they are needed
      if (keep && (p.flags & (EPI_SUM_ADD | EPI_SUM_FIN))) {
#pragma unroll
        for (int q = 0; q < cchunks; ++q)
#pragma unroll
          for (int g = 0; g < G; g += 4)                           // 32 B per position: one 128-byte line holds 4
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sum32 + (((size_t)b * cchunks + q) * (size_t)p.L + (size_t)(t_row + g)) * 8));
      }
      {
        const int tile_n = tile + 2 * (int)gridDim.x;
        if (tile_n < n_live) {
          int bn, mtn;
          tilemap_locate(tile_pre, p.map, p.tiles_per_item, tile_n, bn, mtn);
          const int tn = mtn * valid - p.halo + G * r;
          if (tn >= 0 && tn < p.L) {
#pragma unroll
            for (int q = 0; q < cchunks; ++q)
#pragma unroll
              for (int g = 0; g < G; g += 4)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x32 + (((size_t)bn * cchunks + q) * (size_t)p.L + (size_t)(tn + g)) * 8));
          }
        }
      }
      // ---- the (conv1, conv2) pairs ----
#pragma unroll
      for (int m = 0; m < kGrpMaxPairs; ++m) {
        if (m < n_pairs && ok) {
          const int d = p.dil[m];
          // conv1: TMEM -> + bias -> lrelu -> conv2's input tile in natural order
          ok = mbar_wait_relaxed(bar_acc_full(st), 0u, p.error_flag);     // phases alternate conv1 (0) / conv2 (1): n_convs is even
          tc_fence_after();
          if (ok) {
            const float* bias_c = bias_s + (2 * m) * C;
#pragma unroll
            for (int gh = 0; gh < 2; ++gh) {
            uint32_t rr[2][16];
            __syncwarp();
            tmem_ld16(t_acc1 + (uint32_t)(gh * 32), rr[0]);
            tmem_ld16(t_acc1 + (uint32_t)(gh * 32 + 16), rr[1]);
            tmem_ld_wait();
#pragma unroll
            for (int gi = 2 * gh; gi < 2 * gh + 2; ++gi) {
              const int g = gi / kGroupsPerPos;
              const int ch0 = (gi % kGroupsPerPos) * 16;
              const int tau = (d == 1) ? G * r + g : tauI(m, g);
              const int tt = t0 + tau;
              const bool ins = live && tau < R && tt >= 0 && tt < p.L;
              float lo[8], hi8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                lo[e] = __uint_as_float(rr[gi - 2 * gh][e]) + bias_c[ch0 + e];
                hi8[e] = __uint_as_float(rr[gi - 2 * gh][8 + e]) + bias_c[ch0 + 8 + e];
              }
              if (tau < R)
                store_pos(bufT, tau, (uint32_t)ch0 * 2u, pack8_lrelu(lo, 0.1f, ins, bf16), pack8_lrelu(hi8, 0.1f, ins, bf16));
            }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready(st, 1));
          }
          // conv2: the accumulator IS the residual stream x_{m+1} (minus the biases added here)
          ok = ok && mbar_wait_relaxed(bar_acc_full(st), 1u, p.error_flag);
          tc_fence_after();
          if (ok) {
            const float* cb = cbias_s + m * C;
            uint32_t rr[4][16];
            __syncwarp();
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) tmem_ld16(t_res + (uint32_t)(gi * 16), rr[gi]);
            tmem_ld_wait();
            if (m + 1 < n_pairs) {
              const int dn = p.dil[m + 1 < kGrpMaxPairs ? m + 1 : 0];
#pragma unroll
              for (int gi = 0; gi < 4; ++gi) {
                const int g = gi / kGroupsPerPos;
                const int ch0 = (gi % kGroupsPerPos) * 16;
                float lo[8], hi8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  lo[e] = __uint_as_float(rr[gi][e]) + cb[ch0 + e];
                  hi8[e] = __uint_as_float(rr[gi][8 + e]) + cb[ch0 + 8 + e];
                }
                store_pos(bufA, dn == 1 ? G * r + g : posP(m + 1 < kGrpMaxPairs ? m + 1 : 0, g), (uint32_t)ch0 * 2u,
                          pack8_lrelu(lo, 0.1f, inside, bf16), pack8_lrelu(hi8, 0.1f, inside, bf16));
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_ready(st, 0));
            } else {
              // final epilogue: multi-receptive-field combine (archi.py:82-86) + output streams
              if (keep) {
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                  const int g = gi / kGroupsPerPos;
                  const int ch0 = (gi % kGroupsPerPos) * 16;
                  float v[16];
#pragma unroll
                  for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(rr[gi][e]) + cb[ch0 + e];
                  const size_t i0 = (((size_t)b * cchunks + (ch0 >> 3)) * (size_t)p.L + (size_t)(t_row + g)) * 8;
                  const size_t i1 = i0 + (size_t)p.L * 8;
                  if (p.flags & (EPI_SUM_ADD | EPI_SUM_FIN)) {
                    float4 s0, s1, s2, s3;
                    ldg_f8(p.sum32 + i0, s0, s1);
                    ldg_f8(p.sum32 + i1, s2, s3);
                    v[0] += s0.x; v[1] += s0.y; v[2] += s0.z; v[3] += s0.w; v[4] += s1.x; v[5] += s1.y; v[6] += s1.z; v[7] += s1.w;
                    v[8] += s2.x; v[9] += s2.y; v[10] += s2.z; v[11] += s2.w; v[12] += s3.x; v[13] += s3.y; v[14] += s3.z; v[15] += s3.w;
                  }
                  if (p.flags & EPI_SUM_FIN) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = v[e] / p.n_blocks;
                  }
                  if (p.flags & (EPI_SUM_SET | EPI_SUM_ADD)) { stg_f8(p.sum32 + i0, v); stg_f8(p.sum32 + i1, v + 8); }
                  if (p.flags & EPI_OUT32) { stg_f8(p.out32 + i0, v); stg_f8(p.out32 + i1, v + 8); }
                  if (p.flags & EPI_OUT16) {
                    float lo[8], hi8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { lo[e] = v[e]; hi8[e] = v[8 + e]; }
                    uint8_t* o = static_cast<uint8_t*>(p.out16) + (((size_t)b * (size_t)p.L + (size_t)(t_row + g)) * C + ch0) * 2;
                    stg_u8(o, pack8_lrelu(lo, p.slope_out, true, bf16), pack8_lrelu(hi8, p.slope_out, true, bf16));
                  }
                }
              }
              tc_fence_before();                                 // TMEM reads done before the next tile overwrites the residual
            }
          }
        }
      }
    }
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
