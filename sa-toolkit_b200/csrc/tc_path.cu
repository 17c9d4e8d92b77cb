// Host side of the tensor-core path: weight packing, workspace carving, tensor maps and the
// launch sequence of one generator forward on tcgen05 (kernels in conv_tc.cuh).
//
// Activation buffers (all channel-blocked [B][C/8][L][8]; E = largest activation):
//   XIN16  packed input x                           16-bit  B*T*512
//   P16    lrelu(previous stage / conv_pre output)  16-bit  E      input of the upsampler
//   AX16   lrelu(h), h = upsampler output           16-bit  E      shared by the 3 ResBlocks
//   A16    lrelu(residual stream)                   16-bit  E
//   T16    lrelu(conv1 output)                      16-bit  E
//   X32    h (stage input), later the stage output  fp32    E
//   R32    residual stream of the running ResBlock  fp32    E
//   S32    multi-receptive-field sum                fp32    E
#include "tc_path.cuh"
#include "conv_tc.cuh"
#include "conv_pair_tc.cuh"
#include "resblock_pair_tc.cuh"
#include "chain_tc.cuh"
#include "chain3_tc.cuh"
#include "chain_group_tc.cuh"
#include "chain_group_v1_tc.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace sa {

namespace {

thread_local char g_msg[512];
const char* msgf(const char* fmt, const char* a, const char* b = "") {
  snprintf(g_msg, sizeof(g_msg), fmt, a, b);
  return g_msg;
}
#define TC_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t e_ = (expr);                                                 \
    if (e_ != cudaSuccess) return msgf("%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// byte offset inside a swizzled block whose base is aligned to the swizzle pattern (conv_tc.cuh: swz)
uint32_t swizzle_offset(uint32_t off, uint32_t row_bytes) {
  const uint32_t mask = row_bytes == 128 ? 7u : row_bytes == 64 ? 3u : 1u;
  return off ^ (((off >> 7) & mask) << 4);
}

uint16_t to16(float f, bool bf16) {
  if (bf16) {
    __nv_bfloat16 h = __float2bfloat16_rn(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  }
  __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

// ---- small CUDA-core helper kernels ------------------------------------------------------

// x fp32 [B][Cin][T] -> 16-bit panel-blocked [B][Cpad/pw][T][pw] (channels >= Cin are zero).
__global__ void pack_input_kernel(const float* __restrict__ x, uint4* __restrict__ out, int Cin, int chunks, int T,
                                  int bf16, int pw) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c8 * 8 + e;
    v[e] = c < Cin ? __ldg(x + ((size_t)b * Cin + c) * T + t) : 0.f;
  }
  const int cpp = pw >> 3;                    // chunks per panel row
  out[(((size_t)b * (chunks / cpp) + c8 / cpp) * T + t) * cpp + c8 % cpp] = tc::pack8(v, bf16 != 0);
}

// The same packed input straight from the conditioning parts (Net._forward, hifigan.py:83-97: x = cat(bn, f0, speaker
// one-hot repeated over time)): the [B, 504, T] fp32 tensor is never materialised.  Bit-identical to packing cat(...).
__global__ void pack_parts_kernel(const float* __restrict__ bn, const float* __restrict__ f0, const float* __restrict__ spk,
                                  uint4* __restrict__ out, int n_bn, int n_spk, int chunks, int T, int bf16, int pw) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c8 * 8 + e;
    float x = 0.f;
    if (c < n_bn) x = __ldg(bn + ((size_t)b * n_bn + c) * T + t);
    else if (c == n_bn) x = __ldg(f0 + (size_t)b * T + t);
    else if (c < n_bn + 1 + n_spk) x = __ldg(spk + (size_t)b * n_spk + (c - n_bn - 1));
    v[e] = x;
  }
  const int cpp = pw >> 3;
  out[(((size_t)b * (chunks / cpp) + c8 / cpp) * T + t) * cpp + c8 % cpp] = tc::pack8(v, bf16 != 0);
}

// The same packed input from the COMPACT conditioning: the ASR bottleneck features are rows of a
// VQ codebook (48 codewords, /root/reference/satools/satools/chain/nn.py:427-459), the speaker block is a one-hot that is
// constant in time (hifigan.py:94-97): 1 + 4 bytes per frame + 4 bytes per item cross PCIe instead of 2016 bytes per
// frame.  Bit-identical to packing x when its BN rows are exact codewords (index >= n_codes: the zero padding frames).
__global__ void pack_vq_kernel(const uint8_t* __restrict__ idx, const float* __restrict__ codebook, const float* __restrict__ f0,
                               const int32_t* __restrict__ spk, uint4* __restrict__ out, int n_codes, int n_bn, int n_spk,
                               int chunks, int T, int bf16, int pw) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const int code = idx[(size_t)b * T + t];
  const int sp = spk[b];
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c8 * 8 + e;
    float x = 0.f;
    if (c < n_bn) x = code < n_codes ? __ldg(codebook + (size_t)code * n_bn + c) : 0.f;
    else if (c == n_bn) x = __ldg(f0 + (size_t)b * T + t);
    else if (c < n_bn + 1 + n_spk) x = (c - n_bn - 1 == sp) ? 1.f : 0.f;
    v[e] = x;
  }
  const int cpp = pw >> 3;
  out[(((size_t)b * (chunks / cpp) + c8 / cpp) * T + t) * cpp + c8 % cpp] = tc::pack8(v, bf16 != 0);
}

// fp32 blocked [B][C/8][L][8] -> fp32 [B][C][L]  (debug taps only)
__global__ void unblock_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int L) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = blockIdx.y, b = blockIdx.z;
  if (t >= L) return;
  const float4* src = reinterpret_cast<const float4*>(in + (((size_t)b * (C / 8) + c8) * L + t) * 8);
  const float4 a = src[0], c = src[1];
  float* o = out + ((size_t)b * C + c8 * 8) * L + t;
  o[0] = a.x; o[(size_t)L] = a.y; o[(size_t)2 * L] = a.z; o[(size_t)3 * L] = a.w;
  o[(size_t)4 * L] = c.x; o[(size_t)5 * L] = c.y; o[(size_t)6 * L] = c.z; o[(size_t)7 * L] = c.w;
}

// Tail on the blocked fp32 stage output (archi.py:87-90): lrelu(0.01) -> reflect pad (1,0) ->
// Conv1d(C->1, k, pad (k-1)/2) -> tanh.  w: [C][k] (the fp32 packing [Cin][k][1]).
__global__ void conv_post_blocked_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                         const float* __restrict__ bias, void* y, int C, int L, int k, float slope,
                                         int y_dtype) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < C * k; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int Lout = L + 1;
  if (n >= Lout) return;
  const int pad = (k - 1) / 2;
  float acc = __ldg(bias);
  for (int j = 0; j < k; ++j) {
    const int i = n + j - pad;
    if (i < 0 || i >= Lout) continue;
    const int src = (i == 0) ? 1 : i - 1;
    for (int c8 = 0; c8 < C / 8; ++c8) {
      const float4* p = reinterpret_cast<const float4*>(h + (((size_t)b * (C / 8) + c8) * L + src) * 8);
      const float4 a = __ldg(p), c = __ldg(p + 1);
      const float* wr = ws + (c8 * 8) * k + j;
      acc = fmaf(wr[0], tc::lrelu_f(a.x, slope), acc);
      acc = fmaf(wr[k], tc::lrelu_f(a.y, slope), acc);
      acc = fmaf(wr[2 * k], tc::lrelu_f(a.z, slope), acc);
      acc = fmaf(wr[3 * k], tc::lrelu_f(a.w, slope), acc);
      acc = fmaf(wr[4 * k], tc::lrelu_f(c.x, slope), acc);
      acc = fmaf(wr[5 * k], tc::lrelu_f(c.y, slope), acc);
      acc = fmaf(wr[6 * k], tc::lrelu_f(c.z, slope), acc);
      acc = fmaf(wr[7 * k], tc::lrelu_f(c.w, slope), acc);
    }
  }
  const float v = tanhf(acc);
  const size_t o = (size_t)b * Lout + n;
  if (y_dtype == SA_DTYPE_F32) reinterpret_cast<float*>(y)[o] = v;
  else if (y_dtype == SA_DTYPE_F16) reinterpret_cast<__half*>(y)[o] = __float2half_rn(v);
  else {
    float s = rintf(v * 32767.f);
    s = fminf(fmaxf(s, -32768.f), 32767.f);
    reinterpret_cast<int16_t*>(y)[o] = (int16_t)s;
  }
}

// conv_post over the 16-bit blocked stage output (archi.py:87-90).  The producer has already applied
// lrelu(0.01); ReflectionPad1d((1,0)) is the index map xp[i] = x[i == 0 ? 1 : i - 1]; zero padding of the k = 7
// conv is the range check.  Each thread owns one padded position: it reads its row once (C 16-bit values),
// forms the 7 per-tap partial sums in registers, and the block exchanges them through shared memory, so the
// kernel reads every activation exactly once (249 outputs per 256-thread block).
template <bool BF16>
__global__ void __launch_bounds__(256) conv_post16_k7_kernel(const uint4* __restrict__ h16, const float* __restrict__ w,
                                                             const float* __restrict__ bias, void* y, int C, int PW, int L,
                                                             int y_dtype, const int* __restrict__ frames, int keep_margin, int samples_per_frame) {
  constexpr int K = 7, kOut = 256 - (K - 1);
  extern __shared__ float4 post_smem[];
  float4* ws = post_smem;                                   // [C][2] float4: taps 0-3, taps 4-6 + 0
  float* part = reinterpret_cast<float*>(post_smem + 2 * C);   // [K][256]
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    const float* wr = w + i * K;
    ws[2 * i] = make_float4(wr[0], wr[1], wr[2], wr[3]);
    ws[2 * i + 1] = make_float4(wr[4], wr[5], wr[6], 0.f);
  }
  __syncthreads();
  const int b = blockIdx.y, tid = threadIdx.x;
  const int n0 = blockIdx.x * kOut;
  // ragged batch: samples past the item's true length (+ margin) are never looked at; their inputs were not computed
  if (frames != nullptr && n0 > (frames[b] + keep_margin) * samples_per_frame) return;
  const int i = n0 - (K - 1) / 2 + tid;                     // padded position owned by this thread
  const int Lout = L + 1;
  float p[K];
#pragma unroll
  for (int j = 0; j < K; ++j) p[j] = 0.f;
  if (i >= 0 && i < Lout) {
    const int src = (i == 0) ? 1 : i - 1;
    const int opc = PW >> 3;
    uint4 q2[2];
    for (int c8 = 0; c8 < (C >> 3); ++c8) {
      const int panel = c8 / opc, within = c8 - panel * opc;
      if ((c8 & 1) == 0) {                                  // two chunks = 32 contiguous, 32-byte aligned bytes: one 256-bit load
        const uint4* src16 = h16 + (((size_t)b * (C / PW) + panel) * L + src) * opc + within;
        asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(q2[0].x), "=r"(q2[0].y), "=r"(q2[0].z), "=r"(q2[0].w), "=r"(q2[1].x), "=r"(q2[1].y), "=r"(q2[1].z),
                       "=r"(q2[1].w)
                     : "l"(src16));
      }
      const uint4 q = q2[c8 & 1];
      const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 v;
        if (BF16) v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qw[e]));
        else v = __half22float2(*reinterpret_cast<const __half2*>(&qw[e]));
        const float vv[2] = {v.x, v.y};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 w0 = ws[2 * (c8 * 8 + e * 2 + u)], w1 = ws[2 * (c8 * 8 + e * 2 + u) + 1];
          p[0] = fmaf(w0.x, vv[u], p[0]); p[1] = fmaf(w0.y, vv[u], p[1]);
          p[2] = fmaf(w0.z, vv[u], p[2]); p[3] = fmaf(w0.w, vv[u], p[3]);
          p[4] = fmaf(w1.x, vv[u], p[4]); p[5] = fmaf(w1.y, vv[u], p[5]);
          p[6] = fmaf(w1.z, vv[u], p[6]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < K; ++j) part[j * 256 + tid] = p[j];
  __syncthreads();
  const int n = n0 + tid;
  if (tid >= kOut || n >= Lout) return;
  float acc = __ldg(bias);
#pragma unroll
  for (int j = 0; j < K; ++j) acc += part[j * 256 + tid + j];
  const float v = tanhf(acc);
  const size_t o = (size_t)b * Lout + n;
  if (y_dtype == SA_DTYPE_F32) reinterpret_cast<float*>(y)[o] = v;
  else if (y_dtype == SA_DTYPE_F16) reinterpret_cast<__half*>(y)[o] = __float2half_rn(v);
  else {
    float s = rintf(v * 32767.f);
    s = fminf(fmaxf(s, -32768.f), 32767.f);
    reinterpret_cast<int16_t*>(y)[o] = (int16_t)s;
  }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kStaticSmemReserve = 2048;   // >= static __shared__ of every kernel (tile_pre: (kMaxMapItems + 1) ints)
constexpr int kRaggedMarginFrames = 24;    // > receptive field of the generator in frames (20) + the reflect-pad sample

struct Plan {          // launch geometry of one conv on the tensor cores
  int msub, rows_alloc, box_rows, nseg, k16_per_stage, n_wstages, w_resident, n_abuf;
  size_t smem;
};

int panel_width(int c) { return c >= 64 ? 64 : c; }

constexpr size_t kResidentWeightBytes = 48 * 1024;   // keep the whole filter bank in smem below this

bool make_plan(Plan& pl, int cin_pad, int n, int span, int m_rows, int n_k16_max, int max_smem) {
  const size_t fixed = (size_t)n * 4 + (8 + 2 * tc::kMaxStages) * 8 + 16 + 1024;   // + alignment slack
  const int spp = panel_width(cin_pad) / 16;                    // weight stages hold whole panels
  // 128-row tiles for the widest layers let the A tile and the accumulator double-buffer (256-row tiles of
  // C = 256 fill shared memory and TMEM); SATOOLS_B200_MSUB_WIDE=2 restores 256-row tiles.
  static const int msub_wide = getenv("SATOOLS_B200_MSUB_WIDE") ? atoi(getenv("SATOOLS_B200_MSUB_WIDE")) : 1;
  const int msub_max = (m_rows > 128 ? 2 : 1);
  for (int msub = (cin_pad >= 256 && n >= 256) ? std::min(msub_max, msub_wide) : msub_max; msub >= 1; --msub) {
    const int rows = 128 * msub + span;
    const int nseg = (rows + 255) / 256;
    const int box_rows = (int)align_up((size_t)(rows + nseg - 1) / nseg, 8);
    if (box_rows > 256) continue;
    const int rows_alloc = nseg * box_rows;
    const size_t a_bytes = (size_t)cin_pad * 2 * rows_alloc;
    for (int n_abuf = 2; n_abuf >= 1; --n_abuf) {
      auto accept = [&](int k16, int stages, int resident) {
        pl.msub = msub; pl.rows_alloc = rows_alloc; pl.box_rows = box_rows; pl.nseg = nseg;
        pl.k16_per_stage = k16; pl.n_wstages = stages; pl.w_resident = resident; pl.n_abuf = n_abuf;
        pl.smem = n_abuf * a_bytes + (size_t)stages * k16 * n * 32 + fixed;
      };
      const size_t w_all = (size_t)n_k16_max * n * 32;
      if (w_all <= kResidentWeightBytes) {
        int k16 = std::max((n_k16_max + tc::kMaxStages - 1) / tc::kMaxStages, std::max(1, 8 * 1024 / (n * 32)));
        k16 = (k16 + spp - 1) / spp * spp;
        const int stages = (n_k16_max + k16 - 1) / k16;
        if (n_abuf * a_bytes + (size_t)stages * k16 * n * 32 + fixed <= (size_t)max_smem) {
          accept(k16, stages, 1);
          return true;
        }
      }
      const int k16 = (std::max(1, 16 * 1024 / (n * 32)) + spp - 1) / spp * spp;
      for (int stages = 4; stages >= 2; --stages)
        if (n_abuf * a_bytes + (size_t)stages * k16 * n * 32 + fixed <= (size_t)max_smem) {
          accept(k16, stages, 0);
          return true;
        }
    }
  }
  return false;
}

template <int N, int MSUB, int PW>
cudaError_t launch_one(const tc::ConvParams& p_in, int grid_y, size_t smem, int n_sm, cudaStream_t st) {
  static int occ_cache[16] = {0};
  static size_t occ_smem[16] = {0};
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc::conv_tc_kernel<N, MSUB, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - kStaticSmemReserve);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  if (occ_cache[dev] == 0 || occ_smem[dev] != smem) {
    int occ = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tc::conv_tc_kernel<N, MSUB, PW>, tc::kThreads, smem);
    if (e != cudaSuccess) return e;
    constexpr int acc_cols = N * MSUB * ((2 * N * MSUB <= 512) ? 2 : 1);
    constexpr int tmem_cols = acc_cols <= 32 ? 32 : acc_cols <= 64 ? 64 : acc_cols <= 128 ? 128 : acc_cols <= 256 ? 256 : 512;
    occ_cache[dev] = std::max(1, std::min(occ, 512 / tmem_cols));
    occ_smem[dev] = smem;
  }
  tc::ConvParams p = p_in;
  const int ctas = std::max(1, std::min(p.total_tiles, n_sm * occ_cache[dev] / grid_y));
  dim3 grid((unsigned)ctas, (unsigned)grid_y, 1);
  tc::conv_tc_kernel<N, MSUB, PW><<<grid, tc::kThreads, smem, st>>>(p);
  return cudaGetLastError();
}

cudaError_t dispatch(int n, int msub, int pw, const tc::ConvParams& p, int grid_y, size_t smem, int n_sm, cudaStream_t st) {
#define SA_CASE(NN, PP)                                                                   \
  if (n == NN && pw == PP)                                                                \
    return msub == 2 ? launch_one<NN, 2, PP>(p, grid_y, smem, n_sm, st) : launch_one<NN, 1, PP>(p, grid_y, smem, n_sm, st);
  SA_CASE(256, 64) SA_CASE(128, 64) SA_CASE(64, 64) SA_CASE(32, 64) SA_CASE(16, 64)
  SA_CASE(32, 32) SA_CASE(16, 32) SA_CASE(16, 16)
#undef SA_CASE
  return cudaErrorInvalidValue;
}

// ---- fused ResBlock (chain) launch -------------------------------------------------------
// tc_context::chain_ms_narrow: sub-tiles per CTA for C <= 32 (SATOOLS_B200_CHAIN_MS=3 -> two CTAs per SM)

struct ChainPlan { int ms, wps, n_slots; size_t smem; };

bool chain_plan(ChainPlan& pl, const tc_chain& ch, int max_smem, int ms_narrow) {
  const int C = ch.c;
  const size_t stage = (size_t)ch.k16_per_stage * C * 32;
  auto need = [&](int ms, int slots) {
    const size_t rows = (size_t)ms * 128 + 2 * tc::kChainPad;
    return 2 * rows * (size_t)C * 2 + slots * stage + (size_t)tc::kChainMaxConvs * C * 4 +
           (40 + 2 * tc::kChainMaxSlots) * 8 + 16 + 1024;
  };
  if (ch.k != 3 && ch.k != 7 && ch.k != 11) return false;          // instantiated tap counts
  // C = 64: 3 sub-tiles with 8 epilogue warps each (an epilogue then fits inside one sub-tile's MMA time) and as
  // many weight-tap slots as shared memory allows (>= K); C <= 32: 6 sub-tiles, 4 warps each, 2 whole-conv slots.
  const int ms = (C == 64) ? 3 : ms_narrow;
  const int wps = (C == 64) ? 8 : 4;
  int slots = ch.stages_per_conv == 1 ? 2 : ch.stages_per_conv;
  if (need(ms, slots) > (size_t)max_smem) return false;
  if (C == 64)                                       // sub-tile-major order needs >= one conv of taps; take what fits
    while (slots < tc::kChainMaxSlots && slots < 2 * ch.stages_per_conv && need(ms, slots + 1) <= (size_t)max_smem) ++slots;
  if (ms * 128 - 2 * ch.halo < 64) return false;
  pl.ms = ms; pl.wps = wps; pl.n_slots = slots; pl.smem = need(ms, slots);
  return true;
}

template <int C, int MS, int K, int WPS, bool RT = false>
cudaError_t launch_chain(const tc::ChainParams& p, size_t smem, int n_sm, cudaStream_t st) {
  static bool attr_set[16] = {false};
  static int occ_cache[16] = {0};
  static size_t occ_smem[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc::resblock_chain_kernel<C, MS, K, WPS, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - kStaticSmemReserve);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  if (occ_cache[dev] == 0 || occ_smem[dev] != smem) {
    int occ = 1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tc::resblock_chain_kernel<C, MS, K, WPS, RT>, tc::chain_threads(MS, WPS), smem);
    if (e != cudaSuccess) return e;
    constexpr int need = 2 * MS * C;
    constexpr int cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    occ_cache[dev] = std::max(1, std::min(occ, 512 / cols));
    occ_smem[dev] = smem;
  }
  const int ctas = std::max(1, std::min(p.total_tiles, n_sm * occ_cache[dev]));
  tc::resblock_chain_kernel<C, MS, K, WPS, RT><<<ctas, tc::chain_threads(MS, WPS), smem, st>>>(p);
  return cudaGetLastError();
}

// ---- whole-stage fused launch (three ResBlocks in one kernel, C <= 32) -------------------------
// tc_context::use_chain3: SATOOLS_B200_CHAIN3=0 falls back to one fused kernel per ResBlock

template <int C, int MS>
cudaError_t launch_chain3(const tc::Chain3Params& p, size_t smem, int n_sm, cudaStream_t st) {
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc::stage_chain3_kernel<C, MS, 3, 7, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - kStaticSmemReserve);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int ctas = std::max(1, std::min(p.total_tiles, n_sm));
  tc::stage_chain3_kernel<C, MS, 3, 7, 11><<<ctas, tc::chain_threads(MS, 4), smem, st>>>(p);
  return cudaGetLastError();
}

// ---- grouped (block-Toeplitz) fused ResBlock for C <= 32 (chain_group_tc.cuh) -------------------
// tc_context::use_group: SATOOLS_B200_GROUP=0: C <= 32 on the per-tap kernels (chain_tc / chain3_tc), C = 64 without RT

template <int C, bool BF16, int NS, int MS>
cudaError_t launch_group(const tc::GroupParams& p, size_t smem, int n_sm, cudaStream_t st) {
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc::group_chain_kernel<C, BF16, NS, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - kStaticSmemReserve);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int ctas = std::max(1, std::min((p.total_tiles + NS - 1) / NS, n_sm));     // NS tiles in flight per CTA
  tc::group_chain_kernel<C, BF16, NS, MS><<<ctas, tc::kGrpThreads, smem, st>>>(p);
  return cudaGetLastError();
}

// first epilogue mapping (chain_group_v1_tc.cuh): one thread per row and stream; used for C = 32
template <int C, bool BF16, int NS, int MS>
cudaError_t launch_group_v1(const tc::GroupParams& p, size_t smem, int n_sm, cudaStream_t st) {
  static bool attr_set[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc::v1::group_chain_kernel<C, BF16, NS, MS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - kStaticSmemReserve);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const int ctas = std::max(1, std::min((p.total_tiles + NS - 1) / NS, n_sm));
  tc::v1::group_chain_kernel<C, BF16, NS, MS><<<ctas, tc::kGrpThreads, smem, st>>>(p);
  return cudaGetLastError();
}

// ---- CTA-pair kernel (conv_pair_tc.cuh): plan + launch ------------------------------------
int use_pair() {              // SATOOLS_B200_PAIR=0: wide layers on the single-CTA kernel
  static const int v = getenv("SATOOLS_B200_PAIR") ? atoi(getenv("SATOOLS_B200_PAIR")) : 1;
  return v;
}
bool make_plan_pair(Plan& pl, int cin_pad, int n, int span, int m_rows, int max_smem) {
  const size_t fixed = (size_t)n * 4 + (8 + 2 * tc::kMaxStages) * 8 + 16 + 1024;
  const size_t block_bytes = (size_t)(n / 2) * 128;              // one [N/2][64] weight block (4 K-steps)
  const int bps = (int)std::max<size_t>(1, 16 * 1024 / block_bytes);
  const size_t stage_bytes = bps * block_bytes;
  const int msub_max = (n == 128 && m_rows > 128) ? 2 : 1;       // 2 x N accumulator columns must fit TMEM twice
  for (int msub = msub_max; msub >= 1; --msub) {
    const int rows = 128 * msub + span;
    const int nseg = (rows + 255) / 256;
    const int box_rows = (int)align_up((size_t)(rows + nseg - 1) / nseg, 8);
    if (box_rows > 256) continue;
    const int rows_alloc = nseg * box_rows;
    const size_t a_bytes = (size_t)cin_pad * 2 * rows_alloc;
    // A double-buffered when at least three weight stages still fit (two cannot hide the L2 latency of a 16 KB stage)
    for (int n_abuf = 2; n_abuf >= 1; --n_abuf) {
      if (n_abuf * a_bytes + fixed >= (size_t)max_smem) continue;
      const int stages = (int)std::min<size_t>(tc::kMaxStages, ((size_t)max_smem - fixed - n_abuf * a_bytes) / stage_bytes);
      if (stages < (n_abuf == 2 ? 3 : 2)) continue;
      pl.msub = msub; pl.rows_alloc = rows_alloc; pl.box_rows = box_rows; pl.nseg = nseg;
      pl.k16_per_stage = 4 * bps; pl.n_wstages = stages; pl.w_resident = 0; pl.n_abuf = n_abuf;
      pl.smem = n_abuf * a_bytes + stages * stage_bytes + fixed;
      return true;
    }
  }
  return false;
}

template <int N, int MSUB>
cudaError_t launch_pair(const tc::ConvParams& p, int grid_y, size_t smem, int n_sm, cudaStream_t st) {
  static bool attr_set[16] = {false};
  static int max_clusters[16] = {0};
  static size_t mc_smem[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc::conv_pair_kernel<N, MSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - kStaticSmemReserve);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(tc::kPairThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (max_clusters[dev] == 0 || mc_smem[dev] != smem) {     // co-resident pairs (a GPC with an odd SM count strands one)
    cfg.gridDim = dim3(2 * (unsigned)n_sm, 1, 1);
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, tc::conv_pair_kernel<N, MSUB>, &cfg) != cudaSuccess || nc < 1) nc = n_sm / 2 - 4;
    max_clusters[dev] = nc;
    mc_smem[dev] = smem;
  }
  const int pairs = std::max(1, std::min((p.total_tiles + 1) / 2, max_clusters[dev] / grid_y));
  cfg.gridDim = dim3(2 * (unsigned)pairs, (unsigned)grid_y, 1);
  return cudaLaunchKernelEx(&cfg, tc::conv_pair_kernel<N, MSUB>, p);
}

cudaError_t dispatch_pair(int n, int msub, const tc::ConvParams& p, int grid_y, size_t smem, int n_sm, cudaStream_t st) {
  if (n == 256 && msub == 1) return launch_pair<256, 1>(p, grid_y, smem, n_sm, st);
  if (n == 128 && msub == 2) return launch_pair<128, 2>(p, grid_y, smem, n_sm, st);
  if (n == 128 && msub == 1) return launch_pair<128, 1>(p, grid_y, smem, n_sm, st);
  return cudaErrorInvalidValue;
}

// ---- fused conv1 -> conv2 pair of a C = 128 ResBlock on CTA pairs (resblock_pair_tc.cuh) ----
cudaError_t launch_pair_fused(const tc::PairFuseParams& q, size_t smem, int n_sm, cudaStream_t st) {
  static bool attr_set[16] = {false};
  static int max_clusters[16] = {0};
  static size_t mc_smem[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc::resblock_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - kStaticSmemReserve);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(tc::kPfThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (max_clusters[dev] == 0 || mc_smem[dev] != smem) {
    cfg.gridDim = dim3(2 * (unsigned)n_sm, 1, 1);
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, tc::resblock_pair_kernel, &cfg) != cudaSuccess || nc < 1) nc = n_sm / 2 - 4;
    max_clusters[dev] = nc;
    mc_smem[dev] = smem;
  }
  const int pairs = std::max(1, std::min((q.c2.total_tiles + 1) / 2, max_clusters[dev]));
  cfg.gridDim = dim3(2 * (unsigned)pairs, 1, 1);
  return cudaLaunchKernelEx(&cfg, tc::resblock_pair_kernel, q);
}

struct Epi {
  uint32_t flags = 0;
  const float* res32 = nullptr;
  float* out32 = nullptr;
  float* sum32 = nullptr;
  void* out16 = nullptr;
  float slope_out = 0.1f;
  float n_blocks = 3.f;
};

struct Runner {
  tc_context& ctx;
  const tc_forward_args& a;
  int64_t* launches;
  // cuTensorMapEncodeTiled through the per-handle cache (key: base pointer + dims + box + type + swizzle)
  CUresult tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* ptr, const cuuint64_t* gdim, const cuuint64_t* gstr,
                const cuuint32_t* box, CUtensorMapSwizzle swz) {
    tc_tmap_entry k;
    k.ptr = ptr; k.d0 = gdim[0]; k.d1 = gdim[1]; k.d2 = rank > 2 ? gdim[2] : 0; k.d3 = rank > 3 ? gdim[3] : 0;
    k.b0 = box[0]; k.b1 = box[1]; k.dtype = (int)dt; k.swizzle = (int)swz; k.rank = rank;
    for (const tc_tmap_entry& e : ctx.tmaps)
      if (e.ptr == k.ptr && e.d0 == k.d0 && e.d1 == k.d1 && e.d2 == k.d2 && e.d3 == k.d3 && e.b0 == k.b0 && e.b1 == k.b1 &&
          e.dtype == k.dtype && e.swizzle == k.swizzle && e.rank == k.rank) {
        *out = e.map;
        return CUDA_SUCCESS;
      }
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = reinterpret_cast<encode_tiled_fn>(ctx.encode_fn)(&k.map, dt, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstr, box,
                                                                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return r;
    if (ctx.tmaps.size() >= 512) ctx.tmaps.clear();          // many shapes in one process: start over rather than grow
    ctx.tmaps.push_back(k);
    *out = k.map;
    return CUDA_SUCCESS;
  }
  const int* d_frames = nullptr;                 // device copy of frames_per_item (ragged batches), else nullptr
  tc::TileMapParams tile_map(int rows) const {   // rows = rows per item on the kernel's tile axis
    tc::TileMapParams m;
    m.frames = d_frames; m.n_items = a.B; m.rows_per_frame = rows / a.T; m.margin_frames = kRaggedMarginFrames;
    return m;
  }
  void mark(int tag) { if (a.mark) a.mark(a.mark_ctx, tag, a.stream); }

  // One conv layer (all phases / n-tiles) on the tensor cores.  in16: blocked input with
  // cin_pad channels and l_in rows per item.
  const char* conv(const tc_layer& ly, const void* in16, int l_in, const Epi& e, int tag) {
    const tc_weights& w = *ly.w;
    tc::ConvParams p;
    memset(&p, 0, sizeof(p));
    int span = 0;
    p.tap_step = ly.transposed ? -1 : ly.dil;
    for (int ph = 0; ph < w.n_phases; ++ph) {
      p.n_taps[ph] = w.n_taps[ph];
      p.tap_base[ph] = ly.transposed ? w.tap_base[ph] : -ly.pad;
      const int last = p.tap_base[ph] + (w.n_taps[ph] - 1) * p.tap_step;
      p.row_lo[ph] = std::min(p.tap_base[ph], last);
      span = std::max(span, std::max(p.tap_base[ph], last) - p.row_lo[ph]);
    }
    Plan pl;
    int n_k16_max = 0;
    for (int ph = 0; ph < w.n_phases; ++ph) n_k16_max = std::max(n_k16_max, w.n_taps[ph] * (w.cin_pad / 16));
    if (w.pair) {
      if (!make_plan_pair(pl, w.cin_pad, w.n, span, l_in, ctx.max_smem)) return "conv (CTA pair) does not fit in shared memory";
    } else if (!make_plan(pl, w.cin_pad, w.n, span, l_in, n_k16_max, ctx.max_smem)) {
      return "conv does not fit in shared memory";
    }
    // tensor map over the input activation [B][cin_pad/pw][l_in][pw], swizzle = row bytes
    const int pw = panel_width(w.cin_pad);
    const cuuint64_t rb = (cuuint64_t)pw * 2;
    const cuuint64_t gdim[4] = {(cuuint64_t)pw, (cuuint64_t)l_in, (cuuint64_t)(w.cin_pad / pw), (cuuint64_t)a.B};
    const cuuint64_t gstr[3] = {rb, (cuuint64_t)l_in * rb, (cuuint64_t)(w.cin_pad / pw) * l_in * rb};
    const cuuint32_t box[4] = {(cuuint32_t)pw, (cuuint32_t)pl.box_rows, 1, 1};
    const CUtensorMapSwizzle swz = pw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : pw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                     : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = tmap(&p.tmap, a.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, in16, gdim, gstr, box, swz);
    if (r != CUDA_SUCCESS) {
      snprintf(g_msg, sizeof(g_msg), "cuTensorMapEncodeTiled failed (%d) for L=%d panels=%d box_rows=%d", (int)r, l_in,
               w.cin_pad / pw, pl.box_rows);
      return g_msg;
    }
    if (w.pair) {       // the packed weights as rows of 128 bytes; one box = this CTA's half of one stage
      const cuuint64_t wdim[2] = {64, (cuuint64_t)(w.bytes / 128)};
      const cuuint64_t wstr[1] = {128};
      const cuuint32_t wbox[2] = {64, (cuuint32_t)((pl.k16_per_stage / 4) * (w.n / 2))};
      r = tmap(&p.wmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, w.d_w, wdim, wstr, wbox, CU_TENSOR_MAP_SWIZZLE_NONE);
      if (r != CUDA_SUCCESS) {
        snprintf(g_msg, sizeof(g_msg), "cuTensorMapEncodeTiled failed (%d) for the weights (%zu rows, box %u)", (int)r,
                 w.bytes / 128, wbox[1]);
        return g_msg;
      }
    }
    p.w = w.d_w;
    p.bias = ly.d_bias;
    p.res32 = e.res32; p.out32 = e.out32; p.sum32 = e.sum32; p.out16 = e.out16;
    p.error_flag = ctx.d_error;
    p.timing = (ctx.d_timing && ctx.timing_launches < 64) ? ctx.d_timing + 16 * ctx.timing_launches++ : nullptr;
    p.cin = w.cin_pad;
    p.cout_total = ly.cout;
    p.m_rows = l_in;
    p.out_stride = ly.transposed ? ly.stride : 1;
    p.l_out = l_in * p.out_stride;
    p.n_phases = w.n_phases;
    p.n_tiles = w.n_tiles;
    p.rows_alloc = pl.rows_alloc; p.box_rows = pl.box_rows; p.nseg = pl.nseg;
    p.out_pw = panel_width(ly.cout);
    p.k16_per_stage = pl.k16_per_stage;
    p.n_wstages = pl.n_wstages; p.w_resident = pl.w_resident; p.n_abuf = pl.n_abuf;
    p.m_tiles = (l_in + 128 * pl.msub - 1) / (128 * pl.msub);
    p.total_tiles = p.m_tiles * a.B;
    p.map = tile_map(l_in);
    p.w_tile_bytes = (uint32_t)w.tile_bytes;
    p.flags = e.flags | (a.bf16 ? tc::EPI_BF16 : 0u);
#ifdef SA_DIAG   // diagnostic builds only (nvcc -DSA_DIAG): the shipped library has no switch that changes results
    // diagnostics only (results become wrong): drop epilogue streams to time what each costs
    static const uint32_t dbg_mask = getenv("SATOOLS_B200_DEBUG_EPI_MASK") ? (uint32_t)strtoul(getenv("SATOOLS_B200_DEBUG_EPI_MASK"), nullptr, 16) : 0u;
    if (dbg_mask) {
      static bool warned = false;
      if (!warned) fprintf(stderr, "[satools_b200] SATOOLS_B200_DEBUG_EPI_MASK=%x: epilogue streams dropped, RESULTS ARE WRONG (timing only)\n", dbg_mask);
      warned = true;
      p.flags &= ~dbg_mask;
    }
#endif
    p.slope_out = e.slope_out;
    p.n_blocks = e.n_blocks;
    mark(tag);
    cudaError_t ce = w.pair ? dispatch_pair(w.n, pl.msub, p, w.n_phases * w.n_tiles, pl.smem, a.n_sm, a.stream)
                            : dispatch(w.n, pl.msub, pw, p, w.n_phases * w.n_tiles, pl.smem, a.n_sm, a.stream);
    if (ce != cudaSuccess) return msgf("conv_tc launch: %s", cudaGetErrorString(ce));
    ++*launches;
    return nullptr;
  }

  // conv1 -> lrelu -> conv2 (+ residual and the epilogue streams of conv2) of one dilation step of a C = 128 ResBlock in ONE
  // launch (resblock_pair_tc.cuh; nn.py:169-174).  *done = false: the caller runs the two convs one after the other.
  // SATOOLS_B200_PAIR_FUSE=0 switches it off (read per call, like the other variant switches); so does the per-layer mode
  // (SATOOLS_B200_FUSED=0: a.chains == nullptr).
  const char* pair_fused(const tc_layer& l1, const tc_layer& l2, const void* in16, int L, const Epi& e, int tag, bool* done) {
    *done = false;
    const int on = getenv("SATOOLS_B200_PAIR_FUSE") ? atoi(getenv("SATOOLS_B200_PAIR_FUSE")) : 1;
    const tc_weights& w1 = *l1.w;
    const tc_weights& w2 = *l2.w;
    if (!on || !a.chains || !w1.pair || !w2.pair || w1.n != 128 || w2.n != 128 || w1.n_tiles != 1 || w2.n_tiles != 1) return nullptr;
    if (w1.cin_pad != 128 || w2.cin_pad != 128 || l1.cout != 128 || l2.cout != 128 || l1.transposed || l2.transposed) return nullptr;
    const int k = l1.k;
    // k = 11: the two per-conv launches win (0.545 vs 0.58 ms per step at 64 x 15 s): 7.8 % of both convs is recomputed halo
    // and conv1 alone already runs at the tensor peak.  SATOOLS_B200_PAIR_FUSE_KMAX moves the limit.
    const int kmax = getenv("SATOOLS_B200_PAIR_FUSE_KMAX") ? atoi(getenv("SATOOLS_B200_PAIR_FUSE_KMAX")) : 7;
    if (k > kmax) return nullptr;
    if (l2.k != k || (k & 1) == 0 || k < 3 || (k - 1) / 2 > tc::kPfTPad || l2.dil != 1) return nullptr;
    if (l1.pad != l1.dil * (k - 1) / 2 || l2.pad != (k - 1) / 2) return nullptr;
    if (w1.n_taps[0] != k || w2.n_taps[0] != k || w1.tile_bytes != w2.tile_bytes) return nullptr;
    const int V = tc::kPfRows - (k - 1);
    if (L < 2 * V) return nullptr;                                 // short inputs: the per-layer launches
    const int r1 = l1.dil * (k - 1) / 2;
    const int rows = tc::kPfRows + 2 * r1;
    const int nseg = (rows + 255) / 256;
    const int box_rows = (int)align_up((size_t)(rows + nseg - 1) / nseg, 8);
    if (box_rows > 256) return nullptr;
    const int rows_alloc = nseg * box_rows;
    const size_t a_bytes = (size_t)2 * (2 * rows_alloc * 128);    // two buffers of two panels
    const size_t t_bytes = (size_t)2 * (2 * tc::kPfTRows * 128);   // two buffers of two panels
    const size_t fixed = 2 * 128 * 4 + (12 + 2 * tc::kMaxStages) * 8 + 16 + 1024;
    const size_t block_bytes = 64 * 128;                           // one [N/2][64] weight block
    const int bps = 2;                                             // 16 KB stages; 2 k blocks per conv: whole stages
    const size_t stage_bytes = bps * block_bytes;
    if (a_bytes + t_bytes + fixed + 3 * stage_bytes > (size_t)ctx.max_smem) return nullptr;
    const int stages = (int)std::min<size_t>(tc::kMaxStages, ((size_t)ctx.max_smem - fixed - a_bytes - t_bytes) / stage_bytes);
    tc::PairFuseParams q;
    memset(&q, 0, sizeof(q));
    tc::ConvParams& p = q.c2;
    const cuuint64_t rb = 128;
    const cuuint64_t gdim[4] = {64, (cuuint64_t)L, 2, (cuuint64_t)a.B};
    const cuuint64_t gstr[3] = {rb, (cuuint64_t)L * rb, (cuuint64_t)2 * L * rb};
    const cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    CUresult r = tmap(&p.tmap, a.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, in16, gdim, gstr, box,
                      CU_TENSOR_MAP_SWIZZLE_128B);
    if (r != CUDA_SUCCESS) {
      snprintf(g_msg, sizeof(g_msg), "cuTensorMapEncodeTiled failed (%d) for the fused pair input (L=%d box_rows=%d)", (int)r, L, box_rows);
      return g_msg;
    }
    const cuuint64_t wstr[1] = {128};
    const cuuint32_t wbox[2] = {64, (cuuint32_t)(bps * 64)};
    const cuuint64_t wdim1[2] = {64, (cuuint64_t)(w1.bytes / 128)};
    const cuuint64_t wdim2[2] = {64, (cuuint64_t)(w2.bytes / 128)};
    r = tmap(&q.wmap1, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, w1.d_w, wdim1, wstr, wbox, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r == CUDA_SUCCESS) r = tmap(&p.wmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, w2.d_w, wdim2, wstr, wbox, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(g_msg, sizeof(g_msg), "cuTensorMapEncodeTiled failed (%d) for the fused pair weights", (int)r);
      return g_msg;
    }
    q.bias1 = l1.d_bias;
    q.k = k; q.dil1 = l1.dil; q.V = V;
    p.w = w2.d_w;
    p.bias = l2.d_bias;
    p.res32 = e.res32; p.out32 = e.out32; p.sum32 = e.sum32; p.out16 = e.out16;
    p.error_flag = ctx.d_error;
    p.cin = 128; p.cout_total = 128;
    p.m_rows = L; p.l_out = L; p.out_stride = 1;
    p.n_phases = 1; p.n_tiles = 1;
    p.rows_alloc = rows_alloc; p.box_rows = box_rows; p.nseg = nseg;
    p.out_pw = 64;
    p.k16_per_stage = 4 * bps;
    p.n_wstages = stages;
    p.m_tiles = (L + V - 1) / V;
    p.total_tiles = p.m_tiles * a.B;
    p.map = tile_map(L);
    p.w_tile_bytes = (uint32_t)w2.tile_bytes;
    p.flags = e.flags | (a.bf16 ? tc::EPI_BF16 : 0u);
#ifdef SA_DIAG   // diagnostic builds only: drop epilogue streams to time what each costs (RESULTS ARE WRONG)
    if (const char* dm = getenv("SATOOLS_B200_DEBUG_EPI_MASK")) p.flags &= ~(uint32_t)strtoul(dm, nullptr, 16);
#endif
    p.slope_out = e.slope_out;
    p.n_blocks = e.n_blocks;
    const size_t smem = a_bytes + t_bytes + stages * stage_bytes + fixed;
    mark(tag);
    cudaError_t ce = launch_pair_fused(q, smem, a.n_sm, a.stream);
    if (ce != cudaSuccess) return msgf("resblock_pair launch: %s", cudaGetErrorString(ce));
    ++*launches;
    *done = true;
    return nullptr;
  }

  // One whole ResBlock1 on a narrow stage, fused (chain_tc.cuh).  Returns "" (empty, not an
  // error) when this block must run layer by layer instead.
  // The three ResBlocks of a stage in one kernel (chain3_tc.cuh); *done = false: run them one by one instead.
  const char* chain3(const tc_chain* ch, int nrb, const float* x32, int L, const Epi& e, int tag, bool* done) {
    *done = false;
    if (!ctx.use_chain3 || nrb != 3 || !ch[0].d_w || !ch[1].d_w || !ch[2].d_w) return nullptr;
    const int C = ch[0].c;
    if ((C != 16 && C != 32) || ch[0].k != 3 || ch[1].k != 7 || ch[2].k != 11) return nullptr;
    // C = 32: three per-ResBlock launches (6 sub-tiles each, the k = 11 block split in two) recompute 14 % halo rows
    // instead of 45 % and now win (stage 3: 6.0 -> 5.5 ms); C = 16 stays whole-stage (4.9 vs 5.9 ms).
    static const int c32 = getenv("SATOOLS_B200_CHAIN3_C32") ? atoi(getenv("SATOOLS_B200_CHAIN3_C32")) : 0;
    if (C == 32 && !c32) return nullptr;
    if (ch[0].n_convs != ch[1].n_convs || ch[0].n_convs != ch[2].n_convs) return nullptr;
    // sub-tiles per CTA: more rows per tile amortise the 2 x halo recomputed rows, until registers run out
    static const int ms16 = getenv("SATOOLS_B200_CHAIN3_MS16") ? atoi(getenv("SATOOLS_B200_CHAIN3_MS16")) : 5;   // measured: 4 -> 5.5 ms, 5 -> 5.15, 6 -> 5.2 (spills)
    const int ms = (C == 16) ? ((ms16 >= 4 && ms16 <= 6) ? ms16 : 5) : 3;
    const int halo = std::max(ch[0].halo, std::max(ch[1].halo, ch[2].halo));
    const int valid = ms * 128 - 2 * halo;
    if (valid < 64 || L < 2 * valid) return nullptr;
    const size_t rows = (size_t)ms * 128 + 2 * tc::kChainPad;
    const size_t smem = 6 * rows * (size_t)C * 2 + (size_t)(3 + 7 + 11) * C * C * 2 + 3 * (size_t)tc::kChainMaxConvs * C * 4 + 80 * 8 + 16 + 1024;
    if (smem > (size_t)ctx.max_smem) return nullptr;
    tc::Chain3Params p;
    memset(&p, 0, sizeof(p));
    p.x32 = x32; p.out32 = e.out32; p.out16 = e.out16; p.error_flag = ctx.d_error;
    for (int j = 0; j < 3; ++j) {
      p.w[j] = ch[j].d_w; p.bias[j] = ch[j].d_bias;
      for (int c = 0; c < ch[j].n_convs; ++c) { p.dil[j][c] = ch[j].dil[c]; p.pad[j][c] = ch[j].pad[c]; }
    }
    p.L = L; p.n_convs = ch[0].n_convs; p.halo = halo;
    p.tiles_per_item = (L + valid - 1) / valid;
    p.total_tiles = p.tiles_per_item * a.B;
    p.map = tile_map(L);
    p.flags = (e.flags & (tc::EPI_OUT32 | tc::EPI_OUT16)) | (a.bf16 ? tc::EPI_BF16 : 0u);
    p.slope_out = e.slope_out;
    mark(tag);
    cudaError_t ce = (C == 32) ? launch_chain3<32, 3>(p, smem, a.n_sm, a.stream)
                     : (ms == 6) ? launch_chain3<16, 6>(p, smem, a.n_sm, a.stream)
                     : (ms == 5) ? launch_chain3<16, 5>(p, smem, a.n_sm, a.stream)
                                 : launch_chain3<16, 4>(p, smem, a.n_sm, a.stream);
    if (ce != cudaSuccess) return msgf("stage_chain3 launch: %s", cudaGetErrorString(ce));
    ++*launches;
    *done = true;
    return nullptr;
  }

  // Grouped kernel geometry for convs [c0, c1) of a block: halo (multiple of G), valid positions per tile, ring depth.
  struct GroupPlan { int halo, valid, n_wstages, ns, ms; size_t smem; };
  bool group_plan(GroupPlan& pl, const tc_chain& ch, int L, int c0, int c1) const {
    if (!ctx.use_group || !ch.d_wg || (c0 & 1) || ((c1 - c0) & 1) || c1 <= c0) return false;
    if (ch.k != 3 && ch.k != 7 && ch.k != 11) return false;                  // slice counts the MMA issue loop is instantiated for
    // Two streams of two sub-tiles: 1024 positions per tile for C = 16, 512 for C = 32.  (Four streams of one sub-tile for
    // C = 16 measured slower, 4.67 vs 4.13 ms for stage 4 -- twice the halo share, and the streams run in lockstep on one
    // shared weight ring, so more of them do not decouple anything; the epilogue mapping is now fixed to two streams.)
    pl.ns = 2;
    pl.ms = 4 / pl.ns;
    const int G = 64 / ch.c, R = tc::grp_tile_positions(ch.c, pl.ms);
    int halo = 0;
    for (int c = c0; c < c1; ++c) halo += ch.pad[c];
    halo = (halo + G - 1) / G * G;
    pl.halo = halo;
    pl.valid = R - 2 * halo;
    if (L % G != 0 || pl.valid < R / 2 || L < 2 * pl.valid) return false;    // short sequences: the per-layer path wastes less
    // Little work (single short utterances, the latency path): every launch pays its fixed set-up (zeroing 140 KB of shared
    // memory, barriers, TMEM) and a stage is three launches here instead of one stage_chain3 launch; measured on a 5 s
    // utterance: 1.26 ms grouped vs 0.97 ms per-tap.  SATOOLS_B200_GROUP_MIN_TILES overrides the threshold (tests: 0).
    const int min_tiles_env = getenv("SATOOLS_B200_GROUP_MIN_TILES") ? atoi(getenv("SATOOLS_B200_GROUP_MIN_TILES")) : -1;
    const int min_tiles = min_tiles_env >= 0 ? min_tiles_env : 2 * pl.ns * a.n_sm;
    if (((L + pl.valid - 1) / pl.valid) * a.B < min_tiles) return false;
    const size_t fixed = 2 * (size_t)pl.ns * tc::grp_buf_bytes(pl.ms) + tc::kGrpOnesBytes +
                         (7 * tc::kGrpMaxStreams + 2 * tc::kGrpMaxStages) * 8 + 16 + 1024;
    int stages = std::min(tc::kGrpMaxStages, 2 * ch.g_stages);              // two convs deep: the next conv streams in behind
    while (stages > ch.g_stages && fixed + (size_t)stages * tc::kGrpStageBytes > (size_t)ctx.max_smem) --stages;
    if (stages < ch.g_stages + 1 || fixed + (size_t)stages * tc::kGrpStageBytes > (size_t)ctx.max_smem) return false;
    pl.n_wstages = stages;
    pl.smem = fixed + (size_t)stages * tc::kGrpStageBytes;
    return true;
  }

  // One grouped launch over n_ch ResBlocks of the same stage (n_ch = 1: convs [c0, c1) of one block).  fl[j]: the EPI_*
  // flags of chain j's final epilogue.  *done = false: not applicable, use the per-tap kernels.
  const char* group(const tc_chain* chs, int n_ch, const float* x32, int L, const Epi& e, const uint32_t* fl, int tag, bool* done,
                    int c0 = 0, int c1 = -1, const tc_upgroup* up = nullptr, const void* up_in16 = nullptr) {
    *done = false;
    if (n_ch < 1 || n_ch > tc::kGrpMaxChains) return nullptr;
    if (c1 < 0) c1 = chs[0].n_convs;
    GroupPlan pl, pj;
    if (!group_plan(pl, chs[0], L, c0, c1)) return nullptr;
    int max_stages = chs[0].g_stages;
    for (int j = 1; j < n_ch; ++j) {                              // same tile geometry for all chains: the widest halo
      if (chs[j].c != chs[0].c || chs[j].n_convs != chs[0].n_convs || !group_plan(pj, chs[j], L, c0, c1)) return nullptr;
      for (int c = c0; c < c1; c += 2) if (chs[j].dil[c] != chs[0].dil[c]) return nullptr;
      if (pj.halo > pl.halo) { pl.halo = pj.halo; pl.valid = pj.valid; }
      max_stages = std::max(max_stages, chs[j].g_stages);
    }
    if (n_ch > 1) {                                               // ring deep enough for the widest chain
      const size_t fixed = pl.smem - (size_t)pl.n_wstages * tc::kGrpStageBytes;
      int stages = std::min(tc::kGrpMaxStages, 2 * max_stages);
      while (stages > max_stages && fixed + (size_t)stages * tc::kGrpStageBytes > (size_t)ctx.max_smem) --stages;
      if (stages < max_stages + 1 || fixed + (size_t)stages * tc::kGrpStageBytes > (size_t)ctx.max_smem) return nullptr;
      pl.n_wstages = stages;
      pl.smem = fixed + (size_t)stages * tc::kGrpStageBytes;
      if (L < 2 * pl.valid) return nullptr;
    }
    const tc_chain& ch = chs[0];
    tc::GroupParams p;
    memset(&p, 0, sizeof(p));
    p.x32 = x32; p.sum32 = e.sum32; p.out32 = e.out32; p.out16 = e.out16;
    p.n_chains = n_ch;
    for (int j = 0; j < n_ch; ++j) {
      p.w[j] = static_cast<const uint8_t*>(chs[j].d_wg) + (size_t)c0 * chs[j].g_stages * tc::kGrpStageBytes;
      p.n_slices[j] = chs[j].g_slices; p.stages_per_conv[j] = chs[j].g_stages;
      p.flags[j] = fl[j] | (a.bf16 ? tc::EPI_BF16 : 0u);
    }
    p.error_flag = ctx.d_error;
    p.timing = (ctx.d_timing && ctx.timing_launches < 64) ? ctx.d_timing + 16 * ctx.timing_launches++ : nullptr;
    p.L = L; p.n_convs = c1 - c0;
    for (int c = c0; c < c1; c += 2) p.dil[(c - c0) / 2] = ch.dil[c];
    p.halo = pl.halo;
    if (up && up->d_w && up_in16) {                               // the stage's transposed conv inside this launch
      if (pl.ms != 2 || c0 != 0 || up->stages + 1 > pl.n_wstages) return nullptr;
      // previous stage's output [B][1][L/2][2 C] 16-bit, seen as rows of 128 bytes: L * C * 2 / 128 = L / G rows per item
      const cuuint64_t rows = (cuuint64_t)L / (64 / ch.c);
      const cuuint64_t gdim[3] = {64, rows, (cuuint64_t)a.B};
      const cuuint64_t gstr[2] = {128, rows * 128};
      const cuuint32_t box[3] = {64, (cuuint32_t)(tc::grp_buf_bytes(pl.ms) / 256), 1};
      if (tmap(&p.up_map, a.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, up_in16, gdim, gstr, box,
               CU_TENSOR_MAP_SWIZZLE_128B) != CUDA_SUCCESS)
        return "cuTensorMapEncodeTiled failed for the fused upsampler input";
      p.up_w = up->d_w; p.up_stages = up->stages; p.fuse_up = 1;
    }
    p.n_wstages = pl.n_wstages;
    p.tiles_per_item = (L + pl.valid - 1) / pl.valid;
    p.total_tiles = p.tiles_per_item * a.B;
    p.map = tile_map(L);
    p.slope_out = e.slope_out; p.n_blocks = e.n_blocks;
    mark(tag);
    cudaError_t ce = cudaErrorInvalidValue;
    // Epilogue mapping per channel count: bit 0: C = 16, bit 1: C = 32 on the first mapping (chain_group_v1_tc.cuh: one thread
    // per row and stream).  Default: the half-row mapping for both -- measured on one box after the issue loops went
    // uniform (stage 3: 4.25 vs 4.42-4.55 ms; v1 still wins the k = 3 block, 0.85 vs 0.93 ms, and loses k = 11, 2.25-2.33 vs
    // 2.02 ms).  The two kernels lay the running sum out differently, so a stage never mixes them.
    const int v1_mask = getenv("SATOOLS_B200_GROUP_V1") ? atoi(getenv("SATOOLS_B200_GROUP_V1")) : 0;
    const bool v1 = (v1_mask & (ch.c == 16 ? 1 : 2)) != 0;
#define SA_GROUP(CC, NN, MM)                                                                                          \
    if (ch.c == CC && pl.ns == NN)                                                                                    \
      ce = v1 ? (a.bf16 ? launch_group_v1<CC, true, NN, MM>(p, pl.smem, a.n_sm, a.stream) : launch_group_v1<CC, false, NN, MM>(p, pl.smem, a.n_sm, a.stream)) \
              : (a.bf16 ? launch_group<CC, true, NN, MM>(p, pl.smem, a.n_sm, a.stream) : launch_group<CC, false, NN, MM>(p, pl.smem, a.n_sm, a.stream));
    SA_GROUP(16, 2, 2) SA_GROUP(32, 2, 2)
#undef SA_GROUP
    if (ce != cudaSuccess) return msgf("group_chain launch: %s", cudaGetErrorString(ce));
    ++*launches;
    *done = true;
    return nullptr;
  }

  bool chain_usable(const tc_chain& ch, int L) const {
    ChainPlan pl;
    if (!ch.d_w || !chain_plan(pl, ch, ctx.max_smem, ctx.chain_ms_narrow)) return false;
    return L >= 2 * (pl.ms * 128 - 2 * ch.halo);                // short sequences: the per-layer path wastes less
  }

  // convs [c0, c1) of the block (default: all): a block can be run as two launches to halve the recomputed halo
  const char* chain(const tc_chain& ch, const float* x32, int L, const Epi& e, int tag, bool* done, int c0 = 0, int c1 = -1) {
    *done = false;
    ChainPlan pl;
    if (!chain_usable(ch, L) || !chain_plan(pl, ch, ctx.max_smem, ctx.chain_ms_narrow)) return nullptr;
    if (c1 < 0) c1 = ch.n_convs;
    int halo = 0;
    for (int c = c0; c < c1; ++c) halo += ch.pad[c];
    const int valid = pl.ms * 128 - 2 * halo;
    tc::ChainParams p;
    memset(&p, 0, sizeof(p));
    p.x32 = x32; p.sum32 = e.sum32; p.out32 = e.out32; p.out16 = e.out16;
    p.w = static_cast<const uint8_t*>(ch.d_w) + (size_t)c0 * ch.k * ch.c * ch.c * 2;
    p.bias = ch.d_bias + (size_t)c0 * ch.c; p.error_flag = ctx.d_error;
    p.timing = (ctx.d_timing && ctx.timing_launches < 64) ? ctx.d_timing + 16 * ctx.timing_launches++ : nullptr;
    p.L = L; p.n_convs = c1 - c0; p.ktaps = ch.k;
    for (int c = c0; c < c1; ++c) { p.dil[c - c0] = ch.dil[c]; p.pad[c - c0] = ch.pad[c]; }
    p.halo = halo;
    p.tiles_per_item = (L + valid - 1) / valid;
    p.total_tiles = p.tiles_per_item * a.B;
    p.map = tile_map(L);
    p.k16_per_stage = ch.k16_per_stage; p.stages_per_conv = ch.stages_per_conv; p.n_slots = pl.n_slots;
    p.flags = e.flags | (a.bf16 ? tc::EPI_BF16 : 0u);
#ifdef SA_DIAG
    if (getenv("SATOOLS_B200_DEBUG_FREERUN") && atoi(getenv("SATOOLS_B200_DEBUG_FREERUN"))) p.flags |= 1u << 30;
#endif
    p.slope_out = e.slope_out; p.n_blocks = e.n_blocks;
    mark(tag);
    cudaError_t ce = cudaErrorInvalidValue;
    // C = 64: residual in tensor memory + pipelined tile boundary (chain_tc.cuh, RT); SATOOLS_B200_GROUP=0 / SATOOLS_B200_CHAIN_RT=0
    // keep the form that is bit-identical to the per-layer path
    const int chain_rt = getenv("SATOOLS_B200_CHAIN_RT") ? atoi(getenv("SATOOLS_B200_CHAIN_RT")) : 1;
    const bool rt = chain_rt && ctx.use_group && ch.c == 64;
#define SA_CHAIN(CC, MM, KK, WW) \
    if (ch.c == CC && pl.ms == MM && ch.k == KK && pl.wps == WW) ce = launch_chain<CC, MM, KK, WW>(p, pl.smem, a.n_sm, a.stream);
#define SA_CHAIN_RT(CC, MM, KK, WW) \
    if (ch.c == CC && pl.ms == MM && ch.k == KK && pl.wps == WW)          \
      ce = rt ? launch_chain<CC, MM, KK, WW, true>(p, pl.smem, a.n_sm, a.stream) : launch_chain<CC, MM, KK, WW>(p, pl.smem, a.n_sm, a.stream);
    SA_CHAIN_RT(64, 3, 3, 8) SA_CHAIN_RT(64, 3, 7, 8) SA_CHAIN_RT(64, 3, 11, 8)
    SA_CHAIN(32, 6, 3, 4) SA_CHAIN(32, 6, 7, 4) SA_CHAIN(32, 6, 11, 4)
    SA_CHAIN(16, 6, 3, 4) SA_CHAIN(16, 6, 7, 4) SA_CHAIN(16, 6, 11, 4)
    SA_CHAIN(32, 3, 3, 4) SA_CHAIN(32, 3, 7, 4) SA_CHAIN(32, 3, 11, 4)
    SA_CHAIN(16, 3, 3, 4) SA_CHAIN(16, 3, 7, 4) SA_CHAIN(16, 3, 11, 4)
#undef SA_CHAIN
#undef SA_CHAIN_RT
    if (ce != cudaSuccess) return msgf("resblock_chain launch: %s", cudaGetErrorString(ce));
    ++*launches;
    *done = true;
    return nullptr;
  }
};

}  // namespace

// Whole panels only: the 16-bit activation tensors are [C / PW][L][PW] with PW = min(C, 64), and the kernels, the weight
// packing and the tensor maps all count panels as C / PW.  A channel count that is not 16, 32 or a multiple of 64 is padded
// up to the next such value (conv_pre's input: 257 -> 320, 504 -> 512); padded channels are zero in x and in the weights.
int tc_cin_pad(int cin) { return cin > 64 ? (cin + 63) / 64 * 64 : cin > 32 ? 64 : cin > 16 ? 32 : 16; }

// first_layer: the input is packed by pack_input_kernel (any Cin); otherwise it is the previous layer's output tensor,
// whose channel count must already be a whole number of panels.
bool tc_layer_supported(bool transposed, int cin, int cout, int k, bool first_layer) {
  (void)k;
  if (!first_layer && tc_cin_pad(cin) != cin) return false;
  if (cout != 16 && cout != 32 && cout != 64 && cout != 128 && cout % 256 != 0) return false;   // instantiated N tiles
  if (transposed && cout > 256) return false;
  return true;
}

const char* tc_pack_weights(tc_weights& w, const float* folded, bool transposed, int cin, int cout, int k, int stride,
                            int pad, bool bf16) {
  tc_free_weights(w);
  w.cin_pad = tc_cin_pad(cin);
  w.n = cout > 256 ? 256 : cout;
  w.n_tiles = cout / w.n;
  // Wide layers (N = 256 / 128, 64-channel panels) run on CTA pairs; their weights are packed in half tiles.
  w.pair = (use_pair() && (w.n == 256 || w.n == 128) && w.cin_pad % 64 == 0) ? 1 : 0;
  const int pack_n = w.pair ? w.n / 2 : w.n, pack_tiles = cout / pack_n;
  const int k16_per_tap = w.cin_pad / 16;
  int max_taps = 0;
  if (!transposed) {
    w.n_phases = 1;
    w.n_taps[0] = k;
    w.tap_base[0] = 0;       // -pad applied per launch
    w.tap_step = 1;
    max_taps = k;
  } else {
    if (stride > kTcMaxPhases) return "upsample rate too large for the tensor-core path";
    w.n_phases = stride;
    w.tap_step = -1;
    for (int ph = 0; ph < stride; ++ph) {
      const int j0 = (ph + pad) % stride;
      w.n_taps[ph] = j0 < k ? (k - j0 + stride - 1) / stride : 0;
      w.tap_base[ph] = (ph + pad) / stride;       // tap m reads input row q + tap_base - m
      if (w.n_taps[ph] < 1) return "transposed conv phase without taps";
      max_taps = std::max(max_taps, w.n_taps[ph]);
    }
  }
  // one (phase, n_tile) block = [tap][panel][N rows][pw] with the UMMA K-major swizzle applied
  const int pw = panel_width(w.cin_pad), panels = w.cin_pad / pw, rb = pw * 2;
  const size_t panel_bytes = (size_t)pack_n * rb;
  w.tile_bytes = (size_t)max_taps * panels * panel_bytes;
  w.bytes = w.tile_bytes * w.n_phases * pack_tiles;
  std::vector<uint16_t> host(w.bytes / 2, 0);
  for (int ph = 0; ph < w.n_phases; ++ph)
    for (int t = 0; t < pack_tiles; ++t) {
      uint8_t* tile = reinterpret_cast<uint8_t*>(host.data()) + (size_t)(ph * pack_tiles + t) * w.tile_bytes;
      for (int tap = 0; tap < w.n_taps[ph]; ++tap) {
        const int j = transposed ? ((ph + pad) % stride + stride * tap) : tap;
        for (int pn = 0; pn < panels; ++pn) {
          uint8_t* blk = tile + (size_t)(tap * panels + pn) * panel_bytes;
          for (int r = 0; r < pack_n; ++r)
            for (int kq = 0; kq < pw; ++kq) {
              const int ci = pn * pw + kq, co = t * pack_n + r;
              float v = 0.f;
              if (ci < cin)
                v = transposed ? folded[((size_t)ci * cout + co) * k + j] : folded[((size_t)co * cin + ci) * k + j];
              const uint32_t off = swizzle_offset((uint32_t)r * rb + (uint32_t)kq * 2, rb);
              const uint16_t h = to16(v, bf16);
              memcpy(blk + off, &h, 2);
            }
        }
      }
    }
  TC_CUDA(cudaMalloc(&w.d_w, w.bytes));
  TC_CUDA(cudaMemcpy(w.d_w, host.data(), w.bytes, cudaMemcpyHostToDevice));
  return nullptr;
}

bool tc_chain_supported(int c, int k, int n_convs) {
  if (c != 16 && c != 32 && c != 64) return false;
  if (n_convs < 2 || n_convs > tc::kChainMaxConvs || (n_convs & 1)) return false;
  const int spc = (c == 64) ? k : 1;
  return spc <= tc::kChainMaxSlots && k >= 1;
}

const char* tc_pack_chain(tc_chain& ch, int c, int k, int n_convs, const float* const* folded, const float* const* bias,
                          const int* dil, const int* pad, bool bf16) {
  tc_free_chain(ch);
  if (!tc_chain_supported(c, k, n_convs)) return nullptr;
  ch.c = c; ch.k = k; ch.n_convs = n_convs;
  ch.halo = 0;
  for (int i = 0; i < n_convs; ++i) { ch.dil[i] = dil[i]; ch.pad[i] = pad[i]; ch.halo += pad[i]; }
  for (int i = 0; i < n_convs; ++i) if (pad[i] > 25 || (k - 1) * dil[i] - pad[i] > 25) return nullptr;   // tap reach > slack rows
  const int k16_per_tap = c / 16;
  if (c == 64) {                                                              // kChainTapsPerStage64 taps (16 KB) per stage
    ch.k16_per_stage = tc::kChainTapsPerStage64 * k16_per_tap;
    ch.stages_per_conv = (k + tc::kChainTapsPerStage64 - 1) / tc::kChainTapsPerStage64;
  }
  else { ch.k16_per_stage = k * k16_per_tap; ch.stages_per_conv = 1; }         // the whole conv per stage
  const int rb = 2 * c;                                          // one panel: row = all C channels
  const size_t tap_bytes = (size_t)c * rb, conv_bytes = (size_t)k * tap_bytes;
  std::vector<uint16_t> host(conv_bytes * n_convs / 2);
  for (int cv = 0; cv < n_convs; ++cv)
    for (int tap = 0; tap < k; ++tap) {
      uint8_t* blk = reinterpret_cast<uint8_t*>(host.data()) + cv * conv_bytes + tap * tap_bytes;
      for (int r = 0; r < c; ++r)
        for (int ci = 0; ci < c; ++ci) {
          const uint16_t h = to16(folded[cv][((size_t)r * c + ci) * k + tap], bf16);
          memcpy(blk + swizzle_offset((uint32_t)r * rb + (uint32_t)ci * 2, rb), &h, 2);
        }
    }
  std::vector<float> hb((size_t)n_convs * c);
  for (int cv = 0; cv < n_convs; ++cv) memcpy(hb.data() + (size_t)cv * c, bias[cv], c * sizeof(float));
  TC_CUDA(cudaMalloc(&ch.d_w, host.size() * 2));
  TC_CUDA(cudaMemcpy(ch.d_w, host.data(), host.size() * 2, cudaMemcpyHostToDevice));
  TC_CUDA(cudaMalloc(reinterpret_cast<void**>(&ch.d_bias), hb.size() * sizeof(float)));
  TC_CUDA(cudaMemcpy(ch.d_bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
  // Grouped packing (chain_group_tc.cuh; tools/grouped_chain_model.py is the executable specification): G = 64 / C
  // positions per 128-byte row; slice q = (position offset c = q / (C/16), channel block h = q % (C/16)) holds
  //   B_q[g * C + co][kk] = W[co][h * 16 + kk][j = c - g]   for 0 <= j < k, else 0.
  // Needs every conv1 padded by (k-1)/2 * d and every conv2 undilated (nn.py:96-166), at most kGrpMaxPairs pairs.
  bool grouped = (c == 16 || c == 32) && (k & 1) && n_convs / 2 <= tc::kGrpMaxPairs;
  for (int cv = 0; cv < n_convs && grouped; ++cv)
    grouped = pad[cv] == (k - 1) / 2 * dil[cv] && ((cv & 1) == 0 || dil[cv] == 1);
  if (grouped) {
    const int G = 64 / c, cpp = c / 16;
    ch.g_slices = (G + k - 1) * cpp;
    ch.g_stages = (ch.g_slices + 1 + tc::kGrpSlicesPerStage - 1) / tc::kGrpSlicesPerStage;     // + the bias slice
    const size_t conv_g = (size_t)ch.g_stages * tc::kGrpStageBytes;
    std::vector<uint8_t> hg(conv_g * n_convs, 0);
    for (int cv = 0; cv < n_convs; ++cv)
      for (int q = 0; q < ch.g_slices; ++q) {
        uint8_t* blk = hg.data() + cv * conv_g + (size_t)q * tc::kGrpSliceBytes;
        const int cpos = q / cpp, h = q % cpp;
        for (int g = 0; g < G; ++g) {
          const int j = cpos - g;
          if (j < 0 || j >= k) continue;
          for (int co = 0; co < c; ++co)
            for (int kk = 0; kk < 16; ++kk) {
              const uint16_t hv = to16(folded[cv][((size_t)co * c + h * 16 + kk) * k + j], bf16);
              memcpy(blk + swizzle_offset((uint32_t)(g * c + co) * 32u + (uint32_t)kk * 2u, 32), &hv, 2);
            }
        }
      }
    // the bias slice (its A operand is the kernel's ones tile: 1.0 in K columns 0 and 1): bias = hi + lo in 16 bits each
    for (int cv = 0; cv < n_convs; ++cv) {
      uint8_t* blk = hg.data() + cv * conv_g + (size_t)ch.g_slices * tc::kGrpSliceBytes;
      for (int n = 0; n < 64; ++n) {
        const float bv = bias[cv][n % c];
        const uint16_t hi = to16(bv, bf16);
        float hif;
        if (bf16) { __nv_bfloat16 t; memcpy(&t, &hi, 2); hif = __bfloat162float(t); } else { __half t; memcpy(&t, &hi, 2); hif = __half2float(t); }
        const uint16_t lo = to16(bv - hif, bf16);
        memcpy(blk + swizzle_offset((uint32_t)n * 32u, 32), &hi, 2);
        memcpy(blk + swizzle_offset((uint32_t)n * 32u + 2u, 32), &lo, 2);
      }
    }
    TC_CUDA(cudaMalloc(&ch.d_wg, hg.size()));
    TC_CUDA(cudaMemcpy(ch.d_wg, hg.data(), hg.size(), cudaMemcpyHostToDevice));
  }
  return nullptr;
}

void tc_free_chain(tc_chain& ch) {
  if (ch.d_w) cudaFree(ch.d_w);
  if (ch.d_wg) cudaFree(ch.d_wg);
  if (ch.d_bias) cudaFree(ch.d_bias);
  ch = tc_chain();
}

// Transposed conv (weight [Cin][Cout][k], k = 4, stride u = 2, pad p = 1, Cin = 2 C) as block-Toeplitz slices for the grouped
// kernel.  A 128-byte output row holds G = 64 / C output positions n0 + g'; it reads the input positions i0 - 1 + c,
// c = 0 .. G/2 + 1 (i0 = n0 / 2), each 2 C channels = 2 C / 16 slices; tap j = (n0 + g') + p - u (i0 - 1 + c) = g' + 3 - 2 c.
const char* tc_pack_upgroup(tc_upgroup& u, const float* folded, const float* bias, int cin, int cout, int k, int stride, int pad,
                            bool bf16) {
  tc_free_upgroup(u);
  if ((cout != 16 && cout != 32) || cin != 2 * cout || k != 4 || stride != 2 || pad != 1) return nullptr;
  const int C = cout, G = 64 / C, blocks = cin / 16, n_pos = G / 2 + 2;
  u.slices = n_pos * blocks;
  u.stages = (u.slices + 1 + tc::kGrpSlicesPerStage - 1) / tc::kGrpSlicesPerStage;
  std::vector<uint8_t> h((size_t)u.stages * tc::kGrpStageBytes, 0);
  auto put = [&](uint8_t* blk, int n, int kk, float v) {
    const uint16_t hv = to16(v, bf16);
    memcpy(blk + swizzle_offset((uint32_t)n * 32u + (uint32_t)kk * 2u, 32), &hv, 2);
  };
  for (int q = 0; q < u.slices; ++q) {
    uint8_t* blk = h.data() + (size_t)q * tc::kGrpSliceBytes;
    const int c = q / blocks, hb = q % blocks;
    for (int g = 0; g < G; ++g) {
      const int j = g + pad + stride - stride * c;               // g' + p - u (c - 1)
      if (j < 0 || j >= k) continue;
      for (int co = 0; co < C; ++co)
        for (int kk = 0; kk < 16; ++kk) put(blk, g * C + co, kk, folded[((size_t)(hb * 16 + kk) * cout + co) * k + j]);
    }
  }
  uint8_t* bb = h.data() + (size_t)u.slices * tc::kGrpSliceBytes;  // bias block: hi + lo halves against the kernel's ones tile
  for (int n = 0; n < 64; ++n) {
    const float bv = bias[n % C];
    const uint16_t hi = to16(bv, bf16);
    float hif;
    if (bf16) { __nv_bfloat16 t; memcpy(&t, &hi, 2); hif = __bfloat162float(t); } else { __half t; memcpy(&t, &hi, 2); hif = __half2float(t); }
    put(bb, n, 0, hif);
    put(bb, n, 1, bv - hif);
  }
  TC_CUDA(cudaMalloc(&u.d_w, h.size()));
  TC_CUDA(cudaMemcpy(u.d_w, h.data(), h.size(), cudaMemcpyHostToDevice));
  return nullptr;
}

void tc_free_upgroup(tc_upgroup& u) {
  if (u.d_w) cudaFree(u.d_w);
  u = tc_upgroup();
}

void tc_free_weights(tc_weights& w) {
  if (w.d_w) cudaFree(w.d_w);
  w = tc_weights();
}

const char* tc_init(tc_context& ctx, int device) {
  if (ctx.ready) return nullptr;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  TC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return "cuTensorMapEncodeTiled not available from the driver";
  ctx.encode_fn = fn;
  TC_CUDA(cudaDeviceGetAttribute(&ctx.max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  ctx.max_smem -= kStaticSmemReserve;            // the kernels' static shared memory (live-tile prefix of ragged batches)
  // error flag in mapped pinned host memory: kernels can raise it, the host reads it without a sync
  TC_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx.h_error), sizeof(int), cudaHostAllocMapped));
  *ctx.h_error = 0;
  TC_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx.d_error), ctx.h_error, 0));
  if (const char* env = getenv("SATOOLS_B200_CHAIN_TIMING")) {
    if (atoi(env) != 0) TC_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx.d_timing), 64 * 16 * sizeof(long long)));
  }
  if (const char* env = getenv("SATOOLS_B200_MBAR_NS")) {
    const uint32_t v = (uint32_t)atoi(env);
    TC_CUDA(cudaMemcpyToSymbol(tc::g_mbar_suspend_ns, &v, sizeof(v)));
  }
  ctx.use_chain3 = getenv("SATOOLS_B200_CHAIN3") ? atoi(getenv("SATOOLS_B200_CHAIN3")) : 1;
  ctx.use_group = getenv("SATOOLS_B200_GROUP") ? atoi(getenv("SATOOLS_B200_GROUP")) : 1;
  ctx.chain_ms_narrow = 6;
  if (const char* env = getenv("SATOOLS_B200_CHAIN_MS")) {
    const int v = atoi(env);
    if (v == 3 || v == 6) ctx.chain_ms_narrow = v;
  }
  ctx.ready = true;
  return nullptr;
}

int tc_read_chain_timing(tc_context& ctx, long long* out, int max_launches) {
  if (!ctx.d_timing) return 0;
  const int n = std::min(ctx.timing_launches, max_launches);
  cudaDeviceSynchronize();
  cudaMemcpy(out, ctx.d_timing, (size_t)n * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  return n;
}

bool tc_error_raised(const tc_context& ctx) {
  return ctx.h_error && *static_cast<volatile int*>(ctx.h_error) != 0;
}

namespace {
struct Sizes { size_t e16, e32, xin, frames; };
Sizes sizes(const sa_hifigan_cfg& cfg, int B, int T) {
  int64_t per_frame = cfg.initial_channels, rate = 1;
  for (int i = 0; i < cfg.n_stages; ++i) {
    rate *= cfg.upsample_rates[i];
    per_frame = std::max<int64_t>(per_frame, (int64_t)(cfg.initial_channels >> (i + 1)) * rate);
  }
  const size_t E = (size_t)B * T * per_frame;
  Sizes s;
  s.e16 = align_up(E * 2, 256);
  s.e32 = align_up(E * 4 + 8192, 256);          // + the last partial 32-row block of the grouped kernels' sum layout
  s.xin = align_up((size_t)B * T * tc_cin_pad(cfg.input_dim) * 2, 256);
  s.frames = align_up((size_t)B * sizeof(int32_t), 256);
  return s;
}
}  // namespace

size_t tc_workspace_bytes(const sa_hifigan_cfg& cfg, int B, int T) {
  const Sizes s = sizes(cfg, B, T);
  return s.xin + 4 * s.e16 + 3 * s.e32 + s.frames;
}

const char* tc_forward(tc_context& ctx, const tc_forward_args& a, int64_t* launches) {
  if (!ctx.ready) return "tensor-core context not initialised";
  if (tc_error_raised(ctx)) return "a tcgen05 kernel of an earlier forward timed out on an mbarrier (protocol bug)";
  if (ctx.d_timing) {
    ctx.timing_launches = 0;
    cudaMemsetAsync(ctx.d_timing, 0, 64 * 16 * sizeof(long long), a.stream);
  }
  const sa_hifigan_cfg& cfg = *a.cfg;
  const Sizes s = sizes(cfg, a.B, a.T);
  char* base = static_cast<char*>(a.workspace);
  void* XIN16 = base; base += s.xin;
  void* P16 = base; base += s.e16;
  void* AX16 = base; base += s.e16;
  void* A16 = base; base += s.e16;
  void* T16 = base; base += s.e16;
  float* X32 = reinterpret_cast<float*>(base); base += s.e32;
  float* R32 = reinterpret_cast<float*>(base); base += s.e32;
  float* S32 = reinterpret_cast<float*>(base); base += s.e32;
  // Ragged batch: the true lengths go to the device once; every launch then enumerates only the live tiles.
  const int* d_frames = nullptr;
  if (a.frames_per_item && a.B <= tc::kMaxMapItems && !(a.debug_out)) {
    bool ragged = false;
    for (int b = 0; b < a.B; ++b) ragged = ragged || a.frames_per_item[b] + kRaggedMarginFrames < a.T;
    if (ragged) {
      TC_CUDA(cudaMemcpyAsync(base, a.frames_per_item, (size_t)a.B * sizeof(int32_t), cudaMemcpyHostToDevice, a.stream));
      d_frames = reinterpret_cast<const int*>(base);
    }
  }
  cudaStream_t st = a.stream;
  Runner run{ctx, a, launches};
  run.d_frames = d_frames;
  const int nst = cfg.n_stages, nrb = cfg.n_resblocks, nd = cfg.n_dilations;
  auto L_up = [&](int i) { return 1 + i; };
  auto L_rb = [&](int i, int j, int which, int m) { return 1 + nst + ((i * nrb + j) * 2 + which) * nd + m; };
  const int L_post = 1 + nst + nst * nrb * 2 * nd;
  const char* err;

  float* Hout = X32;                                // where the fp32 output of the running section lives
  auto unblock_tap = [&](int tap, int C, int L) -> const char* {
    if (!a.debug_out || a.debug_tap != tap) return nullptr;
    dim3 g((unsigned)((L + 255) / 256), (unsigned)(C / 8), (unsigned)a.B);
    run.mark(15);
    unblock_kernel<<<g, 256, 0, st>>>(Hout, a.debug_out, C, L);
    ++*launches;
    TC_CUDA(cudaGetLastError());
    return nullptr;
  };

  // pack x -> 16-bit blocked
  {
    const int cpad = tc_cin_pad(cfg.input_dim);
    dim3 g((unsigned)((a.T + 127) / 128), (unsigned)(cpad / 8), (unsigned)a.B);
    run.mark(15);
    if (a.x)
      pack_input_kernel<<<g, 128, 0, st>>>(a.x, reinterpret_cast<uint4*>(XIN16), cfg.input_dim, cpad / 8, a.T, a.bf16 ? 1 : 0,
                                           panel_width(cpad));
    else if (a.bn)
      pack_parts_kernel<<<g, 128, 0, st>>>(a.bn, a.f0, a.spk, reinterpret_cast<uint4*>(XIN16), a.n_bn, a.n_spk, cpad / 8, a.T,
                                           a.bf16 ? 1 : 0, panel_width(cpad));
    else
      pack_vq_kernel<<<g, 128, 0, st>>>(a.vq_idx, a.codebook, a.f0, a.spk_ids, reinterpret_cast<uint4*>(XIN16), a.n_codes, a.n_bn,
                                        a.n_spk, cpad / 8, a.T, a.bf16 ? 1 : 0, panel_width(cpad));
    ++*launches;
    TC_CUDA(cudaGetLastError());
  }
  // conv_pre (archi.py:78); its consumer applies lrelu(0.1) (archi.py:80)
  {
    Epi e;
    e.flags = tc::EPI_OUT16 | ((a.debug_out && a.debug_tap == SA_TAP_CONV_PRE) ? tc::EPI_OUT32 : 0u);
    e.out16 = P16; e.out32 = X32; e.slope_out = 0.1f;
    if ((err = run.conv(a.layers[0], XIN16, a.T, e, 0))) return err;
    if ((err = unblock_tap(SA_TAP_CONV_PRE, cfg.initial_channels, a.T))) return err;
  }
  int L = a.T;
  void* Pcur = P16;                                 // where the 16-bit output of the previous section lives
  // The last stage hands conv_post its lrelu(0.01)-activated output in the 16-bit blocked layout (half the bytes of
  // the fp32 stream, read once); other filter lengths keep the fp32 hand-off.
  const bool post16 = a.layers[L_post].k == 7 && a.layers[L_post].cin % 16 == 0;
  for (int i = 0; i < nst; ++i) {
    const tc_layer& up = a.layers[L_up(i)];
    bool all_fused = a.chains != nullptr;
    bool stage_grouped = a.chains != nullptr;       // C <= 32: every ResBlock of the stage on the grouped kernel
    for (int j = 0; j < nrb && stage_grouped; ++j) {
      Runner::GroupPlan gp;
      stage_grouped = run.group_plan(gp, a.chains[i * nrb + j], L * up.stride, 0, a.chains[i * nrb + j].n_convs);
    }
    // Grouped stages with a k = 4 / stride 2 upsampler run the transposed conv inside the grouped kernel (fuse_up): no
    // upsampler launch and no fp32 stage input in HBM.  The stage then reads Pcur while it writes its own 16-bit output, so
    // that goes to the other buffer.  SATOOLS_B200_GROUP_UP=0 keeps the separate upsampler launch.
    // (the variant switches are read per call, not once per process: a new handle / the tests can change them)
    const int group_up = getenv("SATOOLS_B200_GROUP_UP") ? atoi(getenv("SATOOLS_B200_GROUP_UP")) : 1;
    bool fuse_up = false;
    if (stage_grouped && group_up && a.upg && a.upg[i].d_w) {
      Runner::GroupPlan gp;
      fuse_up = run.group_plan(gp, a.chains[i * nrb], L * up.stride, 0, a.chains[i * nrb].n_convs) && gp.ms == 2 &&
                a.upg[i].stages + 1 <= gp.n_wstages;
    }
    void* const Pin = Pcur;
    void* const Pout = (fuse_up && Pcur == P16) ? T16 : P16;
    if (!fuse_up) {                                 // archi.py:80-81
      // The fused ResBlock kernels read the fp32 stage input only; lrelu(x) in 16 bits is for the per-layer convs.
      for (int j = 0; j < nrb && all_fused && !stage_grouped; ++j) all_fused = run.chain_usable(a.chains[i * nrb + j], L * up.stride);
      Epi e;
      e.flags = tc::EPI_OUT32 | (all_fused ? 0u : tc::EPI_OUT16);
      e.out32 = X32; e.out16 = AX16; e.slope_out = 0.1f;
      if ((err = run.conv(up, Pin, L, e, 16 * (1 + i)))) return err;
    }
    Pcur = Pout;
    const tc_upgroup* upg_i = fuse_up ? &a.upg[i] : nullptr;
    L *= up.stride;
    const bool last_stage = (i == nst - 1);
    const bool tap_here = a.debug_out && a.debug_tap == SA_TAP_STAGE0 + i;
    // The fused kernel reads X32 (with halo) while other CTAs store results, so its fp32 stage
    // output goes to R32; the per-layer path writes it over the then-dead X32.
    const bool last_rb_fused = stage_grouped || (a.chains && run.chain_usable(a.chains[i * nrb + nrb - 1], L));
    float* H32 = last_rb_fused ? R32 : X32;
    Hout = H32;
    // One grouped launch for the whole stage (its ResBlocks one after the other over every tile: x and the running sum stay
    // in L2, 2.5 GB instead of 7.3 GB of DRAM traffic per stage) -- bit 0: C = 16, bit 1: C = 32.  All chains then share the
    // widest halo (60 positions).  Measured: C = 16 (1024-position tiles) 4.00 vs 3.96 ms; C = 32 (512-position tiles: 23 %
    // halo for every chain) 4.94 vs 4.80 ms, so C = 32 keeps one launch per ResBlock.  SATOOLS_B200_GROUP_STAGE overrides.
    const int group_stage = getenv("SATOOLS_B200_GROUP_STAGE") ? atoi(getenv("SATOOLS_B200_GROUP_STAGE")) : 1;
    if (stage_grouped && (group_stage & (up.cout == 16 ? 1 : 2)) && nrb > 1 && nrb <= tc::kGrpMaxChains) {
      Epi fin;
      fin.sum32 = S32; fin.n_blocks = (float)nrb;
      uint32_t fl[tc::kGrpMaxChains];
      for (int j = 0; j < nrb; ++j) fl[j] = (j == 0) ? tc::EPI_SUM_SET : (j == nrb - 1 ? tc::EPI_SUM_FIN : tc::EPI_SUM_ADD);
      if ((last_stage && !post16) || tap_here) { fl[nrb - 1] |= tc::EPI_OUT32; fin.out32 = H32; }
      if (!last_stage || post16) { fl[nrb - 1] |= tc::EPI_OUT16; fin.out16 = Pout; fin.slope_out = last_stage ? 0.01f : 0.1f; }
      bool done = false;
      if ((err = run.group(a.chains + i * nrb, nrb, X32, L, fin, fl, 16 * (1 + i) + 1, &done, 0, -1, upg_i, Pin))) return err;
      if (done) {
        if ((err = unblock_tap(SA_TAP_STAGE0 + i, up.cout, L))) return err;
        continue;
      }
    }
    if (a.chains && !stage_grouped) {               // narrowest stages on the per-tap kernels: the whole stage in one kernel
      Epi fin;
      if ((last_stage && !post16) || tap_here) { fin.flags |= tc::EPI_OUT32; fin.out32 = R32; }
      if (!last_stage || post16) { fin.flags |= tc::EPI_OUT16; fin.out16 = Pout; fin.slope_out = last_stage ? 0.01f : 0.1f; }
      bool done = false;
      if ((err = run.chain3(a.chains + i * nrb, nrb, X32, L, fin, 16 * (1 + i) + 1, &done))) return err;
      if (done) {
        Hout = R32;
        if ((err = unblock_tap(SA_TAP_STAGE0 + i, up.cout, L))) return err;
        continue;
      }
    }
    for (int j = 0; j < nrb; ++j) {
      const int tag = 16 * (1 + i) + 1 + j;
      // multi-receptive-field epilogue of this block's last conv (archi.py:82-86)
      Epi fin;
      fin.sum32 = S32; fin.n_blocks = (float)nrb;
      if (nrb > 1) fin.flags |= (j == 0) ? tc::EPI_SUM_SET : (j == nrb - 1 ? tc::EPI_SUM_FIN : tc::EPI_SUM_ADD);
      if (j == nrb - 1) {
        if ((last_stage && !post16) || tap_here) { fin.flags |= tc::EPI_OUT32; fin.out32 = H32; }
        if (!last_stage || post16) { fin.flags |= tc::EPI_OUT16; fin.out16 = Pout; fin.slope_out = last_stage ? 0.01f : 0.1f; }
      }
      if (a.chains) {                               // narrow stages: the whole ResBlock in one kernel
        bool done = false;
        const tc_chain& ch = a.chains[i * nrb + j];
        if (stage_grouped) {                        // C <= 32: grouped (block-Toeplitz) kernel, one launch per ResBlock
          const uint32_t fl1[1] = {fin.flags};
          if ((err = run.group(&ch, 1, X32, L, fin, fl1, tag, &done, 0, -1, upg_i, Pin))) return err;
          if (done) continue;
          return "grouped ResBlock launch failed";
        }
        // SATOOLS_B200_SPLIT: bit 0 = split the k = 11 blocks, bit 1 = also the k = 7 blocks
        // bit 2 = the k = 11 blocks as three launches (one pair each).  With the pipelined tile boundary of the RT kernel a short
        // launch no longer pays for its boundaries, so less recomputed halo wins: k = 7 split 1.71 -> 1.62 ms (same-box A/B).
        const int split = getenv("SATOOLS_B200_SPLIT") ? atoi(getenv("SATOOLS_B200_SPLIT")) : 3;
        if (all_fused && ch.n_convs == 6 && (split & 4) && ch.k == 11 && !tap_here && !last_stage && run.chain_usable(ch, L)) {
          float* TMP32 = reinterpret_cast<float*>(AX16);        // AX16 + A16 are dead in a fully fused stage; so is R32 here
          Epi mid;
          mid.flags = tc::EPI_OUT32; mid.out32 = TMP32;
          if ((err = run.chain(ch, X32, L, mid, tag, &done, 0, 2))) return err;
          mid.out32 = R32;
          if (done && (err = run.chain(ch, TMP32, L, mid, tag, &done, 2, 4))) return err;
          if (done && (err = run.chain(ch, R32, L, fin, tag, &done, 4, 6))) return err;
          if (done) continue;
          return "split ResBlock launch failed";
        }
        if (all_fused && ch.n_convs == 6 && (((split & 1) && ch.k == 11) || ((split & 2) && ch.k == 7)) && run.chain_usable(ch, L)) {
          // two launches (pairs 0-1 | pair 2): 30 instead of 60 halo rows per side, one extra fp32 round trip
          float* TMP32 = reinterpret_cast<float*>(AX16);        // AX16 + A16 are dead in a fully fused stage
          Epi mid;
          mid.flags = tc::EPI_OUT32; mid.out32 = TMP32;
          if ((err = run.chain(ch, X32, L, mid, tag, &done, 0, 4))) return err;
          if (done && (err = run.chain(ch, TMP32, L, fin, tag, &done, 4, 6))) return err;
          if (done) continue;
          return "split ResBlock launch failed";
        }
        if ((err = run.chain(ch, X32, L, fin, tag, &done))) return err;
        if (done) continue;
      }
      // The fused pair kernel reads its 16-bit input with a halo while other CTAs store this step's 16-bit output, so the
      // steps of a block alternate between A16 and T16 (the per-layer path needs T16 for conv1's output instead).
      void* act_in = AX16;
      for (int m = 0; m < nd; ++m) {               // nn.py:169-174
        Epi e2 = (m < nd - 1) ? Epi() : fin;
        e2.flags |= tc::EPI_RES;
        e2.res32 = (m == 0) ? X32 : R32;
        void* const act_out = (act_in == A16) ? T16 : A16;
        if (m < nd - 1) {
          e2.flags |= tc::EPI_OUT32 | tc::EPI_OUT16;
          e2.out32 = R32; e2.out16 = act_out; e2.slope_out = 0.1f;
        }
        bool done = false;
        if ((err = run.pair_fused(a.layers[L_rb(i, j, 0, m)], a.layers[L_rb(i, j, 1, m)], act_in, L, e2, tag, &done))) return err;
        if (done) { act_in = act_out; continue; }
        void* const tmp16 = (act_in == T16) ? A16 : T16;   // conv1's output; never the buffer this step reads
        if (m < nd - 1) e2.out16 = (tmp16 == T16) ? A16 : T16;
        Epi e1;
        e1.flags = tc::EPI_OUT16; e1.out16 = tmp16; e1.slope_out = 0.1f;
        if ((err = run.conv(a.layers[L_rb(i, j, 0, m)], act_in, L, e1, tag))) return err;
        if ((err = run.conv(a.layers[L_rb(i, j, 1, m)], tmp16, L, e2, tag))) return err;
        if (m < nd - 1) act_in = e2.out16;
      }
    }
    if ((err = unblock_tap(SA_TAP_STAGE0 + i, up.cout, L))) return err;
  }
  {                                                 // archi.py:87-90
    const tc_layer& post = a.layers[L_post];
    const int threads = 256;
    dim3 g((unsigned)((L + 1 + threads - 1) / threads), (unsigned)a.B);
    run.mark(16 * (nst + 1));
    if (post16) {
      dim3 g16((unsigned)((L + 1 + 249) / 250), (unsigned)a.B);
      const size_t sm = (size_t)post.cin * 32 + 7 * 256 * sizeof(float);
      const int pw = panel_width(post.cin);
      if (a.bf16)
        conv_post16_k7_kernel<true><<<g16, 256, sm, st>>>(reinterpret_cast<const uint4*>(Pcur), post.d_w32, post.d_bias, a.y,
                                                          post.cin, pw, L, a.y_dtype, d_frames, kRaggedMarginFrames, L / a.T);
      else
        conv_post16_k7_kernel<false><<<g16, 256, sm, st>>>(reinterpret_cast<const uint4*>(Pcur), post.d_w32, post.d_bias, a.y,
                                                           post.cin, pw, L, a.y_dtype, d_frames, kRaggedMarginFrames, L / a.T);
    } else
    conv_post_blocked_kernel<<<g, threads, post.cin * post.k * sizeof(float), st>>>(Hout, post.d_w32, post.d_bias, a.y,
                                                                                    post.cin, L, post.k, 0.01f, a.y_dtype);
    ++*launches;
    TC_CUDA(cudaGetLastError());
  }
  return nullptr;
}

}  // namespace sa
