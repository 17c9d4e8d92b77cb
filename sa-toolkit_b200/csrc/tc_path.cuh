// Tensor-core (tcgen05) path of the generator: interface used by sa_hifigan.cu.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "../../include/sa_hifigan.h"

namespace sa {

constexpr int kTcMaxPhases = 8;

// Packed 16-bit weights of one conv for the implicit-GEMM kernel (conv_tc.cuh):
// [phase][n_tile][tap][panel][N rows][pw], each [N][pw] block in the UMMA K-major swizzled layout.
struct tc_weights {
  void* d_w = nullptr;
  int n = 0;                 // N per MMA (Cout tile)
  int n_tiles = 0;           // Cout / N
  int pair = 0;              // 1: packed in N/2-row half tiles for the CTA-pair kernel (conv_pair_tc.cuh):
                             //    [phase][2 * n_tile + half][tap][panel][N/2][64]; tile_bytes = one half tile
  int n_phases = 0;          // 1 for Conv1d, stride for ConvTranspose1d
  int cin_pad = 0;           // Cin rounded up to whole panels (tc_cin_pad)
  int n_taps[kTcMaxPhases] = {0};
  int tap_base[kTcMaxPhases] = {0};
  int tap_step = 1;          // informational for transposed convs (-1); convs use the layer dilation
  size_t tile_bytes = 0;     // bytes of one (phase, n_tile) block
  size_t bytes = 0;
};

// One ResBlock1 packed for the fused narrow-stage kernel (chain_tc.cuh): the 2*n_dilations convs in
// execution order (convs1[0], convs2[0], convs1[1], ...), each [tap][C rows][C] swizzled; bias [n_convs][C].
struct tc_chain {
  void* d_w = nullptr;
  float* d_bias = nullptr;
  int c = 0, k = 0, n_convs = 0;
  int dil[8] = {0}, pad[8] = {0};
  int halo = 0;
  int k16_per_stage = 0, stages_per_conv = 0;
  // grouped (block-Toeplitz) packing for C <= 32 (chain_group_tc.cuh): per conv g_stages stages of 8 KB, each four
  // [64 rows (g', co)][16 ci] blocks in the SWIZZLE_32B K-major layout; g_slices = (64 / C + k - 1) * C / 16 are used
  void* d_wg = nullptr;
  int g_slices = 0, g_stages = 0;
};

// Encoded tensor maps are cached per (buffer, geometry): a forward re-uses ~45 of them, and re-encoding them on the host
// for every call showed up in the single-utterance latency path.
struct tc_tmap_entry {
  const void* ptr; uint64_t d0, d1, d2, d3; uint32_t b0, b1; int dtype, swizzle, rank;
  CUtensorMap map;
};

// The stage's transposed conv packed for the grouped kernel (chain_group_tc.cuh, fuse_up): block-Toeplitz slices over the
// input positions a 128-byte output row needs + the bias block, in 8 KB stages.  Only k = 4, stride 2, pad 1, Cin = 2 Cout.
struct tc_upgroup {
  void* d_w = nullptr;
  int slices = 0, stages = 0;
};

struct tc_context {
  // kernel-family switches of THIS handle, read from the environment in tc_init (SATOOLS_B200_GROUP / _CHAIN3 / _CHAIN_MS)
  int use_group = 1, use_chain3 = 1, chain_ms_narrow = 6;
  std::vector<tc_tmap_entry> tmaps;
  bool ready = false;
  void* encode_fn = nullptr;  // cuTensorMapEncodeTiled
  int* h_error = nullptr;     // mapped pinned flag raised by a kernel whose barrier wait timed out
  int* d_error = nullptr;     // device alias of h_error
  int max_smem = 0;
  long long* d_timing = nullptr;   // diagnostics: [64 launches][8] cycle counters of the fused kernels (SATOOLS_B200_CHAIN_TIMING=1)
  int timing_launches = 0;
};

struct tc_layer {
  const tc_weights* w;
  const float* d_w32;        // fp32 packing [Cin][k][Cout] when the layer runs on CUDA cores
  const float* d_bias;
  int cin, cout, k, dil, pad, stride;
  bool transposed;
};

struct tc_forward_args {
  const sa_hifigan_cfg* cfg;
  const float* x;                                             // [B, input_dim, T] fp32, or nullptr: the parts below
  const float* bn = nullptr;                                  // [B, n_bn, T]   channels 0 .. n_bn-1      (hifigan.py:91-93)
  const float* f0 = nullptr;                                  // [B, 1, T]      channel n_bn
  const float* spk = nullptr;                                 // [B, n_spk]     channels n_bn+1 .., constant in time (hifigan.py:94-97)
  int n_bn = 0, n_spk = 0;
  // or (x == nullptr && bn == nullptr) the compact conditioning: VQ code index per frame, F0, speaker id per item
  const uint8_t* vq_idx = nullptr;                            // [B, T]; index >= n_codes: zero vector (padding frames)
  const float* codebook = nullptr;                            // [n_codes, n_bn] device
  const int32_t* spk_ids = nullptr;                           // [B] device; < 0 or >= n_spk: no speaker channel set
  int n_codes = 0;
  int B, T;
  const int32_t* frames_per_item;
  void* y;
  int y_dtype;
  void* workspace;
  cudaStream_t stream;
  int debug_tap;
  float* debug_out;
  bool bf16;
  int n_sm;
  const tc_layer* layers;
  int n_layers;
  const tc_chain* chains = nullptr;                           // [n_stages * n_resblocks]; d_w == NULL: not packed
  const tc_upgroup* upg = nullptr;                            // [n_stages]; d_w == NULL: the upsampler is not fusable
  void* mark_ctx = nullptr;                                   // per-launch profiling hook
  void (*mark)(void* ctx, int tag, cudaStream_t s) = nullptr;
};

int tc_cin_pad(int cin);
bool tc_layer_supported(bool transposed, int cin, int cout, int k, bool first_layer);
const char* tc_pack_weights(tc_weights& w, const float* folded, bool transposed, int cin, int cout, int k,
                            int stride, int pad, bool bf16);
void tc_free_weights(tc_weights& w);
bool tc_chain_supported(int c, int k, int n_convs);
const char* tc_pack_chain(tc_chain& ch, int c, int k, int n_convs, const float* const* folded, const float* const* bias,
                          const int* dil, const int* pad, bool bf16);
void tc_free_chain(tc_chain& ch);
const char* tc_pack_upgroup(tc_upgroup& u, const float* folded, const float* bias, int cin, int cout, int k, int stride, int pad,
                            bool bf16);
void tc_free_upgroup(tc_upgroup& u);
const char* tc_init(tc_context& ctx, int device);
bool tc_error_raised(const tc_context& ctx);
int tc_read_chain_timing(tc_context& ctx, long long* out, int max_launches);
size_t tc_workspace_bytes(const sa_hifigan_cfg& cfg, int B, int T);
const char* tc_forward(tc_context& ctx, const tc_forward_args& a, int64_t* launches);

}  // namespace sa
