// Tensor-core (tcgen05) path of the generator: interface used by sa_hifigan.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/sa_hifigan.h"

namespace sa {

// Packed 16-bit weights of one conv for the implicit-GEMM kernels (see conv_tc.cuh).
struct tc_weights {
  void* d_w = nullptr;       // device, 16-bit, [n_k16][2][N][8] K-major core-matrix order
  int n_k16 = 0;             // number of K=16 MMA steps
  int n_rows = 0;            // N rows per step (Cout, or phases*Cout for the upsamplers)
  size_t bytes = 0;
};

struct tc_context {
  bool ready = false;
  void* encode_fn = nullptr;  // cuTensorMapEncodeTiled
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
  const float* x;
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
  void* mark_ctx = nullptr;                                   // per-launch profiling hook
  void (*mark)(void* ctx, int tag, cudaStream_t s) = nullptr;
};

bool tc_layer_supported(bool transposed, int cin, int cout, int k);
const char* tc_pack_weights(tc_weights& w, const float* folded, bool transposed, int cin, int cout, int k,
                            int stride, int pad, bool bf16);
void tc_free_weights(tc_weights& w);
const char* tc_init(tc_context& ctx, int device);
size_t tc_workspace_bytes(const sa_hifigan_cfg& cfg, int B, int T);
const char* tc_forward(tc_context& ctx, const tc_forward_args& a, int64_t* launches);

}  // namespace sa
