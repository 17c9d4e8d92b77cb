// libsatools_hifigan.so -- C ABI + host orchestration of the B200 HiFi-GAN generator.
// See include/sa_hifigan.h for the contract and the reference entry points each call replaces.
//
// Graph executed (reference: satools/satools/hifigan/archi.py:77-91, nn.py:168-175):
//   h = conv_pre(x)
//   for stage i: h = convT_i(lrelu(h)); h = mean_j ResBlock1_{k_j}(h)
//   y = tanh(conv_post(reflect_pad(lrelu(h, 0.01))))
#include "../../include/sa_hifigan.h"

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "kernels_f32.cuh"
#include "tc_path.cuh"

namespace {

thread_local char g_err[1024] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define SA_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(SA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_),    \
                  __FILE__, __LINE__);                                                    \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// One conv layer: raw state-dict tensors on the host, packed weights on the device.
struct sa_conv {
  std::string name;
  bool transposed = false;   // ConvTranspose1d: weight [Cin,Cout,k]; Conv1d: [Cout,Cin,k]
  int cin = 0, cout = 0, k = 0, dil = 1, pad = 0, stride = 1;
  std::vector<float> g, v, w, bias;  // g/v raw; w folded, in the torch layout
  bool has_g = false, has_v = false, has_w = false, has_bias = false;
  float* d_w32 = nullptr;    // fp32 path: [Cin][k][Cout]
  float* d_bias = nullptr;   // [Cout]
  sa::tc_weights tc;         // tensor-core path packing
};

struct sa_hifigan {
  sa_hifigan_cfg cfg;
  int device = 0;
  std::vector<sa_conv> convs;
  std::map<std::string, int> index;
  int precision = -1;
  bool finalized = false;
  int64_t launches = 0;
  bool use_fused = true;                // SATOOLS_B200_FUSED=0 forces the per-layer tensor-core path
  int debug_tap = -1;
  float* debug_out = nullptr;
  int n_sm = 148;
  sa::tc_context tc;
  std::vector<sa::tc_chain> chains;     // [n_stages * n_resblocks], fused narrow-stage ResBlocks
  std::vector<sa::tc_upgroup> upg;      // [n_stages], transposed convs packed for fusion into the grouped kernel
  float* d_codebook = nullptr;          // [n_codes][code_dim] VQ codebook of the compact conditioning (sa_hifigan_set_codebook)
  int n_codes = 0, code_dim = 0;
  cudaStream_t side_stream = nullptr;   // second stream of sa_hifigan_synthesize_host
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // sa_hifigan_synthesize_host_async: the forwards of all in-flight calls run on ONE compute stream (in submission
  // order), only the copies stay on the callers' streams -- two forwards on two streams would interleave their kernels
  cudaStream_t compute_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_done = nullptr;
  // per-launch profiling: one event before every launch + one closing event
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_tag;
  int prof_n = 0;
  cudaStream_t prof_stream = nullptr;
  void mark(int tag, cudaStream_t st) {
    if (!prof_on) return;
    if (prof_n == (int)prof_ev.size()) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return;
      prof_ev.push_back(e);
      prof_tag.push_back(0);
    }
    prof_tag[prof_n] = tag;
    cudaEventRecord(prof_ev[prof_n++], st);
    prof_stream = st;
  }

  int conv_pre() const { return 0; }
  int up(int i) const { return 1 + i; }
  int rb(int i, int j, int which, int m) const {   // which: 0 convs1, 1 convs2
    const int nrb = cfg.n_resblocks, nd = cfg.n_dilations;
    return 1 + cfg.n_stages + ((i * nrb + j) * 2 + which) * nd + m;
  }
  int conv_post() const { return 1 + cfg.n_stages + cfg.n_stages * cfg.n_resblocks * 2 * cfg.n_dilations; }
  int stage_channels(int i) const { return cfg.initial_channels >> (i + 1); }
  int64_t stage_rate(int i) const {
    int64_t r = 1;
    for (int s = 0; s <= i; ++s) r *= cfg.upsample_rates[s];
    return r;
  }
  // largest C*L of any activation, per frame
  int64_t max_elems_per_frame() const {
    int64_t m = cfg.initial_channels;
    for (int i = 0; i < cfg.n_stages; ++i) m = std::max<int64_t>(m, stage_channels(i) * stage_rate(i));
    return m;
  }
};

extern "C" {

int sa_hifigan_abi_version(void) { return SA_HIFIGAN_ABI_VERSION; }
const char* sa_hifigan_last_error(void) { return g_err; }

int sa_hifigan_default_cfg(sa_hifigan_cfg* cfg) {
  if (!cfg) return fail(SA_ERR_INVALID_ARG, "cfg is NULL");
  memset(cfg, 0, sizeof(*cfg));
  cfg->input_dim = 256 + 1 + 247;   // hifigan.py:45-46 with the 247 LibriTTS speakers
  cfg->initial_channels = 512;
  cfg->n_stages = 5;
  const int rates[5] = {5, 4, 4, 2, 2}, kernels[5] = {11, 8, 8, 4, 4};
  for (int i = 0; i < 5; ++i) { cfg->upsample_rates[i] = rates[i]; cfg->upsample_kernels[i] = kernels[i]; }
  cfg->n_resblocks = 3;
  const int rk[3] = {3, 7, 11}, rd[3] = {1, 3, 5};
  cfg->n_dilations = 3;
  for (int j = 0; j < 3; ++j) {
    cfg->resblock_kernels[j] = rk[j];
    for (int m = 0; m < 3; ++m) cfg->resblock_dilations[j][m] = rd[m];
  }
  cfg->device = -1;
  return SA_OK;
}

int sa_hifigan_create(const sa_hifigan_cfg* cfg, sa_hifigan** out) {
  if (!cfg || !out) return fail(SA_ERR_INVALID_ARG, "cfg/out is NULL");
  *out = nullptr;
  if (cfg->n_stages < 1 || cfg->n_stages > SA_HIFIGAN_MAX_STAGES || cfg->n_resblocks < 1 ||
      cfg->n_resblocks > SA_HIFIGAN_MAX_RB || cfg->n_dilations < 1 || cfg->n_dilations > SA_HIFIGAN_MAX_RB ||
      cfg->input_dim < 1 || cfg->initial_channels < (1 << cfg->n_stages))
    return fail(SA_ERR_INVALID_ARG, "bad generator configuration");
  if (cfg->initial_channels % (1 << cfg->n_stages) != 0)
    return fail(SA_ERR_INVALID_ARG, "initial_channels must be divisible by 2^n_stages");
  for (int i = 0; i < cfg->n_stages; ++i) {
    const int u = cfg->upsample_rates[i], k = cfg->upsample_kernels[i];
    if (u < 1 || k < u || ((k - u) & 1))
      return fail(SA_ERR_UNSUPPORTED, "upsample stage %d: need kernel >= rate and (kernel - rate) even", i);
  }
  for (int j = 0; j < cfg->n_resblocks; ++j)
    if (cfg->resblock_kernels[j] < 1 || !(cfg->resblock_kernels[j] & 1))
      return fail(SA_ERR_UNSUPPORTED, "resblock kernel sizes must be odd");

  int dev = cfg->device;
  if (dev < 0) SA_CUDA(cudaGetDevice(&dev));
  int n_dev = 0;
  SA_CUDA(cudaGetDeviceCount(&n_dev));
  if (dev >= n_dev) return fail(SA_ERR_INVALID_ARG, "device %d out of range (%d devices)", dev, n_dev);
  cudaDeviceProp prop;
  SA_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(SA_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                dev, prop.major, prop.minor);

  sa_hifigan* h = new sa_hifigan();
  h->cfg = *cfg;
  h->device = dev;
  if (const char* env = getenv("SATOOLS_B200_FUSED")) h->use_fused = atoi(env) != 0;
  h->n_sm = prop.multiProcessorCount;

  auto add = [&](const std::string& name, bool tr, int cin, int cout, int k, int dil, int pad, int stride) {
    sa_conv c;
    c.name = name; c.transposed = tr; c.cin = cin; c.cout = cout; c.k = k; c.dil = dil; c.pad = pad; c.stride = stride;
    h->index[name] = (int)h->convs.size();
    h->convs.push_back(std::move(c));
  };
  add("conv_pre", false, cfg->input_dim, cfg->initial_channels, 7, 1, 3, 1);                  // archi.py:40-42
  for (int i = 0; i < cfg->n_stages; ++i) {                                                    // archi.py:47-59
    const int u = cfg->upsample_rates[i], k = cfg->upsample_kernels[i];
    add("ups." + std::to_string(i), true, cfg->initial_channels >> i, cfg->initial_channels >> (i + 1), k, 1, (k - u) / 2, u);
  }
  for (int i = 0; i < cfg->n_stages; ++i)                                                      // archi.py:61-67
    for (int j = 0; j < cfg->n_resblocks; ++j) {
      const int ch = h->stage_channels(i), k = cfg->resblock_kernels[j];
      const std::string base = "resblocks." + std::to_string(i * cfg->n_resblocks + j);
      for (int m = 0; m < cfg->n_dilations; ++m) {                                             // nn.py:96-131
        const int d = cfg->resblock_dilations[j][m];
        add(base + ".convs1." + std::to_string(m), false, ch, ch, k, d, (k * d - d) / 2, 1);
      }
      for (int m = 0; m < cfg->n_dilations; ++m)                                               // nn.py:133-166
        add(base + ".convs2." + std::to_string(m), false, ch, ch, k, 1, (k - 1) / 2, 1);
    }
  add("conv_post", false, h->stage_channels(cfg->n_stages - 1), 1, 7, 1, 3, 1);                // archi.py:72
  *out = h;
  return SA_OK;
}

static void free_device_weights(sa_hifigan* h) {
  for (auto& c : h->convs) {
    if (c.d_w32) cudaFree(c.d_w32);
    if (c.d_bias) cudaFree(c.d_bias);
    c.d_w32 = nullptr; c.d_bias = nullptr;
    sa::tc_free_weights(c.tc);
  }
  for (auto& ch : h->chains) sa::tc_free_chain(ch);
  h->chains.clear();
  for (auto& u : h->upg) sa::tc_free_upgroup(u);
  h->upg.clear();
  h->tc.tmaps.clear();                  // cached tensor maps point at the freed weight buffers
}

void sa_hifigan_destroy(sa_hifigan* h) {
  if (!h) return;
  int cur = -1;
  if (cudaGetDevice(&cur) == cudaSuccess) {
    cudaSetDevice(h->device);
    free_device_weights(h);
    if (h->d_codebook) cudaFree(h->d_codebook);
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->compute_stream) cudaStreamDestroy(h->compute_stream);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    cudaSetDevice(cur);
  }
  delete h;
}

int sa_hifigan_set_weight(sa_hifigan* h, const char* key, const void* data, const int64_t* shape,
                          int32_t ndim, int32_t dtype) {
  if (!h || !key || !data || !shape) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  const std::string k(key);
  const size_t dot = k.rfind('.');
  if (dot == std::string::npos) return fail(SA_ERR_BAD_KEY, "unknown key '%s'", key);
  const std::string layer = k.substr(0, dot), leaf = k.substr(dot + 1);
  auto it = h->index.find(layer);
  if (it == h->index.end()) return fail(SA_ERR_BAD_KEY, "unknown layer in key '%s'", key);
  sa_conv& c = h->convs[it->second];
  const int64_t d0 = c.transposed ? c.cin : c.cout, d1 = c.transposed ? c.cout : c.cin;

  std::vector<float>* dst = nullptr;
  bool* flag = nullptr;
  int64_t expect[3] = {0, 0, 0};
  int expect_nd = 0;
  if (leaf == "weight_v" || leaf == "weight") {
    expect[0] = d0; expect[1] = d1; expect[2] = c.k; expect_nd = 3;
    dst = (leaf == "weight") ? &c.w : &c.v;
    flag = (leaf == "weight") ? &c.has_w : &c.has_v;
  } else if (leaf == "weight_g") {
    expect[0] = d0; expect[1] = 1; expect[2] = 1; expect_nd = 3;   // weight_norm dim=0
    dst = &c.g; flag = &c.has_g;
  } else if (leaf == "bias") {
    expect[0] = c.cout; expect_nd = 1;
    dst = &c.bias; flag = &c.has_bias;
  } else {
    return fail(SA_ERR_BAD_KEY, "unknown parameter '%s' in key '%s'", leaf.c_str(), key);
  }
  if (ndim != expect_nd) return fail(SA_ERR_BAD_SHAPE, "%s: expected %d dims, got %d", key, expect_nd, ndim);
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] != expect[i])
      return fail(SA_ERR_BAD_SHAPE, "%s: dim %d is %lld, expected %lld", key, i, (long long)shape[i], (long long)expect[i]);
    n *= (size_t)shape[i];
  }
  size_t esz = 0;
  switch (dtype) {
    case SA_DTYPE_F32: esz = 4; break;
    case SA_DTYPE_F16: case SA_DTYPE_BF16: esz = 2; break;
    case SA_DTYPE_F64: esz = 8; break;
    default: return fail(SA_ERR_INVALID_ARG, "%s: unsupported dtype %d", key, dtype);
  }
  std::vector<unsigned char> raw(n * esz);
  SA_CUDA(cudaMemcpy(raw.data(), data, n * esz, cudaMemcpyDefault));   // host or device source
  dst->resize(n);
  for (size_t i = 0; i < n; ++i) {
    float f;
    switch (dtype) {
      case SA_DTYPE_F32: f = reinterpret_cast<const float*>(raw.data())[i]; break;
      case SA_DTYPE_F64: f = (float)reinterpret_cast<const double*>(raw.data())[i]; break;
      case SA_DTYPE_F16: f = __half2float(reinterpret_cast<const __half*>(raw.data())[i]); break;
      default: f = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(raw.data())[i]); break;
    }
    (*dst)[i] = f;
  }
  *flag = true;
  if (leaf == "weight") { c.has_g = c.has_v = false; }
  else if (leaf == "weight_g" || leaf == "weight_v") c.has_w = false;
  h->finalized = false;
  return SA_OK;
}

int sa_hifigan_finalize(sa_hifigan* h, int32_t precision) {
  if (!h) return fail(SA_ERR_INVALID_ARG, "NULL handle");
  if (precision != SA_PRECISION_FP32 && precision != SA_PRECISION_FP16 && precision != SA_PRECISION_BF16)
    return fail(SA_ERR_INVALID_ARG, "unknown precision %d", precision);
  SA_CUDA(cudaSetDevice(h->device));
  // 1. fold weight-norm on the host: w = g * v / ||v||_2, norm over all dims but 0.
  for (auto& c : h->convs) {
    if (!c.has_bias) return fail(SA_ERR_MISSING_WEIGHT, "%s.bias was never set", c.name.c_str());
    if (c.has_g && c.has_v) {
      const int64_t d0 = c.transposed ? c.cin : c.cout;
      const size_t inner = c.v.size() / (size_t)d0;
      c.w.resize(c.v.size());
      for (int64_t r = 0; r < d0; ++r) {
        double ss = 0.0;
        const float* vr = c.v.data() + r * inner;
        for (size_t i = 0; i < inner; ++i) ss += (double)vr[i] * vr[i];
        const float scale = (float)((double)c.g[r] / std::sqrt(ss));
        for (size_t i = 0; i < inner; ++i) c.w[r * inner + i] = vr[i] * scale;
      }
    } else if (!c.has_w) {
      return fail(SA_ERR_MISSING_WEIGHT, "%s: need weight_g + weight_v, or weight", c.name.c_str());
    }
  }
  free_device_weights(h);
  // 2. pack + upload.
  for (auto& c : h->convs) {
    SA_CUDA(cudaMalloc(&c.d_bias, c.cout * sizeof(float)));
    SA_CUDA(cudaMemcpy(c.d_bias, c.bias.data(), c.cout * sizeof(float), cudaMemcpyHostToDevice));
  }
  for (size_t li = 0; li < h->convs.size(); ++li) {
    sa_conv& c = h->convs[li];
    const bool tc_ok = sa::tc_layer_supported(c.transposed, c.cin, c.cout, c.k, (int)li == h->conv_pre());
    // The tensor-core forward runs every layer but conv_post through the tcgen05 kernels: a layer they cannot take
    // (channel counts that are not whole panels / instantiated tiles) is refused here, not discovered at launch time.
    if (precision != SA_PRECISION_FP32 && (int)li != h->conv_post() && !tc_ok)
      return fail(SA_ERR_UNSUPPORTED, "%s: Cin=%d Cout=%d is not supported by the fp16/bf16 tensor-core path (channels after "
                  "conv_pre must be 16, 32, 64, 128 or a multiple of 256); use precision fp32", c.name.c_str(), c.cin, c.cout);
    const bool need_f32 = precision == SA_PRECISION_FP32 || (int)li == h->conv_post();
    if (need_f32) {
      // [Cin][k][Cout] from [Cout][Cin][k] (Conv1d) or [Cin][Cout][k] (ConvTranspose1d)
      std::vector<float> p((size_t)c.cin * c.k * c.cout);
      for (int ci = 0; ci < c.cin; ++ci)
        for (int j = 0; j < c.k; ++j)
          for (int co = 0; co < c.cout; ++co) {
            const size_t src = c.transposed ? ((size_t)ci * c.cout + co) * c.k + j
                                            : ((size_t)co * c.cin + ci) * c.k + j;
            p[((size_t)ci * c.k + j) * c.cout + co] = c.w[src];
          }
      SA_CUDA(cudaMalloc(&c.d_w32, p.size() * sizeof(float)));
      SA_CUDA(cudaMemcpy(c.d_w32, p.data(), p.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (precision != SA_PRECISION_FP32 && !need_f32) {
      const char* err = sa::tc_pack_weights(c.tc, c.w.data(), c.transposed, c.cin, c.cout, c.k, c.stride, c.pad,
                                            precision == SA_PRECISION_BF16);
      if (err) return fail(SA_ERR_CUDA, "%s: %s", c.name.c_str(), err);
    }
  }
  if (precision != SA_PRECISION_FP32) {
    // fused ResBlock packing for the narrow stages (C <= 64)
    const sa_hifigan_cfg& cfg = h->cfg;
    h->chains.assign((size_t)cfg.n_stages * cfg.n_resblocks, sa::tc_chain());
    for (int i = 0; i < cfg.n_stages; ++i)
      for (int j = 0; j < cfg.n_resblocks; ++j) {
        const int C = h->stage_channels(i), k = cfg.resblock_kernels[j], nc = 2 * cfg.n_dilations;
        if (!sa::tc_chain_supported(C, k, nc)) continue;
        const float* wp[8]; const float* bp[8]; int dil[8], pad[8];
        for (int m = 0; m < cfg.n_dilations; ++m)
          for (int which = 0; which < 2; ++which) {
            const sa_conv& c = h->convs[h->rb(i, j, which, m)];
            wp[2 * m + which] = c.w.data(); bp[2 * m + which] = c.bias.data();
            dil[2 * m + which] = c.dil; pad[2 * m + which] = c.pad;
          }
        const char* err = sa::tc_pack_chain(h->chains[(size_t)i * cfg.n_resblocks + j], C, k, nc, wp, bp, dil, pad,
                                            precision == SA_PRECISION_BF16);
        if (err) return fail(SA_ERR_CUDA, "resblocks.%d: %s", i * cfg.n_resblocks + j, err);
      }
    h->upg.assign((size_t)cfg.n_stages, sa::tc_upgroup());
    for (int i = 0; i < cfg.n_stages; ++i) {
      const sa_conv& c = h->convs[h->up(i)];
      const char* uerr = sa::tc_pack_upgroup(h->upg[i], c.w.data(), c.bias.data(), c.cin, c.cout, c.k, c.stride, c.pad,
                                             precision == SA_PRECISION_BF16);
      if (uerr) return fail(SA_ERR_CUDA, "ups.%d: %s", i, uerr);
    }
    const char* err = sa::tc_init(h->tc, h->device);
    if (err) return fail(SA_ERR_CUDA, "tensor-core path init: %s", err);
  }
  h->precision = precision;
  h->finalized = true;
  return SA_OK;
}

int64_t sa_hifigan_output_length(const sa_hifigan* h, int64_t T) {
  if (!h) return -1;
  return h->stage_rate(h->cfg.n_stages - 1) * T + 1;
}

// Workspace: five fp32 activation buffers of the largest activation (fp32 path), or the
// tensor-core path's own arena.
static size_t buf_bytes(const sa_hifigan* h, int B, int T) {
  return align_up((size_t)B * (size_t)T * (size_t)h->max_elems_per_frame() * sizeof(float), 256);
}

size_t sa_hifigan_workspace_bytes(const sa_hifigan* h, int32_t B, int32_t T) {
  if (!h || B < 1 || T < 1) return 0;
  if (h->finalized && h->precision != SA_PRECISION_FP32) return sa::tc_workspace_bytes(h->cfg, B, T);
  return 5 * buf_bytes(h, B, T);
}

int sa_hifigan_set_debug_tap(sa_hifigan* h, int32_t tap, float* out) {
  if (!h) return fail(SA_ERR_INVALID_ARG, "NULL handle");
  if (out && (tap < 0 || tap > h->cfg.n_stages)) return fail(SA_ERR_INVALID_ARG, "tap %d out of range", tap);
  h->debug_tap = out ? tap : -1;
  h->debug_out = out;
  return SA_OK;
}

int64_t sa_hifigan_last_launch_count(const sa_hifigan* h) { return h ? h->launches : -1; }

}  // extern "C"

// ---------------------------------------------------------------------------------------
// fp32 path
// ---------------------------------------------------------------------------------------
namespace {

constexpr int kTT = 128;
constexpr int kCIB = 8;

int launch_conv_f32(sa_hifigan* h, const sa_conv& c, const float* x, const float* res, float* y, int B, int64_t L,
                    float slope_in, cudaStream_t st, int tag) {
  h->mark(tag, st);
  const int halo = (c.k - 1) * c.dil;
  dim3 grid((unsigned)((L + kTT - 1) / kTT), 1, (unsigned)B);
  if (c.cout % 32 == 0) {
    grid.y = c.cout / 32;
    const size_t smem = (size_t)(kCIB * (kTT + halo) + kCIB * c.k * 32) * sizeof(float);
    sa::conv1d_f32_kernel<32, kTT, kCIB><<<grid, kTT, smem, st>>>(x, c.d_w32, c.d_bias, res, y, c.cin, c.cout, (int)L,
                                                                  c.k, c.dil, c.pad, slope_in);
  } else if (c.cout % 16 == 0) {
    grid.y = c.cout / 16;
    const size_t smem = (size_t)(kCIB * (kTT + halo) + kCIB * c.k * 16) * sizeof(float);
    sa::conv1d_f32_kernel<16, kTT, kCIB><<<grid, kTT, smem, st>>>(x, c.d_w32, c.d_bias, res, y, c.cin, c.cout, (int)L,
                                                                  c.k, c.dil, c.pad, slope_in);
  } else if (c.cout % 4 == 0) {
    grid.y = c.cout / 4;
    const size_t smem = (size_t)(kCIB * (kTT + halo) + kCIB * c.k * 4) * sizeof(float);
    sa::conv1d_f32_kernel<4, kTT, kCIB><<<grid, kTT, smem, st>>>(x, c.d_w32, c.d_bias, res, y, c.cin, c.cout, (int)L,
                                                                 c.k, c.dil, c.pad, slope_in);
  } else {
    return fail(SA_ERR_UNSUPPORTED, "%s: Cout=%d must be a multiple of 4", c.name.c_str(), c.cout);
  }
  h->launches++;
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

int launch_convt_f32(sa_hifigan* h, const sa_conv& c, const float* x, float* y, int B, int64_t Lin, float slope_in,
                     cudaStream_t st, int tag) {
  h->mark(tag, st);
  const int u = c.stride, taps = (c.k + u - 1) / u;
  const int64_t Lout = Lin * u;
  dim3 grid((unsigned)((Lout + kTT - 1) / kTT), 1, (unsigned)B);
  const int xw_max = kTT / u + taps + 2;
  if (c.cout % 16 == 0) {
    grid.y = c.cout / 16;
    const size_t smem = (size_t)(kCIB * xw_max + kCIB * c.k * 16) * sizeof(float);
    sa::convt1d_f32_kernel<16, kTT, kCIB><<<grid, kTT, smem, st>>>(x, c.d_w32, c.d_bias, y, c.cin, c.cout, (int)Lin,
                                                                   c.k, u, c.pad, slope_in);
  } else if (c.cout % 4 == 0) {
    grid.y = c.cout / 4;
    const size_t smem = (size_t)(kCIB * xw_max + kCIB * c.k * 4) * sizeof(float);
    sa::convt1d_f32_kernel<4, kTT, kCIB><<<grid, kTT, smem, st>>>(x, c.d_w32, c.d_bias, y, c.cin, c.cout, (int)Lin,
                                                                  c.k, u, c.pad, slope_in);
  } else {
    return fail(SA_ERR_UNSUPPORTED, "%s: Cout=%d must be a multiple of 4", c.name.c_str(), c.cout);
  }
  h->launches++;
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

int launch_mrf(sa_hifigan* h, float* s, const float* r, float* out, size_t n, int mode, int nrb, cudaStream_t st,
               int tag) {
  h->mark(tag, st);
  const int threads = 256;
  const unsigned blocks = (unsigned)std::min<size_t>((n + threads - 1) / threads, (size_t)h->n_sm * 16);
  sa::mrf_combine_f32_kernel<<<blocks, threads, 0, st>>>(s, r, out, n, mode, (float)nrb);
  h->launches++;
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

int copy_tap(sa_hifigan* h, int tap, const float* src, size_t n, cudaStream_t st) {
  if (h->debug_out && h->debug_tap == tap)
    SA_CUDA(cudaMemcpyAsync(h->debug_out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return SA_OK;
}

int forward_f32(sa_hifigan* h, const float* x, int B, int T, void* y, int y_dtype, void* ws, cudaStream_t st) {
  const sa_hifigan_cfg& cfg = h->cfg;
  const size_t bb = buf_bytes(h, B, T);
  float* buf[5];
  for (int i = 0; i < 5; ++i) buf[i] = reinterpret_cast<float*>(static_cast<char*>(ws) + i * bb);
  float *P = buf[0], *X = buf[1], *R = buf[2], *Tm = buf[3], *S = buf[4];
  int rc;

  if ((rc = launch_conv_f32(h, h->convs[h->conv_pre()], x, nullptr, P, B, T, 1.0f, st, 0))) return rc;       // archi.py:78
  if ((rc = copy_tap(h, SA_TAP_CONV_PRE, P, (size_t)B * cfg.initial_channels * T, st))) return rc;
  int64_t L = T;
  for (int i = 0; i < cfg.n_stages; ++i) {
    const sa_conv& up = h->convs[h->up(i)];
    if ((rc = launch_convt_f32(h, up, P, X, B, L, 0.1f, st, 16 * (1 + i)))) return rc;                                   // archi.py:80-81
    L *= up.stride;
    const size_t n = (size_t)B * up.cout * L;
    for (int j = 0; j < cfg.n_resblocks; ++j) {
      const float* src = X;
      for (int m = 0; m < cfg.n_dilations; ++m) {                                                          // nn.py:169-174
        if ((rc = launch_conv_f32(h, h->convs[h->rb(i, j, 0, m)], src, nullptr, Tm, B, L, 0.1f, st, 16 * (1 + i) + 1 + j))) return rc;
        if ((rc = launch_conv_f32(h, h->convs[h->rb(i, j, 1, m)], Tm, src, R, B, L, 0.1f, st, 16 * (1 + i) + 1 + j))) return rc;
        src = R;
      }
      const int mode = (j == cfg.n_resblocks - 1) ? 2 : (j == 0 ? 0 : 1);                                  // archi.py:82-86
      if (cfg.n_resblocks == 1) {
        SA_CUDA(cudaMemcpyAsync(P, R, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
      } else if ((rc = launch_mrf(h, S, R, P, n, mode, cfg.n_resblocks, st, 16 * (1 + i) + 15))) return rc;
    }
    if ((rc = copy_tap(h, SA_TAP_STAGE0 + i, P, n, st))) return rc;
  }
  const sa_conv& post = h->convs[h->conv_post()];
  {
    const int threads = 256;
    dim3 grid((unsigned)((L + 1 + threads - 1) / threads), (unsigned)B);
    h->mark(16 * (cfg.n_stages + 1), st);
    sa::conv_post_f32_kernel<<<grid, threads, 0, st>>>(P, post.d_w32, post.d_bias, y, post.cin, (int)L, post.k, 0.01f,
                                                       y_dtype);                                          // archi.py:87-90
    h->launches++;
    SA_CUDA(cudaGetLastError());
  }
  return SA_OK;
}

}  // namespace

extern "C" {

// x, or (x == nullptr) the conditioning parts bn / f0 / spk
static int forward_impl(sa_hifigan* h, const float* x, const float* bn, const float* f0, const float* spk, int32_t n_bn,
                        int32_t n_spk, int32_t B, int32_t T, const int32_t* frames_per_item, void* y, int32_t y_dtype,
                        void* workspace, size_t workspace_bytes, void* stream, const uint8_t* vq_idx = nullptr,
                        const int32_t* spk_ids = nullptr) {
  if (!h || !y || !workspace) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  if (vq_idx) {                                     // compact conditioning: code index + F0 per frame, speaker id per item
    if (!f0 || !spk_ids) return fail(SA_ERR_INVALID_ARG, "NULL argument");
    if (!h->d_codebook) return fail(SA_ERR_NOT_FINALIZED, "call sa_hifigan_set_codebook first");
    if (h->finalized && h->precision == SA_PRECISION_FP32)
      return fail(SA_ERR_UNSUPPORTED, "the fp32 parity mode takes the assembled x (sa_hifigan_forward)");
    n_bn = h->code_dim; n_spk = h->cfg.input_dim - 1 - h->code_dim;
    if (n_spk < 0) return fail(SA_ERR_INVALID_ARG, "codebook dimension %d + 1 exceeds input_dim %d", h->code_dim, h->cfg.input_dim);
    x = nullptr; bn = nullptr;
  } else if (!x && (!bn || !f0 || !spk)) {
    return fail(SA_ERR_INVALID_ARG, "NULL argument");
  }
  if (!x && !vq_idx) {
    if (n_bn < 1 || n_spk < 0 || n_bn + 1 + n_spk != h->cfg.input_dim)
      return fail(SA_ERR_INVALID_ARG, "n_bn + 1 + n_spk = %d + 1 + %d != input_dim %d", n_bn, n_spk, h->cfg.input_dim);
    if (h->finalized && h->precision == SA_PRECISION_FP32)
      return fail(SA_ERR_UNSUPPORTED, "the fp32 parity mode takes the assembled x (sa_hifigan_forward)");
    if ((reinterpret_cast<uintptr_t>(bn) | reinterpret_cast<uintptr_t>(f0) | reinterpret_cast<uintptr_t>(spk)) & 3)
      return fail(SA_ERR_INVALID_ARG, "bn / f0 / spk must be 4-byte aligned");
    x = nullptr;
  }
  if (!h->finalized) return fail(SA_ERR_NOT_FINALIZED, "call sa_hifigan_finalize first");
  if (B < 1 || T < 1) return fail(SA_ERR_INVALID_ARG, "need B >= 1 and T >= 1 frames (got B=%d T=%d)", B, T);
  if (y_dtype != SA_DTYPE_F32 && y_dtype != SA_DTYPE_F16 && y_dtype != SA_DTYPE_PCM16)
    return fail(SA_ERR_INVALID_ARG, "y_dtype must be F32, F16 or PCM16");
  if ((x && (reinterpret_cast<uintptr_t>(x) & 15)) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return fail(SA_ERR_INVALID_ARG, "x must be 16-byte and workspace 256-byte aligned");
  const size_t need = sa_hifigan_workspace_bytes(h, B, T);
  if (workspace_bytes < need)
    return fail(SA_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
  if ((int64_t)T * h->stage_rate(h->cfg.n_stages - 1) + 1 > 0x7fffff00LL)
    return fail(SA_ERR_INVALID_ARG, "T=%d too long: chunk the utterance", T);
  if (frames_per_item)
    for (int b = 0; b < B; ++b)
      if (frames_per_item[b] < 1 || frames_per_item[b] > T)
        return fail(SA_ERR_INVALID_ARG, "frames_per_item[%d]=%d outside [1,%d]", b, frames_per_item[b], T);
  int cur = -1;
  SA_CUDA(cudaGetDevice(&cur));
  if (cur != h->device) SA_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  h->launches = 0;
  h->prof_n = 0;
  int rc;
  if (h->precision == SA_PRECISION_FP32) {
    rc = forward_f32(h, x, B, T, y, y_dtype, workspace, st);
  } else {
    sa::tc_forward_args a;
    a.cfg = &h->cfg; a.x = x; a.B = B; a.T = T; a.frames_per_item = frames_per_item; a.y = y; a.y_dtype = y_dtype;
    a.bn = bn; a.f0 = f0; a.spk = spk; a.n_bn = n_bn; a.n_spk = n_spk;
    a.vq_idx = vq_idx; a.spk_ids = spk_ids; a.codebook = h->d_codebook; a.n_codes = h->n_codes;
    a.workspace = workspace; a.stream = st; a.debug_tap = h->debug_tap; a.debug_out = h->debug_out;
    a.bf16 = h->precision == SA_PRECISION_BF16; a.n_sm = h->n_sm;
    std::vector<sa::tc_layer> layers(h->convs.size());
    for (size_t i = 0; i < h->convs.size(); ++i) {
      const sa_conv& c = h->convs[i];
      layers[i] = sa::tc_layer{&c.tc, c.d_w32, c.d_bias, c.cin, c.cout, c.k, c.dil, c.pad, c.stride, c.transposed};
    }
    a.layers = layers.data();
    a.n_layers = (int)layers.size();
    a.chains = (h->use_fused && !h->chains.empty()) ? h->chains.data() : nullptr;
    a.upg = (h->use_fused && !h->upg.empty()) ? h->upg.data() : nullptr;
    a.mark_ctx = h;
    a.mark = h->prof_on ? +[](void* ctx, int tag, cudaStream_t s) { static_cast<sa_hifigan*>(ctx)->mark(tag, s); }
                        : nullptr;
    const char* err = sa::tc_forward(h->tc, a, &h->launches);
    rc = err ? fail(SA_ERR_CUDA, "%s", err) : SA_OK;
  }
  h->mark(-1, st);
  if (cur != h->device) cudaSetDevice(cur);
  return rc;
}

int sa_hifigan_forward(sa_hifigan* h, const float* x, int32_t B, int32_t T, const int32_t* frames_per_item, void* y,
                       int32_t y_dtype, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  return forward_impl(h, x, nullptr, nullptr, nullptr, 0, 0, B, T, frames_per_item, y, y_dtype, workspace, workspace_bytes,
                      stream);
}

int sa_hifigan_forward_parts(sa_hifigan* h, const float* bn, int32_t n_bn, const float* f0, const float* spk, int32_t n_spk,
                             int32_t B, int32_t T, const int32_t* frames_per_item, void* y, int32_t y_dtype, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (!bn || !f0 || !spk) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  return forward_impl(h, nullptr, bn, f0, spk, n_bn, n_spk, B, T, frames_per_item, y, y_dtype, workspace, workspace_bytes,
                      stream);
}

int sa_hifigan_set_codebook(sa_hifigan* h, const float* codebook, int32_t n_codes, int32_t dim) {
  if (!h || !codebook) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  if (n_codes < 1 || n_codes > 255 || dim < 1 || dim + 1 > h->cfg.input_dim)
    return fail(SA_ERR_INVALID_ARG, "need 1 <= n_codes <= 255 and dim + 1 <= input_dim (got %d codes of %d)", n_codes, dim);
  SA_CUDA(cudaSetDevice(h->device));
  if (h->d_codebook) cudaFree(h->d_codebook);
  h->d_codebook = nullptr;
  SA_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->d_codebook), (size_t)n_codes * dim * sizeof(float)));
  SA_CUDA(cudaMemcpy(h->d_codebook, codebook, (size_t)n_codes * dim * sizeof(float), cudaMemcpyDefault));
  h->n_codes = n_codes; h->code_dim = dim;
  return SA_OK;
}

int sa_hifigan_forward_vq(sa_hifigan* h, const uint8_t* vq_idx, const float* f0, const int32_t* spk_ids, int32_t B, int32_t T,
                          const int32_t* frames_per_item, void* y, int32_t y_dtype, void* workspace, size_t workspace_bytes,
                          void* stream) {
  if (!vq_idx || !f0 || !spk_ids) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  return forward_impl(h, nullptr, nullptr, f0, nullptr, 0, 0, B, T, frames_per_item, y, y_dtype, workspace, workspace_bytes,
                      stream, vq_idx, spk_ids);
}

// ---- N3, last step of extract_bn: nearest-codeword assignment (VectorQuantizerEMA.forward in eval mode) -------------
// chain/nn.py:423-436 computes |x|^2 + |e|^2 - 2 x.e for every row against every code and takes the first argmin; :448-456
// returns inputs + (codeword - inputs).  HBM bound: one read of the row (4 dim bytes), one index byte (+ 4 dim bytes when
// the quantised rows are wanted).  One warp per row, dims over the lanes; the codebook and |e|^2 sit in shared memory.
constexpr int kVqMaxDimRegs = 16;                                         // dim <= 32 * 16
__global__ void __launch_bounds__(256) vq_assign_kernel(const float* __restrict__ bn, const float* __restrict__ codebook,
                                                        int64_t n_rows, int n_codes, int dim, int in_smem,
                                                        uint8_t* __restrict__ vq_idx, float* __restrict__ quantized) {
  extern __shared__ float vq_smem[];
  float* ee = vq_smem;                                                    // [n_codes]
  const float* cb = codebook;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  if (in_smem) {
    float* s = vq_smem + n_codes;
    for (int i = threadIdx.x; i < n_codes * dim; i += blockDim.x) s[i] = codebook[i];
    cb = s;
    __syncthreads();
  }
  for (int c = warp; c < n_codes; c += n_warps) {
    float a = 0.f;
    for (int d = lane; d < dim; d += 32) a = fmaf(cb[(size_t)c * dim + d], cb[(size_t)c * dim + d], a);
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) ee[c] = a;
  }
  __syncthreads();
  for (int64_t r = (int64_t)blockIdx.x * n_warps + warp; r < n_rows; r += (int64_t)gridDim.x * n_warps) {
    const float* x = bn + r * dim;
    float xr[kVqMaxDimRegs];
    float xx = 0.f;
#pragma unroll
    for (int j = 0; j < kVqMaxDimRegs; ++j) {
      const int d = lane + 32 * j;
      xr[j] = d < dim ? __ldg(x + d) : 0.f;
      xx = fmaf(xr[j], xr[j], xx);
    }
    for (int o = 16; o > 0; o >>= 1) xx += __shfl_xor_sync(0xffffffffu, xx, o);
    int best = 0;
    float best_d = INFINITY;
    for (int c = 0; c < n_codes; ++c) {
      const float* e = cb + (size_t)c * dim;
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < kVqMaxDimRegs; ++j) {
        const int d = lane + 32 * j;
        if (d < dim) dot = fmaf(xr[j], e[d], dot);
      }
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      const float dist = __fsub_rn(__fadd_rn(xx, ee[c]), __fmul_rn(2.f, dot));
      if (dist < best_d || c == 0) { best_d = dist; best = c; }          // first minimum, as torch.argmin
    }
    if (lane == 0) vq_idx[r] = (uint8_t)best;
    if (quantized) {
      const float* e = cb + (size_t)best * dim;
#pragma unroll
      for (int j = 0; j < kVqMaxDimRegs; ++j) {
        const int d = lane + 32 * j;
        if (d < dim) quantized[r * dim + d] = __fadd_rn(xr[j], __fsub_rn(e[d], xr[j]));
      }
    }
  }
}

// The same assignment as a register-tiled fp32 product (dim % 4 == 0): 24.6 kFLOP per 1 KB row in exact fp32 makes the step
// bound by the fp32 FMA pipe, not by HBM, so the warp-per-row form above (a shuffle tree per code) is kept for odd dims only.
// A CTA of 4 warps takes 128 rows; 16-wide slices of the dimension of the rows and of 48 codewords flow through a three-stage
// cp.async ring in shared memory (row-major, 20-float pitch: the 16-byte reads of a quarter warp fall on distinct banks).  A
// thread owns rows lane + 32 i (i < 4) x 12 codes: four 16-byte row reads + twelve broadcast 16-byte code reads per 192
// FMAs; warp w owns codes 12w .. 12w + 11 of each 48-wide group.  |e|^2 once per CTA, |x|^2 beside the products.
constexpr int kVqRows = 128, kVqCodes = 48, kVqK = 16, kVqPitch = 20, kVqStages = 3, kVqThreads = 128;
__device__ __forceinline__ void vq_cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__global__ void __launch_bounds__(kVqThreads, 3) vq_assign_tiled_kernel(const float* __restrict__ bn, const float* __restrict__ codebook,
                                                                        int64_t n_rows, int n_codes, int dim,
                                                                        uint8_t* __restrict__ vq_idx, float* __restrict__ quantized) {
  __shared__ __align__(16) float xs[kVqStages][kVqRows][kVqPitch];
  __shared__ __align__(16) float es[kVqStages][kVqCodes][kVqPitch];
  __shared__ float ees[256];
  __shared__ float best_d[4][kVqRows];
  __shared__ int best_i[4][kVqRows];
  __shared__ int win[kVqRows];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = warp * 12;
  for (int c = warp; c < n_codes; c += 4) {                                // |e|^2 of every code
    float a = 0.f;
    for (int d = lane; d < dim; d += 32) { const float v = __ldg(codebook + (size_t)c * dim + d); a = fmaf(v, v, a); }
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) ees[c] = a;
  }
  const int n_tiles = (int)((n_rows + kVqRows - 1) / kVqRows);
  const int n_chunks = (dim + kVqK - 1) / kVqK;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * kVqRows;
    float bd[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    int bi[4] = {0, 0, 0, 0};
    for (int g0 = 0; g0 < n_codes; g0 += kVqCodes) {
      auto load = [&](int chunk) {                                         // one slice of the rows and of the codes; always commits
        if (chunk < n_chunks) {
          const int st = chunk % kVqStages, k0 = chunk * kVqK;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = tid + kVqThreads * i, r = e >> 2, kq = (e & 3) * 4;
            const bool ok = row0 + r < n_rows && k0 + kq < dim;
            vq_cp_async16(&xs[st][r][kq], ok ? bn + (row0 + r) * dim + k0 + kq : bn, ok ? 16 : 0);
          }
          for (int e = tid; e < kVqCodes * 4; e += kVqThreads) {
            const int c = e >> 2, kq = (e & 3) * 4;
            const bool ok = g0 + c < n_codes && k0 + kq < dim;
            vq_cp_async16(&es[st][c][kq], ok ? codebook + (size_t)(g0 + c) * dim + k0 + kq : codebook, ok ? 16 : 0);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      float acc[4][12];
      float xx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 12; ++j) acc[i][j] = 0.f;
      __syncthreads();                                                     // the ring is free (previous group / tile done)
      load(0);
      load(1);
      for (int chunk = 0; chunk < n_chunks; ++chunk) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();                                                   // slice `chunk` landed; slice chunk - 1 consumed by all
        load(chunk + 2);
        const int st = chunk % kVqStages;
#pragma unroll
        for (int kk = 0; kk < kVqK; kk += 4) {
          float4 xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            xv[i] = *reinterpret_cast<const float4*>(&xs[st][lane + 32 * i][kk]);
            xx[i] = fmaf(xv[i].x, xv[i].x, xx[i]); xx[i] = fmaf(xv[i].y, xv[i].y, xx[i]);
            xx[i] = fmaf(xv[i].z, xv[i].z, xx[i]); xx[i] = fmaf(xv[i].w, xv[i].w, xx[i]);
          }
#pragma unroll
          for (int j = 0; j < 12; ++j) {
            const float4 ev = *reinterpret_cast<const float4*>(&es[st][c0 + j][kk]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[i][j] = fmaf(xv[i].x, ev.x, acc[i][j]); acc[i][j] = fmaf(xv[i].y, ev.y, acc[i][j]);
              acc[i][j] = fmaf(xv[i].z, ev.z, acc[i][j]); acc[i][j] = fmaf(xv[i].w, ev.w, acc[i][j]);
            }
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {                                     // ascending codes, strict <: the first minimum
          const int c = g0 + c0 + j;
          if (c < n_codes) {
            const float dist = __fsub_rn(__fadd_rn(xx[i], ees[c]), __fmul_rn(2.f, acc[i][j]));
            if (dist < bd[i] || c == 0) { bd[i] = dist; bi[i] = c; }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { best_d[warp][lane + 32 * i] = bd[i]; best_i[warp][lane + 32 * i] = bi[i]; }
    __syncthreads();
    {
      // the first minimum over the four warps' code subsets: the lowest index among the smallest distances
      float d = best_d[0][tid];
      int c = best_i[0][tid];
      for (int w = 1; w < 4; ++w) {
        const float dw = best_d[w][tid];
        const int cw = best_i[w][tid];
        if (dw < d || (dw == d && cw < c)) { d = dw; c = cw; }
      }
      win[tid] = c;
      if (row0 + tid < n_rows) vq_idx[row0 + tid] = (uint8_t)c;
    }
    if (quantized) {                                                       // inputs + (codeword - inputs), chain/nn.py:456
      __syncthreads();
      const int q4 = dim >> 2;
      for (int e = tid; e < kVqRows * q4; e += kVqThreads) {
        const int r = e / q4, k = (e - r * q4) * 4;
        if (row0 + r >= n_rows) break;
        const float4 x = __ldg(reinterpret_cast<const float4*>(bn + (row0 + r) * dim + k));
        const float4 c = __ldg(reinterpret_cast<const float4*>(codebook + (size_t)win[r] * dim + k));
        float4 o;
        o.x = __fadd_rn(x.x, __fsub_rn(c.x, x.x)); o.y = __fadd_rn(x.y, __fsub_rn(c.y, x.y));
        o.z = __fadd_rn(x.z, __fsub_rn(c.z, x.z)); o.w = __fadd_rn(x.w, __fsub_rn(c.w, x.w));
        *reinterpret_cast<float4*>(quantized + (row0 + r) * dim + k) = o;
      }
    }
    __syncthreads();
  }
}

int sa_hifigan_vq_assign(sa_hifigan* h, const float* bn, int64_t n_rows, uint8_t* vq_idx, float* quantized, void* stream) {
  if (!h) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  if (!h->d_codebook) return fail(SA_ERR_NOT_FINALIZED, "call sa_hifigan_set_codebook first");
  if (n_rows == 0) return SA_OK;
  if (!bn || !vq_idx) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  if (n_rows < 0) return fail(SA_ERR_INVALID_ARG, "n_rows %lld < 0", (long long)n_rows);
  if (n_rows == 0) return SA_OK;
  SA_CUDA(cudaSetDevice(h->device));
  if (h->code_dim % 4 == 0 && (reinterpret_cast<uintptr_t>(bn) & 15) == 0 && (!quantized || (reinterpret_cast<uintptr_t>(quantized) & 15) == 0)) {
    int n_sm = 0;
    SA_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device));
    const int64_t tiles = (n_rows + kVqRows - 1) / kVqRows;
    vq_assign_tiled_kernel<<<(int)std::min<int64_t>(tiles, (int64_t)n_sm * 3), kVqThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        bn, h->d_codebook, n_rows, h->n_codes, h->code_dim, vq_idx, quantized);
    SA_CUDA(cudaGetLastError());
    return SA_OK;
  }
  if (h->code_dim > 32 * kVqMaxDimRegs) return fail(SA_ERR_UNSUPPORTED, "codebook dimension %d > %d", h->code_dim, 32 * kVqMaxDimRegs);
  const size_t full = ((size_t)h->n_codes * h->code_dim + h->n_codes) * sizeof(float);
  const int in_smem = full <= 200 * 1024;
  const size_t smem = in_smem ? full : (size_t)h->n_codes * sizeof(float);
  if (smem > 48 * 1024) SA_CUDA(cudaFuncSetAttribute(vq_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 0;
  SA_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
  const int64_t blocks_needed = (n_rows + 7) / 8;
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 50 * 1024 ? 2 : 4);
  const int grid = (int)std::min<int64_t>(blocks_needed, (int64_t)sms * per_sm);
  vq_assign_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(bn, h->d_codebook, n_rows, h->n_codes, h->code_dim,
                                                                          in_smem, vq_idx, quantized);
  SA_CUDA(cudaGetLastError());
  return SA_OK;
}

// Shared tail of the stream-ordered host entries with TRIMMED output: after the forward, only the kept
// samples of every item, 320 * frames_per_item[b] + 1, go back to the host, packed one item after the other
// (pipeline.py:156 trims to the original length on the host after copying the whole padded batch).
static int d2h_trimmed(sa_hifigan* h, const void* yd, void* y_host, int32_t B, int32_t T, const int32_t* frames, size_t esz,
                       cudaStream_t st) {
  const int64_t Lout = sa_hifigan_output_length(h, T);
  size_t off = 0;
  for (int b = 0; b < B; ++b) {
    const size_t n = (size_t)sa_hifigan_output_length(h, frames[b]) * esz;
    SA_CUDA(cudaMemcpyAsync(static_cast<char*>(y_host) + off, static_cast<const char*>(yd) + (size_t)b * Lout * esz, n,
                            cudaMemcpyDeviceToHost, st));
    off += n;
  }
  return SA_OK;
}

static int host_async_common_checks(sa_hifigan* h, int32_t B, int32_t T, const int32_t* frames, int32_t y_dtype, void* dev_scratch,
                                    size_t dev_scratch_bytes) {
  if (!h->finalized) return fail(SA_ERR_NOT_FINALIZED, "call sa_hifigan_finalize first");
  if (B < 1 || T < 1) return fail(SA_ERR_INVALID_ARG, "need B >= 1 and T >= 1");
  if (!frames) return fail(SA_ERR_INVALID_ARG, "the trimmed entries need frames_per_item");
  for (int b = 0; b < B; ++b)
    if (frames[b] < 1 || frames[b] > T) return fail(SA_ERR_INVALID_ARG, "frames_per_item[%d]=%d outside [1,%d]", b, frames[b], T);
  if (y_dtype != SA_DTYPE_F32 && y_dtype != SA_DTYPE_F16 && y_dtype != SA_DTYPE_PCM16)
    return fail(SA_ERR_INVALID_ARG, "y_dtype must be F32, F16 or PCM16");
  const size_t need = sa_hifigan_host_scratch_bytes(h, B, T, y_dtype);
  if (dev_scratch_bytes < need) return fail(SA_ERR_WORKSPACE, "dev_scratch too small: %zu < %zu", dev_scratch_bytes, need);
  if (reinterpret_cast<uintptr_t>(dev_scratch) & 255) return fail(SA_ERR_INVALID_ARG, "dev_scratch must be 256-byte aligned");
  return SA_OK;
}

int sa_hifigan_synthesize_host_trimmed_async(sa_hifigan* h, const float* x_host, int32_t B, int32_t T,
                                             const int32_t* frames_per_item, void* y_host, int32_t y_dtype, void* dev_scratch,
                                             size_t dev_scratch_bytes, void* stream) {
  if (!h || !x_host || !y_host || !dev_scratch) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  int rc = host_async_common_checks(h, B, T, frames_per_item, y_dtype, dev_scratch, dev_scratch_bytes);
  if (rc != SA_OK) return rc;
  const size_t esz = (y_dtype == SA_DTYPE_F32) ? 4 : 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cur = -1;
  SA_CUDA(cudaGetDevice(&cur));
  if (cur != h->device) SA_CUDA(cudaSetDevice(h->device));
  char* base = static_cast<char*>(dev_scratch);
  const size_t x_bytes = (size_t)B * h->cfg.input_dim * T * sizeof(float);
  const size_t y_bytes = (size_t)B * (size_t)sa_hifigan_output_length(h, T) * esz;
  float* xd = reinterpret_cast<float*>(base);
  void* yd = base + align_up(x_bytes, 256);
  void* ws = base + align_up(x_bytes, 256) + align_up(y_bytes, 256);
  if (!h->compute_stream) {
    SA_CUDA(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  }
  SA_CUDA(cudaMemcpyAsync(xd, x_host, x_bytes, cudaMemcpyHostToDevice, st));
  SA_CUDA(cudaEventRecord(h->ev_in, st));
  SA_CUDA(cudaStreamWaitEvent(h->compute_stream, h->ev_in, 0));
  rc = sa_hifigan_forward(h, xd, B, T, frames_per_item, yd, y_dtype, ws, sa_hifigan_workspace_bytes(h, B, T), h->compute_stream);
  SA_CUDA(cudaEventRecord(h->ev_done, h->compute_stream));
  SA_CUDA(cudaStreamWaitEvent(st, h->ev_done, 0));
  if (rc == SA_OK) rc = d2h_trimmed(h, yd, y_host, B, T, frames_per_item, esz, st);
  if (cur != h->device) cudaSetDevice(cur);
  return rc;
}

int sa_hifigan_synthesize_host_vq_trimmed_async(sa_hifigan* h, const uint8_t* vq_host, const float* f0_host,
                                                const int32_t* spk_ids_host, int32_t B, int32_t T,
                                                const int32_t* frames_per_item, void* y_host, int32_t y_dtype,
                                                void* dev_scratch, size_t dev_scratch_bytes, void* stream) {
  if (!h || !vq_host || !f0_host || !spk_ids_host || !y_host || !dev_scratch) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  int rc = host_async_common_checks(h, B, T, frames_per_item, y_dtype, dev_scratch, dev_scratch_bytes);
  if (rc != SA_OK) return rc;
  const size_t esz = (y_dtype == SA_DTYPE_F32) ? 4 : 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cur = -1;
  SA_CUDA(cudaGetDevice(&cur));
  if (cur != h->device) SA_CUDA(cudaSetDevice(h->device));
  char* base = static_cast<char*>(dev_scratch);
  const size_t x_bytes = (size_t)B * h->cfg.input_dim * T * sizeof(float);      // the staging region of the dense entry
  const size_t y_bytes = (size_t)B * (size_t)sa_hifigan_output_length(h, T) * esz;
  const size_t f0_bytes = (size_t)B * T * sizeof(float), idx_bytes = (size_t)B * T, spk_bytes = (size_t)B * sizeof(int32_t);
  float* f0d = reinterpret_cast<float*>(base);
  int32_t* spkd = reinterpret_cast<int32_t*>(base + align_up(f0_bytes, 256));
  uint8_t* idxd = reinterpret_cast<uint8_t*>(base + align_up(f0_bytes, 256) + align_up(spk_bytes, 256));
  if (align_up(f0_bytes, 256) + align_up(spk_bytes, 256) + idx_bytes > align_up(x_bytes, 256))
    return fail(SA_ERR_WORKSPACE, "compact conditioning does not fit the x staging region");
  void* yd = base + align_up(x_bytes, 256);
  void* ws = base + align_up(x_bytes, 256) + align_up(y_bytes, 256);
  if (!h->compute_stream) {
    SA_CUDA(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  }
  SA_CUDA(cudaMemcpyAsync(f0d, f0_host, f0_bytes, cudaMemcpyHostToDevice, st));
  SA_CUDA(cudaMemcpyAsync(spkd, spk_ids_host, spk_bytes, cudaMemcpyHostToDevice, st));
  SA_CUDA(cudaMemcpyAsync(idxd, vq_host, idx_bytes, cudaMemcpyHostToDevice, st));
  SA_CUDA(cudaEventRecord(h->ev_in, st));
  SA_CUDA(cudaStreamWaitEvent(h->compute_stream, h->ev_in, 0));
  rc = sa_hifigan_forward_vq(h, idxd, f0d, spkd, B, T, frames_per_item, yd, y_dtype, ws, sa_hifigan_workspace_bytes(h, B, T),
                             h->compute_stream);
  SA_CUDA(cudaEventRecord(h->ev_done, h->compute_stream));
  SA_CUDA(cudaStreamWaitEvent(st, h->ev_done, 0));
  if (rc == SA_OK) rc = d2h_trimmed(h, yd, y_host, B, T, frames_per_item, esz, st);
  if (cur != h->device) cudaSetDevice(cur);
  return rc;
}

int sa_hifigan_check(sa_hifigan* h, void* stream) {
  if (!h) return fail(SA_ERR_INVALID_ARG, "NULL handle");
  SA_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  SA_CUDA(cudaGetLastError());
  if (sa::tc_error_raised(h->tc))
    return fail(SA_ERR_CUDA, "a tcgen05 kernel timed out waiting on an mbarrier (protocol bug); results are invalid");
  return SA_OK;
}

int sa_hifigan_chain_timing(sa_hifigan* h, int64_t* out, int32_t max_launches) {
  if (!h || !out) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  return sa::tc_read_chain_timing(h->tc, reinterpret_cast<long long*>(out), max_launches);
}

int sa_hifigan_set_profiling(sa_hifigan* h, int32_t enable) {
  if (!h) return fail(SA_ERR_INVALID_ARG, "NULL handle");
  h->prof_on = enable != 0;
  h->prof_n = 0;
  return SA_OK;
}

int sa_hifigan_get_profile(sa_hifigan* h, float* ms, int32_t* tags, int32_t max_n) {
  if (!h) return fail(SA_ERR_INVALID_ARG, "NULL handle");
  if (h->prof_n < 2) return 0;
  SA_CUDA(cudaEventSynchronize(h->prof_ev[h->prof_n - 1]));
  const int n = h->prof_n - 1;
  for (int i = 0; i < n && i < max_n; ++i) {
    float t = 0.f;
    SA_CUDA(cudaEventElapsedTime(&t, h->prof_ev[i], h->prof_ev[i + 1]));
    if (ms) ms[i] = t;
    if (tags) tags[i] = h->prof_tag[i];
  }
  return n;
}

// Host-buffer entry: the batch is cut into two halves that run on two internal streams, so the H2D / D2H copies
// of one half overlap the kernels of the other (the reference does audio.to(device) -> convert -> .cpu() strictly
// one after the other, bin/pipeline.py:104-149).  Scratch layout: [x | y | workspace] per half.
static int host_halves(int B) { return B >= 8 ? 2 : 1; }

static size_t host_part_bytes(const sa_hifigan* h, int B, int T, int y_dtype) {
  const size_t esz = (y_dtype == SA_DTYPE_F32) ? 4 : 2;
  return align_up((size_t)B * h->cfg.input_dim * T * sizeof(float), 256) +
         align_up((size_t)B * (size_t)sa_hifigan_output_length(h, T) * esz, 256) + sa_hifigan_workspace_bytes(h, B, T);
}

size_t sa_hifigan_host_scratch_bytes(const sa_hifigan* h, int32_t B, int32_t T, int32_t y_dtype) {
  if (!h || B < 1 || T < 1) return 0;
  const int nh = host_halves(B);
  size_t total = 0;
  for (int i = 0; i < nh; ++i) total += align_up(host_part_bytes(h, (B + nh - 1 - i) / nh, T, y_dtype), 256);
  return std::max(total, align_up(host_part_bytes(h, B, T, y_dtype), 256));   // the _async entry does not split
}

// Stream-ordered variant: H2D copy, forward and D2H copy are enqueued on `stream` and the call returns.  A caller
// that alternates two streams (with their own pinned buffers and scratch) overlaps the copies of one batch with
// the kernels of the other across calls; see satools_b200/synth.py.
int sa_hifigan_synthesize_host_async(sa_hifigan* h, const float* x_host, int32_t B, int32_t T,
                                     const int32_t* frames_per_item, void* y_host, int32_t y_dtype, void* dev_scratch,
                                     size_t dev_scratch_bytes, void* stream) {
  if (!h || !x_host || !y_host || !dev_scratch) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  if (!h->finalized) return fail(SA_ERR_NOT_FINALIZED, "call sa_hifigan_finalize first");
  if (B < 1 || T < 1) return fail(SA_ERR_INVALID_ARG, "need B >= 1 and T >= 1");
  if (y_dtype != SA_DTYPE_F32 && y_dtype != SA_DTYPE_F16 && y_dtype != SA_DTYPE_PCM16)
    return fail(SA_ERR_INVALID_ARG, "y_dtype must be F32, F16 or PCM16");
  const size_t need = align_up(host_part_bytes(h, B, T, y_dtype), 256);
  if (dev_scratch_bytes < need) return fail(SA_ERR_WORKSPACE, "dev_scratch too small: %zu < %zu", dev_scratch_bytes, need);
  if (reinterpret_cast<uintptr_t>(dev_scratch) & 255) return fail(SA_ERR_INVALID_ARG, "dev_scratch must be 256-byte aligned");
  const size_t esz = (y_dtype == SA_DTYPE_F32) ? 4 : 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cur = -1;
  SA_CUDA(cudaGetDevice(&cur));
  if (cur != h->device) SA_CUDA(cudaSetDevice(h->device));
  const int64_t Lout = sa_hifigan_output_length(h, T);
  char* base = static_cast<char*>(dev_scratch);
  const size_t x_bytes = (size_t)B * h->cfg.input_dim * T * sizeof(float);
  const size_t y_bytes = (size_t)B * (size_t)Lout * esz;
  float* xd = reinterpret_cast<float*>(base);
  void* yd = base + align_up(x_bytes, 256);
  void* ws = base + align_up(x_bytes, 256) + align_up(y_bytes, 256);
  if (!h->compute_stream) {
    SA_CUDA(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  }
  // copies on the caller's stream, kernels on the shared compute stream: H2D -> [ev_in] -> forward -> [ev_done] -> D2H.
  // (An event may be re-recorded by the next call: a wait refers to the record that preceded it.)
  SA_CUDA(cudaMemcpyAsync(xd, x_host, x_bytes, cudaMemcpyHostToDevice, st));
  SA_CUDA(cudaEventRecord(h->ev_in, st));
  SA_CUDA(cudaStreamWaitEvent(h->compute_stream, h->ev_in, 0));
  int rc = sa_hifigan_forward(h, xd, B, T, frames_per_item, yd, y_dtype, ws, sa_hifigan_workspace_bytes(h, B, T),
                              h->compute_stream);
  SA_CUDA(cudaEventRecord(h->ev_done, h->compute_stream));
  SA_CUDA(cudaStreamWaitEvent(st, h->ev_done, 0));
  if (rc == SA_OK) SA_CUDA(cudaMemcpyAsync(y_host, yd, y_bytes, cudaMemcpyDeviceToHost, st));
  if (cur != h->device) cudaSetDevice(cur);
  return rc;
}

// Parts variant of the stream-ordered host entry: H2D of bn + f0 + spk (49 % of the bytes of x for the 504-channel
// model), forward_parts on the compute stream, D2H.  Same scratch size and layout as the x variant.
int sa_hifigan_synthesize_host_parts_async(sa_hifigan* h, const float* bn_host, int32_t n_bn, const float* f0_host,
                                           const float* spk_host, int32_t n_spk, int32_t B, int32_t T,
                                           const int32_t* frames_per_item, void* y_host, int32_t y_dtype, void* dev_scratch,
                                           size_t dev_scratch_bytes, void* stream) {
  if (!h || !bn_host || !f0_host || !spk_host || !y_host || !dev_scratch) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  if (!h->finalized) return fail(SA_ERR_NOT_FINALIZED, "call sa_hifigan_finalize first");
  if (B < 1 || T < 1) return fail(SA_ERR_INVALID_ARG, "need B >= 1 and T >= 1");
  if (n_bn < 1 || n_spk < 0 || n_bn + 1 + n_spk != h->cfg.input_dim)
    return fail(SA_ERR_INVALID_ARG, "n_bn + 1 + n_spk = %d + 1 + %d != input_dim %d", n_bn, n_spk, h->cfg.input_dim);
  if (y_dtype != SA_DTYPE_F32 && y_dtype != SA_DTYPE_F16 && y_dtype != SA_DTYPE_PCM16)
    return fail(SA_ERR_INVALID_ARG, "y_dtype must be F32, F16 or PCM16");
  const size_t need = align_up(host_part_bytes(h, B, T, y_dtype), 256);
  if (dev_scratch_bytes < need) return fail(SA_ERR_WORKSPACE, "dev_scratch too small: %zu < %zu", dev_scratch_bytes, need);
  if (reinterpret_cast<uintptr_t>(dev_scratch) & 255) return fail(SA_ERR_INVALID_ARG, "dev_scratch must be 256-byte aligned");
  const size_t esz = (y_dtype == SA_DTYPE_F32) ? 4 : 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cur = -1;
  SA_CUDA(cudaGetDevice(&cur));
  if (cur != h->device) SA_CUDA(cudaSetDevice(h->device));
  const int64_t Lout = sa_hifigan_output_length(h, T);
  char* base = static_cast<char*>(dev_scratch);
  const size_t x_bytes = (size_t)B * h->cfg.input_dim * T * sizeof(float);
  const size_t y_bytes = (size_t)B * (size_t)Lout * esz;
  const size_t bn_bytes = (size_t)B * n_bn * T * sizeof(float), f0_bytes = (size_t)B * T * sizeof(float);
  const size_t spk_bytes = (size_t)B * n_spk * sizeof(float);
  float* bnd = reinterpret_cast<float*>(base);                   // the three parts share the x staging region
  float* f0d = reinterpret_cast<float*>(base + align_up(bn_bytes, 256));
  float* spkd = reinterpret_cast<float*>(base + align_up(bn_bytes, 256) + align_up(f0_bytes, 256));
  if (align_up(bn_bytes, 256) + align_up(f0_bytes, 256) + spk_bytes > align_up(x_bytes, 256))
    return fail(SA_ERR_WORKSPACE, "parts do not fit the x staging region");
  void* yd = base + align_up(x_bytes, 256);
  void* ws = base + align_up(x_bytes, 256) + align_up(y_bytes, 256);
  if (!h->compute_stream) {
    SA_CUDA(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  }
  SA_CUDA(cudaMemcpyAsync(bnd, bn_host, bn_bytes, cudaMemcpyHostToDevice, st));
  SA_CUDA(cudaMemcpyAsync(f0d, f0_host, f0_bytes, cudaMemcpyHostToDevice, st));
  if (spk_bytes) SA_CUDA(cudaMemcpyAsync(spkd, spk_host, spk_bytes, cudaMemcpyHostToDevice, st));
  SA_CUDA(cudaEventRecord(h->ev_in, st));
  SA_CUDA(cudaStreamWaitEvent(h->compute_stream, h->ev_in, 0));
  int rc = sa_hifigan_forward_parts(h, bnd, n_bn, f0d, spkd, n_spk, B, T, frames_per_item, yd, y_dtype, ws,
                                    sa_hifigan_workspace_bytes(h, B, T), h->compute_stream);
  SA_CUDA(cudaEventRecord(h->ev_done, h->compute_stream));
  SA_CUDA(cudaStreamWaitEvent(st, h->ev_done, 0));
  if (rc == SA_OK) SA_CUDA(cudaMemcpyAsync(y_host, yd, y_bytes, cudaMemcpyDeviceToHost, st));
  if (cur != h->device) cudaSetDevice(cur);
  return rc;
}

int sa_hifigan_synthesize_host(sa_hifigan* h, const float* x_host, int32_t B, int32_t T,
                               const int32_t* frames_per_item, void* y_host, int32_t y_dtype, void* dev_scratch,
                               size_t dev_scratch_bytes, void* stream) {
  if (!h || !x_host || !y_host || !dev_scratch) return fail(SA_ERR_INVALID_ARG, "NULL argument");
  if (!h->finalized) return fail(SA_ERR_NOT_FINALIZED, "call sa_hifigan_finalize first");
  if (B < 1 || T < 1) return fail(SA_ERR_INVALID_ARG, "need B >= 1 and T >= 1");
  const size_t need = sa_hifigan_host_scratch_bytes(h, B, T, y_dtype);
  if (dev_scratch_bytes < need) return fail(SA_ERR_WORKSPACE, "dev_scratch too small: %zu < %zu", dev_scratch_bytes, need);
  if (reinterpret_cast<uintptr_t>(dev_scratch) & 255) return fail(SA_ERR_INVALID_ARG, "dev_scratch must be 256-byte aligned");
  const size_t esz = (y_dtype == SA_DTYPE_F32) ? 4 : 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cur = -1;
  SA_CUDA(cudaGetDevice(&cur));
  if (cur != h->device) SA_CUDA(cudaSetDevice(h->device));
  const int nh = host_halves(B);
  if (nh > 1 && !h->side_stream) {
    SA_CUDA(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    SA_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  if (nh > 1) {                                   // side stream starts after the caller's stream reaches this point
    SA_CUDA(cudaEventRecord(h->ev_fork, st));
    SA_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
  }
  const int64_t Lout = sa_hifigan_output_length(h, T);
  char* base = static_cast<char*>(dev_scratch);
  int rc = SA_OK, b0 = 0;
  int64_t launches = 0;
  for (int i = 0; i < nh && rc == SA_OK; ++i) {
    const int Bi = (B + nh - 1 - i) / nh;
    cudaStream_t si = (i == 0) ? st : h->side_stream;
    const size_t x_bytes = (size_t)Bi * h->cfg.input_dim * T * sizeof(float);
    const size_t y_bytes = (size_t)Bi * (size_t)Lout * esz;
    float* xd = reinterpret_cast<float*>(base);
    void* yd = base + align_up(x_bytes, 256);
    void* ws = base + align_up(x_bytes, 256) + align_up(y_bytes, 256);
    const size_t ws_bytes = sa_hifigan_workspace_bytes(h, Bi, T);
    SA_CUDA(cudaMemcpyAsync(xd, x_host + (size_t)b0 * h->cfg.input_dim * T, x_bytes, cudaMemcpyHostToDevice, si));
    rc = sa_hifigan_forward(h, xd, Bi, T, frames_per_item ? frames_per_item + b0 : nullptr, yd, y_dtype, ws, ws_bytes, si);
    launches += h->launches;
    if (rc == SA_OK)
      SA_CUDA(cudaMemcpyAsync(static_cast<char*>(y_host) + (size_t)b0 * Lout * esz, yd, y_bytes, cudaMemcpyDeviceToHost, si));
    base += align_up(host_part_bytes(h, Bi, T, y_dtype), 256);
    b0 += Bi;
  }
  h->launches = launches;
  if (nh > 1) {
    SA_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
    SA_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
  }
  if (rc == SA_OK) SA_CUDA(cudaStreamSynchronize(st));
  if (cur != h->device) cudaSetDevice(cur);
  return rc;
}

}  // extern "C"
