// YAAPT front end on the GPU (include/sa_yaapt.h): padding, squared signal, the two torchaudio biquads and the NLFER
// frame energies / voiced flags of a whole batch -- what `_yaapt` (satools/satools/hifigan/yaapt.py:873-899) does per
// utterance on one CPU thread before its trackers run.
//
// Kernels (all HBM / latency bound; nothing here is GEMM shaped):
//   biquad_chunk_kernel   one thread per 1024-sample chunk of one signal.  An IIR recursion is sequential, but its impulse
//                         response decays (pole radius 0.986 for the 50 Hz low-pass, 0.66 for the 1500 Hz high-pass), so a
//                         chunk can start from a zero state W samples early: after W = ln(1e-18) / ln(r) samples (3000 /
//                         100) what is left of the unknown true state is below double rounding.  Chunks that start at
//                         sample 0 have the exact zero state of torchaudio's lfilter.  Recursion in double (the reference's
//                         float32 recursion carries ~5e-5 relative noise on this signal; double is the value it approximates),
//                         one DFMA on the critical path per sample, output clamped to [-1, 1] and stored as float32 like
//                         the reference's tensors.
//   nlfer_frame_kernel    one block per frame: the Hann-windowed frame in shared memory, one thread per DFT bin of the F0
//                         band (bins 60..204 of 8192 for the defaults): direct DFT with an exact twiddle every 64 samples
//                         (sincospi of the reduced integer phase) and a complex rotation in between, |X| summed over the band.
//   nlfer_normalize_kernel one block per utterance: mean over its frames, energy / mean, voiced = energy > threshold.
#include "../../include/sa_yaapt.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <string.h>

namespace {

thread_local char g_err[256] = "";
int fail(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}

struct Biquad { double b0, b1, b2, a1, a2; int warm; };

// torchaudio lowpass_biquad / highpass_biquad for a float32 waveform: every step in float32 (filtering.py), then
// lfilter's b / a0 and a / a0, also float32.
Biquad design(bool lowpass, double fs, double cutoff) {
  const float w0 = (float)(2.0 * M_PI) * (float)cutoff / (float)(int)fs;
  const float alpha = sinf(w0) / 2.0f / 0.707f;
  const float c = cosf(w0);
  const float b0 = lowpass ? (1.0f - c) / 2.0f : (1.0f + c) / 2.0f;
  const float b1 = lowpass ? 1.0f - c : -1.0f - c;
  const float a0 = 1.0f + alpha, a1 = -2.0f * c, a2 = 1.0f - alpha;
  Biquad q;
  q.b0 = (double)(b0 / a0); q.b1 = (double)(b1 / a0); q.b2 = q.b0;
  q.a1 = (double)(a1 / a0); q.a2 = (double)(a2 / a0);
  // warm-up: the state decays like r^n, r = sqrt(a2) for complex poles (else the larger real pole)
  const double disc = q.a1 * q.a1 - 4.0 * q.a2;
  double r = disc < 0 ? sqrt(q.a2) : fmax(fabs((-q.a1 + sqrt(disc)) / 2), fabs((-q.a1 - sqrt(disc)) / 2));
  r = fmin(fmax(r, 0.05), 0.99995);
  q.warm = (int)ceil(log(1e-18) / log(r));
  q.warm = (q.warm + 3) / 4 * 4;
  return q;
}

struct Geometry { int pad, frame_size, frame_jump, nfft, bin_lo, bin_hi; };

// Direct DFTs below: the twiddle of a bin advances by a complex rotation per sample and is recomputed exactly (sincospi of the
// integer phase reduced mod nfft) every kResync samples: 64 rotations accumulate ~4e-6 of relative error, and the exact
// twiddle costs ~45 instructions (every 16 samples it was a quarter of the issue slots of these issue-bound kernels).
constexpr int kResync = 64;

bool geometry(const sa_yaapt_params* p, Geometry& g) {
  if (!p || p->sr < 1000 || p->frame_length <= 0 || p->frame_space <= 0 || p->fft_length < 16) return false;
  g.pad = (int)(p->frame_length / 1000 * (int)p->sr) / 2;
  g.frame_size = (int)floor(p->frame_length * p->sr / 1000);
  g.frame_jump = (int)floor(p->frame_space * p->sr / 1000);
  g.nfft = (int)p->fft_length;
  // torch.round (half to even) of float32 products, yaapt.py:153-154
  const float lo = (float)(p->f0_min * 2 / p->sr) * (float)g.nfft, hi = (float)(p->f0_max / p->sr) * (float)g.nfft;
  g.bin_lo = (int)nearbyintf(lo) - 1;
  g.bin_hi = (int)nearbyintf(hi);
  return g.frame_size > 15 && g.frame_size < 2048 && g.frame_jump >= 1 && g.bin_lo >= 0 && g.bin_hi > g.bin_lo &&
         g.bin_hi <= g.nfft / 2 + 1 && g.frame_size <= g.nfft;
}

int64_t frames_of(const Geometry& g, int64_t n_samples) {
  const int64_t size = n_samples + 2 * g.pad, half = g.frame_size / 2;
  const int64_t span = size - half - half;                       // len(range(half, size - half, jump))
  return span <= 0 ? 0 : (span + g.frame_jump - 1) / g.frame_jump;
}


// src_mode 0: input = padded (and for odd signals squared) waveform; 1: input = `in` [2B, stride]
__global__ void biquad_chunk_kernel(const float* __restrict__ wav, const float* __restrict__ in, float* __restrict__ out_a,
                                    float* __restrict__ out_b, const int* __restrict__ lengths, int64_t n_max, int64_t stride,
                                    int pad, int n_chunks, int chunk_len, Biquad q, int src_mode, int B) {
  const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
  const int sig = blockIdx.y;                                     // 2 b + (0: signal, 1: squared signal)
  if (chunk >= n_chunks) return;
  const int b = sig >> 1, squared = sig & 1;
  const int64_t len = lengths ? lengths[b] : n_max;
  const int64_t size = len + 2 * pad;                             // samples of this item's padded signal
  float* out = (squared ? out_b : out_a);
  const int64_t o0 = (int64_t)chunk * chunk_len, o1 = min(o0 + chunk_len, stride);
  if (out) out += (src_mode == 0 ? (int64_t)sig : (int64_t)b) * stride;
  if (o0 >= size) {                                               // beyond the item: zeros
    if (out) for (int64_t i = o0; i < o1; ++i) out[i] = 0.f;
    return;
  }
  const float* w = wav ? wav + (int64_t)b * n_max : nullptr;
  const float* src = in ? in + (int64_t)sig * stride : nullptr;
  auto x_at = [&](int64_t i) -> double {
    if (src_mode == 0) {
      const int64_t j = i - pad;
      if (j < 0 || j >= len) return 0.0;
      const float v = w[j];
      return squared ? (double)(v * v) : (double)v;              // signal.data ** 2 is a float32 product (yaapt.py:877)
    }
    return (double)src[i];
  };
  const int64_t start = max((int64_t)0, o0 - q.warm);
  double x1 = start > 0 ? x_at(start - 1) : 0.0, x2 = start > 1 ? x_at(start - 2) : 0.0, y1 = 0.0, y2 = 0.0;
  const int64_t end = min(o1, size);
  for (int64_t i0 = start; i0 < end; i0 += 8) {                   // inputs of eight steps first: their loads overlap
    double xs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) xs[j] = i0 + j < end ? x_at(i0 + j) : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t i = i0 + j;
      const double x0 = xs[j];
      const double t = fma(q.b2, x2, fma(q.b1, x1, q.b0 * x0)) - q.a2 * y2;   // off the critical path
      const double y = fma(-q.a1, y1, t);
      x2 = x1; x1 = x0; y2 = y1; y1 = y;
      if (i >= o0 && i < end && out) out[i] = (float)fmin(1.0, fmax(-1.0, y));
    }
  }
  if (out) for (int64_t i = end; i < o1; ++i) out[i] = 0.f;
}

// The second biquad reads the first one's CLAMPED float32 output; both outputs of pass 2 go to separate user buffers, so
// the destination row is the item index there (out_a / out_b each [B, stride]) and the signal index in pass 1 ([2B, stride]).

__global__ void nlfer_frame_kernel(const float* __restrict__ filtered, float* __restrict__ frame_energy,
                                   const int* __restrict__ lengths, int64_t n_max, int64_t stride, int f_max, Geometry g) {
  extern __shared__ float frame[];                                // [frame_size] windowed samples, then [warps] partial sums
  const int f = blockIdx.x, b = blockIdx.y;
  const int64_t len = lengths ? lengths[b] : n_max;
  const int64_t size = len + 2 * g.pad, half = g.frame_size / 2;
  const int64_t span = size - half - half;
  const int64_t n_frames = span <= 0 ? 0 : (span + g.frame_jump - 1) / g.frame_jump;
  if (f >= n_frames) {
    if (threadIdx.x == 0) frame_energy[(int64_t)b * f_max + f] = 0.f;
    return;
  }
  const float* x = filtered + (int64_t)b * stride + (int64_t)f * g.frame_jump;
  for (int n = threadIdx.x; n < g.frame_size; n += blockDim.x) {
    // torch.hann_window(frame_size + 2)[1:-1] (periodic): 0.5 - 0.5 cos(2 pi (n + 1) / (frame_size + 2))
    const float wgt = (float)(0.5 - 0.5 * cospi(2.0 * (double)(n + 1) / (double)(g.frame_size + 2)));
    frame[n] = x[n] * wgt;
  }
  for (int n = g.frame_size + threadIdx.x; n < 4 * ((g.frame_size + 3) / 4); n += blockDim.x) frame[n] = 0.f;
  __syncthreads();
  float mag = 0.f;
  const int k = g.bin_lo + (int)threadIdx.x;
  if (k < g.bin_hi) {
    double sd, cd;
    sincospi(-2.0 * (double)k / (double)g.nfft, &sd, &cd);        // one-sample rotation e^{-2 pi i k / nfft}
    const float rc = (float)cd, rs = (float)sd;
    // Four quarter frames share every twiddle: X[k] = sum_j w^{k j Q} sum_{n < Q} x[n + j Q] w^{k n} -- one rotation per four
    // samples instead of one per sample (7 -> 4 instructions per sample); the samples past the frame are zeros.
    const int Q = (g.frame_size + 3) / 4;
    float pr[4] = {0.f, 0.f, 0.f, 0.f}, pi[4] = {0.f, 0.f, 0.f, 0.f};
    for (int n0 = 0; n0 < Q; n0 += kResync) {
      const int idx = (int)(((int64_t)k * n0) % g.nfft);          // exact phase of sample n0
      float ws, wc;
      sincospif(-2.0f * (float)idx / (float)g.nfft, &ws, &wc);
      const int n1 = min(n0 + kResync, Q);
      for (int n = n0; n < n1; ++n) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v = frame[n + j * Q];
          pr[j] = fmaf(v, wc, pr[j]);
          pi[j] = fmaf(v, ws, pi[j]);
        }
        const float t = wc * rc - ws * rs;
        ws = wc * rs + ws * rc;
        wc = t;
      }
    }
    float re = pr[0], im = pi[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) {
      float ws, wc;
      sincospif(-2.0f * (float)(int)(((int64_t)k * j * Q) % g.nfft) / (float)g.nfft, &ws, &wc);
      re += pr[j] * wc - pi[j] * ws;
      im += pr[j] * ws + pi[j] * wc;
    }
    mag = sqrtf(re * re + im * im);
  }
  // block sum (fixed order: lanes by shuffle, warps in order)
  for (int d = 16; d > 0; d >>= 1) mag += __shfl_xor_sync(0xffffffffu, mag, d);
  __syncthreads();
  float* part = frame;
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = mag;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) s += part[w];
    frame_energy[(int64_t)b * f_max + f] = s;
  }
}

__global__ void nlfer_normalize_kernel(const float* __restrict__ frame_energy, float* __restrict__ energy, uint8_t* __restrict__ vuv,
                                       float* __restrict__ mean_energy, const int* __restrict__ lengths, int64_t n_max, int f_max,
                                       Geometry g, float threshold) {
  __shared__ double part[32];
  const int b = blockIdx.x;
  const int64_t len = lengths ? lengths[b] : n_max;
  const int64_t size = len + 2 * g.pad, half = g.frame_size / 2;
  const int64_t span = size - half - half;
  const int n_frames = (int)(span <= 0 ? 0 : (span + g.frame_jump - 1) / g.frame_jump);
  const float* e = frame_energy + (int64_t)b * f_max;
  double s = 0.0;
  for (int f = threadIdx.x; f < n_frames; f += blockDim.x) s += (double)e[f];
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x + 31) / 32 ? part[threadIdx.x] : 0.0;
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (threadIdx.x == 0) part[0] = s;
  }
  __syncthreads();
  const float mean = n_frames > 0 ? (float)(part[0] / (double)n_frames) : 0.f;     // torch.mean of float32 energies
  if (threadIdx.x == 0 && mean_energy) mean_energy[b] = mean;
  for (int f = threadIdx.x; f < f_max; f += blockDim.x) {
    const float v = f < n_frames ? e[f] / mean : 0.f;
    if (energy) energy[(int64_t)b * f_max + f] = v;
    if (vuv) vuv[(int64_t)b * f_max + f] = (f < n_frames && v > threshold) ? 1 : 0;
  }
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// ---- spec_track's SHC (yaapt.py:184-231) ----------------------------------------------------------------------
struct ShcGeometry { int nframe, wl, half, min_shc, max_shc, n_harm, bin_lo, n_bins; };
// `peaks` (yaapt.py:383-497)
struct PeakParams { int maxpeaks, center, min_lag, max_lag; float t1, t2, inv_t1, f0_double, f0_half, merit_extra; double delta; };
constexpr int kMaxPeakList = 96;
constexpr int kMaxPeaksOut = 8;

bool shc_geometry(const sa_yaapt_params* p, const Geometry& g, ShcGeometry& s) {
  if (p->shc_numharms < 0 || p->shc_numharms > 7 || p->shc_window <= 0 || p->shc_pwidth < 0) return false;
  const double delta = p->sr / g.nfft;
  s.nframe = 2 * g.frame_size;
  s.wl = (int)floor(p->shc_window / delta);
  s.half = (int)floor((double)s.wl / 2);
  if (s.wl % 2 == 0) s.wl += 1;
  s.max_shc = (int)floor((p->f0_max + p->shc_pwidth * 2) / delta);
  s.min_shc = (int)ceil(p->f0_min / delta);
  s.n_harm = (int)p->shc_numharms + 1;
  s.bin_lo = s.min_shc - s.half < 0 ? 0 : s.min_shc - s.half;
  const int bin_hi = s.max_shc * s.n_harm + s.wl - 1 - s.half;          // last bin any product reads
  s.n_bins = bin_hi - s.bin_lo + 1;
  return s.min_shc >= 1 && s.max_shc >= s.min_shc && bin_hi <= g.nfft / 2 && s.nframe <= g.nfft && s.wl <= 64;
}

// torch.kaiser_window(n, periodic=True, beta): I0(beta sqrt(1 - (2 i / n - 1)^2)) / I0(beta)
__global__ void kaiser_kernel(float* w, int n, double beta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto i0 = [](double x) {
    double t = 1.0, sum = 1.0;
    for (int k = 1; k < 40; ++k) { t *= (x / 2) / k; sum += t * t; if (t * t < 1e-20 * sum) break; }
    return sum;
  };
  const double r = 2.0 * i / n - 1.0;
  w[i] = (float)(i0(beta * sqrt(fmax(0.0, 1.0 - r * r))) / i0(beta));
}

// One block per (frame, item).  smem: [nframe] windowed, mean-free samples | [n_bins] magnitudes | [32] partial sums.
__global__ void __launch_bounds__(1024, 1) shc_frame_kernel(const float* __restrict__ filtered_nl, const uint8_t* __restrict__ vuv, const float* __restrict__ window,
                                 float* __restrict__ shc, float* __restrict__ cand_pitch, float* __restrict__ cand_merit,
                                 const int* __restrict__ lengths, int64_t n_max, int64_t stride, int f_max, Geometry g, ShcGeometry s,
                                 PeakParams pk, int split_p) {
  // split_p = nfft / 64 when the two-level DFT below is used (0: the direct loop).  Its layout: [nfp] samples, zero padded to
  // a multiple of 64 | [64][P] inner sums | [P] e^{-2 pi i m / P} | [64] e^{-2 pi i m / nfft} | magnitudes | partial sums.
  extern __shared__ __align__(16) float sm[];
  float* frame = sm;
  const int nfp = split_p ? (s.nframe + 63) & ~63 : s.nframe;
  float2* S = reinterpret_cast<float2*>(sm + nfp);
  float2* WP = S + 64 * split_p;
  float2* WF = WP + split_p;
  float* mag = split_p ? reinterpret_cast<float*>(WF + 64) : sm + s.nframe;
  float* part = mag + s.n_bins;
  const int f = blockIdx.x, b = blockIdx.y;
  const int64_t len = lengths ? lengths[b] : n_max;
  const int64_t size = len + 2 * g.pad, half = g.frame_size / 2;
  const int64_t span = size - half - half;
  const int64_t n_frames = span <= 0 ? 0 : (span + g.frame_jump - 1) / g.frame_jump;
  float* out = shc ? shc + ((int64_t)b * f_max + f) * s.max_shc : nullptr;
  if (f >= n_frames || !vuv[(int64_t)b * f_max + f]) {
    if (out) for (int k = threadIdx.x; k < s.max_shc; k += blockDim.x) out[k] = 0.f;
    if (cand_pitch && (int)threadIdx.x < pk.maxpeaks) {                    // spec_track's initial values (yaapt.py:204-205)
      cand_pitch[((int64_t)b * pk.maxpeaks + threadIdx.x) * f_max + f] = 0.f;
      cand_merit[((int64_t)b * pk.maxpeaks + threadIdx.x) * f_max + f] = 1.f;
    }
    return;
  }
  // data[f jump : f jump + nframe] * window, zero beyond the signal (the reference pads `data` with zeros, yaapt.py:208-212)
  const float* x = filtered_nl + (int64_t)b * stride;
  const int64_t i0 = (int64_t)f * g.frame_jump;
  float psum = 0.f;
  for (int n = threadIdx.x; n < s.nframe; n += blockDim.x) {
    const int64_t i = i0 + n;
    const float v = (i < size && i < stride) ? x[i] * window[n] : 0.f;
    frame[n] = v;
    psum += v;
  }
  for (int d = 16; d > 0; d >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, d);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = psum;
  __syncthreads();
  float total = 0.f;
  for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) total += part[w];
  const float mean = total / (float)s.nframe;
  __syncthreads();
  for (int n = threadIdx.x; n < s.nframe; n += blockDim.x) frame[n] -= mean;
  if (split_p) {
    // Two-level DFT: n = 64 a + b gives X[k] = sum_b w^{k b} S[b][k mod P], S[b][r] = sum_a x[64 a + b] e^{-2 pi i r a / P}
    // with P = nfft / 64 -- the inner sums depend on k only through k mod P, so they are formed once for P residues
    // (~0.14 M real x complex products) and every bin costs 64 complex products instead of nframe (~1.1 M in all for
    // 1014 bins x 1120 samples).  Inner twiddles come exact from a table; the outer ones rotate 16 steps from an exact
    // product of two table entries.
    const int P = split_p;
    for (int n = s.nframe + threadIdx.x; n < nfp; n += blockDim.x) frame[n] = 0.f;
    for (int m = threadIdx.x; m < P + 64; m += blockDim.x) {
      double sd, cd;
      if (m < P) { sincospi(-2.0 * (double)m / (double)P, &sd, &cd); WP[m] = make_float2((float)cd, (float)sd); }
      else { sincospi(-2.0 * (double)(m - P) / (double)g.nfft, &sd, &cd); WF[m - P] = make_float2((float)cd, (float)sd); }
    }
    __syncthreads();
    const int A = nfp / 64;
    for (int task = threadIdx.x; task < 4 * P; task += blockDim.x) {     // (residue r, 16 of the 64 offsets b); r over the lanes
      const int r = task % P, bg = task / P;
      float ar[16], ai[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { ar[j] = 0.f; ai[j] = 0.f; }
      for (int a = 0; a < A; ++a) {
        const float2 tw = WP[(r * a) % P];
        const float4* xp = reinterpret_cast<const float4*>(frame + 64 * a + 16 * bg);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = xp[q];
          ar[4 * q + 0] = fmaf(v.x, tw.x, ar[4 * q + 0]); ai[4 * q + 0] = fmaf(v.x, tw.y, ai[4 * q + 0]);
          ar[4 * q + 1] = fmaf(v.y, tw.x, ar[4 * q + 1]); ai[4 * q + 1] = fmaf(v.y, tw.y, ai[4 * q + 1]);
          ar[4 * q + 2] = fmaf(v.z, tw.x, ar[4 * q + 2]); ai[4 * q + 2] = fmaf(v.z, tw.y, ai[4 * q + 2]);
          ar[4 * q + 3] = fmaf(v.w, tw.x, ar[4 * q + 3]); ai[4 * q + 3] = fmaf(v.w, tw.y, ai[4 * q + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) S[(16 * bg + j) * P + r] = make_float2(ar[j], ai[j]);
    }
    __syncthreads();
    auto tw_at = [&](int m) {                                              // e^{-2 pi i m / nfft}, 0 <= m < nfft
      const float2 a = WP[m >> 6], c = WF[m & 63];
      return make_float2(a.x * c.x - a.y * c.y, a.x * c.y + a.y * c.x);
    };
    for (int kb = threadIdx.x; kb < s.n_bins; kb += blockDim.x) {
      const int k = s.bin_lo + kb, r = k % P;
      const float2 rot = tw_at(k % g.nfft);
      float re = 0.f, im = 0.f;
      for (int b0 = 0; b0 < 64; b0 += 16) {
        float2 tw = tw_at((int)(((int64_t)k * b0) % g.nfft));
#pragma unroll
        for (int b = b0; b < b0 + 16; ++b) {
          const float2 sv = S[b * P + r];
          re = fmaf(sv.x, tw.x, re); re = fmaf(-sv.y, tw.y, re);
          im = fmaf(sv.x, tw.y, im); im = fmaf(sv.y, tw.x, im);
          const float t = tw.x * rot.x - tw.y * rot.y;
          tw.y = tw.x * rot.y + tw.y * rot.x;
          tw.x = t;
        }
      }
      mag[kb] = sqrtf(re * re + im * im);
    }
  }
  __syncthreads();
  // the direct loop: two bins per thread (kb and kb + half_bins) share every sample load
  const int half_bins = split_p ? 0 : (s.n_bins + 1) / 2;
  for (int kb = threadIdx.x; kb < half_bins; kb += blockDim.x) {
    const int ka = s.bin_lo + kb, kc = min(ka + half_bins, s.bin_lo + s.n_bins - 1);
    double sd, cd;
    sincospi(-2.0 * (double)ka / (double)g.nfft, &sd, &cd);
    const float rca = (float)cd, rsa = (float)sd;
    sincospi(-2.0 * (double)kc / (double)g.nfft, &sd, &cd);
    const float rcc = (float)cd, rsc = (float)sd;
    float rea = 0.f, ima = 0.f, rec = 0.f, imc = 0.f;
    for (int n0 = 0; n0 < s.nframe; n0 += kResync) {
      float wsa, wca, wsc, wcc;
      sincospif(-2.0f * (float)(int)(((int64_t)ka * n0) % g.nfft) / (float)g.nfft, &wsa, &wca);
      sincospif(-2.0f * (float)(int)(((int64_t)kc * n0) % g.nfft) / (float)g.nfft, &wsc, &wcc);
      const int n1 = min(n0 + kResync, s.nframe);
#pragma unroll 8
      for (int n = n0; n < n1; ++n) {
        const float v = frame[n];
        rea = fmaf(v, wca, rea);
        ima = fmaf(v, wsa, ima);
        rec = fmaf(v, wcc, rec);
        imc = fmaf(v, wsc, imc);
        const float ta = wca * rca - wsa * rsa;
        wsa = wca * rsa + wsa * rca;
        wca = ta;
        const float tc = wcc * rcc - wsc * rsc;
        wsc = wcc * rsc + wsc * rcc;
        wcc = tc;
      }
    }
    mag[kb] = sqrtf(rea * rea + ima * ima);
    if (kb + half_bins < s.n_bins) mag[kb + half_bins] = sqrtf(rec * rec + imc * imc);
  }
  __syncthreads();
  const int rows = s.max_shc - s.min_shc + 1;
  for (int k = threadIdx.x; k < s.max_shc; k += blockDim.x) {
    const int r = k - (s.min_shc - 1);
    float acc = 0.f;
    if (r >= 0 && r < rows) {
      for (int c = 0; c < s.wl; ++c) {
        float prod = 1.f;
        for (int h = 1; h <= s.n_harm; ++h) {
          const int bin = (s.min_shc + r) * h + c - s.half;                // magnitude[half + bin] of the reference
          prod *= (bin >= s.bin_lo) ? mag[bin - s.bin_lo] : 0.f;
        }
        acc += prod;
      }
    }
    if (out) out[k] = acc;
    frame[k] = acc;                                                        // the samples are no longer needed: `data` of peaks
  }
  if (!cand_pitch) return;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  // ---- peaks(SHC) by warp 0 (yaapt.py:383-497); float32 arithmetic where the reference has float32 tensors ----
  const int lane = threadIdx.x;
  float* data = frame;
  float* lst_p = frame + s.max_shc;                                        // candidates in increasing n
  float* lst_m = lst_p + kMaxPeakList;
  float vmax = -INFINITY;
  for (int n = pk.min_lag + lane; n <= pk.max_lag; n += 32) vmax = fmaxf(vmax, data[n]);
  for (int d = 16; d > 0; d >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
  if (vmax > 1e-14f)
    for (int n = lane; n < s.max_shc; n += 32) data[n] = data[n] / vmax;
  __syncwarp();
  float sum = 0.f;
  for (int n = pk.min_lag + lane; n <= pk.max_lag; n += 32) sum += data[n];
  for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
  const float avg = sum / (float)(pk.max_lag - pk.min_lag + 1);
  float* op = cand_pitch + (int64_t)b * pk.maxpeaks * f_max + f;
  float* om = cand_merit + (int64_t)b * pk.maxpeaks * f_max + f;
  auto unvoiced = [&]() {
    if (lane < pk.maxpeaks) { op[(int64_t)lane * f_max] = 0.f; om[(int64_t)lane * f_max] = 1.f; }
  };
  if (avg > pk.inv_t1) { unvoiced(); return; }
  // Step 1: strict local maxima above thresh2 * avg that are the first maximum of their +-center window, in increasing n
  const int lo = pk.min_lag + pk.center + 1, hi = pk.max_lag - pk.center + 1;
  const float floor_v = pk.t2 * avg;
  int count = 0;
  for (int n0 = lo; n0 < hi; n0 += 32) {
    const int n = n0 + lane;
    bool is = false;
    if (n < hi) {
      const float v = data[n];
      is = v > data[n - 1] && v > data[n + 1] && v > floor_v;
      if (is) {
        for (int j = n - pk.center; j < n; ++j) is = is && v > data[j];          // argmax returns the FIRST maximum
        for (int j = n + 1; j <= n + pk.center; ++j) is = is && v >= data[j];
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, is);
    if (is) {
      const int at = count + __popc(m & ((1u << lane) - 1u));
      if (at < kMaxPeakList) { lst_p[at] = (float)((double)n * pk.delta); lst_m[at] = data[n]; }
    }
    count += __popc(m);
  }
  count = min(count, kMaxPeakList);
  __syncwarp();
  if (lane != 0) return;
  // Step 2: the reference's merit vector has at least maxpeaks entries (zeros when fewer peaks were found)
  float best = count < pk.maxpeaks ? 0.f : -INFINITY;
  for (int i = 0; i < count; ++i) best = fmaxf(best, lst_m[i]);
  auto unvoiced0 = [&]() {
    for (int i = 0; i < pk.maxpeaks; ++i) { op[(int64_t)i * f_max] = 0.f; om[(int64_t)i * f_max] = 1.f; }
  };
  if (best / avg < pk.t1) { unvoiced0(); return; }
  // Step 3: the maxpeaks largest merits, descending (ties keep their order)
  float pit[kMaxPeaksOut], mer[kMaxPeaksOut];
  int numpeaks = min(count, pk.maxpeaks);
  for (int i = 0; i < pk.maxpeaks; ++i) {
    pit[i] = 0.f; mer[i] = 0.f;
    if (i >= numpeaks) continue;
    int arg = -1;
    for (int j = 0; j < count; ++j)
      if (lst_m[j] >= 0.f && (arg < 0 || lst_m[j] > lst_m[arg])) arg = j;
    pit[i] = lst_p[arg]; mer[i] = lst_m[arg];
    lst_m[arg] = -1.f;                                                    // taken (merits are >= 0)
  }
  // Step 4
  if (numpeaks == 0) { unvoiced0(); return; }
  if (pit[0] > pk.f0_double) {
    numpeaks = min(numpeaks + 1, pk.maxpeaks);
    pit[numpeaks - 1] = pit[0] / 2.0f; mer[numpeaks - 1] = pk.merit_extra;
  }
  if (pit[0] < pk.f0_half) {
    numpeaks = min(numpeaks + 1, pk.maxpeaks);
    pit[numpeaks - 1] = pit[0] * 2.0f; mer[numpeaks - 1] = pk.merit_extra;
  }
  for (int i = numpeaks; i < pk.maxpeaks; ++i) { pit[i] = pit[0]; mer[i] = mer[0]; }
  for (int i = 0; i < pk.maxpeaks; ++i) { op[(int64_t)i * f_max] = pit[i]; om[(int64_t)i * f_max] = mer[i]; }
}

// ---- the rest of spec_track (yaapt.py:233-316): one warp per utterance; utterances run in parallel --------------------
struct TrackParams { int maxpeaks, median_k; float f0_min, dp5_k1, min_std; };

__device__ float median_of(float* v, int k) {                  // k <= 9, sorts in place
  for (int i = 1; i < k; ++i) {
    const float x = v[i];
    int j = i - 1;
    while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; --j; }
    v[j + 1] = x;
  }
  return v[(k - 1) / 2];
}
__global__ void spec_track_finish_kernel(const float* __restrict__ cand_pitch, const float* __restrict__ cand_merit,
                                         float* __restrict__ spec_pitch, float* __restrict__ pitch_std, float* __restrict__ work,
                                         const int* __restrict__ lengths, int64_t n_max, int f_max, Geometry g, TrackParams tp) {
  // One warp per utterance.  Per-frame steps run with the frames spread over the lanes (ordered compactions by ballot + prefix
  // count); the dynamic programming is a chain over the voiced frames: lane a owns candidate row a, transition costs by shuffle.
  const int b = blockIdx.x, lane = threadIdx.x;
  const unsigned below = (1u << lane) - 1u;
  const int64_t len = lengths ? lengths[b] : n_max;
  const int64_t size = len + 2 * g.pad, half = g.frame_size / 2;
  const int64_t span = size - half - half;
  const int F = (int)(span <= 0 ? 0 : (span + g.frame_jump - 1) / g.frame_jump);
  float* out = spec_pitch + (int64_t)b * f_max;
  for (int f = lane; f < f_max; f += 32) out[f] = 0.f;
  if (F < 4) { if (lane == 0) pitch_std[b] = nanf(""); return; }
  const int M = tp.maxpeaks;
  const float* cp = cand_pitch + (int64_t)b * M * f_max;
  const float* cm = cand_merit + (int64_t)b * M * f_max;
  // per-item scratch: [M][F] vcp | [M][F] vcm | [F] a | [F] c | [F] voiced | [F] spec | [F] idx | [F] sel | [M][F] pred
  float* base = work + (int64_t)b * (size_t)(3 * M + 6) * f_max;
  float* vcp = base;
  float* vcm = vcp + (size_t)M * f_max;
  float* ta = vcm + (size_t)M * f_max;
  float* tc = ta + f_max;
  float* voiced = tc + f_max;
  float* spec = voiced + f_max;
  int* idx = reinterpret_cast<int*>(spec + f_max);
  int* sel = idx + f_max;
  int* pred = sel + f_max;
  int num = 0;
  for (int f0 = 0; f0 < F; f0 += 32) {                                   // voiced frames (cand_pitch[0] > 0), in order
    const int f = f0 + lane;
    const bool v = f < F && cp[f] > 0.f;
    if (f < F) spec[f] = cp[f];
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (v) {
      const int at = num + __popc(m & below);
      idx[at] = f;
      for (int r = 0; r < M; ++r) { vcp[(size_t)r * f_max + at] = cp[(size_t)r * f_max + f]; vcm[(size_t)r * f_max + at] = cm[(size_t)r * f_max + f]; }
    }
    num += __popc(m);
  }
  __syncwarp();
  auto wsum = [&](float v) { for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d); return v; };
  auto medfilt_w = [&](const float* xin, float* yout, int n) {           // medfilt (yaapt.py:54-70), frames over the lanes
    const int pad = tp.median_k / 2;
    for (int i = lane; i < n; i += 32) {
      float w[9];
      for (int j = 0; j < tp.median_k; ++j) { const int u = i - pad + j; w[j] = (u >= 0 && u < n) ? xin[u] : 0.f; }
      yout[i] = median_of(w, tp.median_k);
    }
    __syncwarp();
  };
  int n_voiced_out = num;
  if (num > 2) {
    float s = 0.f;
    for (int i = lane; i < num; i += 32) s += vcp[i];
    const float avg = wsum(s) / (float)num;
    float q = 0.f;
    for (int i = lane; i < num; i += 32) { const float d = vcp[i] - avg; q += d * d; }
    const float sd = sqrtf(wsum(q) / (float)(num - 1));
    // the candidate closest (merit weighted) to 0.8 avg of every frame, median smoothed (yaapt.py:243-255)
    for (int i = lane; i < num; i += 32) {
      int arg = 0;
      float bestv = INFINITY;
      for (int m = 0; m < M; ++m) {
        const float d1 = fabsf(vcp[(size_t)m * f_max + i] - 0.8f * avg) * (3.f - vcm[(size_t)m * f_max + i]);
        if (d1 < bestv) { bestv = d1; arg = m; }
      }
      sel[i] = arg;
      ta[i] = vcp[(size_t)arg * f_max + i];
    }
    __syncwarp();
    medfilt_w(ta, tc, num);
    for (int i = lane; i < num; i += 32) vcp[(size_t)sel[i] * f_max + i] = tc[i];
    __syncwarp();
    // dynamic5 / path1 (yaapt.py:506-569): trans[a, c, t] = k1 (0.05 d + d^2), d = |p[c, t] - p[a, t - 1]| / f0_min
    const float k1 = tp.dp5_k1 * sd / avg;
    const bool row = lane < M;
    const size_t mine = (size_t)(row ? lane : 0) * f_max;
    float pcost = row ? 1.f - vcm[mine] : INFINITY;
    float prv = row ? vcp[mine] : 0.f;
    auto trans = [&](float c, float pv) {
      const float d = fabsf(c - pv) / tp.f0_min;
      return k1 * (0.05f * d + d * d);
    };
    for (int t = 1; t < num; ++t) {
      const float cur = row ? vcp[mine + t] : 0.f;
      const float loc = row ? 1.f - vcm[mine + t] : 0.f;
      int kk = 0;
      float bst = INFINITY;
      for (int c = 0; c < M; ++c) {                                      // the LAST minimum (flip / argmin idiom)
        const float v = __shfl_sync(0xffffffffu, pcost, c) + trans(__shfl_sync(0xffffffffu, cur, c), prv);
        if (v <= bst) { bst = v; kk = c; }
      }
      const float ck = __shfl_sync(0xffffffffu, pcost, kk) + trans(cur, __shfl_sync(0xffffffffu, prv, kk)) + loc;
      if (row) pred[mine + t] = kk;
      pcost = row ? ck : INFINITY;
      prv = cur;
    }
    int last = 0;
    {
      float bc = INFINITY;
      for (int a = 0; a < M; ++a) {
        const float v = __shfl_sync(0xffffffffu, pcost, a);
        if (v <= bc) { bc = v; last = a; }
      }
    }
    __syncwarp();
    if (lane == 0) {
      int pth = last;                                                    // P[-1] = p_small[-1]
      for (int t = num - 1; t >= 0; --t) {
        ta[t] = vcp[(size_t)pth * f_max + t];
        if (t > 0) pth = pred[(size_t)pth * f_max + t];
      }
    }
    __syncwarp();
    medfilt_w(ta, voiced, num);
  } else {
    n_voiced_out = num > 0 ? num : 1;
    for (int i = lane; i < n_voiced_out; i += 32) voiced[i] = 150.f;
    __syncwarp();
  }
  float s = 0.f;
  for (int i = lane; i < n_voiced_out; i += 32) s += voiced[i];
  const float pavg = wsum(s) / (float)n_voiced_out;
  float sd = nanf("");
  if (n_voiced_out > 1) {
    float q = 0.f;
    for (int i = lane; i < n_voiced_out; i += 32) { const float d = voiced[i] - pavg; q += d * d; }
    sd = sqrtf(wsum(q) / (float)(n_voiced_out - 1));
  }
  if (lane == 0) pitch_std[b] = (sd != sd) ? sd : fmaxf(sd, pavg * tp.min_std);      // torch.maximum propagates NaN
  for (int i = lane; i < num; i += 32) spec[idx[i]] = voiced[i];
  __syncwarp();
  if (lane == 0) {
    if (spec[0] < pavg / 2.f) spec[0] = pavg;
    if (spec[F - 1] < pavg / 2.f) spec[F - 1] = pavg;
  }
  __syncwarp();
  int nz = 0;
  for (int f0 = 0; f0 < F; f0 += 32) {                                   // the non-zero values, in order
    const int f = f0 + lane;
    const bool v = f < F && spec[f] != 0.f;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    if (v) ta[nz + __popc(m & below)] = spec[f];
    nz += __popc(m);
  }
  __syncwarp();
  // F.interpolate(mode='linear', align_corners=False) of the nz non-zero values to F samples
  const float scale = (float)nz / (float)F;
  for (int f = lane; f < F; f += 32) {
    const float src = fmaxf(0.f, scale * ((float)f + 0.5f) - 0.5f);
    const int i0 = min((int)src, nz - 1), i1 = min(i0 + 1, nz - 1);
    const float lam = src - (float)i0;
    out[f] = (1.f - lam) * ta[i0] + lam * ta[i1];
  }
  __syncwarp();
  if (lane == 0) { out[0] = out[2]; out[1] = out[3]; }
}

// ---- time_track / refine / dynamic (yaapt.py:577-787, 321-372) -------------------------------------------------------
struct TdaParams { int tda_len, maxcands, center; float fs, f0_min, f0_max, t1, t2, boost; };

__device__ __forceinline__ int tda_frames(int64_t size, const TdaParams& q, int jump, int n_frames) {
  const int t = (int)((size - (q.tda_len - jump)) / jump);          // int((len(data) - noverlap) / frame_jump)
  return max(0, min(t, n_frames));
}

// crs_corr's `data -= mean(data)` acts on a view of the signal buffer: frame f + 1 sees what frame f left behind.  One block
// per (utterance, signal) walks the frames in order on a private copy of the signal.
__global__ void tda_mean_kernel(const float* __restrict__ fa, const float* __restrict__ fb, float* __restrict__ copy,
                                const float* __restrict__ pitch_std, const int* __restrict__ lengths, int64_t n_max, int64_t stride,
                                Geometry g, TdaParams q) {
  // One WARP per (utterance, signal): the frames are a chain (each mean sees the previous shifts), so what counts is the
  // latency of one step; with a block of 128 threads two __syncthreads per frame made it 1.6 ms per batch, a warp needs 0.3.
  const int b = blockIdx.x, sig = blockIdx.y, lane = threadIdx.x;
  const float* src = (sig ? fb : fa) + (int64_t)b * stride;
  float* x = copy + ((int64_t)sig * gridDim.x + b) * stride;
  for (int64_t i = threadIdx.x; i < stride; i += blockDim.x) x[i] = src[i];      // the private copy: the whole block
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int64_t len = lengths ? lengths[b] : n_max;
  const int64_t size = len + 2 * g.pad, half = g.frame_size / 2;
  const int64_t span = size - half - half;
  const int n_frames = (int)(span <= 0 ? 0 : (span + g.frame_jump - 1) / g.frame_jump);
  const float sd = pitch_std[b];
  if (sd != sd) return;                                              // NaN range: time_track skips every frame (yaapt.py:714)
  const int F = tda_frames(size, q, g.frame_jump, n_frames);
  for (int f = 0; f < F; ++f) {
    float* fr = x + (int64_t)f * g.frame_jump;
    float s = 0.f;
    for (int n = lane; n < q.tda_len; n += 32) s += fr[n];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    const float mean = s / (float)q.tda_len;
    for (int n = lane; n < q.tda_len; n += 32) fr[n] -= mean;
    __syncwarp();
  }
}

// The same result without the chain over the frames, for tda_len < 2 frame_jump (only frame f - 1 overlaps frame f, in its
// first c = tda_len - frame_jump samples): the sum frame f sees is S_f - c mean_{f-1} with S_f the sum of the ORIGINAL samples,
// so the S_f are formed in parallel, the means follow from a scalar recurrence (one thread, F steps of two operations) and the
// shifts are applied to every sample in parallel, in the reference's order (mean_{f-1} first, then mean_f).  The means differ
// from the chained ones by the rounding of c mean_{f-1} (1e-7 relative).  1.5 ms -> 0.05 ms per 64 x 10-15 s batch.
__global__ void tda_mean_split_kernel(const float* __restrict__ fa, const float* __restrict__ fb, float* __restrict__ copy,
                                      const float* __restrict__ pitch_std, const int* __restrict__ lengths, int64_t n_max,
                                      int64_t stride, Geometry g, TdaParams q) {
  extern __shared__ float mean_s[];                                  // [F]: S_f, then mean_f
  const int b = blockIdx.x, sig = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const float* src = (sig ? fb : fa) + (int64_t)b * stride;
  float* x = copy + ((int64_t)sig * gridDim.x + b) * stride;
  const int64_t len = lengths ? lengths[b] : n_max;
  const int64_t size = len + 2 * g.pad, half = g.frame_size / 2;
  const int64_t span = size - half - half;
  const int n_frames = (int)(span <= 0 ? 0 : (span + g.frame_jump - 1) / g.frame_jump);
  const float sd = pitch_std[b];
  const int F = sd != sd ? 0 : tda_frames(size, q, g.frame_jump, n_frames);   // NaN range: time_track skips every frame
  const int L = q.tda_len, J = g.frame_jump, c = L - J;
  for (int f = warp; f < F; f += n_warps) {
    const float* fr = src + (int64_t)f * J;
    float s = 0.f;
    for (int n = lane; n < L; n += 32) s += fr[n];
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) mean_s[f] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float prev = 0.f;
    for (int f = 0; f < F; ++f) {
      const float s = f > 0 ? fmaf(-(float)c, prev, mean_s[f]) : mean_s[f];
      prev = s / (float)L;
      mean_s[f] = prev;
    }
  }
  __syncthreads();
  for (int64_t i = threadIdx.x; i < stride; i += blockDim.x) {
    float v = src[i];
    const int f = (int)(i / J), off = (int)(i - (int64_t)f * J);
    if (f >= 1 && f - 1 < F && off < c) v -= mean_s[f - 1];
    if (f < F && off < L) v -= mean_s[f];
    x[i] = v;
  }
}

// One block per (frame, utterance, signal): NCCF over the lag range the spectral track allows, then cmp_rate and the merit
// weighting at the end of time_track.  tracks [2][B][maxcands][2 (pitch, merit)][F_max].
__global__ void nccf_frame_kernel(const float* __restrict__ copy, const float* __restrict__ spec_pitch, const float* __restrict__ pitch_std,
                                  float* __restrict__ tracks, const int* __restrict__ lengths, int64_t n_max, int64_t stride, int f_max,
                                  Geometry g, TdaParams q, int B) {
  extern __shared__ float sm[];
  float* x = sm;                          // [tda_len]
  float* phi = sm + q.tda_len;            // [tda_len]
  __shared__ float red[32];
  const int f = blockIdx.x, b = blockIdx.y, sig = blockIdx.z;
  float* tp = tracks + (((int64_t)sig * B + b) * q.maxcands) * 2 * f_max;
  auto put = [&](int row, float pitch, float merit) { tp[((int64_t)row * 2) * f_max + f] = pitch; tp[((int64_t)row * 2 + 1) * f_max + f] = merit; };
  const int64_t len = lengths ? lengths[b] : n_max;
  const int64_t size = len + 2 * g.pad, half = g.frame_size / 2;
  const int64_t span = size - half - half;
  const int n_frames = (int)(span <= 0 ? 0 : (span + g.frame_jump - 1) / g.frame_jump);
  const int F = tda_frames(size, q, g.frame_jump, n_frames);
  const float sd = pitch_std[b];
  const float sp = f < n_frames ? spec_pitch[(int64_t)b * f_max + f] : 0.f;
  const float lo = fmaxf(sp - 2.0f * sd, q.f0_min), hi = fminf(sp + 2.0f * sd, q.f0_max);
  const float qa = q.fs / hi, qb = q.fs / lo;
  if (f >= F || sd != sd || qa != qa || qb != qb) {                  // beyond the tracker's frames (zero padding of _yaapt) or skipped
    if ((int)threadIdx.x < q.maxcands) put(threadIdx.x, 0.f, 0.f);
    return;
  }
  const int lag_min = (int)floorf(qa) - q.center, lag_max = (int)floorf(qb) + q.center;
  const int N = q.tda_len - lag_max;                                 // > 0: checked on the host for the widest range
  const float* src = copy + ((int64_t)sig * B + b) * stride + (int64_t)f * g.frame_jump;
  for (int n = threadIdx.x; n < q.tda_len; n += blockDim.x) { x[n] = src[n]; phi[n] = 0.f; }
  __syncthreads();
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s += x[n] * x[n];
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float pw = 0.f;
  for (int w = 0; w < (int)blockDim.x / 32; ++w) pw += red[w];
  for (int lag = lag_min + (int)threadIdx.x; lag < lag_max; lag += blockDim.x) {
    if (lag < 0) continue;
    float nu = 0.f, de = 0.f;
    for (int j = 0; j < N; ++j) { const float v = x[lag + j]; nu = fmaf(v, x[j], nu); de = fmaf(v, v, de); }
    phi[lag] = nu / sqrtf(de * pw);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  // cmp_rate: the first strict local maximum above thresh1 in [lag_min + center, lag_max - center]
  float pitch = 0.f, merit = 0.f;
  int n0 = -1;
  for (int n = lag_min + q.center; n <= lag_max - q.center; ++n)
    if (phi[n] > phi[n - 1] && phi[n] > phi[n + 1] && phi[n] > q.t1) { n0 = n; break; }
  if (n0 >= 0) {
    float amax = 0.f;                                                  // phi is zero outside [lag_min, lag_max)
    for (int n = max(lag_min, 0); n < lag_max; ++n) amax = fmaxf(amax, phi[n]);
    bool take = amax > q.t2;
    if (!take) {                                                       // first maximum of its +-center window
      take = true;
      for (int j = n0 - q.center; j < n0; ++j) take = take && phi[n0] > phi[j];
      for (int j = n0 + 1; j <= n0 + q.center; ++j) take = take && phi[n0] >= phi[j];
    }
    if (take) { pitch = (float)((double)q.fs / (double)(n0 + 1)); merit = phi[n0]; }
  }
  if (merit > 1.0f) merit = merit / merit;
  const float thr = 5.0f * sd;
  for (int r = 0; r < q.maxcands; ++r) {
    const float pr = r == 0 ? pitch : 0.f, mr = r == 0 ? merit : 0.f;
    const float diff = fabsf(pr - sp);
    const float match = (1.0f - diff / thr) * (diff < thr ? 1.f : 0.f);
    put(r, pr, (q.boost * mr) * match);
  }
}

struct RefineParams { int maxcands, median_k; float thresh2, pivot, w1, w2, w3, w4; };

// refine + dynamic: one warp per utterance.  C = 2 maxcands candidate rows.  The per-frame work of refine (sorting the candidates
// of both trackers, the median filter, the rule table) runs with the frames spread over the lanes; the dynamic programming of
// `dynamic` / `path1` is a chain over the frames, so there lane a owns candidate row a (its path cost, its previous and current
// pitch) and the C x C transition costs of a frame are formed with shuffles: ~100 cycles per frame instead of a single thread
// re-reading the candidate matrices from global memory (6.4 -> 0.3 ms per 64 x 10-15 s batch).
__global__ void refine_dynamic_kernel(const float* __restrict__ tracks, const float* __restrict__ energy, const uint8_t* __restrict__ vuv,
                                      const float* __restrict__ spec_pitch, float* __restrict__ final_pitch, float* __restrict__ work,
                                      const int* __restrict__ lengths, int64_t n_max, int f_max, Geometry g, RefineParams rp, int B) {
  const int b = blockIdx.x, lane = threadIdx.x;
  float* out = final_pitch + (int64_t)b * f_max;
  for (int f = lane; f < f_max; f += 32) out[f] = 0.f;
  const int64_t len = lengths ? lengths[b] : n_max;
  const int64_t size = len + 2 * g.pad, half = g.frame_size / 2;
  const int64_t span = size - half - half;
  const int F = (int)(span <= 0 ? 0 : (span + g.frame_jump - 1) / g.frame_jump);
  if (F < 4) return;
  const int mc = rp.maxcands, C = 2 * mc;
  // scratch per item: P [C][F] | M [C][F] | best [F] | row0 [F] | pred (int) [C][F]
  float* P = work + (int64_t)b * (size_t)(3 * C + 2) * f_max;
  float* M = P + (size_t)C * f_max;
  float* best = M + (size_t)C * f_max;
  float* row0 = best + f_max;
  int* pred = reinterpret_cast<int*>(row0 + f_max);
  const float* e = energy + (int64_t)b * f_max;
  const float* sp = spec_pitch + (int64_t)b * f_max;
  const uint8_t* vv = vuv + (int64_t)b * f_max;
  for (int t = lane; t < F; t += 32) {                                 // candidates of both trackers, merits descending (stable)
    float pp[2 * kMaxPeaksOut], mm[2 * kMaxPeaksOut];
    for (int sgn = 0; sgn < 2; ++sgn)
      for (int r = 0; r < mc; ++r) {
        const float* tp = tracks + (((int64_t)sgn * B + b) * mc + r) * 2 * f_max;
        pp[sgn * mc + r] = tp[t]; mm[sgn * mc + r] = tp[f_max + t];
      }
    for (int i = 1; i < C; ++i) {
      const float pv = pp[i], mv = mm[i];
      int j = i - 1;
      while (j >= 0 && mm[j] < mv) { pp[j + 1] = pp[j]; mm[j + 1] = mm[j]; --j; }
      pp[j + 1] = pv; mm[j + 1] = mv;
    }
    for (int r = 0; r < C; ++r) { P[(size_t)r * f_max + t] = pp[r]; M[(size_t)r * f_max + t] = mm[r]; }
    row0[t] = pp[0];
  }
  __syncwarp();
  const int pad = rp.median_k / 2;
  for (int t = lane; t < F; t += 32) {                                 // medfilt(time_pitch[0], median_value) * vuv
    float w[9];
    for (int j = 0; j < rp.median_k; ++j) { const int u = t - pad + j; w[j] = (u >= 0 && u < F) ? row0[u] : 0.f; }
    best[t] = median_of(w, rp.median_k) * (vv[t] ? 1.f : 0.f);
  }
  __syncwarp();
  float s = 0.f;
  int cnt = 0;
  for (int t = lane; t < F; t += 32) {                                 // the rule table of refine (yaapt.py:757-785), in its order
    const float en = e[t], p0 = row0[t];
    const bool i1 = en <= rp.thresh2, i2 = en > rp.thresh2 && p0 > 0.f, i3 = en > rp.thresh2 && p0 <= 0.f;
    bool zero_mid[2 * kMaxPeaksOut];
    for (int r = 1; r < C - 1; ++r) zero_mid[r] = i2 && P[(size_t)r * f_max + t] == 0.f;
    if (i1) for (int r = 0; r < C; ++r) { P[(size_t)r * f_max + t] = 0.f; M[(size_t)r * f_max + t] = rp.pivot; }
    if (i2) { P[(size_t)(C - 1) * f_max + t] = 0.f; M[(size_t)(C - 1) * f_max + t] = 1.0f - M[t]; }
    for (int r = 1; r < C - 1; ++r) if (zero_mid[r]) M[(size_t)r * f_max + t] = 0.f;
    if (i3) {
      P[t] = sp[t];
      M[t] = fminf(1.0f, en / 2.0f);
      for (int r = 1; r < C; ++r) { P[(size_t)r * f_max + t] = 0.f; M[(size_t)r * f_max + t] = 1.0f - M[t]; }
    }
    P[(size_t)(C - 2) * f_max + t] = best[t];
    M[(size_t)(C - 2) * f_max + t] = best[t] > 0.f ? M[t] : 1.0f - fminf(1.0f, en / 2.0f);
    P[(size_t)(C - 3) * f_max + t] = sp[t];
    M[(size_t)(C - 3) * f_max + t] = en / 5.0f;
    if (best[t] > 0.f) { s += best[t]; ++cnt; }
  }
  for (int d = 16; d > 0; d >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, d); cnt += __shfl_xor_sync(0xffffffffu, cnt, d); }
  const float mean_pitch = s / (float)cnt;                              // mean of the non-zero best pitches
  __syncwarp();
  // dynamic / path1 (yaapt.py:321-372, 530-569): lane a = candidate row a
  const bool row = lane < C;
  const size_t mine = (size_t)(row ? lane : 0) * f_max;
  float pcost = row ? 1.0f - M[mine] : INFINITY;
  // The chain reads its per-frame inputs (and keeps its predecessors) in shared memory, staged 32 frames at a time by all
  // lanes: a dependent global load per step made the chain and the backtrack 1.2 ms per batch, staged they take 0.2.
  // The C x C transition costs of a frame do not depend on the chain, so lane i forms them for frame t0 + i of the chunk
  // (all lanes busy) and the chain itself is left with a shuffle, a shared-memory read and an add per pair: a single warp
  // exposes every instruction's latency, and the two divisions per cost made the chain ~1500 cycles per frame.
  constexpr int kC = 2 * kMaxPeaksOut;
  __shared__ float sP[kC][33], sM[kC][32], sE[33], sOut[32], sT[kC * kC][33];
  __shared__ int sPred[kC][32];
  for (int t0 = 1; t0 < F; t0 += 32) {
    const int nt = min(32, F - t0);
    if (lane < nt) {
      for (int r = 0; r < C; ++r) { sP[r][lane + 1] = P[(size_t)r * f_max + t0 + lane]; sM[r][lane] = M[(size_t)r * f_max + t0 + lane]; }
      sE[lane + 1] = e[t0 + lane];
    }
    if (lane == 0) {
      sE[0] = e[t0 - 1];
      for (int r = 0; r < C; ++r) sP[r][0] = P[(size_t)r * f_max + t0 - 1];
    }
    __syncwarp();
    if (lane < nt) {
      const float benefit = fminf(1.0f, fabsf(sE[lane] - sE[lane + 1]));
      auto trans = [&](float c, float pv) {                            // current pitch c, previous pitch pv
        float v = 1.0f;
        if (c > 0.f && pv > 0.f) v = rp.w1 * (fabsf(c - pv) / mean_pitch);
        else if ((c == 0.f && pv > 0.f) || (c > 0.f && pv == 0.f)) v = rp.w2 * (1.0f - benefit);
        else if (c == 0.f && pv == 0.f) v = rp.w3;
        return v / rp.w4;
      };
      for (int pr = 0; pr < C; ++pr)                                   // sT[pr * C + c]: previous pitch of row pr, current of row c
        for (int c = 0; c < C; ++c) sT[pr * C + c][lane] = trans(sP[c][lane + 1], sP[pr][lane]);
    }
    __syncwarp();
    const int a = row ? lane : 0;
    for (int i = 0; i < nt; ++i) {
      const float loc = row ? 1.0f - sM[a][i] : 0.f;
      int kk = 0;
      float bst = INFINITY;
      for (int c = 0; c < C; ++c) {                                    // aux[a, c] = PCOST[c] + trans[a, c, t]; the LAST minimum
        const float v = __shfl_sync(0xffffffffu, pcost, c) + sT[a * C + c][i];
        if (v <= bst) { bst = v; kk = c; }
      }
      const float ck = __shfl_sync(0xffffffffu, pcost, kk) + sT[kk * C + a][i] + loc;
      if (row) sPred[lane][i] = kk;
      pcost = row ? ck : INFINITY;
    }
    __syncwarp();
    if (lane < nt)
      for (int r = 0; r < C; ++r) pred[(size_t)r * f_max + t0 + lane] = sPred[r][lane];
    __syncwarp();
  }
  int last = 0;                                                        // the LAST minimum of the final costs
  {
    float bc = INFINITY;
    for (int a = 0; a < C; ++a) {
      const float v = __shfl_sync(0xffffffffu, pcost, a);
      if (v <= bc) { bc = v; last = a; }
    }
  }
  __syncwarp();
  int pth = last;
  for (int t1 = F; t1 > 0; t1 -= 32) {                                 // backtrack, frames [t0c, t1) per pass
    const int t0c = max(0, t1 - 32), nt = t1 - t0c;
    if (lane < nt)
      for (int r = 0; r < C; ++r) {
        sP[r][lane] = P[(size_t)r * f_max + t0c + lane];
        sPred[r][lane] = t0c + lane > 0 ? pred[(size_t)r * f_max + t0c + lane] : 0;
      }
    __syncwarp();
    if (lane == 0)
      for (int i = nt - 1; i >= 0; --i) {
        sOut[i] = sP[pth][i];
        if (t0c + i > 0) pth = sPred[pth][i];
      }
    pth = __shfl_sync(0xffffffffu, pth, 0);
    __syncwarp();
    if (lane < nt) out[t0c + lane] = sOut[lane];
    __syncwarp();
  }
}

}  // namespace

extern "C" {

const char* sa_yaapt_last_error(void) { return g_err; }

int sa_yaapt_default_params(sa_yaapt_params* p) {
  if (!p) return fail("sa_yaapt_default_params: NULL params");
  p->sr = 16000.0; p->frame_length = 35.0; p->frame_space = 10.0; p->f0_min = 60.0; p->f0_max = 400.0;
  p->fft_length = 8192.0; p->bp_low = 50.0; p->bp_high = 1500.0; p->nlfer_thresh1 = 0.75;
  p->shc_numharms = 3.0; p->shc_window = 40.0; p->shc_pwidth = 50.0;
  p->shc_maxpeaks = 4.0; p->shc_thresh1 = 5.0; p->shc_thresh2 = 1.25; p->f0_double = 150.0; p->f0_half = 150.0; p->merit_extra = 0.4;
  p->median_value = 7.0; p->dp5_k1 = 11.0; p->spec_pitch_min_std = 0.05;
  p->tda_frame_length = 35.0; p->nccf_thresh1 = 0.3; p->nccf_thresh2 = 0.9; p->nccf_maxcands = 3.0; p->nccf_pwidth = 5.0;
  p->merit_boost = 0.2; p->nlfer_thresh2 = 0.1; p->merit_pivot = 0.99; p->dp_w1 = 0.15; p->dp_w2 = 0.5; p->dp_w3 = 0.1; p->dp_w4 = 0.9;
  return 0;
}

int64_t sa_yaapt_padded_length(const sa_yaapt_params* p, int64_t n_samples) {
  Geometry g;
  if (!geometry(p, g) || n_samples < 0) { fail("sa_yaapt_padded_length: bad parameters"); return -1; }
  return n_samples + 2 * g.pad;
}

int64_t sa_yaapt_num_frames(const sa_yaapt_params* p, int64_t n_samples) {
  Geometry g;
  if (!geometry(p, g) || n_samples < 0) { fail("sa_yaapt_num_frames: bad parameters"); return -1; }
  return frames_of(g, n_samples);
}

size_t sa_yaapt_frontend_workspace_bytes(const sa_yaapt_params* p, int32_t B, int64_t n_max) {
  Geometry g;
  if (!geometry(p, g) || B <= 0 || n_max <= 0) return 0;
  const size_t stride = (size_t)(n_max + 2 * g.pad);
  // [lengths B][pass-1 output 2B x stride][pass-2 outputs 2 x B x stride (when the caller keeps none)][frame energies]
  return align256((size_t)B * 4) + 2 * align256((size_t)2 * B * stride * 4) + align256((size_t)B * (size_t)frames_of(g, n_max) * 4) + 256;
}

int sa_yaapt_frontend(const sa_yaapt_params* p, const float* wav, int32_t B, int64_t n_max, const int32_t* lengths,
                      float* filtered, float* filtered_nl, float* energy, uint8_t* vuv, float* mean_energy, void* workspace,
                      size_t workspace_bytes, void* stream) {
  Geometry g;
  if (!geometry(p, g)) return fail("sa_yaapt_frontend: bad parameters (frame_length must give 16..2047 samples, band inside the FFT)");
  if (!wav || B <= 0 || n_max <= 0) return fail("sa_yaapt_frontend: NULL waveform or empty batch");
  if (B > 32767) return fail("sa_yaapt_frontend: at most 32767 utterances per call");
  if (!workspace || workspace_bytes < sa_yaapt_frontend_workspace_bytes(p, B, n_max)) return fail("sa_yaapt_frontend: workspace too small");
  if (lengths)
    for (int b = 0; b < B; ++b)
      if (lengths[b] < 0 || lengths[b] > n_max) return fail("sa_yaapt_frontend: lengths[b] outside [0, n_max]");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t stride = n_max + 2 * g.pad;
  const int f_max = (int)frames_of(g, n_max);
  uint8_t* base = static_cast<uint8_t*>(workspace);
  int* d_len = reinterpret_cast<int*>(base);
  base += align256((size_t)B * 4);
  float* tmp1 = reinterpret_cast<float*>(base);                   // [2B, stride]
  base += align256((size_t)2 * B * stride * 4);
  float* own = reinterpret_cast<float*>(base);                    // [2, B, stride] when the caller keeps no filtered output
  base += align256((size_t)2 * B * stride * 4);
  float* d_fe = reinterpret_cast<float*>(base);                   // [B, f_max]
  if (lengths) {
    cudaError_t e = cudaMemcpyAsync(d_len, lengths, (size_t)B * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  }
  const int* dl = lengths ? d_len : nullptr;
  const Biquad lp = design(true, p->sr, p->bp_low), hp = design(false, p->sr, p->bp_high);
  // Chunk length: every thread also runs `warm` samples of warm-up, and the recursion is a latency chain.  Measured at
  // 64 x 10-15 s (whole front end): 128 -> 5.5 ms, 256 -> 3.8, 512 -> 3.2, 1024 -> 2.9, 2048 -> 3.0.  SATOOLS_B200_YAAPT_CHUNK overrides.
  auto chunk_for = [](const Biquad&) {
    if (const char* e = getenv("SATOOLS_B200_YAAPT_CHUNK")) { const int v = atoi(e); if (v >= 16 && v <= 65536) return v; }
    return 1024;
  };
  float* fa = filtered ? filtered : own;
  float* fb = filtered_nl ? filtered_nl : own + (size_t)B * stride;
  {
    const int cl = chunk_for(lp), n_chunks = (int)((stride + cl - 1) / cl);
    const dim3 grid((unsigned)((n_chunks + 127) / 128), (unsigned)(2 * B));
    biquad_chunk_kernel<<<grid, 128, 0, st>>>(wav, nullptr, tmp1, tmp1, dl, n_max, stride, g.pad, n_chunks, cl, lp, 0, B);
  }
  {
    const int cl = chunk_for(hp), n_chunks = (int)((stride + cl - 1) / cl);
    const dim3 grid((unsigned)((n_chunks + 127) / 128), (unsigned)(2 * B));
    biquad_chunk_kernel<<<grid, 128, 0, st>>>(nullptr, tmp1, fa, fb, dl, n_max, stride, g.pad, n_chunks, cl, hp, 1, B);
  }
  if (f_max > 0 && (energy || vuv || mean_energy)) {
    const int bins = g.bin_hi - g.bin_lo;
    const int threads = (bins + 31) / 32 * 32;
    if (threads > 1024) return fail("sa_yaapt_frontend: more than 1024 bins in the F0 band");
    const size_t smem = (size_t)(g.frame_size > 32 ? g.frame_size + 4 : 36) * sizeof(float);
    nlfer_frame_kernel<<<dim3((unsigned)f_max, (unsigned)B), threads, smem, st>>>(fa, d_fe, dl, n_max, stride, f_max, g);
    nlfer_normalize_kernel<<<B, 256, 0, st>>>(d_fe, energy, vuv, mean_energy, dl, n_max, f_max, g, (float)p->nlfer_thresh1);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  return 0;
}

int64_t sa_yaapt_shc_length(const sa_yaapt_params* p) {
  Geometry g;
  ShcGeometry sg;
  if (!geometry(p, g) || !shc_geometry(p, g, sg)) { fail("sa_yaapt_shc_length: bad parameters"); return -1; }
  return sg.max_shc;
}

size_t sa_yaapt_shc_workspace_bytes(const sa_yaapt_params* p, int32_t B, int64_t n_max) {
  Geometry g;
  ShcGeometry sg;
  if (!geometry(p, g) || !shc_geometry(p, g, sg) || B <= 0 || n_max <= 0) return 0;
  return align256((size_t)B * 4) + align256((size_t)sg.nframe * 4) + 256;
}

int sa_yaapt_shc(const sa_yaapt_params* p, const float* filtered_nl, int32_t B, int64_t n_max, const int32_t* lengths,
                 const uint8_t* vuv, float* shc, float* cand_pitch, float* cand_merit, void* workspace, size_t workspace_bytes,
                 void* stream) {
  Geometry g;
  ShcGeometry sg;
  if (!geometry(p, g) || !shc_geometry(p, g, sg)) return fail("sa_yaapt_shc: bad parameters (the SHC products must stay inside the FFT)");
  if (!filtered_nl || !vuv || B <= 0 || n_max <= 0 || B > 32767) return fail("sa_yaapt_shc: NULL argument or bad batch size");
  if ((cand_pitch == nullptr) != (cand_merit == nullptr)) return fail("sa_yaapt_shc: cand_pitch and cand_merit go together");
  if (!shc && !cand_pitch) return fail("sa_yaapt_shc: no output requested");
  PeakParams pk;
  {                                                                       // peaks, yaapt.py:388-410
    const double delta = p->sr / g.nfft;
    int width = (int)floor(p->shc_pwidth / delta);
    if (width % 2 == 0) width += 1;
    pk.center = (int)ceil(width / 2.0);
    pk.min_lag = (int)floor(p->f0_min / delta - pk.center);
    pk.max_lag = (int)floor(p->f0_max / delta + pk.center);
    if (pk.min_lag < 1) pk.min_lag = 1;
    if (pk.max_lag > sg.max_shc - width) pk.max_lag = sg.max_shc - width;
    pk.maxpeaks = (int)p->shc_maxpeaks;
    pk.t1 = (float)p->shc_thresh1; pk.t2 = (float)p->shc_thresh2; pk.inv_t1 = (float)(1.0 / p->shc_thresh1);
    pk.f0_double = (float)p->f0_double; pk.f0_half = (float)p->f0_half; pk.merit_extra = (float)p->merit_extra;
    pk.delta = delta;
    if (cand_pitch && (pk.maxpeaks < 1 || pk.maxpeaks > kMaxPeaksOut || pk.max_lag - pk.center <= pk.min_lag + pk.center + 1 ||
                       pk.max_lag + 1 >= sg.max_shc || sg.max_shc + 2 * kMaxPeakList > sg.nframe))
      return fail("sa_yaapt_shc: peak-picking parameters out of range");
  }
  if (!workspace || workspace_bytes < sa_yaapt_shc_workspace_bytes(p, B, n_max)) return fail("sa_yaapt_shc: workspace too small");
  if (lengths)
    for (int b = 0; b < B; ++b)
      if (lengths[b] < 0 || lengths[b] > n_max) return fail("sa_yaapt_shc: lengths[b] outside [0, n_max]");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t stride = n_max + 2 * g.pad;
  const int f_max = (int)frames_of(g, n_max);
  if (f_max == 0) return 0;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  int* d_len = reinterpret_cast<int*>(base);
  base += align256((size_t)B * 4);
  float* d_win = reinterpret_cast<float*>(base);
  if (lengths) {
    cudaError_t e = cudaMemcpyAsync(d_len, lengths, (size_t)B * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  }
  kaiser_kernel<<<(sg.nframe + 255) / 256, 256, 0, st>>>(d_win, sg.nframe, 0.5);
  const int half_bins = (sg.n_bins + 1) / 2;
  const int threads = half_bins >= 1024 ? 1024 : (std::max(half_bins, sg.max_shc > 256 ? 256 : sg.max_shc) + 31) / 32 * 32;
  size_t smem = (size_t)(sg.nframe + sg.n_bins + 32) * sizeof(float);
  // two-level DFT (see the kernel) when nfft = 64 P and the P x 64 inner sums fit beside the frame: 76 KB for nfft = 8192
  int split_p = 0;
  if (g.nfft % 64 == 0 && !getenv("SATOOLS_B200_YAAPT_DIRECT_DFT")) {
    const int P = g.nfft / 64;
    const size_t need = (size_t)(((sg.nframe + 63) & ~63) + sg.n_bins + 32 + 2 * (64 * P + P + 64)) * sizeof(float);
    if (need <= 100 * 1024) { split_p = P; smem = need; }
  }
  if (!split_p && smem > 48 * 1024) return fail("sa_yaapt_shc: frame + spectrum do not fit in 48 KB of shared memory");
  if (smem > 48 * 1024) {
    cudaError_t ea = cudaFuncSetAttribute(shc_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ea != cudaSuccess) return fail(cudaGetErrorString(ea));
  }
  shc_frame_kernel<<<dim3((unsigned)f_max, (unsigned)B), split_p ? 512 : threads, smem, st>>>(
      filtered_nl, vuv, d_win, shc, cand_pitch, cand_merit, lengths ? d_len : nullptr, n_max, stride, f_max, g, sg, pk, split_p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  return 0;
}

size_t sa_yaapt_spec_track_workspace_bytes(const sa_yaapt_params* p, int32_t B, int64_t n_max) {
  Geometry g;
  if (!geometry(p, g) || B <= 0 || n_max <= 0 || p->shc_maxpeaks < 1 || p->shc_maxpeaks > kMaxPeaksOut) return 0;
  const size_t f_max = (size_t)frames_of(g, n_max), M = (size_t)p->shc_maxpeaks;
  return align256((size_t)B * 4) + align256((size_t)B * (3 * M + 6) * f_max * 4) + 256;
}

int sa_yaapt_spec_track(const sa_yaapt_params* p, const float* cand_pitch, const float* cand_merit, int32_t B, int64_t n_max,
                        const int32_t* lengths, float* spec_pitch, float* pitch_std, void* workspace, size_t workspace_bytes,
                        void* stream) {
  Geometry g;
  if (!geometry(p, g)) return fail("sa_yaapt_spec_track: bad parameters");
  if (!cand_pitch || !cand_merit || !spec_pitch || !pitch_std || B <= 0 || n_max <= 0) return fail("sa_yaapt_spec_track: NULL argument or empty batch");
  const int k = (int)p->median_value - 2 < 1 ? 1 : (int)p->median_value - 2;
  if (p->shc_maxpeaks < 1 || p->shc_maxpeaks > kMaxPeaksOut || k > 9 || k % 2 == 0) return fail("sa_yaapt_spec_track: maxpeaks must be 1..8, median_value - 2 odd and <= 9");
  if (!workspace || workspace_bytes < sa_yaapt_spec_track_workspace_bytes(p, B, n_max)) return fail("sa_yaapt_spec_track: workspace too small");
  if (lengths)
    for (int b = 0; b < B; ++b)
      if (lengths[b] < 0 || lengths[b] > n_max) return fail("sa_yaapt_spec_track: lengths[b] outside [0, n_max]");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int f_max = (int)frames_of(g, n_max);
  if (f_max == 0) return 0;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  int* d_len = reinterpret_cast<int*>(base);
  base += align256((size_t)B * 4);
  if (lengths) {
    cudaError_t e = cudaMemcpyAsync(d_len, lengths, (size_t)B * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  }
  TrackParams tp;
  tp.maxpeaks = (int)p->shc_maxpeaks; tp.median_k = k; tp.f0_min = (float)p->f0_min; tp.dp5_k1 = (float)p->dp5_k1;
  tp.min_std = (float)p->spec_pitch_min_std;
  spec_track_finish_kernel<<<B, 32, 0, st>>>(cand_pitch, cand_merit, spec_pitch, pitch_std, reinterpret_cast<float*>(base),
                                             lengths ? d_len : nullptr, n_max, f_max, g, tp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  return 0;
}

size_t sa_yaapt_track_workspace_bytes(const sa_yaapt_params* p, int32_t B, int64_t n_max) {
  Geometry g;
  if (!geometry(p, g) || B <= 0 || n_max <= 0 || p->nccf_maxcands < 1 || p->nccf_maxcands > kMaxPeaksOut) return 0;
  const size_t f_max = (size_t)frames_of(g, n_max), mc = (size_t)p->nccf_maxcands, stride = (size_t)(n_max + 2 * g.pad);
  return align256((size_t)B * 4) + align256((size_t)2 * B * stride * 4) + align256((size_t)2 * B * mc * 2 * f_max * 4) +
         align256((size_t)B * (3 * 2 * mc + 2) * f_max * 4) + 256;
}

int sa_yaapt_track(const sa_yaapt_params* p, const float* filtered, const float* filtered_nl, const float* energy, const uint8_t* vuv,
                   const float* spec_pitch, const float* pitch_std, int32_t B, int64_t n_max, const int32_t* lengths,
                   float* final_pitch, void* workspace, size_t workspace_bytes, void* stream) {
  Geometry g;
  if (!geometry(p, g)) return fail("sa_yaapt_track: bad parameters");
  if (!filtered || !filtered_nl || !energy || !vuv || !spec_pitch || !pitch_std || !final_pitch || B <= 0 || n_max <= 0 || B > 32767)
    return fail("sa_yaapt_track: NULL argument or bad batch size");
  TdaParams q;
  q.tda_len = (int)(p->tda_frame_length * p->sr / 1000);
  q.maxcands = (int)p->nccf_maxcands;
  q.center = (int)floor(p->nccf_pwidth / 2.0);
  q.fs = (float)p->sr; q.f0_min = (float)p->f0_min; q.f0_max = (float)p->f0_max;
  q.t1 = (float)p->nccf_thresh1; q.t2 = (float)p->nccf_thresh2; q.boost = (float)(1.0 + p->merit_boost);
  const int k = (int)p->median_value;
  if (q.maxcands < 1 || q.maxcands > kMaxPeaksOut || k < 1 || k > 9 || k % 2 == 0 || q.tda_len < 16 || q.tda_len > 4096)
    return fail("sa_yaapt_track: nccf_maxcands must be 1..8, median_value odd and <= 9, tda_frame_length 16..4096 samples");
  // crs_corr asserts N = frame - lag_max > 0 (yaapt.py:583-587); the widest lag range is floor(fs / f0_min) + center
  if ((int)floor(p->sr / p->f0_min) + q.center + 2 >= q.tda_len || (int)floor(p->sr / p->f0_max) - q.center - 1 < 1)
    return fail("sa_yaapt_track: negative index in the cross correlation: increase tda_frame_length (as the reference asks)");
  if (!workspace || workspace_bytes < sa_yaapt_track_workspace_bytes(p, B, n_max)) return fail("sa_yaapt_track: workspace too small");
  if (lengths)
    for (int b = 0; b < B; ++b)
      if (lengths[b] < 0 || lengths[b] > n_max) return fail("sa_yaapt_track: lengths[b] outside [0, n_max]");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t stride = n_max + 2 * g.pad;
  const int f_max = (int)frames_of(g, n_max);
  if (f_max == 0) return 0;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  int* d_len = reinterpret_cast<int*>(base);
  base += align256((size_t)B * 4);
  float* copy = reinterpret_cast<float*>(base);
  base += align256((size_t)2 * B * stride * 4);
  float* tracks = reinterpret_cast<float*>(base);
  base += align256((size_t)2 * B * q.maxcands * 2 * f_max * 4);
  float* work = reinterpret_cast<float*>(base);
  if (lengths) {
    cudaError_t e = cudaMemcpyAsync(d_len, lengths, (size_t)B * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  }
  const int* dl = lengths ? d_len : nullptr;
  const int tda_overlap = q.tda_len - g.frame_jump;
  if (tda_overlap >= 0 && tda_overlap < g.frame_jump && (size_t)f_max * sizeof(float) <= 40 * 1024 && !getenv("SATOOLS_B200_YAAPT_TDA_CHAIN"))
    tda_mean_split_kernel<<<dim3((unsigned)B, 2), 512, (size_t)f_max * sizeof(float), st>>>(filtered, filtered_nl, copy, pitch_std, dl, n_max,
                                                                                       stride, g, q);
  else
    tda_mean_kernel<<<dim3((unsigned)B, 2), 256, 0, st>>>(filtered, filtered_nl, copy, pitch_std, dl, n_max, stride, g, q);
  const int lag_threads = 128;
  nccf_frame_kernel<<<dim3((unsigned)f_max, (unsigned)B, 2), lag_threads, (size_t)2 * q.tda_len * sizeof(float), st>>>(
      copy, spec_pitch, pitch_std, tracks, dl, n_max, stride, f_max, g, q, B);
  RefineParams rp;
  rp.maxcands = q.maxcands; rp.median_k = k; rp.thresh2 = (float)p->nlfer_thresh2; rp.pivot = (float)p->merit_pivot;
  rp.w1 = (float)p->dp_w1; rp.w2 = (float)p->dp_w2; rp.w3 = (float)p->dp_w3; rp.w4 = (float)p->dp_w4;
  refine_dynamic_kernel<<<B, 32, 0, st>>>(tracks, energy, vuv, spec_pitch, final_pitch, work, dl, n_max, f_max, g, rp, B);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
