// fp32 CUDA-core kernels of the generator: the SA_PRECISION_FP32 parity mode.
//
// Activations are fp32 [B][C][L] (time contiguous), exactly the reference's tensor layout.
// Every conv is a direct convolution, fp32 operands, fp32 FMA accumulation, with the
// elementwise neighbours of the reference graph fused in:
//   leaky_relu on the input   (archi.py:80, nn.py:170,172)
//   bias                      (torch Conv1d / ConvTranspose1d)
//   residual add  x = xt + x  (nn.py:174)
// These kernels bound the precision of the whole path (SNR > 120 dB against the fp64
// oracle); the tensor-core kernels in conv_tc.cuh are the throughput path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace sa {

__device__ __forceinline__ float lrelu(float v, float slope) { return v >= 0.f ? v : v * slope; }

// ---------------------------------------------------------------------------------------
// Conv1d, stride 1, zero padding, dilation `dil`   (nn.py:170-173, archi.py:78)
//   y[b,co,t] = bias[co] + sum_ci sum_j w[co,ci,j] * lrelu(x[b,ci,t + j*dil - pad], slope_in)
//               (+ res[b,co,t])
// wt is the folded weight repacked as [Cin][k][Cout] so one (ci, j) row of COB output
// channels is contiguous.  Block = TT threads, one per output time step; each thread keeps
// COB output channels in registers.  grid = (ceil(L/TT), Cout/COB, B).
// ---------------------------------------------------------------------------------------
template <int COB, int TT, int CIB>
__global__ void __launch_bounds__(TT)
conv1d_f32_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                  const float* __restrict__ bias, const float* res, float* y,
                  int Cin, int Cout, int L, int k, int dil, int pad, float slope_in) {
  extern __shared__ float smem[];
  const int halo = (k - 1) * dil;
  const int xw = TT + halo;                 // staged input width
  float* xs = smem;                         // [CIB][xw]
  float* ws = smem + CIB * xw;              // [CIB][k][COB]

  const int t0 = blockIdx.x * TT;
  const int co0 = blockIdx.y * COB;
  const int b = blockIdx.z;
  const int tid = threadIdx.x;
  const float* xb = x + (size_t)b * Cin * L;

  float acc[COB];
#pragma unroll
  for (int c = 0; c < COB; ++c) acc[c] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += CIB) {
    const int nci = min(CIB, Cin - ci0);
    for (int i = tid; i < nci * xw; i += TT) {
      const int ci = i / xw, o = i - ci * xw;
      const int t = t0 + o - pad;
      float v = 0.f;
      if (t >= 0 && t < L) v = lrelu(__ldg(xb + (size_t)(ci0 + ci) * L + t), slope_in);
      xs[ci * xw + o] = v;
    }
    for (int i = tid; i < nci * k * COB; i += TT) {
      const int c = i % COB, r = i / COB;   // r = ci*k + j
      ws[i] = __ldg(wt + ((size_t)ci0 * k + r) * Cout + co0 + c);
    }
    __syncthreads();
    for (int ci = 0; ci < nci; ++ci) {
      const float* xrow = xs + ci * xw + tid;
      const float* wrow = ws + ci * k * COB;
      for (int j = 0; j < k; ++j) {
        const float xv = xrow[j * dil];
        const float4* w4 = reinterpret_cast<const float4*>(wrow + j * COB);
#pragma unroll
        for (int c = 0; c < COB / 4; ++c) {
          const float4 w = w4[c];
          acc[4 * c + 0] = fmaf(w.x, xv, acc[4 * c + 0]);
          acc[4 * c + 1] = fmaf(w.y, xv, acc[4 * c + 1]);
          acc[4 * c + 2] = fmaf(w.z, xv, acc[4 * c + 2]);
          acc[4 * c + 3] = fmaf(w.w, xv, acc[4 * c + 3]);
        }
      }
    }
    __syncthreads();
  }

  const int t = t0 + tid;
  if (t < L) {
#pragma unroll
    for (int c = 0; c < COB; ++c) {
      const size_t o = ((size_t)b * Cout + co0 + c) * L + t;
      float v = acc[c] + __ldg(bias + co0 + c);
      if (res != nullptr) v += res[o];
      y[o] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// ConvTranspose1d(stride u, padding p) as a polyphase filter bank   (archi.py:47-59,80-81)
//   y[b,co,n] = bias[co] + sum_ci sum_m lrelu(x[b,ci,i0-m], slope_in) * w[ci,co,j0+u*m]
//   with j0 = (n+p) % u, i0 = (n+p) / u  (only taps with (n+p-j) % u == 0 contribute).
// wt repacked as [Cin][k][Cout].  One thread per output sample n, COB channels each.
// grid = (ceil(Lout/TT), Cout/COB, B).
// ---------------------------------------------------------------------------------------
template <int COB, int TT, int CIB>
__global__ void __launch_bounds__(TT)
convt1d_f32_kernel(const float* __restrict__ x, const float* __restrict__ wt,
                   const float* __restrict__ bias, float* y,
                   int Cin, int Cout, int Lin, int k, int u, int p, float slope_in) {
  extern __shared__ float smem[];
  const int Lout = Lin * u;
  const int n0 = blockIdx.x * TT;
  const int co0 = blockIdx.y * COB;
  const int b = blockIdx.z;
  const int tid = threadIdx.x;
  const int taps = (k + u - 1) / u;
  // input rows needed by this tile: [ilo, ihi]
  const int ilo = (n0 + p) / u - (taps - 1);
  const int ihi = (n0 + TT - 1 + p) / u;
  const int xw = ihi - ilo + 1;             // <= TT/u + taps + 1
  float* xs = smem;                         // [CIB][xw_max]
  const int xw_max = TT / u + taps + 2;
  float* ws = smem + CIB * xw_max;          // [CIB][k][COB]
  const float* xb = x + (size_t)b * Cin * Lin;

  const int n = n0 + tid;
  const int j0 = (n + p) % u;
  const int i0 = (n + p) / u - ilo;         // row inside the staged tile

  float acc[COB];
#pragma unroll
  for (int c = 0; c < COB; ++c) acc[c] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += CIB) {
    const int nci = min(CIB, Cin - ci0);
    for (int i = tid; i < nci * xw; i += TT) {
      const int ci = i / xw, o = i - ci * xw;
      const int t = ilo + o;
      float v = 0.f;
      if (t >= 0 && t < Lin) v = lrelu(__ldg(xb + (size_t)(ci0 + ci) * Lin + t), slope_in);
      xs[ci * xw_max + o] = v;
    }
    for (int i = tid; i < nci * k * COB; i += TT) {
      const int c = i % COB, r = i / COB;
      ws[i] = __ldg(wt + ((size_t)ci0 * k + r) * Cout + co0 + c);
    }
    __syncthreads();
    for (int ci = 0; ci < nci; ++ci) {
      const float* xrow = xs + ci * xw_max;
      const float* wrow = ws + ci * k * COB;
      int m = 0;
      for (int j = j0; j < k; j += u, ++m) {
        const float xv = xrow[i0 - m];
        const float* w = wrow + j * COB;
#pragma unroll
        for (int c = 0; c < COB; ++c) acc[c] = fmaf(w[c], xv, acc[c]);
      }
    }
    __syncthreads();
  }
  if (n < Lout) {
#pragma unroll
    for (int c = 0; c < COB; ++c)
      y[((size_t)b * Cout + co0 + c) * Lout + n] = acc[c] + __ldg(bias + co0 + c);
  }
}

// ---------------------------------------------------------------------------------------
// Multi-receptive-field combine   (archi.py:82-86): xs = 0; xs += r0; xs += r1; xs += r2;
// x = xs / 3.  mode 0: s = r;  mode 1: s += r;  mode 2: out = (s + r) / n.
// ---------------------------------------------------------------------------------------
__global__ void mrf_combine_f32_kernel(float* s, const float* __restrict__ r, float* out,
                                       size_t n_elem, int mode, float n_blocks) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n_elem; i += stride) {
    const float rv = r[i];
    if (mode == 0) s[i] = rv;
    else if (mode == 1) s[i] += rv;
    else out[i] = (s[i] + rv) / n_blocks;
  }
}

// ---------------------------------------------------------------------------------------
// Tail   (archi.py:87-90): leaky_relu(h, 0.01) -> ReflectionPad1d((1,0)) -> Conv1d(C->1,k7,p3)
// -> tanh.  hp[0] = h[1], hp[i] = h[i-1]; output n in [0, L+1) reads hp[n-3 .. n+3], zero
// outside [0, L+1).  Output converted to y_dtype (0 f32, 1 f16, 4 pcm16).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void store_wave(void* y, size_t o, float v, int y_dtype) {
  if (y_dtype == 0) reinterpret_cast<float*>(y)[o] = v;
  else if (y_dtype == 1) reinterpret_cast<__half*>(y)[o] = __float2half_rn(v);
  else {
    float s = rintf(v * 32767.f);
    s = fminf(fmaxf(s, -32768.f), 32767.f);
    reinterpret_cast<int16_t*>(y)[o] = (int16_t)s;
  }
}

__global__ void conv_post_f32_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                     const float* __restrict__ bias, void* y,
                                     int C, int L, int k, float slope, int y_dtype) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int Lout = L + 1;
  if (n >= Lout) return;
  const int pad = (k - 1) / 2;
  const float* hb = h + (size_t)b * C * L;
  float acc = __ldg(bias);
  for (int ci = 0; ci < C; ++ci) {
    const float* hr = hb + (size_t)ci * L;
    for (int j = 0; j < k; ++j) {
      const int i = n + j - pad;            // index into the reflect-padded signal
      if (i < 0 || i >= Lout) continue;
      const int src = (i == 0) ? 1 : i - 1;
      acc = fmaf(__ldg(w + ci * k + j), lrelu(__ldg(hr + src), slope), acc);
    }
  }
  store_wave(y, (size_t)b * Lout + n, tanhf(acc), y_dtype);
}

}  // namespace sa
