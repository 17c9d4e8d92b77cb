// Microbenchmark 3: tcgen05.ld (TMEM -> registers) throughput per SM, alone and under concurrent MMAs.
#include <cstdio>
#include "../sa-toolkit_b200/csrc/conv_tc.cuh"
using namespace sa::tc;

// nw epilogue warps each read `iters` x 16 columns of their lane group; optional MMA stream in warp 0.
__global__ void __launch_bounds__(1024, 1) tmem_rate(int iters, int nw, int with_mma, int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  long long t0 = clock64();
  if (warp == 0) {
    if (with_mma) {
      const bool leader = elect_one();
      const uint32_t idesc = make_idesc(64, false);
      const uint32_t hi = desc_hi(128);
      uint32_t a = desc_lo(smem_u32(smem)), b = desc_lo(smem_u32(smem) + 64 * 1024);
      for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (leader) umma_f16(tmem + 256 + g * 64, desc64(a + g * 1024, hi), desc64(b, hi), idesc, i > 0 ? 1u : 0u);
        a ^= 2u; b ^= 2u;
      }
      if (leader) umma_commit(smem_u32(&bar));
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0, nullptr);
      if (lane == 0) out[1] = clock64() - t0;
    }
  } else if (warp <= nw) {
    const int lg = warp & 3;
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)((i * 16) & 255), r);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) acc += r[e];
    }
    if (acc == 0x12345678u) out[7] = acc;
    if (warp == 1 && lane == 0) out[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * sizeof(long long));
  cudaFuncSetAttribute(tmem_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  for (int with_mma : {0, 1})
    for (int nw : {4, 8, 16, 24}) {
      cudaMemset(d, 0, 64);
      tmem_rate<<<148, 32 * (nw + 1), 140 * 1024>>>(iters, nw, with_mma, 16384, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[8];
      cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
      const double bytes = (double)nw * iters * 32 * 16 * 4;
      printf("warps=%2d mma=%d : ld %8lld cycles -> %6.1f B/cycle/SM (%.1f cycles per x16 ld per warp)   mma %lld cycles/16384 = %.1f per MMA  %s\n",
             nw, with_mma, h[0], bytes / h[0], (double)h[0] / iters, h[1], h[1] / 16384.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
