// Microbenchmark 6: several MMA-issuing warps.  mma_bench5 shows that whatever the issuing thread does between its
// tcgen05.mma instructions (try_wait + fence ~90 cycles, commit ~50, __syncwarp / bookkeeping ~60 per iteration) is NOT hidden
// behind the queued MMAs: N = 64 runs at 64-85 cycles/MMA instead of 48.  Here W warps issue independent MMA streams (own
// accumulator, own barriers) with that per-block overhead: if the pipe takes MMAs from one warp while another does its
// bookkeeping, the aggregate rate returns to 48 cycles/MMA.
#include <cstdio>
#include <cstdlib>
#include "../sa-toolkit_b200/csrc/conv_tc.cuh"
using namespace sa::tc;

__global__ void __launch_bounds__(256, 1) mma_multi(int iters, int W, int blk, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[4], dummy[4][4], done;
  __shared__ uint32_t holder;
  __shared__ long long t_end[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&done), 1);
    for (int w = 0; w < 4; ++w) { mbar_init(smem_u32(&bar[w]), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&dummy[w][i]), 1); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (threadIdx.x == 255) mbar_arrive(smem_u32(&done));
  __syncthreads();
  const long long t0 = clock64();
  if (warp < W) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(64, false);
    constexpr uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a0 = desc_lo(smem_u32(smem) + 8 * 1024 + warp * 16 * 1024);
    const uint32_t b0 = desc_lo(smem_u32(smem) + 96 * 1024 + warp * 8 * 1024);
    int k = 0;
    for (int i = 0; i < iters; i += blk) {
      mbar_wait(smem_u32(&done), 0, nullptr);                     // "activations ready" (already complete)
      tc_fence_after();
      for (int j = 0; j < blk; j += 4) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (leader) umma_f16(tmem + (uint32_t)warp * 64u, desc64(a0 + (uint32_t)g * 2u, hi), desc64(b0 + (uint32_t)g * 2u, hi), idesc, (i | j | g) ? 1u : 0u);
      }
      if (leader) umma_commit(smem_u32(&dummy[warp][k]));         // "accumulator full"
      k = (k + 1) & 3;
      __syncwarp();
    }
    if (leader) umma_commit(smem_u32(&bar[warp]));
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0, nullptr);
    if (lane == 0) t_end[warp] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    long long mx = 0;
    for (int w = 0; w < W; ++w) mx = t_end[w] > mx ? t_end[w] : mx;
    out_cycles[blockIdx.x] = mx - t0;
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

void run(int W, int blk) {
  const int iters = 8064, ctas = 148;                             // multiple of 4, 12, 28
  long long* d;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaFuncSetAttribute(mma_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) mma_multi<<<ctas, 256, 180 * 1024>>>(iters, W, blk, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=64  %d issuing warp(s), blocks of %2d MMAs (wait + fence, MMAs, commit, syncwarp): %6.1f cycles/MMA aggregate  %s\n", W, blk,
         (double)mx / ((double)iters * W), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int blk : {4, 12, 28}) for (int W : {1, 2, 4}) run(W, blk);
  return 0;
}
