// Microbenchmark: raw tcgen05.mma issue/execute rate for the operand patterns of conv_tc.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_bench tools/mma_bench.cu
// Each CTA (one per SM) issues `iters` MMAs (M=128, N, K=16, fp16) from static smem and reports
// cycles per MMA.  mode 0: same descriptors every time; mode 1: A start walks over rows (tap shifts)
// and K halves like the conv kernel; msub: accumulators alternated.
#include <cstdio>
#include <cstdlib>
#include "../sa-toolkit_b200/csrc/conv_tc.cuh"
using namespace sa::tc;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate(int iters, int mode, int msub, int row_bytes, int use_commit_every,
                                                    long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(N, false);
    const uint32_t hi = desc_hi(row_bytes);
    const uint32_t row16 = row_bytes >> 4;
    const uint32_t a0 = desc_lo(smem_u32(smem));
    const uint32_t b0 = desc_lo(smem_u32(smem) + 64 * 1024);
    const long long t0 = clock64();
    uint32_t parity = 0;
    for (int i = 0; i < iters; ++i) {
      uint32_t a = a0, b = b0;
      if (mode == 1) {
        a += (uint32_t)((i % 11) * 3) * row16 + (uint32_t)((i & 3) * 2);   // tap row shift + K step inside the row
        b += (uint32_t)(i & 3) * 2 + (uint32_t)((i >> 2) & 1) * (uint32_t)(N * row16);
      }
      const uint32_t d = tmem + (uint32_t)((msub > 1 ? (i % msub) : 0) * N);
      if (leader) umma_f16(d, desc64(a, hi), desc64(b, hi), idesc, i >= msub ? 1u : 0u);
      if (use_commit_every > 0 && (i % use_commit_every) == use_commit_every - 1) {
        if (leader) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), parity, nullptr);
        parity ^= 1;
      }
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), parity, nullptr);
    const long long t1 = clock64();
    if (lane == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N>
void run(int iters, int mode, int msub, int row_bytes, int commit_every, int ctas) {
  long long* d;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaFuncSetAttribute(mma_rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) mma_rate<N><<<ctas, 128, 180 * 1024>>>(iters, mode, msub, row_bytes, commit_every, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d mode=%d msub=%d row_bytes=%3d commit_every=%3d ctas=%3d : %7.1f cycles/MMA (ideal %d)  %s\n", N, mode, msub,
         row_bytes, commit_every, ctas, (double)mx / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int iters = 4096;
  for (int ctas : {1, 148}) {
    run<256>(iters, 0, 1, 128, 0, ctas);
    run<256>(iters, 1, 1, 128, 0, ctas);
    run<256>(iters, 1, 2, 128, 0, ctas);
    run<256>(iters, 1, 2, 128, 8, ctas);
    run<128>(iters, 1, 2, 128, 0, ctas);
    run<64>(iters, 1, 2, 128, 0, ctas);
    run<32>(iters, 1, 2, 64, 0, ctas);
    run<16>(iters, 1, 2, 32, 0, ctas);
    run<16>(iters, 0, 1, 32, 0, ctas);
  }
  return 0;
}
