// Microbenchmark 2: MMA rate vs number of independent accumulators (G) with a lean issue loop,
// and the effect of concurrent bulk copies into shared memory (weight streaming).
#include <cstdio>
#include <cstdlib>
#include "../sa-toolkit_b200/csrc/conv_tc.cuh"
using namespace sa::tc;

template <int N, int G>
__global__ void __launch_bounds__(128, 1) mma_rate(int iters, int row_bytes, int stream_kb, const uint8_t* gsrc,
                                                    long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, wbar[2];
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 176 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&wbar[0]), 1); mbar_init(smem_u32(&wbar[1]), 1); fence_barrier_init(); }
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
    uint32_t a = desc_lo(smem_u32(smem));
    uint32_t b = desc_lo(smem_u32(smem) + 64 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += G) {
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (leader) umma_f16(tmem + (uint32_t)(g * N), desc64(a + (uint32_t)g * 128u * (row_bytes >> 4), hi), desc64(b, hi), idesc, i > 0 ? 1u : 0u);
      a ^= 2u; b ^= 2u;                                   // wiggle the K-half like a real loop
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr);
    const long long t1 = clock64();
    if (lane == 0) out_cycles[blockIdx.x] = t1 - t0;
  } else if (warp == 1 && stream_kb > 0) {
    // concurrent weight streaming: bulk copies of 16 KB into a scratch smem region while MMAs run
    const bool leader = elect_one();
    const int n = stream_kb / 16;
    uint32_t par[2] = {0, 0};
    for (int i = 0; i < n; ++i) {
      const int s = i & 1;
      if (i >= 2) { mbar_wait(smem_u32(&wbar[s]), par[s], nullptr); par[s] ^= 1; }
      if (leader) {
        mbar_arrive_expect_tx(smem_u32(&wbar[s]), 16384);
        bulk_load(smem_u32(smem) + 136 * 1024 + s * 16384, gsrc + (size_t)((blockIdx.x * 7 + i) % 64) * 16384, 16384, smem_u32(&wbar[s]));
      }
      __syncwarp();
    }
    for (int s = 0; s < 2 && s < n; ++s) mbar_wait(smem_u32(&wbar[s]), par[s], nullptr);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int G>
void run(int iters, int row_bytes, int stream_kb, int ctas, const uint8_t* gsrc) {
  long long* d;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaFuncSetAttribute(mma_rate<N, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) mma_rate<N, G><<<ctas, 128, 180 * 1024>>>(iters, row_bytes, stream_kb, gsrc, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d G=%d row_bytes=%3d stream=%5d KB ctas=%3d : %6.1f cycles/MMA (tensor ideal %d)  %s\n", N, G, row_bytes,
         stream_kb, ctas, (double)mx / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  uint8_t* gsrc;
  cudaMalloc(&gsrc, 64 * 16384);
  cudaMemset(gsrc, 0, 64 * 16384);
  const int it = 4096;
  run<256, 1>(it, 128, 0, 148, gsrc);
  run<256, 2>(it, 128, 0, 148, gsrc);
  run<256, 2>(it, 128, it * 4, 148, gsrc);      // 8 KB of weights per 2 MMAs, like stage 0
  run<128, 1>(it, 128, 0, 148, gsrc);
  run<128, 2>(it, 128, 0, 148, gsrc);
  run<128, 4>(it, 128, 0, 148, gsrc);
  run<128, 2>(it, 128, it * 2, 148, gsrc);
  run<64, 1>(it, 128, 0, 148, gsrc);
  run<64, 2>(it, 128, 2736, 148, gsrc);        // 8 KB of weights per 12 MMAs, like the C = 64 chain kernel (k = 3 taps)
  run<64, 2>(it, 128, 8192, 148, gsrc);        // 8 KB per 4 MMAs
  run<64, 2>(it, 128, 32768, 148, gsrc);       // 8 KB per MMA
  run<64, 2>(it, 128, 0, 148, gsrc);
  run<64, 4>(it, 128, 0, 148, gsrc);
  run<64, 8>(it, 128, 0, 148, gsrc);
  run<32, 1>(it, 64, 0, 148, gsrc);
  run<32, 2>(it, 64, 0, 148, gsrc);
  run<32, 4>(it, 64, 0, 148, gsrc);
  run<32, 8>(it, 64, 0, 148, gsrc);
  run<16, 1>(it, 32, 0, 148, gsrc);
  run<16, 2>(it, 32, 0, 148, gsrc);
  run<16, 4>(it, 32, 0, 148, gsrc);
  run<16, 8>(it, 32, 0, 148, gsrc);
  run<16, 16>(it, 32, 0, 148, gsrc);
  return 0;
}
