// Microbenchmark 4: does the 48-cycle N = 64 MMA of mma_bench2/3 survive when no operand repeats?  There the A / B
// descriptors toggled between 1-4 values; here they walk over `na` distinct A operands (128 rows x 32 B slices of SWIZZLE_128B
// tiles) and `nb` distinct B operands.  In the fused kernels every MMA of a conv has a different A slice and weight slice.
#include <cstdio>
#include <cstdlib>
#include "../sa-toolkit_b200/csrc/conv_tc.cuh"
using namespace sa::tc;

template <int N>
__global__ void __launch_bounds__(128, 1) mma_walk(int iters, int na, int nb, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
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
    constexpr uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);     // SWIZZLE_128B, K-major
    const uint32_t a0 = desc_lo(smem_u32(smem));                               // A tiles: 16 KB each (128 rows x 128 B), 4 K-slices per tile
    const uint32_t b0 = desc_lo(smem_u32(smem) + 128 * 1024);                  // B tiles: N rows x 128 B, 4 K-slices per tile
    const uint32_t b_tile16 = (uint32_t)(N * 128) >> 4;
    int ia = 0, ib = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t a = a0 + (uint32_t)(ia >> 2) * 1024u + (uint32_t)(ia & 3) * 2u;
        const uint32_t b = b0 + (uint32_t)(ib >> 2) * b_tile16 + (uint32_t)(ib & 3) * 2u;
        if (leader) umma_f16(tmem, desc64(a, hi), desc64(b, hi), idesc, i > 0 ? 1u : 0u);
        if (++ia == na) ia = 0;
        if (++ib == nb) ib = 0;
      }
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr);
    const long long t1 = clock64();
    if (lane == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N>
void run(int na, int nb) {
  const int iters = 8192, ctas = 148;
  long long* d;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaFuncSetAttribute(mma_walk<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  for (int rep = 0; rep < 2; ++rep) mma_walk<N><<<ctas, 128, 202 * 1024>>>(iters, na, nb, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d distinct A=%2d B=%2d : %6.1f cycles/MMA (tensor ideal %d, operand bytes %d)  %s\n", N, na, nb, (double)mx / iters, N / 2,
         4096 + N * 32, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  // A region 128 KB = 8 tiles = 32 slices; B region: N = 64: 8 KB per tile -> up to 8 tiles (32 slices) in 64 KB
  run<64>(1, 1); run<64>(2, 2); run<64>(4, 4); run<64>(8, 8); run<64>(16, 16); run<64>(32, 32);
  run<64>(32, 1); run<64>(1, 32); run<64>(32, 4); run<64>(4, 32);
  run<32>(1, 1); run<32>(32, 32); run<16>(1, 1); run<16>(32, 32);
  run<128>(1, 1); run<128>(32, 16); run<256>(1, 1); run<256>(32, 8);
  return 0;
}
