// Microbenchmark 3: tcgen05.mma rate (M = 128, N = 64, K = 16, SWIZZLE_128B A operand) when the A descriptor starts in the
// middle of an 8-row swizzle atom (tap = row shift, conv_tc / chain_tc) or in the middle of a 128-byte row (slice = 32-byte
// shift, chain_group_tc), and with concurrent generic-proxy shared-memory stores from other warps (epilogue traffic).
#include <cstdio>
#include <cstdlib>
#include "../sa-toolkit_b200/csrc/conv_tc.cuh"
using namespace sa::tc;

__global__ void __launch_bounds__(640, 1) mma_shift(int iters, int shift_rows, int shift_bytes, int vary, int sts_warps, long long* out_cycles, int random_data) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t holder;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0x3c003c00u;
    if (random_data) {                                            // fp16 values in (-2, 2) with random mantissas: realistic switching activity
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      v = (h & 0x83ff83ffu) | 0x38003800u | ((h >> 4) & 0x04000400u);
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(64, false);
    constexpr uint32_t hiA = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);     // SWIZZLE_128B
    constexpr uint32_t hiB = ((8u * 32u) >> 4) | (1u << 14) | (6u << 29);      // SWIZZLE_32B (grouped kernel's weight blocks)
    const uint32_t a0 = desc_lo(smem_u32(smem) + 8 * 1024 + shift_rows * 128 + shift_bytes);
    const uint32_t b0 = desc_lo(smem_u32(smem) + 96 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        // vary: walk the shift like a real tap / slice loop (g-th tap = g rows or g * 32 bytes further)
        const uint32_t a = a0 + (vary == 1 ? (uint32_t)g * 8u : vary == 2 ? (uint32_t)g * 2u : 0u);
        if (leader) umma_f16(tmem + (uint32_t)((i / 4 & 1) * 64), desc64(a, hiA), desc64(b0 + (uint32_t)g * 128u, hiB), idesc, i > 0 ? 1u : 0u);
      }
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr);
    const long long t1 = clock64();
    if (lane == 0) out_cycles[blockIdx.x] = t1 - t0;
    stop = 1;
  } else if (warp - 1 < sts_warps) {
    // epilogue-like traffic: every thread stores 128 bytes (8 x 16 B, swizzled rows) per round into a scratch tile
    uint8_t* t = smem + 112 * 1024 + (uint32_t)((warp - 1) & 7) * 4096 + lane * 128;
    uint4 v = make_uint4(warp, lane, 0, 0);
    while (!stop) {
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(t + ((c * 16) ^ ((lane & 7) << 4))) = v;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      v.z++;
      __nanosleep(20);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

void run(int shift_rows, int shift_bytes, int vary, int sts_warps, int random_data = 0, int iters = 4096) {
  const int ctas = 148;
  long long* d;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaFuncSetAttribute(mma_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) mma_shift<<<ctas, 640, 180 * 1024>>>(iters, shift_rows, shift_bytes, vary, sts_warps, d, random_data);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=64 shift_rows=%d shift_bytes=%3d vary=%d sts_warps=%2d random=%d iters=%7d : %6.1f cycles/MMA  %s\n", shift_rows, shift_bytes, vary, sts_warps, random_data, iters,
         (double)mx / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run(0, 0, 0, 0, 0, 4096);
  run(0, 0, 0, 0, 1, 4096);
  run(0, 0, 0, 0, 0, 1 << 20);            // ~30 ms: long enough for the power management to react
  run(0, 0, 0, 0, 1, 1 << 20);
  run(0, 0, 1, 0, 1, 1 << 20);
  run(0, 0, 1, 8, 1, 1 << 20);
  for (int r = 1; r <= 2; ++r) run(r, 0, 0, 0);
  for (int b = 32; b <= 96; b += 32) run(0, b, 0, 0);
  run(3, 64, 0, 0);
  run(0, 0, 1, 0);            // taps: rows 0..3
  run(0, 0, 2, 0);            // slices: bytes 0, 32, 64, 96
  for (int w = 4; w <= 16; w += 4) { run(0, 0, 0, w); run(1, 0, 0, w); }
  return 0;
}
