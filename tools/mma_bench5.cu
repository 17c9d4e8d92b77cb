// Microbenchmark 5: what does a tcgen05.commit cost the MMA stream?  N = 64 MMAs (48 cycles each back to back, mma_bench2)
// with a commit to an mbarrier every `ci` MMAs (nobody waits on it), and optionally a try_wait on an already-complete
// barrier + tcgen05.fence::after_thread_sync every `wi` MMAs (what the fused kernels' issue loops do per sub-tile / stage).
#include <cstdio>
#include <cstdlib>
#include "../sa-toolkit_b200/csrc/conv_tc.cuh"
using namespace sa::tc;

__global__ void __launch_bounds__(128, 1) mma_commit(int iters, int ci, int wi, int same_acc, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, dummy[4], done;
  __shared__ uint32_t holder;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&done), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&dummy[i]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (threadIdx.x == 32) mbar_arrive(smem_u32(&done));                      // phase 0 of `done` is complete: try_wait(done, 0) succeeds at once
  __syncthreads();
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(64, false);
    constexpr uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a0 = desc_lo(smem_u32(smem) + 8 * 1024);
    const uint32_t b0 = desc_lo(smem_u32(smem) + 96 * 1024);
    int nc = 0, nw = 0, k = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (wi > 0 && (nw += 4) >= wi) {
        nw = 0;
        mbar_wait(smem_u32(&done), 0, nullptr);
        tc_fence_after();
      }
      const uint32_t acc = same_acc ? 0u : (uint32_t)((i >> 2) & 3) * 64u;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (leader) umma_f16(tmem + acc, desc64(a0 + (uint32_t)g * 2u, hi), desc64(b0 + (uint32_t)g * 2u, hi), idesc, i > 0 ? 1u : 0u);
      if (ci > 0 && (nc += 4) >= ci) {
        nc = 0;
        if (leader) umma_commit(smem_u32(&dummy[k]));
        k = (k + 1) & 3;
      }
      __syncwarp();
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

void run(int ci, int wi, int same_acc) {
  const int iters = 8192, ctas = 148;
  long long* d;
  cudaMalloc(&d, ctas * sizeof(long long));
  cudaFuncSetAttribute(mma_commit, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) mma_commit<<<ctas, 128, 180 * 1024>>>(iters, ci, wi, same_acc, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=64 commit every %3d MMAs, try_wait+fence every %3d MMAs, %s accumulator : %6.1f cycles/MMA  %s\n", ci, wi,
         same_acc ? "one " : "four", (double)mx / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run(0, 0, 1); run(0, 0, 0);
  for (int ci : {4, 8, 12, 16, 28, 44}) { run(ci, 0, 1); run(ci, 0, 0); }
  for (int wi : {4, 12, 28}) { run(0, wi, 1); run(0, wi, 0); }
  run(12, 12, 1); run(12, 12, 0); run(4, 12, 0);
  return 0;
}
