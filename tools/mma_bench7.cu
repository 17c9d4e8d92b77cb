// Microbenchmark 7: the MMA issue structure of resblock_chain_kernel<64, 3, K, 8> (chain_tc.cuh) without anything else:
// blocks of K taps x 4 K-slices, MS sub-tiles per conv, two accumulator sets by conv parity, weights in ring slots, commits.
// In the real kernel (epilogue switched off) these MMAs run at ~89 cycles, in mma_bench2/3 at 48: which feature costs it?
// flags: 1 commits (per stage on the last sub-tile + per block), 2 accumulate = 0 on a block's first MMA, 4 accumulator column
// varies with (conv parity, sub-tile), 8 A start shifts by `dil` rows per tap and 128 rows per sub-tile, 16 B walks over the ring.
#include <cstdio>
#include <cstdlib>
#include "../sa-toolkit_b200/csrc/conv_tc.cuh"
using namespace sa::tc;

template <int K>
__global__ void __launch_bounds__(128, 1) chain_issue(int n_tiles, int n_convs, int dil, int flags, long long* out_cycles, long long* out_mmas) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar, dummy[8];
  __shared__ uint32_t holder;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int MS = 3, N = 64, K16 = 4, NSLOTS = 6;
  constexpr uint32_t RB = 128, row16 = RB >> 4, kTapBytes = N * RB, stage16 = (2 * kTapBytes) >> 4, tap16 = kTapBytes >> 4;
  for (int i = threadIdx.x; i < 214 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&dummy[i]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&holder), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = holder;
  if (warp == 3) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(N, false);
    constexpr uint32_t hi = ((8u * RB) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t bufA = desc_lo(smem_u32(smem)), bufT = desc_lo(smem_u32(smem) + 448 * 128), b_lo0 = desc_lo(smem_u32(smem) + 2 * 448 * 128);
    const uint32_t dil16 = (flags & 8) ? (uint32_t)dil * row16 : 0u;
    long long n_mma = 0;
    int slot0 = 0, k = 0;
    const long long t0 = clock64();
    for (int tile = 0; tile < n_tiles; ++tile)
      for (int c = 0; c < n_convs; ++c) {
        const uint32_t in_lo0 = ((c & 1) ? bufT : bufA) + (uint32_t)(32 - dil * (K / 2)) * row16;
        int slot_end = slot0;
#pragma unroll
        for (int s = 0; s < MS; ++s) {
          tc_fence_after();
          const uint32_t d_tmem = tmem + ((flags & 4) ? (uint32_t)(((c & 1) * MS + s) * N) : 0u);
          uint32_t a_tap = in_lo0 + ((flags & 8) ? (uint32_t)(s * 128) * row16 : 0u);
          int slot = slot0;
          uint32_t b_tap = b_lo0 + ((flags & 16) ? (uint32_t)slot * stage16 : 0u);
#pragma unroll
          for (int tap = 0; tap < K; ++tap) {
            const bool stage_end = ((tap + 1) % 2) == 0 || tap == K - 1;
#pragma unroll
            for (int kk = 0; kk < K16; ++kk)
              if (leader) umma_f16(d_tmem, desc64(a_tap + 2u * kk, hi), desc64(b_tap + 2u * kk, hi), idesc, ((flags & 2) ? (tap | kk) : 1) ? 1u : 0u);
            a_tap += dil16;
            if ((flags & 1) && s == MS - 1 && stage_end) { if (leader) umma_commit(smem_u32(&dummy[k & 7])); ++k; }
            if (stage_end) {
              if (++slot == NSLOTS) slot = 0;
              b_tap = b_lo0 + ((flags & 16) ? (uint32_t)slot * stage16 : 0u);
            } else if (flags & 16) {
              b_tap += tap16;
            }
          }
          slot_end = slot;
          n_mma += K * K16;
          if (flags & 1) { if (leader) umma_commit(smem_u32(&dummy[k & 7])); ++k; }
          __syncwarp();
        }
        slot0 = slot_end;
      }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, nullptr);
    const long long t1 = clock64();
    if (lane == 0) { out_cycles[blockIdx.x] = t1 - t0; out_mmas[blockIdx.x] = n_mma; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int K>
void run(int dil, int flags, int n_tiles = 8) {
  const int ctas = 148, n_convs = 6;
  long long *d, *m;
  cudaMalloc(&d, ctas * sizeof(long long)); cudaMalloc(&m, ctas * sizeof(long long));
  cudaFuncSetAttribute(chain_issue<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int rep = 0; rep < 2; ++rep) chain_issue<K><<<ctas, 128, 216 * 1024>>>(n_tiles, n_convs, dil, flags, d, m);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148], hm[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hm, m, sizeof(hm), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("K=%2d dil=%d tiles=%6d flags=%2d [%s%s%s%s%s]: %6.1f cycles/MMA  %s\n", K, dil, n_tiles, flags, (flags & 1) ? "commit " : "", (flags & 2) ? "acc0 " : "",
         (flags & 4) ? "dcol " : "", (flags & 8) ? "ashift " : "", (flags & 16) ? "bwalk" : "", (double)mx / (double)hm[0],
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(m);
}

int main() {
  for (int f : {0, 1, 2, 4, 8, 16, 31}) run<3>(1, f);
  run<3>(1, 31, 4000); run<3>(1, 31, 40000); run<3>(1, 31, 40000);   // 40000 tiles: ~0.3 s per launch, the power management reacts
  run<3>(3, 31); run<7>(1, 0); run<7>(1, 31); run<7>(5, 31); run<11>(1, 31);
  return 0;
}
