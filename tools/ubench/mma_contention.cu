// Microbenchmark (B200): does background activity of other warps slow tcgen05.mma execution?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../modaltune_b200/csrc/sm100_ptx.cuh"
using namespace mt::sm100;

template <int BG>
__global__ void __launch_bounds__(32 * 17, 1) k(int reps, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); stop = 0; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 16) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  if (warp == 16) {
    constexpr uint32_t id48 = umma_idesc_bf16(128, 48, 0, 1);
    constexpr uint32_t id64 = umma_idesc_bf16(128, 64, 0, 0);
    const uint64_t v0 = umma_smem_desc(sbase + 65536, 16384, 1024);
    const uint64_t q0 = umma_smem_desc(sbase + 98304, 16, 1024);
    long long t0 = clock64(), t1 = 0;
    if (BG == 5) {   // dQ-type: SS, A MN-major [key rows][128 queries] smem, B MN-major K tile, N = 48
      constexpr uint32_t iddq = umma_idesc_bf16(128, 48, 1, 1);
      const uint64_t a0 = umma_smem_desc(sbase, 16384, 1024), b0 = umma_smem_desc(sbase + 65536, 16384, 1024);
      if (elect_one()) {
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss(tm + 384, umma_desc_adv(a0, kk * 2048), umma_desc_adv(b0, kk * 2048), iddq, 1);
        }
        t1 = clock64();
        umma_commit(smem_u32(&bar));
      }
    } else if (BG == 6) {   // forward S-type: SS K-major N=128, three k-steps
      constexpr uint32_t ids = umma_idesc_bf16(128, 128, 0, 0);
      const uint64_t a0 = umma_smem_desc(sbase, 16, 1024), b0 = umma_smem_desc(sbase + 65536, 16, 1024);
      if (elect_one()) {
        for (int r = 0; r < reps; r += 3) {
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) umma_ss(tm, umma_desc_adv(a0, kk * 32), umma_desc_adv(b0, kk * 32), ids, 1);
        }
        t1 = clock64();
        umma_commit(smem_u32(&bar));
      }
    } else if (elect_one()) {
      for (int r = 0; r < reps; r += 14) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ts(tm + 256 + (kk >> 2) * 64, tm + (kk & 3) * 16, umma_desc_adv(v0, kk * 2048), id48, 1);
#pragma unroll
        for (int kk = 0; kk < 6; ++kk) umma_ts(tm + 128 * (kk / 3 & 1), tm + 448 + (kk % 3) * 8, umma_desc_adv(q0, (kk % 3) * 32), id64, 1);
      }
      t1 = clock64();
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    stop = 1;
    if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else {
    // background warps
    float acc = (float)threadIdx.x * 1e-3f;
    uint32_t* sm32 = reinterpret_cast<uint32_t*>(smem);
    const uint32_t t_lane = (uint32_t)((warp & 3) * 32) << 16;
    while (!stop) {
      if (BG == 1) {        // MUFU + FMA
#pragma unroll
        for (int i = 0; i < 32; ++i) acc = ex2(acc * 0.5f - 1.0f) + 0.25f;
      } else if (BG == 2) { // shared-memory stores + loads (conflict-free 16-byte accesses)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 v = *reinterpret_cast<uint4*>(smem + ((warp * 2048 + i * 512 + lane * 16) & 0xffff));
          v.x += 1;
          *reinterpret_cast<uint4*>(smem + ((warp * 2048 + i * 512 + lane * 16) & 0xffff)) = v;
          acc += __uint_as_float(v.y);
        }
      } else if (BG == 3) { // TMEM loads
        float v[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) { tmem_ld16(tm + t_lane + 320 + (i & 1) * 16, v); tmem_ld_wait(); acc += v[3]; }
      } else if (BG == 4) { // mbarrier polling (try_wait on a barrier that never completes in this phase)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc += mbar_try_wait(smem_u32(&bar), 1) ? 1.f : 0.f;
      } else {
        __nanosleep(200);
      }
    }
    if (acc == 12345.678f) sink[threadIdx.x] = acc + sm32[lane];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tm, 512);
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  float* sink; cudaMalloc(&sink, 4096);
  const int reps = 14 * 256;
  auto run = [&](auto kern, const char* name) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    kern<<<148, 32 * 17, 131072>>>(reps, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    printf("%-44s issue %.1f clk/MMA, complete %.1f clk/MMA\n", name, (double)out[0] / reps, (double)out[1] / reps);
  };
  run(k<0>, "background: idle (nanosleep)");
  run(k<1>, "background: 16 warps of MUFU/FMA");
  run(k<2>, "background: 16 warps of smem ld/st");
  run(k<3>, "background: 16 warps of tcgen05.ld");
  run(k<4>, "background: 16 warps polling an mbarrier");
  run(k<5>, "dQ-type SS (A, B MN-major, N=48), idle background");
  run(k<6>, "S-type SS (K-major, N=128), idle background");
  return 0;
}
