// Microbenchmark (B200): cost of ISSUING tcgen05.mma from one thread, by code shape.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../modaltune_b200/csrc/sm100_ptx.cuh"
using namespace mt::sm100;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    const uint64_t a0 = umma_smem_desc(sbase, 16, 1024), b0 = umma_smem_desc(sbase + 32768, 16, 1024);
    long long t0 = clock64(), t1 = 0;
    if (MODE == 0) {          // lane-0 branch, descriptors rebuilt per MMA (what the kernels did)
      if (lane == 0) {
        for (int r = 0; r < reps; r += 3) {
#pragma unroll
          for (int kk = 0; kk < 3; ++kk)
            umma_ss(tm, umma_smem_desc(sbase + kk * 32, 16, 1024), umma_smem_desc(sbase + 32768 + kk * 32, 16, 1024), idesc, 1);
        }
        t1 = clock64();
        umma_commit(smem_u32(&bar));
      }
    } else if (MODE == 1) {   // lane-0 branch, precomputed descriptors + constant advance
      if (lane == 0) {
        for (int r = 0; r < reps; r += 3) {
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) umma_ss(tm, a0 + 2 * kk, b0 + 2 * kk, idesc, 1);
        }
        t1 = clock64();
        umma_commit(smem_u32(&bar));
      }
    } else if (MODE == 2) {   // elect.sync predicate, precomputed descriptors
      for (int r = 0; r < reps; r += 3) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) umma_ss(tm, a0 + 2 * kk, b0 + 2 * kk, idesc, 1);
        }
        __syncwarp();
      }
      t1 = clock64();
      if (elect_one()) umma_commit(smem_u32(&bar));
    } else if (MODE == 4) {   // TS N=48 (24-clk MMAs), precomputed descriptor, 8 unrolled k-steps per group
      constexpr uint32_t id48 = umma_idesc_bf16(128, 48, 0, 1);
      const uint64_t v0 = umma_smem_desc(sbase + 32768, 16384, 1024);
      if (lane == 0) {
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ts(tm + 128, tm + 256 + kk * 8, v0 + 128 * kk, id48, 1);
        }
        t1 = clock64();
        umma_commit(smem_u32(&bar));
      }
    } else if (MODE == 5) {   // same, but the stage (descriptor base + TMEM column) is chosen at run time per group
      constexpr uint32_t id48 = umma_idesc_bf16(128, 48, 0, 1);
      if (lane == 0) {
        for (int r = 0; r < reps; r += 8) {
          const int st = (r >> 3) & 1;
          const uint32_t vb = sbase + 32768 + st * 16384;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_ts(tm + 128 + st * 64, tm + 256 + kk * 8, umma_smem_desc(vb + kk * 2048, 16384, 1024), id48, (r > 0) || (kk > 0));
        }
        t1 = clock64();
        umma_commit(smem_u32(&bar));
      }
    } else if (MODE == 6) {   // run-time stage, descriptor bases precomputed for both stages
      constexpr uint32_t id48 = umma_idesc_bf16(128, 48, 0, 1);
      const uint64_t vd[2] = {umma_smem_desc(sbase + 32768, 16384, 1024), umma_smem_desc(sbase + 49152, 16384, 1024)};
      if (lane == 0) {
        for (int r = 0; r < reps; r += 8) {
          const int st = (r >> 3) & 1;
          const uint64_t v0 = st ? vd[1] : vd[0];
          const uint32_t dcol = tm + 128 + st * 64;
          umma_ts(dcol, tm + 256, v0, id48, r > 0);
#pragma unroll
          for (int kk = 1; kk < 8; ++kk) umma_ts(dcol, tm + 256 + kk * 8, v0 + 128 * kk, id48, 1);
        }
        t1 = clock64();
        umma_commit(smem_u32(&bar));
      }
    } else {                  // elect.sync once around the whole loop
      if (elect_one()) {
        for (int r = 0; r < reps; r += 3) {
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) umma_ss(tm, a0 + 2 * kk, b0 + 2 * kk, idesc, 1);
        }
        t1 = clock64();
        umma_commit(smem_u32(&bar));
      }
    }
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  const int reps = 3072;
  auto run = [&](auto kern, const char* name) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    kern<<<148, 128, 65536>>>(reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    printf("%-60s issue %.1f clk/MMA, complete %.1f clk/MMA\n", name, (double)out[0] / reps, (double)out[1] / reps);
  };
  run(k<0>, "lane==0, descriptors rebuilt per MMA");
  run(k<1>, "lane==0, precomputed descriptors");
  run(k<2>, "elect.sync per group of 3, precomputed descriptors");
  run(k<3>, "elect.sync around the loop, precomputed descriptors");
  run(k<4>, "TS N=48: precomputed descriptor, constant k advance");
  run(k<5>, "TS N=48: run-time stage, descriptors rebuilt, run-time accumulate flag");
  run(k<6>, "TS N=48: run-time stage, precomputed bases, constant flags");
  return 0;
}
