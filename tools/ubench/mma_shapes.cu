// Microbenchmark (B200): cycles per tcgen05.mma (M = 128, K = 16, bf16) by operand kind and N, descriptors prebuilt,
// 8 unrolled MMAs per loop trip issued from one elected thread, one CTA per SM on all SMs.
//   A: S = shared memory (K-major), M = shared memory MN-major, T = tensor memory
//   B: K = K-major, M = MN-major          D: same accumulator or NROT rotating accumulators
#include <cstdio>
#include <cuda_runtime.h>
#include "../../modaltune_b200/csrc/sm100_ptx.cuh"
using namespace mt::sm100;

template <int AKIND, int BMN, int N, int NROT>
__global__ void __launch_bounds__(128, 1) k(int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tptr;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, AKIND == 1 ? 1 : 0, BMN);
    // A: K-major: k-step = 32 B inside the swizzle atom; MN-major: k-step = 16 rows = 2048 B
    const uint64_t a0 = AKIND == 1 ? umma_smem_desc(sbase, 16384, 1024) : umma_smem_desc(sbase, 16, 1024);
    const uint64_t b0 = BMN ? umma_smem_desc(sbase + 32768, 16384, 1024) : umma_smem_desc(sbase + 32768, 16, 1024);
    const uint32_t astep = AKIND == 1 ? 128 : 2, bstep = BMN ? 128 : 2;   // descriptor units of 16 B
    long long t0 = clock64(), t1 = 0;
    if (elect_one()) {
      for (int r = 0; r < reps; r += 8) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t d = tm + (NROT > 1 ? ((kk % NROT) * (N <= 64 ? 64 : 128)) : 0);
          const int ks = AKIND == 1 || BMN ? kk : (kk & 3);
          if (AKIND == 2) umma_ts(d, tm + 384 + (kk & 3) * 8, b0 + bstep * (BMN ? kk : (kk & 3)), idesc, 1);
          else umma_ss(d, a0 + astep * (AKIND == 1 ? kk : (kk & 3)), b0 + bstep * (BMN ? kk : (kk & 3)), idesc, 1);
          (void)ks;
        }
      }
      t1 = clock64();
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0 && t1 != 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int AKIND, int BMN, int N, int NROT>
void run(const char* name, long long* out) {
  const int reps = 4096;
  cudaFuncSetAttribute(k<AKIND, BMN, N, NROT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  k<AKIND, BMN, N, NROT><<<148, 128, 96 * 1024>>>(reps, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
  printf("%-44s N=%3d rot=%d : issue %.1f clk/MMA, complete %.1f clk/MMA (math floor %d)\n", name, N, NROT,
         (double)out[0] / reps, (double)out[1] / reps, N / 2);
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  run<0, 0, 128, 1>("SS A K-major, B K-major (fwd S=QK^T)", out);
  run<0, 0, 64, 1>("SS A K-major, B K-major", out);
  run<0, 0, 256, 1>("SS A K-major, B K-major", out);
  run<2, 0, 64, 1>("TS A TMEM, B K-major (bwd S^T half)", out);
  run<2, 0, 128, 1>("TS A TMEM, B K-major (bwd S^T full)", out);
  run<2, 0, 256, 1>("TS A TMEM, B K-major", out);
  run<2, 1, 48, 1>("TS A TMEM, B MN-major (PV, dV, dK)", out);
  run<2, 1, 48, 2>("TS A TMEM, B MN-major (dV / dK alternating)", out);
  run<2, 1, 64, 1>("TS A TMEM, B MN-major", out);
  run<2, 1, 64, 2>("TS A TMEM, B MN-major", out);
  run<2, 1, 32, 1>("TS A TMEM, B MN-major", out);
  run<2, 1, 16, 1>("TS A TMEM, B MN-major", out);
  run<1, 1, 48, 1>("SS A MN-major, B MN-major (dQ = dS K)", out);
  run<1, 1, 64, 1>("SS A MN-major, B MN-major", out);
  run<0, 1, 48, 1>("SS A K-major, B MN-major", out);
  run<0, 1, 64, 1>("SS A K-major, B MN-major", out);
  return 0;
}
