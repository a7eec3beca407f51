// Microbenchmark (B200): cycles per tcgen05.mma for the operand layouts the attention kernels use.
// One CTA per SM, one thread issues `reps` MMAs of one kind back to back, commits, waits; prints cycles per MMA.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../modaltune_b200/csrc/sm100_ptx.cuh"
using namespace mt::sm100;

__global__ void __launch_bounds__(128, 1) k(int kind, int reps, long long* out) {
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
  if (threadIdx.x == 0) {
    const uint32_t A = sbase, B = sbase + 32768;
    uint32_t idesc; 
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int kk = r & 7;
      switch (kind) {
        case 0:  // SS, A K-major, B K-major, N=128 (S = Q K^T)
          idesc = umma_idesc_bf16(128, 128, 0, 0);
          umma_ss(tm, umma_smem_desc(A + (kk % 3) * 32, 16, 1024), umma_smem_desc(B + (kk % 3) * 32, 16, 1024), idesc, 1); break;
        case 1:  // SS, A K-major (P from smem), B MN-major N=48 (old forward P V)
          idesc = umma_idesc_bf16(128, 48, 0, 1);
          umma_ss(tm + 128, umma_smem_desc(A + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024), umma_smem_desc(B + kk * 2048, 16384, 1024), idesc, 1); break;
        case 2:  // TS, A TMEM, B MN-major N=48 (forward P V, backward dV / dK)
          idesc = umma_idesc_bf16(128, 48, 0, 1);
          umma_ts(tm + 128, tm + 256 + kk * 8, umma_smem_desc(B + kk * 2048, 16384, 1024), idesc, 1); break;
        case 3:  // SS, A MN-major, B MN-major N=48 (backward v1 dV / dK, v2 dQ)
          idesc = umma_idesc_bf16(128, 48, 1, 1);
          umma_ss(tm + 128, umma_smem_desc(A + kk * 2048, 16384, 1024), umma_smem_desc(B + kk * 2048, 16384, 1024), idesc, 1); break;
        case 4:  // TS, A TMEM, B K-major N=64 (backward v2 S^T)
          idesc = umma_idesc_bf16(128, 64, 0, 0);
          umma_ts(tm, tm + 256 + (kk % 3) * 8, umma_smem_desc(B + (kk % 3) * 32, 16, 1024), idesc, 1); break;
        case 5:  // TS, A TMEM, B K-major N=128
          idesc = umma_idesc_bf16(128, 128, 0, 0);
          umma_ts(tm, tm + 256 + (kk % 3) * 8, umma_smem_desc(B + (kk % 3) * 32, 16, 1024), idesc, 1); break;
        case 6:  // SS, A K-major, B K-major N=64
          idesc = umma_idesc_bf16(128, 64, 0, 0);
          umma_ss(tm, umma_smem_desc(A + (kk % 3) * 32, 16, 1024), umma_smem_desc(B + (kk % 3) * 32, 16, 1024), idesc, 1); break;
        case 7:  // TS, A TMEM, B MN-major N=64
          idesc = umma_idesc_bf16(128, 64, 0, 1);
          umma_ts(tm + 128, tm + 256 + kk * 8, umma_smem_desc(B + kk * 2048, 16384, 1024), idesc, 1); break;
        case 8:  // kind 2 with FOUR independent accumulators (rotating D)
          idesc = umma_idesc_bf16(128, 48, 0, 1);
          umma_ts(tm + (r & 3) * 64, tm + 256 + kk * 8, umma_smem_desc(B + kk * 2048, 16384, 1024), idesc, 1); break;
        case 9:  // kind 0 with two independent accumulators
          idesc = umma_idesc_bf16(128, 128, 0, 0);
          umma_ss(tm + (r & 1) * 128, umma_smem_desc(A + (kk % 3) * 32, 16, 1024), umma_smem_desc(B + (kk % 3) * 32, 16, 1024), idesc, 1); break;
        case 10:  // SS N=256
          idesc = umma_idesc_bf16(128, 256, 0, 0);
          umma_ss(tm, umma_smem_desc(A + (kk % 3) * 32, 16, 1024), umma_smem_desc(B + (kk % 3) * 32, 16, 1024), idesc, 1); break;
        case 11:  // kind 0, overwrite (accumulate = 0)
          idesc = umma_idesc_bf16(128, 128, 0, 0);
          umma_ss(tm, umma_smem_desc(A + (kk % 3) * 32, 16, 1024), umma_smem_desc(B + (kk % 3) * 32, 16, 1024), idesc, 0); break;
      }
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* out; cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const char* names[] = {"SS  A=K-major B=K-major  N=128", "SS  A=K-major B=MN-major N=48 ", "TS  A=TMEM    B=MN-major N=48 ",
                         "SS  A=MN-maj  B=MN-major N=48 ", "TS  A=TMEM    B=K-major  N=64 ", "TS  A=TMEM    B=K-major  N=128",
                         "SS  A=K-major B=K-major  N=64 ", "TS  A=TMEM    B=MN-major N=64 ", "TS  N=48, 4 rotating D tiles   ",
                         "SS  N=128, 2 rotating D tiles  ", "SS  A=K-major B=K-major  N=256", "SS  N=128 accumulate=0         "};
  for (int grid : {148}) for (int kind = 0; kind < 12; ++kind) {
    const int reps = 2048;
    k<<<grid, 128, 65536>>>(kind, reps, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kind %d: %s\n", kind, cudaGetErrorString(e)); return 1; }
    printf("grid %3d  %s : issue %.1f clk/MMA, complete %.1f clk/MMA\n", grid, names[kind], (double)out[0] / reps, (double)out[1] / reps);
  }
  return 0;
}
