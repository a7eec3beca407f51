// Microbenchmark (B200): throughput of the three ways a [128 rows][48 floats] fp32 gradient tile can be reduce-added into
// a row-major [n_rows][2304] fp32 matrix (row stride 9216 B, 192 contiguous bytes per row), all 148 SMs busy:
//   0: red.global.add.v4.f32 from registers, one row per thread (32 lines per warp instruction)
//   1: red.global.add.v4.f32 coalesced (12 lanes cover the 192 bytes of a row)
//   2: cp.reduce.async.bulk.tensor (TMA reduce-add) of a [128][32] + [128][16] staging tile in shared memory
// Prints cycles per tile per SM and the aggregate reduce bandwidth.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../modaltune_b200/csrc/sm100_ptx.cuh"
using namespace mt::sm100;

constexpr int LD = 2304, ROWS = 10240;

__global__ void __launch_bounds__(128, 1) k(int mode, int tiles, float* dst, const __grid_constant__ CUtensorMap m32,
                                             const __grid_constant__ CUtensorMap m16, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < 24576 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.f;
  fence_proxy_async_smem();
  __syncthreads();
  const int row = threadIdx.x;
  long long t0 = clock64();
  for (int t = 0; t < tiles; ++t) {
    // tile t of this CTA: 128 consecutive rows, head column block (pseudo-random walk over the matrix)
    const unsigned idx = (blockIdx.x * 7919u + t * 104729u);
    const int r0 = (idx % (ROWS / 128)) * 128, c0 = ((idx / 97) % 48) * 48;
    if (mode == 0) {
      float* p = dst + (size_t)(r0 + row) * LD + c0;
#pragma unroll
      for (int c = 0; c < 12; ++c) red_add_v4(p + c * 4, 1.f, 1.f, 1.f, 1.f);
    } else if (mode == 1) {
      // 12 lanes per row: thread -> (row in group, chunk); 128 threads cover 10 rows (120 lanes) per pass... use 8 lanes x 16 B
      // = 128 B per row-part: pass A covers columns 0..31 (8 chunks), pass B columns 32..47 (4 chunks)
#pragma unroll
      for (int pass = 0; pass < 8; ++pass) {   // 128 rows x 8 chunks / 128 threads
        const int rr = pass * 16 + (threadIdx.x >> 3), ch = threadIdx.x & 7;
        red_add_v4(dst + (size_t)(r0 + rr) * LD + c0 + ch * 4, 1.f, 1.f, 1.f, 1.f);
      }
#pragma unroll
      for (int pass = 0; pass < 4; ++pass) {   // 128 rows x 4 chunks / 128 threads
        const int rr = pass * 32 + (threadIdx.x >> 2), ch = threadIdx.x & 3;
        red_add_v4(dst + (size_t)(r0 + rr) * LD + c0 + 32 + ch * 4, 1.f, 1.f, 1.f, 1.f);
      }
    } else {
      if (threadIdx.x == 0) {
        tma_reduce_add_3d(&m32, sbase, c0, 0, r0);
        tma_reduce_add_3d(&m16, sbase + 16384, c0 + 32, 0, r0);
        bulk_commit_group();
        bulk_wait_group_read<0>();
      }
      __syncthreads();
    }
  }
  if (mode == 2 && threadIdx.x == 0) bulk_wait_group_all();
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  float* dst; cudaMalloc(&dst, (size_t)ROWS * LD * 4); cudaMemset(dst, 0, (size_t)ROWS * LD * 4);
  long long* out; cudaMallocManaged(&out, 16);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap m32, m16;
  for (int w = 0; w < 2; ++w) {
    cuuint64_t dims[3] = {LD, 1, ROWS};
    cuuint64_t strides[2] = {(cuuint64_t)LD * 4, (cuuint64_t)LD * 4};
    cuuint32_t box[3] = {w == 0 ? 32u : 16u, 1, 128};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult rc = enc(w == 0 ? &m32 : &m16, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dst, dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, w == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return 1; }
  }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  const char* names[] = {"red.v4 row-per-thread (uncoalesced)", "red.v4 coalesced                   ", "TMA reduce-add (2 boxes)           "};
  const int tiles = 256;
  for (int mode = 0; mode < 3; ++mode) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, 128, 32768>>>(mode, 8, dst, m32, m16, out);
    cudaEventRecord(e0);
    k<<<148, 128, 32768>>>(mode, tiles, dst, m32, m16, out);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%s : %.0f clk per tile per SM, %.3f ms for %d tiles x 148 SMs = %.2f TB/s of fp32 adds\n", names[mode],
           (double)out[0] / tiles, ms, tiles, 148.0 * tiles * 24576 / ms / 1e9);
  }
  return 0;
}
