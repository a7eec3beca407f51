// Shared helpers for the modaltune_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "../../include/modaltune_b200.h"

namespace mt {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define MT_REQUIRE(cond, ...)         \
  do {                                \
    if (!(cond)) {                    \
      mt::set_error(__VA_ARGS__);     \
      return MT_E_BADARG;             \
    }                                 \
  } while (0)

#define MT_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t e__ = (call);                                                \
    if (e__ != cudaSuccess) {                                                \
      mt::set_error("%s failed: %s", #call, cudaGetErrorString(e__));        \
      return (int)e__;                                                       \
    }                                                                        \
  } while (0)

static constexpr int kNumSMs = 148;

// ---- dtype-generic 8-element chunk loads/stores (16 B for bf16, 2 x 16 B for float) ---------------------------------
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float to_float(float x) { return x; }
__device__ __forceinline__ float to_float(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_float(float x);
template <> __device__ __forceinline__ float from_float<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- dilated-attention geometry shared by the SIMT and tcgen05 kernels ----------------------------------------------
struct BranchGeom {
  int r;         // dilation
  int g;         // segment length (clamped to N)
  int m;         // sparse slots per (segment, head) = ceil(g / r)
  int n_seg;     // ceil(N / g)
  int hpb;       // heads that own one position = H / r (>= 1)
  int pow2;      // 1 when r and hpb are powers of two (always with 16 heads): ownership is shifts and masks
  int log2_hpb;
  float inv_g;   // 1 / g: p % g by a float reciprocal and one correction step (positions < 2^24)
  int64_t o_off;    // element offset of this branch in o_br   (layout [N][hpb*D])
  int64_t lse_off;  // element offset of this branch in lse_br (layout [N][hpb])
};

struct DilatedGeom {
  int N, H, D, nb;
  BranchGeom b[MT_MAX_BRANCHES];
};

// fills `out`; returns 0 or MT_E_*
int make_dilated_geom(const mt_dilated_geometry* g, DilatedGeom* out);

// does head h own position p in branch b?  (p % g) % r == floor(h*r/H); also gives the compact head slot
// The merge kernels evaluate this for every (position, head, branch): with run-time divisors the three integer
// divisions were most of their instruction stream, so the common case avoids them.
__device__ __forceinline__ bool branch_owns(const BranchGeom& bg, int H, int p, int h, int* slot) {
  if (bg.pow2) {
    const int q = __float2int_rz(__int2float_rn(p) * bg.inv_g);
    int local = p - q * bg.g;                       // q is floor(p / g) or one off: one correction step is exact
    if (local < 0) local += bg.g;
    else if (local >= bg.g) local -= bg.g;
    const int o = h >> bg.log2_hpb;                 // (h * r) / H = h / (H / r)
    *slot = h & (bg.hpb - 1);
    return (local & (bg.r - 1)) == o;
  }
  int o = (h * bg.r) / H;
  int local = p % bg.g;
  *slot = h - o * bg.hpb;
  return (local % bg.r) == o;
}

}  // namespace mt
