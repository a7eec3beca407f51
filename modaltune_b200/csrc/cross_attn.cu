// Injector / Extractor cross-attention core: softmax(q k^T / sqrt(16)) v with 12 heads x 16 in the 192-d compressed
// space (models/vitadapter/adapter_modules.py:225-229 -> nn.MultiheadAttention).  The reference materialises the
// [12, Lq, Lk] probabilities (and their head average, which it discards); here nothing of that size is ever written.
//
//   Injector  (Lq = tiles ~1e4, Lk = modal tokens ~66): one thread per query, K/V tiles broadcast from shared memory.
//   Extractor (Lq ~66, Lk = tiles ~1e4): split-K over the keys (grid.z), partial (m, l, acc) per split, LSE combine.
// Backward is two kernels, each accumulating in registers along its own loop (plain stores when the loop is not split,
// 16-byte fp32 reduce-adds when it is):
//   dq-kernel (thread per query, loop over keys) and dkv-kernel (thread per key, loop over queries).
#include "mt_common.cuh"

namespace mt {

static constexpr int HD = 16;
static constexpr int XT = 128;  // threads per CTA = rows (queries or keys) per CTA
static constexpr int XC = 64;   // staged rows per chunk

// Two fp32 FMAs in one instruction (FFMA2, new with sm_100): the kernels below are bound by instruction issue (one
// shared-memory operand per four FMAs plus the exponentials), so halving the FMA instruction count is a direct win.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b);
  uint64_t rc = *reinterpret_cast<uint64_t*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
// dot product of a register row (8 float2) with a 16-float shared-memory row, two accumulator chains
__device__ __forceinline__ float dot16(const float2 (&a)[HD / 2], const float* __restrict__ row) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
  float2 d0 = make_float2(0.f, 0.f), d1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    const float4 k = r4[c];
    d0 = ffma2(a[2 * c], make_float2(k.x, k.y), d0);
    d1 = ffma2(a[2 * c + 1], make_float2(k.z, k.w), d1);
  }
  return (d0.x + d1.x) + (d0.y + d1.y);
}
// register-register dot product in EXACTLY the order of dot16: delta = dO.o must equal dO.v bit for bit when a query
// has a single key (p = 1, o = v), so that dS is exactly zero there
__device__ __forceinline__ float dot16v(const float (&a)[HD], const float (&b)[HD]) {
  float2 d0 = make_float2(0.f, 0.f), d1 = make_float2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    d0 = ffma2(make_float2(a[4 * c], a[4 * c + 1]), make_float2(b[4 * c], b[4 * c + 1]), d0);
    d1 = ffma2(make_float2(a[4 * c + 2], a[4 * c + 3]), make_float2(b[4 * c + 2], b[4 * c + 3]), d1);
  }
  return (d0.x + d1.x) + (d0.y + d1.y);
}
// acc += w * row
__device__ __forceinline__ void axpy16(float2 (&acc)[HD / 2], float w, const float* __restrict__ row) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
  const float2 w2 = make_float2(w, w);
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    const float4 k = r4[c];
    acc[2 * c] = ffma2(w2, make_float2(k.x, k.y), acc[2 * c]);
    acc[2 * c + 1] = ffma2(w2, make_float2(k.z, k.w), acc[2 * c + 1]);
  }
}
// acc += w * r (register row)
__device__ __forceinline__ void axpy16r(float2 (&acc)[HD / 2], float w, const float2 (&r)[HD / 2]) {
  const float2 w2 = make_float2(w, w);
#pragma unroll
  for (int c = 0; c < HD / 2; ++c) acc[c] = ffma2(w2, r[c], acc[c]);
}
__device__ __forceinline__ void to_pairs(const float (&v)[HD], float scale, float2 (&o)[HD / 2]) {
#pragma unroll
  for (int c = 0; c < HD / 2; ++c) o[c] = make_float2(v[2 * c] * scale, v[2 * c + 1] * scale);
}

template <typename T>
__device__ __forceinline__ void load16(const T* p, float (&v)[HD]) {
  float a[8], b[8];
  load8(p, a);
  load8(p + 8, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = a[j];
    v[8 + j] = b[j];
  }
}

// stage rows [r0, r0+XC) of a [L, heads*HD] matrix (head h) into smem [XC][HD] (fp32); rows >= lim zero-filled
template <typename T>
__device__ __forceinline__ void stage16(float* dst, const T* __restrict__ src, int64_t ld, int h, int64_t r0,
                                        int64_t lim) {
  // XC rows x 2 chunks of 8 = 128 chunk loads = one per thread
  const int row = threadIdx.x >> 1, ch = threadIdx.x & 1;
  float v[8];
  if (r0 + row < lim) {
    load8(src + (r0 + row) * ld + h * HD + ch * 8, v);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) dst[row * HD + ch * 8 + j] = v[j];
}

template <typename T>
__global__ void __launch_bounds__(XT) cross_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                       const T* __restrict__ v, T* __restrict__ o,
                                                       float* __restrict__ lse, float* __restrict__ ws, int64_t lq,
                                                       int64_t lk, int heads, int64_t keys_per_split) {
  __shared__ float Ks[XC * HD], Vs[XC * HD];
  const int h = blockIdx.y, split = blockIdx.z, nsplit = gridDim.z;
  const int64_t qi = (int64_t)blockIdx.x * XT + threadIdx.x;
  const int64_t ld = (int64_t)heads * HD;
  const bool live = qi < lq;
  float qv[HD];
  if (live) load16(q + qi * ld + h * HD, qv);
  else
#pragma unroll
    for (int j = 0; j < HD; ++j) qv[j] = 0.f;
  float2 q2[HD / 2], acc2[HD / 2];
  to_pairs(qv, 0.25f, q2);  // 1/sqrt(16)
  float m = -INFINITY, l = 0.f;
#pragma unroll
  for (int j = 0; j < HD / 2; ++j) acc2[j] = make_float2(0.f, 0.f);
  const int64_t kbeg = split * keys_per_split, kend = min(lk, kbeg + keys_per_split);
  for (int64_t k0 = kbeg; k0 < kend; k0 += XC) {
    __syncthreads();
    stage16(Ks, k, ld, h, k0, kend);
    stage16(Vs, v, ld, h, k0, kend);
    __syncthreads();
    const int cnt = (int)min((int64_t)XC, kend - k0);
    for (int j0 = 0; j0 < cnt; j0 += 8) {
      float s[8], mx = m;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float d = dot16(q2, Ks + (j0 + jj) * HD);
        s[jj] = (j0 + jj < cnt) ? d : -INFINITY;
        mx = fmaxf(mx, s[jj]);
      }
      const float alpha = expf(m - mx);
      l *= alpha;
#pragma unroll
      for (int e = 0; e < HD / 2; ++e) acc2[e] = make_float2(acc2[e].x * alpha, acc2[e].y * alpha);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const float p = expf(s[jj] - mx);
        l += p;
        axpy16(acc2, p, Vs + (j0 + jj) * HD);
      }
      m = mx;
    }
  }
  if (!live) return;
  float acc[HD];
#pragma unroll
  for (int e = 0; e < HD / 2; ++e) {
    acc[2 * e] = acc2[e].x;
    acc[2 * e + 1] = acc2[e].y;
  }
  if (nsplit == 1) {
    const float inv = 1.f / l;
    float out[HD];
#pragma unroll
    for (int e = 0; e < HD; ++e) out[e] = acc[e] * inv;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = out[j];
      b[j] = out[8 + j];
    }
    store8(o + qi * ld + h * HD, a);
    store8(o + qi * ld + h * HD + 8, b);
    lse[qi * heads + h] = m + logf(l);
  } else {
    float* w = ws + (((int64_t)split * lq + qi) * heads + h) * (HD + 2);
    w[0] = m;
    w[1] = l;
#pragma unroll
    for (int e = 0; e < HD; ++e) w[2 + e] = acc[e];
  }
}

template <typename T>
__global__ void __launch_bounds__(128) cross_combine_kernel(const float* __restrict__ ws, T* __restrict__ o,
                                                            float* __restrict__ lse, int64_t lq, int heads,
                                                            int nsplit) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= lq * heads) return;
  float m = -INFINITY;
  for (int s = 0; s < nsplit; ++s) m = fmaxf(m, ws[((int64_t)s * lq * heads + idx) * (HD + 2)]);
  float l = 0.f, acc[HD];
#pragma unroll
  for (int e = 0; e < HD; ++e) acc[e] = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float* w = ws + ((int64_t)s * lq * heads + idx) * (HD + 2);
    const float f = (w[0] == -INFINITY) ? 0.f : expf(w[0] - m);
    l = fmaf(f, w[1], l);
#pragma unroll
    for (int e = 0; e < HD; ++e) acc[e] = fmaf(f, w[2 + e], acc[e]);
  }
  const float inv = 1.f / l;
  const int64_t qi = idx / heads;
  const int h = (int)(idx % heads);
#pragma unroll
  for (int e = 0; e < HD; ++e) o[qi * heads * HD + h * HD + e] = from_float<T>(acc[e] * inv);
  lse[idx] = m + logf(l);
}

// dq[qi] += sum_k exp(s - lse) * (dO.v_k - delta) * k_k / 4
template <typename T>
__global__ void __launch_bounds__(XT) cross_bwd_dq_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                          const T* __restrict__ v, const T* __restrict__ o,
                                                          const T* __restrict__ d_o, const float* __restrict__ lse,
                                                          float* __restrict__ dq, int64_t lq, int64_t lk, int heads,
                                                          int64_t keys_per_split) {
  __shared__ float Ks[XC * HD], Vs[XC * HD];
  const int h = blockIdx.y, split = blockIdx.z;
  const int64_t qi = (int64_t)blockIdx.x * XT + threadIdx.x;
  const int64_t ld = (int64_t)heads * HD;
  const bool live = qi < lq;
  float qv[HD], gv[HD], acc[HD], delta = 0.f, L = INFINITY;
#pragma unroll
  for (int j = 0; j < HD; ++j) qv[j] = gv[j] = acc[j] = 0.f;
  if (live) {
    float ov[HD];
    load16(q + qi * ld + h * HD, qv);
    load16(d_o + qi * ld + h * HD, gv);
    load16(o + qi * ld + h * HD, ov);
    delta = dot16v(gv, ov);
    L = lse[qi * heads + h];
  }
  float2 q2[HD / 2], g2[HD / 2], acc2[HD / 2];
  to_pairs(qv, 0.25f, q2);
  to_pairs(gv, 1.f, g2);
#pragma unroll
  for (int j = 0; j < HD / 2; ++j) acc2[j] = make_float2(0.f, 0.f);
  const int64_t kbeg = split * keys_per_split, kend = min(lk, kbeg + keys_per_split);
  for (int64_t k0 = kbeg; k0 < kend; k0 += XC) {
    __syncthreads();
    stage16(Ks, k, ld, h, k0, kend);
    stage16(Vs, v, ld, h, k0, kend);
    __syncthreads();
    const int cnt = (int)min((int64_t)XC, kend - k0);
#pragma unroll 2
    for (int j = 0; j < cnt; ++j) {
      const float s = dot16(q2, Ks + j * HD), dp = dot16(g2, Vs + j * HD);
      const float ds = expf(s - L) * (dp - delta) * 0.25f;
      axpy16(acc2, ds, Ks + j * HD);
    }
  }
#pragma unroll
  for (int e = 0; e < HD / 2; ++e) {
    acc[2 * e] = acc2[e].x;
    acc[2 * e + 1] = acc2[e].y;
  }
  if (live) {
    float* dst = dq + qi * ld + h * HD;
    if (gridDim.z == 1) {   // the only contribution to this row: plain 16-byte stores, no zero fill needed
#pragma unroll
      for (int e = 0; e < HD; e += 4)
        *reinterpret_cast<float4*>(dst + e) = make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]);
    } else {
#pragma unroll
      for (int e = 0; e < HD; e += 4)
        atomicAdd(reinterpret_cast<float4*>(dst + e), make_float4(acc[e], acc[e + 1], acc[e + 2], acc[e + 3]));
    }
  }
}

// thread per key: dv[k] += sum_q p dO_q ; dk[k] += sum_q p (dO_q.v_k - delta_q) q_q / 4
template <typename T>
__global__ void __launch_bounds__(XT) cross_bwd_dkv_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                           const T* __restrict__ v, const T* __restrict__ o,
                                                           const T* __restrict__ d_o, const float* __restrict__ lse,
                                                           float* __restrict__ dk, float* __restrict__ dv, int64_t lq,
                                                           int64_t lk, int heads, int64_t q_per_split) {
  __shared__ float Qs[XC * HD], Gs[XC * HD], Ls[XC], Ds[XC];
  const int h = blockIdx.y, split = blockIdx.z;
  const int64_t ki = (int64_t)blockIdx.x * XT + threadIdx.x;
  const int64_t ld = (int64_t)heads * HD;
  const bool live = ki < lk;
  float kv[HD], vv[HD], dka[HD], dva[HD];
#pragma unroll
  for (int j = 0; j < HD; ++j) kv[j] = vv[j] = dka[j] = dva[j] = 0.f;
  if (live) {
    load16(k + ki * ld + h * HD, kv);
    load16(v + ki * ld + h * HD, vv);
  }
  float2 k2[HD / 2], v2[HD / 2], dk2[HD / 2], dv2[HD / 2];
  to_pairs(kv, 1.f, k2);
  to_pairs(vv, 1.f, v2);
#pragma unroll
  for (int j = 0; j < HD / 2; ++j) dk2[j] = dv2[j] = make_float2(0.f, 0.f);
  const int64_t qbeg = split * q_per_split, qend = min(lq, qbeg + q_per_split);
  for (int64_t q0 = qbeg; q0 < qend; q0 += XC) {
    __syncthreads();
    if (threadIdx.x < XC) {
      const int64_t qi = q0 + threadIdx.x;
      float a[HD], g[HD], ov[HD], de = 0.f;
      if (qi < qend) {
        load16(q + qi * ld + h * HD, a);
        load16(d_o + qi * ld + h * HD, g);
        load16(o + qi * ld + h * HD, ov);
        de = dot16v(g, ov);
        Ls[threadIdx.x] = lse[qi * heads + h];
      } else {
#pragma unroll
        for (int e = 0; e < HD; ++e) a[e] = g[e] = 0.f;
        Ls[threadIdx.x] = INFINITY;
      }
      Ds[threadIdx.x] = de;
#pragma unroll
      for (int e = 0; e < HD; ++e) {
        Qs[threadIdx.x * HD + e] = a[e] * 0.25f;
        Gs[threadIdx.x * HD + e] = g[e];
      }
    }
    __syncthreads();
    const int cnt = (int)min((int64_t)XC, qend - q0);
#pragma unroll 2
    for (int j = 0; j < cnt; ++j) {
      const float s = dot16(k2, Qs + j * HD), dp = dot16(v2, Gs + j * HD);
      const float p = expf(s - Ls[j]);
      const float ds = p * (dp - Ds[j]);
      axpy16(dv2, p, Gs + j * HD);
      axpy16(dk2, ds, Qs + j * HD);  // Qs already carries the 1/4
    }
  }
#pragma unroll
  for (int e = 0; e < HD / 2; ++e) {
    dka[2 * e] = dk2[e].x;
    dka[2 * e + 1] = dk2[e].y;
    dva[2 * e] = dv2[e].x;
    dva[2 * e + 1] = dv2[e].y;
  }
  if (live) {
    float* dstk = dk + ki * ld + h * HD;
    float* dstv = dv + ki * ld + h * HD;
    if (gridDim.z == 1) {
#pragma unroll
      for (int e = 0; e < HD; e += 4) {
        *reinterpret_cast<float4*>(dstk + e) = make_float4(dka[e], dka[e + 1], dka[e + 2], dka[e + 3]);
        *reinterpret_cast<float4*>(dstv + e) = make_float4(dva[e], dva[e + 1], dva[e + 2], dva[e + 3]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < HD; e += 4) {
        atomicAdd(reinterpret_cast<float4*>(dstk + e), make_float4(dka[e], dka[e + 1], dka[e + 2], dka[e + 3]));
        atomicAdd(reinterpret_cast<float4*>(dstv + e), make_float4(dva[e], dva[e + 1], dva[e + 2], dva[e + 3]));
      }
    }
  }
}

static int pick_splits(int64_t rows_parallel, int64_t loop_len, int heads) {
  const int64_t ctas = ((rows_parallel + XT - 1) / XT) * heads;
  if (ctas >= 2 * kNumSMs) return 1;
  int64_t want = (4 * kNumSMs + ctas - 1) / ctas;
  int64_t maxs = (loop_len + 2 * XC - 1) / (2 * XC);
  if (want > maxs) want = maxs;
  if (want > 64) want = 64;
  return (int)(want < 1 ? 1 : want);
}

}  // namespace mt

using namespace mt;

extern "C" int64_t mt_cross_attn_workspace_floats(int64_t lq, int64_t lk, int heads, int head_dim) {
  (void)head_dim;
  const int ns = pick_splits(lq, lk, heads);
  return ns == 1 ? 0 : (int64_t)ns * lq * heads * (HD + 2);
}

extern "C" int mt_cross_attn_fwd(const void* q, const void* k, const void* v, int dtype, void* o, float* lse,
                                 int64_t lq, int64_t lk, int heads, int head_dim, float* workspace,
                                 int64_t workspace_floats, void* stream) {
  MT_REQUIRE(head_dim == HD, "cross_attn: head_dim must be 16 (got %d)", head_dim);
  MT_REQUIRE(lq > 0 && lk > 0 && heads > 0 && heads <= 65535, "cross_attn: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int ns = pick_splits(lq, lk, heads);
  MT_REQUIRE(ns == 1 || (workspace != nullptr && workspace_floats >= (int64_t)ns * lq * heads * (HD + 2)),
             "cross_attn: workspace too small");
  const int64_t kps = ((lk + ns - 1) / ns + XC - 1) / XC * XC;
  dim3 grid((unsigned)((lq + XT - 1) / XT), (unsigned)heads, (unsigned)ns);
  if (dtype == MT_F32) {
    cross_fwd_kernel<float><<<grid, XT, 0, st>>>((const float*)q, (const float*)k, (const float*)v, (float*)o, lse,
                                                 workspace, lq, lk, heads, kps);
    if (ns > 1)
      cross_combine_kernel<float><<<(unsigned)((lq * heads + 127) / 128), 128, 0, st>>>(workspace, (float*)o, lse, lq, heads, ns);
  } else if (dtype == MT_BF16) {
    using bf = __nv_bfloat16;
    cross_fwd_kernel<bf><<<grid, XT, 0, st>>>((const bf*)q, (const bf*)k, (const bf*)v, (bf*)o, lse, workspace, lq, lk,
                                              heads, kps);
    if (ns > 1)
      cross_combine_kernel<bf><<<(unsigned)((lq * heads + 127) / 128), 128, 0, st>>>(workspace, (bf*)o, lse, lq, heads, ns);
  } else {
    set_error("cross_attn: bad dtype %d", dtype);
    return MT_E_BADARG;
  }
  return check_launch("cross_fwd_kernel");
}

extern "C" int mt_cross_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o,
                                 const float* lse, int dtype, float* dq_f32, float* dk_f32, float* dv_f32, int64_t lq,
                                 int64_t lk, int heads, int head_dim, void* stream) {
  MT_REQUIRE(head_dim == HD, "cross_attn: head_dim must be 16 (got %d)", head_dim);
  MT_REQUIRE(lq > 0 && lk > 0 && heads > 0 && heads <= 65535, "cross_attn: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ld = (size_t)heads * HD;
  const int nsk = pick_splits(lq, lk, heads), nsq = pick_splits(lk, lq, heads);
  MT_REQUIRE(((reinterpret_cast<uintptr_t>(dq_f32) | reinterpret_cast<uintptr_t>(dk_f32) |
               reinterpret_cast<uintptr_t>(dv_f32)) & 15) == 0, "cross_attn_bwd: gradients must be 16-byte aligned");
  // split loops accumulate with 16-byte reduce-adds into zero-filled gradients; an unsplit loop owns its rows and stores
  if (nsk > 1) MT_CUDA(cudaMemsetAsync(dq_f32, 0, sizeof(float) * lq * ld, st));
  if (nsq > 1) {
    MT_CUDA(cudaMemsetAsync(dk_f32, 0, sizeof(float) * lk * ld, st));
    MT_CUDA(cudaMemsetAsync(dv_f32, 0, sizeof(float) * lk * ld, st));
  }
  const int64_t kps = ((lk + nsk - 1) / nsk + XC - 1) / XC * XC;
  const int64_t qps = ((lq + nsq - 1) / nsq + XC - 1) / XC * XC;
  dim3 gq((unsigned)((lq + XT - 1) / XT), (unsigned)heads, (unsigned)nsk);
  dim3 gk((unsigned)((lk + XT - 1) / XT), (unsigned)heads, (unsigned)nsq);
  if (dtype == MT_F32) {
    using f = float;
    cross_bwd_dq_kernel<f><<<gq, XT, 0, st>>>((const f*)q, (const f*)k, (const f*)v, (const f*)o, (const f*)d_o, lse,
                                              dq_f32, lq, lk, heads, kps);
    cross_bwd_dkv_kernel<f><<<gk, XT, 0, st>>>((const f*)q, (const f*)k, (const f*)v, (const f*)o, (const f*)d_o, lse,
                                               dk_f32, dv_f32, lq, lk, heads, qps);
  } else if (dtype == MT_BF16) {
    using f = __nv_bfloat16;
    cross_bwd_dq_kernel<f><<<gq, XT, 0, st>>>((const f*)q, (const f*)k, (const f*)v, (const f*)o, (const f*)d_o, lse,
                                              dq_f32, lq, lk, heads, kps);
    cross_bwd_dkv_kernel<f><<<gk, XT, 0, st>>>((const f*)q, (const f*)k, (const f*)v, (const f*)o, (const f*)d_o, lse,
                                               dk_f32, dv_f32, lq, lk, heads, qps);
  } else {
    set_error("cross_attn: bad dtype %d", dtype);
    return MT_E_BADARG;
  }
  return check_launch("cross_bwd kernels");
}
