// LongNet dilated attention on the 5th-generation tensor cores (sm_100a): TMA-staged Q/K/V tiles, tcgen05 MMAs with
// TMEM accumulators, online softmax in fp32.
//
// One CTA = one (branch, segment, head, 128-slot query tile); it streams the 128-slot key/value tiles of the same
// (branch, segment, head).  The dilated gather of the reference (dilated_attention.py:22-37, 82-111) is the TMA box
// itself: the qkv buffer [n_alloc, 3*768] is described per branch as a 3-D tensor (column, residue o, slot j) with
// position = j * r + o, so a box of (64 columns, 1, 128 slots) IS the sparse tile of one head -- no sparse copy exists.
// Rows >= n_tokens are zero in memory and rows >= n_alloc are zero-filled by TMA: these are the reference's zero
// keys (score 0, value 0, counted in the softmax denominator).  Slots past the segment's own m = ceil(g / r) belong to
// the next segment and are masked to -inf.
//
// head_dim = 48: the box is 64 columns wide (one 128-byte swizzle atom per row), the MMAs only consume K = 48 (three
// k-steps of 16) for S = Q K^T and N = 48 for O = P V, so no tensor-core work is spent on the padding columns.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2-5 = softmax (one query
// row per thread = one TMEM lane).  P is written back to TMEM as packed bf16 (tcgen05.st) and is the A operand of the
// P V MMA straight from TMEM (no shared-memory round trip); each P V result lands in a TMEM tile of its own and is
// folded into fp32 register accumulators with the usual online-softmax rescale, so no accumulator is ever
// read-modify-written in TMEM.
#include <cuda.h>

#include <type_traits>

#include "mt_common.cuh"
#include "sm100_ptx.cuh"

namespace mt {
using namespace sm100;

// experiment-only phase tracing (compile with -DMT_DEBUG_TRACE): block 0 prints clock64 stamps of its first iterations
#ifdef MT_DEBUG_TRACE
#ifndef MT_TRACE_BLOCK
#define MT_TRACE_BLOCK 0
#endif
#define MT_TRACE_DECL long long tr_[96]; int trn_ = 0; const bool tron_ = (blockIdx.x == MT_TRACE_BLOCK);
#define MT_TRACE(tag) do { if (tron_ && trn_ < 94) { tr_[trn_++] = (long long)(tag); tr_[trn_++] = clock64(); } } while (0)
#define MT_TRACE_DUMP(who) do { if (tron_) for (int q_ = 0; q_ + 1 < trn_; q_ += 2) printf("%s %lld %lld\n", who, tr_[q_], tr_[q_ + 1] & 0xffffffffll); } while (0)
#else
#define MT_TRACE_DECL
#define MT_TRACE(tag)
#define MT_TRACE_DUMP(who)
#endif

#ifndef MT_FWD_POLY
#define MT_FWD_POLY 0                     // exponentials per 8 computed on the FMA pipes in the forward (ex2_poly)
#endif
static constexpr int DH = 48;            // head dim
static constexpr int BT = 128;           // slots per tile (queries and keys)
static constexpr int TILE_BYTES = BT * 128;  // [128 rows][128 B]: 64 bf16 columns per row, SWIZZLE_128B
static constexpr int KV_STAGES = 2;
static constexpr int FWD_THREADS = 192;
static constexpr uint32_t TMEM_COLS = 256;  // S: [0,128)  P (bf16 pairs): [128,192)  O tile: [192,256)

struct TensorMaps {
  CUtensorMap m[MT_MAX_BRANCHES];
};

struct Sm100Params {
  DilatedGeom geo;
  int item_prefix[MT_MAX_BRANCHES + 1];  // first CTA of each launch-order branch
  int order[MT_MAX_BRANCHES];            // launch order -> branch (longest key loops first)
  int tiles[MT_MAX_BRANCHES];            // 128-slot tiles per (segment, head), indexed by branch
  float scale_log2;                      // head_dim^-0.5 * log2(e)
  float scale;
};

// smem carve-up (dynamic, base must be 1024-byte aligned)
struct FwdSmem {
  static constexpr int Q = 0;
  static constexpr int K = Q + TILE_BYTES;
  static constexpr int V = K + KV_STAGES * TILE_BYTES;
  static constexpr int BAR = V + KV_STAGES * TILE_BYTES;
  // barriers (8 B each): q_full, kv_full[2], kv_empty[2], s_full, s_free, p_full, o_full[2]; then the TMEM pointer
  static constexpr int NBAR = 10;
  static constexpr int TMEM_PTR = BAR + NBAR * 8;
  static constexpr int XCH = TMEM_PTR + 16;          // forward with two threads per row: 3 x [2][128] floats
  static constexpr int TOTAL = XCH + 3 * 2 * 128 * 4;
};

__global__ void __launch_bounds__(FWD_THREADS, 2)
dilated_fwd_sm100_kernel(const __grid_constant__ TensorMaps maps, const Sm100Params P, __nv_bfloat16* __restrict__ o_br,
                         float* __restrict__ lse_br, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  // roles by warp id: the scheduler favours high warp ids, so the latency-critical single-thread roles sit last
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = 4, W_MMA = 5;
  if ((sbase & 1023u) != 0) {  // SWIZZLE_128B tiles need 1024-byte alignment; never expected, but fail loudly
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  // ---- which tile -------------------------------------------------------------------------------------------------
  int oi = 0;
  while (oi + 1 < P.geo.nb && (int)blockIdx.x >= P.item_prefix[oi + 1]) ++oi;
  const int b = P.order[oi];
  const BranchGeom bg = P.geo.b[b];
  int local = blockIdx.x - P.item_prefix[oi];
  const int qt = local % P.tiles[b];
  local /= P.tiles[b];
  const int h = local % P.geo.H;
  const int s = local / P.geo.H;
  const int H = P.geo.H, N = P.geo.N, E = H * DH;
  const int off = (h * bg.r) / H;                    // residue of the positions this head owns
  const int jseg = (s * bg.g) / bg.r;                // first slot of the segment in the branch's j axis
  const int n_kv = P.tiles[b];
  const int q0 = qt * BT;

  const uint32_t bar_q_full = sbase + FwdSmem::BAR + 0;
  const uint32_t bar_kv_full = sbase + FwdSmem::BAR + 8;    // [2]
  const uint32_t bar_kv_empty = sbase + FwdSmem::BAR + 24;  // [2]
  const uint32_t bar_s_full = sbase + FwdSmem::BAR + 40;
  const uint32_t bar_s_free = sbase + FwdSmem::BAR + 48;
  const uint32_t bar_p_full = sbase + FwdSmem::BAR + 56;
  const uint32_t bar_o_full = sbase + FwdSmem::BAR + 64;    // [2]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + FwdSmem::TMEM_PTR);

  if (threadIdx.x == 0) {
    mbar_init(bar_q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(bar_kv_full + 8 * i, 1);
      mbar_init(bar_kv_empty + 8 * i, 1);
      mbar_init(bar_o_full + 8 * i, 1);
    }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_free, 128);
    mbar_init(bar_p_full, 128);
    fence_barrier_init();
    tma_prefetch_desc(&maps.m[b]);
  }
  if (warp == W_MMA) {
    tmem_alloc(smem_u32((const void*)tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_s = tmem;
  const uint32_t tmem_p = tmem + 128;  // P as the A operand of P V: lane = query row, column k/2 holds keys (k, k+1)
  const uint32_t tmem_o = tmem + 192;

  if (warp == W_TMA) {
    // ===== TMA producer ===============================================================================================
    if (lane == 0) {
      const void* map = &maps.m[b];
      mbar_expect_tx(bar_q_full, TILE_BYTES);
      tma_load_3d(sbase + FwdSmem::Q, map, bar_q_full, h * DH, off, jseg + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1, use = j >> 1;
        mbar_wait(bar_kv_empty + 8 * st, (use & 1) ^ 1);
        mbar_expect_tx(bar_kv_full + 8 * st, 2 * TILE_BYTES);
        tma_load_3d(sbase + FwdSmem::K + st * TILE_BYTES, map, bar_kv_full + 8 * st, E + h * DH, off, jseg + j * BT);
        tma_load_3d(sbase + FwdSmem::V + st * TILE_BYTES, map, bar_kv_full + 8 * st, 2 * E + h * DH, off, jseg + j * BT);
      }
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer =================================================================================================
    constexpr uint32_t IDESC_QK = umma_idesc_bf16(BT, BT, 0, 0);
    constexpr uint32_t IDESC_PV = umma_idesc_bf16(BT, DH, 0, 1);
    // descriptors are built once; inside the loops an MMA costs one UTCHMMA (+ a constant descriptor advance)
    const uint64_t q_desc = umma_smem_desc(sbase + FwdSmem::Q, 16, 1024);
    const uint64_t k_desc0 = umma_smem_desc(sbase + FwdSmem::K, 16, 1024);
    const uint64_t k_desc1 = umma_smem_desc(sbase + FwdSmem::K + TILE_BYTES, 16, 1024);
    const uint64_t v_desc0 = umma_smem_desc(sbase + FwdSmem::V, TILE_BYTES, 1024);
    const uint64_t v_desc1 = umma_smem_desc(sbase + FwdSmem::V + TILE_BYTES, TILE_BYTES, 1024);
    auto issue_qk = [&](int j) {  // S = Q K_j^T : three k-steps of 16 inside the 128-byte swizzle atom
      if (elect_one()) {
        const uint64_t kd = (j & 1) ? k_desc1 : k_desc0;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_ss(tmem_s, umma_desc_adv(q_desc, k * 32), umma_desc_adv(kd, k * 32), IDESC_QK, k > 0);
        umma_commit(bar_s_full);
      }
      __syncwarp();
    };
    MT_TRACE_DECL
    mbar_wait(bar_q_full, 0);
    mbar_wait(bar_kv_full, 0);
    tc_fence_after();
    MT_TRACE(0);
    issue_qk(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) {
        mbar_wait(bar_kv_full + 8 * ((j + 1) & 1), ((j + 1) >> 1) & 1);
        MT_TRACE(100 + j);
        mbar_wait(bar_s_free, j & 1);  // the softmax threads have read S_j out of TMEM
        tc_fence_after();
        MT_TRACE(200 + j);
        issue_qk(j + 1);
      }
      mbar_wait(bar_p_full, j & 1);    // P_j is in TMEM (and the O tile of P_{j-1} V_{j-1} has been folded)
      tc_fence_after();
      MT_TRACE(300 + j);
      if (elect_one()) {
        // O_tile = P_j V_j : A = P straight from TMEM (16 keys = 8 packed columns per k-step), B = V in place as an
        // MN-major operand (keys are the rows of the tile: 16 rows = 2048 B per k-step)
        const uint64_t vd = (j & 1) ? v_desc1 : v_desc0;
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)
          umma_ts(tmem_o, tmem_p + k * 8, umma_desc_adv(vd, k * 2048), IDESC_PV, k > 0);
        umma_commit(bar_kv_empty + 8 * (j & 1));
        umma_commit(bar_o_full);
      }
      __syncwarp();
    }
    if (lane == 0) { MT_TRACE_DUMP("mma"); }
  } else {
    // ===== softmax: one query row per thread ==========================================================================
    const int lane_grp = warp & 3;                    // TMEM lanes this warp may touch: [32*lane_grp, +32)
    const int row = lane_grp * 32 + lane;
    const uint32_t t_lane = (uint32_t)(lane_grp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o_acc[i] = 0.f;

    auto fold = [&](int j) {  // o_acc += O_tile(j)
      mbar_wait(bar_o_full, j & 1);
      tc_fence_after();
      float t[16];
#pragma unroll
      for (int c = 0; c < DH / 16; ++c) {
        tmem_ld16(tmem_o + t_lane + c * 16, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[c * 16 + i] += t[i];
      }
    };

    const float scale_log2 = P.scale_log2;
    MT_TRACE_DECL
    // one key tile: MASK = the tile holds slots past the segment's m (only ever the last tile of the loop)
    auto tile = [&](int j, auto mask_tag) {
      constexpr bool MASK = decltype(mask_tag)::value;
      const int kvalid = bg.m - j * BT;  // key slots of this tile that belong to the segment (>= 1)
      MT_TRACE(1000 + j);
      mbar_wait(bar_s_full, j & 1);
      tc_fence_after();
      MT_TRACE(1100 + j);
      // pass 1: row maximum
      float mx = -INFINITY;
      float sv[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(tmem_s + t_lane + c * 32, sv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (!MASK || c * 32 + i < kvalid) ? sv[i] : -INFINITY);
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2((m_run - m_new) * scale_log2);
      MT_TRACE(1200 + j);
      if (j > 0) fold(j - 1);  // O tile j-1 is relative to m_run; P's columns are free once that MMA has completed
      MT_TRACE(1300 + j);
#pragma unroll
      for (int i = 0; i < DH; ++i) o_acc[i] *= alpha;
      l_run *= alpha;
      m_run = m_new;
      const float mb = m_new * scale_log2;
      // pass 2: p = exp2(s * scale_log2 - mb), row sum, bf16 P into TMEM
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(tmem_s + t_lane + c * 32, sv);
        tmem_ld_wait();
        if (c == 3) {  // S_j is fully in registers: the MMA warp may overwrite it with S_{j+1}
          tc_fence_before();
          mbar_arrive(bar_s_free);
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = ex2(fmaf(sv[i], scale_log2, -mb));
          float p1 = ex2(fmaf(sv[i + 1], scale_log2, -mb));
          if (MASK) {
            p0 = (c * 32 + i < kvalid) ? p0 : 0.f;
            p1 = (c * 32 + i + 1 < kvalid) ? p1 : 0.f;
          }
          rs += p0 + p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
        tmem_st16(tmem_p + t_lane + c * 16, pk);  // 32 keys = 16 packed columns
      }
      l_run += rs;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p_full);
      MT_TRACE(1400 + j);
    };
    for (int j = 0; j < n_kv; ++j) {
      if (bg.m - j * BT >= BT) tile(j, std::false_type{});
      else tile(j, std::true_type{});
    }
    fold(n_kv - 1);
    if (warp == 0 && lane == 0) { MT_TRACE_DUMP("smx"); }
    // ---- epilogue: normalise and write the compact per-branch output ---------------------------------------------
    const int slot = q0 + row;
    const int pos = s * bg.g + off + slot * bg.r;
    const int seg_end = min(N, (s + 1) * bg.g);
    if (slot < bg.m && pos < seg_end) {
      const float inv = 1.f / l_run;
      const int slot_h = h - off * bg.hpb;
      __nv_bfloat16* dst = o_br + bg.o_off + ((int64_t)pos * bg.hpb + slot_h) * DH;
#pragma unroll
      for (int c = 0; c < DH / 8; ++c) {
        uint4 u;
        u.x = pack_bf16(o_acc[c * 8 + 0] * inv, o_acc[c * 8 + 1] * inv);
        u.y = pack_bf16(o_acc[c * 8 + 2] * inv, o_acc[c * 8 + 3] * inv);
        u.z = pack_bf16(o_acc[c * 8 + 4] * inv, o_acc[c * 8 + 5] * inv);
        u.w = pack_bf16(o_acc[c * 8 + 6] * inv, o_acc[c * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + c * 8) = u;
      }
      lse_br[bg.lse_off + (int64_t)pos * bg.hpb + slot_h] = m_run * P.scale + logf(l_run);
    }
  }
  // ---- teardown ------------------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// forward, second version: O accumulated in TMEM with a lazily raised row maximum (no per-tile fold / rescale in
// registers), the whole 128-column score row loaded once, 3-input maxima.
// TPR = threads per query row: 1 (impl 2) or 2 (impl 3: eight softmax warps, warps w and w + 4 share the 32 TMEM lanes of
// their rows and each own 64 of the 128 score columns and 24 of the 48 output columns; the row maximum is exchanged
// through shared memory once per tile, the row sums once per CTA).  With four softmax warps the kernel is bound by the
// latency of each warp's dependent instruction stream (issue slots 44 %, MUFU 63 % busy, and moving exponentials to the
// FMA pipes did not help, MT_FWD_POLY); eight warps give every scheduler four streams to interleave.
// ---------------------------------------------------------------------------------------------------------------------
// 64-thread named barrier of the two softmax warps that share TMEM lane group g (compile-time ids 1..4)
__device__ __forceinline__ void pair_bar_sync(int g) {
  switch (g) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

template <int TPR>
__global__ void __launch_bounds__(64 + 128 * TPR, 2)
dilated_fwd2_sm100_kernel(const __grid_constant__ TensorMaps maps, const Sm100Params P, __nv_bfloat16* __restrict__ o_br,
                         float* __restrict__ lse_br, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  // roles by warp id: the scheduler favours high warp ids, so the latency-critical single-thread roles sit last
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = 4 * TPR, W_MMA = 4 * TPR + 1;
  if ((sbase & 1023u) != 0) {  // SWIZZLE_128B tiles need 1024-byte alignment; never expected, but fail loudly
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  // ---- which tile -------------------------------------------------------------------------------------------------
  int oi = 0;
  while (oi + 1 < P.geo.nb && (int)blockIdx.x >= P.item_prefix[oi + 1]) ++oi;
  const int b = P.order[oi];
  const BranchGeom bg = P.geo.b[b];
  int local = blockIdx.x - P.item_prefix[oi];
  const int qt = local % P.tiles[b];
  local /= P.tiles[b];
  const int h = local % P.geo.H;
  const int s = local / P.geo.H;
  const int H = P.geo.H, N = P.geo.N, E = H * DH;
  const int off = (h * bg.r) / H;                    // residue of the positions this head owns
  const int jseg = (s * bg.g) / bg.r;                // first slot of the segment in the branch's j axis
  const int q0 = qt * BT;
  // Zero padding (positions >= N in the last segment, dilated_attention.py:82-111) in whole tiles is never computed:
  // a query tile without a real position writes nothing, and key tiles without a real position are all zero keys
  // (score 0, value 0), whose only effect -- n_zero_tail * exp(0 - max) in the softmax denominator -- is added in
  // closed form in the epilogue.  At 32k tiles 30 % of the padded tile pairs of the reference disappear this way.
  const int seg_lo = s * bg.g + off;
  const int c_real = min(N, (s + 1) * bg.g) > seg_lo ? (min(N, (s + 1) * bg.g) - seg_lo + bg.r - 1) / bg.r : 0;
  if (q0 >= c_real) return;
  const int n_kv = min(P.tiles[b], (c_real + BT - 1) / BT);
  const int n_zero_tail = max(0, bg.m - n_kv * BT);

  const uint32_t bar_q_full = sbase + FwdSmem::BAR + 0;
  const uint32_t bar_kv_full = sbase + FwdSmem::BAR + 8;    // [2]
  const uint32_t bar_kv_empty = sbase + FwdSmem::BAR + 24;  // [2]
  const uint32_t bar_s_full = sbase + FwdSmem::BAR + 40;
  const uint32_t bar_s_free = sbase + FwdSmem::BAR + 48;
  const uint32_t bar_p_full = sbase + FwdSmem::BAR + 56;
  const uint32_t bar_o_full = sbase + FwdSmem::BAR + 64;    // [2]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + FwdSmem::TMEM_PTR);

  if (threadIdx.x == 0) {
    mbar_init(bar_q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      mbar_init(bar_kv_full + 8 * i, 1);
      mbar_init(bar_kv_empty + 8 * i, 1);
      mbar_init(bar_o_full + 8 * i, 1);
    }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_free, 128 * TPR);
    mbar_init(bar_p_full, 128 * TPR);
    fence_barrier_init();
    tma_prefetch_desc(&maps.m[b]);
  }
  if (warp == W_MMA) {
    tmem_alloc(smem_u32((const void*)tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_s = tmem;
  const uint32_t tmem_p = tmem + 128;  // P as the A operand of P V: lane = query row, column k/2 holds keys (k, k+1)
  const uint32_t tmem_o = tmem + 192;

  if (warp == W_TMA) {
    // ===== TMA producer ===============================================================================================
    if (lane == 0) {
      const void* map = &maps.m[b];
      mbar_expect_tx(bar_q_full, TILE_BYTES);
      tma_load_3d(sbase + FwdSmem::Q, map, bar_q_full, h * DH, off, jseg + q0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1, use = j >> 1;
        mbar_wait(bar_kv_empty + 8 * st, (use & 1) ^ 1);
        mbar_expect_tx(bar_kv_full + 8 * st, 2 * TILE_BYTES);
        tma_load_3d(sbase + FwdSmem::K + st * TILE_BYTES, map, bar_kv_full + 8 * st, E + h * DH, off, jseg + j * BT);
        tma_load_3d(sbase + FwdSmem::V + st * TILE_BYTES, map, bar_kv_full + 8 * st, 2 * E + h * DH, off, jseg + j * BT);
      }
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer =================================================================================================
    constexpr uint32_t IDESC_QK = umma_idesc_bf16(BT, BT, 0, 0);
    constexpr uint32_t IDESC_PV = umma_idesc_bf16(BT, DH, 0, 1);
    // descriptors are built once; inside the loops an MMA costs one UTCHMMA (+ a constant descriptor advance)
    const uint64_t q_desc = umma_smem_desc(sbase + FwdSmem::Q, 16, 1024);
    const uint64_t k_desc0 = umma_smem_desc(sbase + FwdSmem::K, 16, 1024);
    const uint64_t k_desc1 = umma_smem_desc(sbase + FwdSmem::K + TILE_BYTES, 16, 1024);
    const uint64_t v_desc0 = umma_smem_desc(sbase + FwdSmem::V, TILE_BYTES, 1024);
    const uint64_t v_desc1 = umma_smem_desc(sbase + FwdSmem::V + TILE_BYTES, TILE_BYTES, 1024);
    auto issue_qk = [&](int j) {  // S = Q K_j^T : three k-steps of 16 inside the 128-byte swizzle atom
      if (elect_one()) {
        const uint64_t kd = (j & 1) ? k_desc1 : k_desc0;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_ss(tmem_s, umma_desc_adv(q_desc, k * 32), umma_desc_adv(kd, k * 32), IDESC_QK, k > 0);
        umma_commit(bar_s_full);
      }
      __syncwarp();
    };
    MT_TRACE_DECL
    mbar_wait(bar_q_full, 0);
    mbar_wait(bar_kv_full, 0);
    tc_fence_after();
    MT_TRACE(0);
    issue_qk(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) {
        mbar_wait(bar_kv_full + 8 * ((j + 1) & 1), ((j + 1) >> 1) & 1);
        MT_TRACE(100 + j);
        mbar_wait(bar_s_free, j & 1);  // the softmax threads have read S_j out of TMEM
        tc_fence_after();
        MT_TRACE(200 + j);
        issue_qk(j + 1);
      }
      mbar_wait(bar_p_full, j & 1);    // P_j is in TMEM (and the O tile of P_{j-1} V_{j-1} has been folded)
      tc_fence_after();
      MT_TRACE(300 + j);
      if (elect_one()) {
        // O_tile = P_j V_j : A = P straight from TMEM (16 keys = 8 packed columns per k-step), B = V in place as an
        // MN-major operand (keys are the rows of the tile: 16 rows = 2048 B per k-step)
        const uint64_t vd = (j & 1) ? v_desc1 : v_desc0;
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)
          umma_ts(tmem_o, tmem_p + k * 8, umma_desc_adv(vd, k * 2048), IDESC_PV, (j > 0) || (k > 0));
        umma_commit(bar_kv_empty + 8 * (j & 1));
        umma_commit(bar_o_full);
      }
      __syncwarp();
    }
    if (lane == 0) { MT_TRACE_DUMP("mma"); }
  } else {
    // ===== softmax: one query row per thread; O accumulates in TMEM over the whole key loop =========================
    // The row maximum used for the exponentials is only raised when the new maximum exceeds it by more than 8 in the
    // log2 domain (P <= 2^8 stays exact enough in bf16, l is fp32): the O accumulator then never needs the per-tile
    // rescale, and in the rare tile where a row does move, its warp rescales its 32 rows of O in TMEM in place.
    const int lane_grp = warp & 3;                    // TMEM lanes this warp may touch: [32*lane_grp, +32)
    const int half = (TPR == 2) ? (warp >> 2) : 0;    // which half of the score / output columns this thread owns
    constexpr int NC = BT / TPR;                      // score columns per thread
    constexpr int OC = DH / TPR;                      // output columns per thread
    float* xch = reinterpret_cast<float*>(smem + FwdSmem::XCH);   // [3][2 halves][128 rows]: maxima (x2), row sums
    const int row = lane_grp * 32 + lane;
    const uint32_t t_lane = (uint32_t)(lane_grp * 32) << 16;
    float m_used = -INFINITY, l_run = 0.f;
    const float scale_log2 = P.scale_log2;
    MT_TRACE_DECL
    auto tile = [&](int j, auto mask_tag) {
      constexpr bool MASK = decltype(mask_tag)::value;
      const int kvalid = bg.m - j * BT;  // key slots of this tile that belong to the segment (>= 1)
      MT_TRACE(1000 + j);
      mbar_wait(bar_s_full, j & 1);
      tc_fence_after();
      MT_TRACE(1100 + j);
      float sv[NC];
      tmem_ld64(tmem_s + t_lane + half * NC, sv);
      if (TPR == 1) tmem_ld64(tmem_s + t_lane + 64, sv + (TPR == 1 ? 64 : 0));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_s_free);          // S_j is in registers: the MMA warp may overwrite it with S_{j+1}
      if (MASK) {
#pragma unroll
        for (int i = 0; i < NC; ++i) sv[i] = (half * NC + i < kvalid) ? sv[i] : -INFINITY;
      }
      float mx = fmax3(sv[0], sv[1], sv[2]);
#pragma unroll
      for (int i = 3; i + 1 < NC; i += 2) mx = fmax3(mx, sv[i], sv[i + 1]);
      mx = fmaxf(mx, sv[NC - 1]);
      if (TPR == 2) {   // the two threads of a row agree on its maximum (double-buffered slot, 64-thread named barrier)
        float* x = xch + (j & 1) * 256;
        x[half * 128 + row] = mx;
        pair_bar_sync(lane_grp);
        mx = fmaxf(mx, x[(half ^ 1) * 128 + row]);
      }
      MT_TRACE(1200 + j);
      bool waited = false;
      if (j == 0) {
        m_used = mx;
      } else {
        const bool need = (mx - m_used) * scale_log2 > 8.f;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(bar_o_full, (j - 1) & 1);   // P V of tile j-1 has completed: O may be touched
          tc_fence_after();
          waited = true;
          const float alpha = need ? ex2((m_used - mx) * scale_log2) : 1.f;
          float t[8];
#pragma unroll
          for (int c = 0; c < OC / 8; ++c) {
            tmem_ld8(tmem_o + t_lane + half * OC + c * 8, t);
            tmem_ld_wait();
            uint32_t u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(t[i] * alpha);
            tmem_st8(tmem_o + t_lane + half * OC + c * 8, u);
          }
          if (need) {
            l_run *= alpha;
            m_used = mx;
          }
        }
      }
      const float mb = m_used * scale_log2;
      if (j > 0 && !waited) {             // P_j overwrites P_{j-1}: its P V must have read it
        mbar_wait(bar_o_full, (j - 1) & 1);
        tc_fence_after();
      }
      MT_TRACE(1300 + j);
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < NC / 32; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          // MT_FWD_POLY of every 8 exponentials run as an FMA-pipe polynomial (full tiles only: masked slots need
          // ex2(-inf) = 0 exactly)
          const float x0 = fmaf(sv[c * 32 + i], scale_log2, -mb);
          const float x1 = fmaf(sv[c * 32 + i + 1], scale_log2, -mb);
#ifndef MT_EXP_NO_MUFU  // experiment builds only: the loop without the exponentials (wrong results)
          const float p0 = (!MASK && (i & 7) < MT_FWD_POLY) ? ex2_poly(x0) : ex2(x0);
          const float p1 = (!MASK && ((i + 1) & 7) < MT_FWD_POLY) ? ex2_poly(x1) : ex2(x1);
#else
          const float p0 = x0 * 0.01f, p1 = x1 * 0.01f;
#endif
          rs += p0 + p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
#ifndef MT_EXP_NO_PST   // experiment builds only: time the loop without the P stores (wrong results)
        tmem_st16(tmem_p + t_lane + half * (NC / 2) + c * 16, pk);  // 32 keys = 16 packed columns
#else
        if (pk[0] == 0x12345678u && pk[15] == 0x9abcdef0u) l_run += 1.f;
#endif
      }
      l_run += rs;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p_full);
      MT_TRACE(1400 + j);
    };
    for (int j = 0; j < n_kv; ++j) {
      if (bg.m - j * BT >= BT) tile(j, std::false_type{});
      else tile(j, std::true_type{});
    }
    mbar_wait(bar_o_full, (n_kv - 1) & 1);
    tc_fence_after();
    if (warp == 0 && lane == 0) { MT_TRACE_DUMP("smx"); }
    // ---- epilogue: normalise and write the compact per-branch output ---------------------------------------------
    const int slot = q0 + row;
    const int pos = s * bg.g + off + slot * bg.r;
    const int seg_end = min(N, (s + 1) * bg.g);
    float o_acc[OC];
#pragma unroll
    for (int c = 0; c < OC / 8; ++c) {
      float t[8];
      tmem_ld8(tmem_o + t_lane + half * OC + c * 8, t);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) o_acc[c * 8 + i] = t[i];
    }
    if (TPR == 2) {   // row sum = sum of the two column halves
      float* x = xch + 512;
      x[half * 128 + row] = l_run;
      pair_bar_sync(lane_grp);
      l_run += x[(half ^ 1) * 128 + row];
    }
    if (n_zero_tail > 0) {   // the zero keys of the tiles that were skipped
      const float m_fin = fmaxf(m_used, 0.f);
      const float alpha = ex2((m_used - m_fin) * scale_log2);
      l_run = l_run * alpha + (float)n_zero_tail * ex2(-m_fin * scale_log2);
#pragma unroll
      for (int i = 0; i < OC; ++i) o_acc[i] *= alpha;
      m_used = m_fin;
    }
    if (slot < bg.m && pos < seg_end) {
      const float inv = 1.f / l_run;
      const int slot_h = h - off * bg.hpb;
      __nv_bfloat16* dst = o_br + bg.o_off + ((int64_t)pos * bg.hpb + slot_h) * DH + half * OC;
#pragma unroll
      for (int c = 0; c < OC / 8; ++c) {
        uint4 u;
        u.x = pack_bf16(o_acc[c * 8 + 0] * inv, o_acc[c * 8 + 1] * inv);
        u.y = pack_bf16(o_acc[c * 8 + 2] * inv, o_acc[c * 8 + 3] * inv);
        u.z = pack_bf16(o_acc[c * 8 + 4] * inv, o_acc[c * 8 + 5] * inv);
        u.w = pack_bf16(o_acc[c * 8 + 6] * inv, o_acc[c * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + c * 8) = u;
      }
      if (half == 0) lse_br[bg.lse_off + (int64_t)pos * bg.hpb + slot_h] = m_used * P.scale + logf(l_run);
    }
  }
  // ---- teardown ------------------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 3-D view of a row-major [n_alloc, ld] bf16 matrix for dilation r: (column, residue o, slot j) -> row j*r + o
static int encode_branch_map(CUtensorMap* map, const void* base, int64_t ld, int64_t n_alloc, int r) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return MT_E_UNSUPPORTED;
  }
  cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)r, (cuuint64_t)(n_alloc / r)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)r};
  cuuint32_t box[3] = {64, 1, (cuuint32_t)BT};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for dilation %d, ld %lld, rows %lld", (int)rc, r, (long long)ld,
              (long long)n_alloc);
    return MT_E_BADARG;
  }
  return 0;
}

// fp32 [n_alloc, ld] gradient buffer, same (column, residue, slot) view, for the TMA reduce-add of dQ tiles:
// box = [128 slots][cols] floats; cols = 48 unswizzled (dense staging tile), or 32 / 16 with the 128 B / 64 B swizzle
// (row pitch = swizzle span, so that one-row-per-thread staging stores are bank-conflict free)
static int encode_branch_map_f32(CUtensorMap* map, const void* base, int64_t ld, int64_t n_alloc, int r, int cols = DH) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return MT_E_UNSUPPORTED;
  }
  cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)r, (cuuint64_t)(n_alloc / r)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * 4 * (cuuint64_t)r};
  cuuint32_t box[3] = {(cuuint32_t)cols, 1, (cuuint32_t)BT};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle swz = cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32, %d columns) failed (%d) for dilation %d", cols, (int)rc, r);
    return MT_E_BADARG;
  }
  return 0;
}

static int make_sm100_params(const mt_dilated_geometry* geom, Sm100Params* P) {
  int rc = make_dilated_geom(geom, &P->geo);
  if (rc) return rc;
  if (P->geo.D != DH || P->geo.H != 16) {
    set_error("tcgen05 dilated attention is built for 16 heads x 48 (got %d x %d)", P->geo.H, P->geo.D);
    return MT_E_UNSUPPORTED;
  }
  const int nb = P->geo.nb;
  for (int b = 0; b < nb; ++b) {
    P->tiles[b] = (P->geo.b[b].m + BT - 1) / BT;
    P->order[b] = b;
  }
  for (int i = 0; i < nb; ++i)  // longest key loops first
    for (int j = i + 1; j < nb; ++j)
      if (P->tiles[P->order[j]] > P->tiles[P->order[i]]) {
        int t = P->order[i];
        P->order[i] = P->order[j];
        P->order[j] = t;
      }
  int total = 0;
  for (int i = 0; i < nb; ++i) {
    const int b = P->order[i];
    P->item_prefix[i] = total;
    total += P->geo.b[b].n_seg * P->geo.H * P->tiles[b];
  }
  P->item_prefix[nb] = total;
  P->scale = 1.0f / sqrtf((float)DH);
  P->scale_log2 = P->scale * 1.4426950408889634f;
  return 0;
}

static int* error_flag() {  // one device word, allocated once (reported through the return code of the next call)
  static int* flag = nullptr;
  if (flag == nullptr) {
    if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(flag, 0, sizeof(int));
  }
  return flag;
}

int dilated_attn_fwd_sm100(const mt_dilated_geometry* geom, const void* qkv, int64_t qkv_ld, int64_t n_alloc,
                           void* o_br, float* lse_br, int impl, cudaStream_t st) {
  Sm100Params P;
  int rc = make_sm100_params(geom, &P);
  if (rc) return rc;
  MT_REQUIRE(n_alloc >= P.geo.N && n_alloc % 128 == 0, "dilated_attn_fwd: n_alloc must be a multiple of 128 >= n_tokens");
  MT_REQUIRE(qkv_ld % 8 == 0 && ((uintptr_t)qkv & 15) == 0, "dilated_attn_fwd: qkv must be 16-byte aligned");
  TensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int b = 0; b < P.geo.nb; ++b) {
    rc = encode_branch_map(&maps.m[b], qkv, qkv_ld, n_alloc, P.geo.b[b].r);
    if (rc) return rc;
  }
  int* flag = error_flag();
  MT_REQUIRE(flag != nullptr, "dilated_attn_fwd: cannot allocate the error flag");
  if (impl == 2) {
    MT_CUDA(cudaFuncSetAttribute(dilated_fwd2_sm100_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    dilated_fwd2_sm100_kernel<1><<<P.item_prefix[P.geo.nb], FWD_THREADS, FwdSmem::TOTAL, st>>>(
        maps, P, (__nv_bfloat16*)o_br, lse_br, flag);
    return check_launch("dilated_fwd2_sm100_kernel<1>");
  }
  if (impl == 3) {
    MT_CUDA(cudaFuncSetAttribute(dilated_fwd2_sm100_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
    dilated_fwd2_sm100_kernel<2><<<P.item_prefix[P.geo.nb], 64 + 256, FwdSmem::TOTAL, st>>>(
        maps, P, (__nv_bfloat16*)o_br, lse_br, flag);
    return check_launch("dilated_fwd2_sm100_kernel<2>");
  }
  MT_CUDA(cudaFuncSetAttribute(dilated_fwd_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::TOTAL));
  dilated_fwd_sm100_kernel<<<P.item_prefix[P.geo.nb], FWD_THREADS, FwdSmem::TOTAL, st>>>(
      maps, P, (__nv_bfloat16*)o_br, lse_br, flag);
  return check_launch("dilated_fwd_sm100_kernel");
}

// =====================================================================================================================
// backward
// =====================================================================================================================
// One CTA = one (branch, segment, head, 128-slot KEY tile); it streams the 128-slot query tiles of the same
// (branch, segment, head).  Per (query tile i, key tile j):
//     S  = Q_i K_j^T                  dP = dO_i V_j^T                      (tcgen05, K = 48, into TMEM)
//     P  = exp(S * scale - lse_i)     dS = P * (dP - delta_i) * scale      (fp32, 512 threads, bf16 into smem)
//     dV_j += P^T dO_i                dK_j += dS^T Q_i                     (tcgen05, accumulate in TMEM over i)
//     dQ_i  = dS K_j                                                        (tcgen05, fresh TMEM tile, then
//                                                                            red.global.add.v4.f32 into dqkv)
// lse is the MERGED log-sum-exp over the branches and delta_i = dO_i . o_b,i the per-branch row dot: with them
// P = w_b p_b and the formula above is the reference's backward with detached merge weights.
// P and dS are stored once as [query][key] swizzled tiles and consumed both as K-major A (dQ) and as MN-major A
// (P^T, dS^T); Q, dO, K are consumed in place as MN-major B operands where the contraction runs over tile rows.
static constexpr int BWD_COMPUTE_WARPS = 16;
static constexpr int BWD_THREADS = 64 + 32 * BWD_COMPUTE_WARPS;  // 576
static constexpr uint32_t BWD_TMEM_COLS = 512;  // S [0,128) dP [128,256) dV [256,320) dK [320,384) dQ [384,448) [448,512)

struct BwdSmem {
  static constexpr int K = 0;
  static constexpr int V = K + TILE_BYTES;
  static constexpr int Q = V + TILE_BYTES;                 // [2]
  static constexpr int DO = Q + 2 * TILE_BYTES;            // [2]
  static constexpr int P = DO + 2 * TILE_BYTES;            // two 64-key blocks
  static constexpr int DS = P + 2 * TILE_BYTES;
  static constexpr int DQ = DS + 2 * TILE_BYTES;           // [2] fp32 [128][48] staging tiles of the dQ TMA reduce
  static constexpr int DQ_BYTES = BT * DH * 4;
  static constexpr int BAR = DQ + 2 * DQ_BYTES;
  // kv_full, qdo_full[2], qdo_empty[2], s_full, s_free, pds_full, dq_full[2], dq_free[2]
  static constexpr int NBAR = 12;
  static constexpr int TMEM_PTR = BAR + NBAR * 8;
  static constexpr int TOTAL = TMEM_PTR + 16;
};

__global__ void __launch_bounds__(BWD_THREADS, 1)
dilated_bwd_sm100_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ TensorMaps do_maps,
                         const __grid_constant__ TensorMaps dq_maps, const Sm100Params P, const float* __restrict__ lse, const float* __restrict__ delta_br,
                         float* __restrict__ dqkv, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  int oi = 0;
  while (oi + 1 < P.geo.nb && (int)blockIdx.x >= P.item_prefix[oi + 1]) ++oi;
  const int b = P.order[oi];
  const BranchGeom bg = P.geo.b[b];
  int local = blockIdx.x - P.item_prefix[oi];
  const int kt = local % P.tiles[b];
  local /= P.tiles[b];
  const int h = local % P.geo.H;
  const int s = local / P.geo.H;
  const int H = P.geo.H, N = P.geo.N, E = H * DH;
  const int off = (h * bg.r) / H;
  const int jseg = (s * bg.g) / bg.r;
  const int n_q = P.tiles[b];
  const int k0 = kt * BT;
  const int seg_end = min(N, (s + 1) * bg.g);
  const int slot_h = h - off * bg.hpb;

  const uint32_t bar_kv_full = sbase + BwdSmem::BAR + 0;
  const uint32_t bar_qdo_full = sbase + BwdSmem::BAR + 8;    // [2]
  const uint32_t bar_qdo_empty = sbase + BwdSmem::BAR + 24;  // [2]
  const uint32_t bar_s_full = sbase + BwdSmem::BAR + 40;
  const uint32_t bar_s_free = sbase + BwdSmem::BAR + 48;
  const uint32_t bar_pds_full = sbase + BwdSmem::BAR + 56;
  const uint32_t bar_dq_full = sbase + BwdSmem::BAR + 64;    // [2]
  const uint32_t bar_dq_free = sbase + BwdSmem::BAR + 80;    // [2]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + BwdSmem::TMEM_PTR);
  constexpr int NCOMP = 32 * BWD_COMPUTE_WARPS;

  if (threadIdx.x == 0) {
    mbar_init(bar_kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qdo_full + 8 * i, 1);
      mbar_init(bar_qdo_empty + 8 * i, 1);
      mbar_init(bar_dq_full + 8 * i, 1);
      mbar_init(bar_dq_free + 8 * i, NCOMP);
    }
    mbar_init(bar_s_full, 1);
    mbar_init(bar_s_free, NCOMP);
    mbar_init(bar_pds_full, NCOMP);
    fence_barrier_init();
    tma_prefetch_desc(&maps.m[b]);
    tma_prefetch_desc(&do_maps.m[b]);
    tma_prefetch_desc(&dq_maps.m[b]);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32((const void*)tmem_slot), BWD_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_s = tmem, tm_dp = tmem + 128, tm_dv = tmem + 256, tm_dk = tmem + 320, tm_dq = tmem + 384;

  if (warp == 0) {
    // ===== TMA producer ===============================================================================================
    if (lane == 0) {
      const void* map = &maps.m[b];
      const void* dmap = &do_maps.m[b];
      mbar_expect_tx(bar_kv_full, 2 * TILE_BYTES);
      tma_load_3d(sbase + BwdSmem::K, map, bar_kv_full, E + h * DH, off, jseg + k0);
      tma_load_3d(sbase + BwdSmem::V, map, bar_kv_full, 2 * E + h * DH, off, jseg + k0);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1, use = i >> 1;
        mbar_wait(bar_qdo_empty + 8 * st, (use & 1) ^ 1);
        mbar_expect_tx(bar_qdo_full + 8 * st, 2 * TILE_BYTES);
        tma_load_3d(sbase + BwdSmem::Q + st * TILE_BYTES, map, bar_qdo_full + 8 * st, h * DH, off, jseg + i * BT);
        tma_load_3d(sbase + BwdSmem::DO + st * TILE_BYTES, dmap, bar_qdo_full + 8 * st, h * DH, off, jseg + i * BT);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =================================================================================================
    constexpr uint32_t IDESC_S = umma_idesc_bf16(BT, BT, 0, 0);     // A K-major, B K-major
    constexpr uint32_t IDESC_T = umma_idesc_bf16(BT, DH, 1, 1);     // A MN-major (P^T, dS^T), B MN-major (dO, Q)
    constexpr uint32_t IDESC_DQ = umma_idesc_bf16(BT, DH, 0, 1);    // A K-major (dS), B MN-major (K)
    const uint32_t sK = sbase + BwdSmem::K, sV = sbase + BwdSmem::V, sP = sbase + BwdSmem::P, sDS = sbase + BwdSmem::DS;
    auto issue_s_dp = [&](int i) {
      if (lane == 0) {
        const uint32_t q = sbase + BwdSmem::Q + (i & 1) * TILE_BYTES, g = sbase + BwdSmem::DO + (i & 1) * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_ss(tm_s, umma_smem_desc(q + k * 32, 16, 1024), umma_smem_desc(sK + k * 32, 16, 1024), IDESC_S, k > 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_ss(tm_dp, umma_smem_desc(g + k * 32, 16, 1024), umma_smem_desc(sV + k * 32, 16, 1024), IDESC_S, k > 0);
        umma_commit(bar_s_full);
      }
      __syncwarp();
    };
    MT_TRACE_DECL
    mbar_wait(bar_kv_full, 0);
    mbar_wait(bar_qdo_full, 0);
    tc_fence_after();
    MT_TRACE(0);
    issue_s_dp(0);
    for (int i = 0; i < n_q; ++i) {
      mbar_wait(bar_s_free, i & 1);  // S_i / dP_i are in registers
      MT_TRACE(100 + i);
      if (i + 1 < n_q) {
        mbar_wait(bar_qdo_full + 8 * ((i + 1) & 1), ((i + 1) >> 1) & 1);
        tc_fence_after();
        MT_TRACE(200 + i);
        issue_s_dp(i + 1);
      }
      mbar_wait(bar_pds_full, i & 1);  // P_i, dS_i are in shared memory
      MT_TRACE(300 + i);
      if (i >= 2) mbar_wait(bar_dq_free + 8 * (i & 1), ((i - 2) >> 1) & 1);  // dQ tile i&1 has been drained
      tc_fence_after();
      MT_TRACE(400 + i);
      if (lane == 0) {
        const uint32_t q = sbase + BwdSmem::Q + (i & 1) * TILE_BYTES, g = sbase + BwdSmem::DO + (i & 1) * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)  // contraction over the 128 queries (tile rows): 16 rows = 2048 B per step
          umma_ss(tm_dv, umma_smem_desc(sP + k * 2048, TILE_BYTES, 1024), umma_smem_desc(g + k * 2048, TILE_BYTES, 1024),
                  IDESC_T, (i > 0) || (k > 0));
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)
          umma_ss(tm_dk, umma_smem_desc(sDS + k * 2048, TILE_BYTES, 1024), umma_smem_desc(q + k * 2048, TILE_BYTES, 1024),
                  IDESC_T, (i > 0) || (k > 0));
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)  // contraction over the 128 keys: dS K-major (two 64-key blocks)
          umma_ss(tm_dq + (i & 1) * 64, umma_smem_desc(sDS + (k >> 2) * TILE_BYTES + (k & 3) * 32, 16, 1024),
                  umma_smem_desc(sK + k * 2048, TILE_BYTES, 1024), IDESC_DQ, k > 0);
        umma_commit(bar_qdo_empty + 8 * (i & 1));
        umma_commit(bar_dq_full + 8 * (i & 1));
      }
      __syncwarp();
    }
    if (lane == 0) { MT_TRACE_DUMP("mma"); }
  } else {
    // ===== compute: thread = (query row, 32-key quarter) ===============================================================
    const int cw = warp - 2;
    const int lane_grp = warp & 3;               // TMEM lanes of this warp
    const int quarter = cw >> 2;                 // key columns [32*quarter, +32)
    const int row = lane_grp * 32 + lane;
    const uint32_t t_lane = (uint32_t)(lane_grp * 32) << 16;
    const int sw = row & 7;
    const int kvalid = bg.m - k0;
    const float LOG2E = 1.4426950408889634f;
    const float scale_log2 = P.scale_log2, sc = P.scale;
    uint8_t* p_row = smem + BwdSmem::P + (quarter >> 1) * TILE_BYTES + row * 128;
    uint8_t* ds_row = smem + BwdSmem::DS + (quarter >> 1) * TILE_BYTES + row * 128;

    // dQ tile of pair i -> global: TMEM -> fp32 staging tile in shared memory -> ONE TMA reduce-add (the element-wise
    // fp32 add happens in L2).  Per-thread red.global of the 128 x 48 tile caps the whole kernel at the LSU atomic
    // rate (~1.2 floats / clk / SM, measured: 5 200 clk per tile pair); the bulk reduce does not.
    const bool issuer = (cw == 0 && lane == 0);
    auto drain_dq = [&](int i) {
      mbar_wait(bar_dq_full + 8 * (i & 1), (i >> 1) & 1);
      tc_fence_after();
      float v[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) tmem_ld4(tm_dq + t_lane + (i & 1) * 64 + quarter * 12 + c * 4, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_dq_free + 8 * (i & 1));
      if (issuer) bulk_wait_group_read<1>();   // the reduce of pair i-2 has finished reading staging tile i&1
      named_bar_sync(1, NCOMP);
      float* stage = reinterpret_cast<float*>(smem + BwdSmem::DQ + (i & 1) * BwdSmem::DQ_BYTES) + row * DH + quarter * 12;
#pragma unroll
      for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(stage + c * 4) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
      fence_proxy_async_smem();
      named_bar_sync(1, NCOMP);
      if (issuer) {
        // rows of slots >= m (next segment) and of padded positions are exact zeros (P = 0 there); rows >= n_alloc clip
        tma_reduce_add_3d(&dq_maps.m[b], sbase + BwdSmem::DQ + (i & 1) * BwdSmem::DQ_BYTES, h * DH, off, jseg + i * BT);
        bulk_commit_group();
      }
    };

    // per-row statistics (merged lse, per-branch delta) of query tile i: strided 4-byte global loads with L2 latency,
    // so tile i+1's values are requested one iteration ahead and only consumed after a full tile of math
    auto load_stats = [&](int i, float& l_raw, float& d_raw) {
      const int slot = i * BT + row;
      const int pos = s * bg.g + off + slot * bg.r;
      const bool qok = i < n_q && slot < bg.m && pos < seg_end;
      l_raw = qok ? lse[(int64_t)pos * H + h] : INFINITY;   // +inf: P = exp(S - inf) = 0 for rows that do not exist
      d_raw = qok ? delta_br[bg.lse_off + (int64_t)pos * bg.hpb + slot_h] : 0.f;
    };
    float l_next, d_next;
    load_stats(0, l_next, d_next);
    MT_TRACE_DECL
    for (int i = 0; i < n_q; ++i) {
      const float l2 = l_next * LOG2E;
      const float de = d_next;
      load_stats(i + 1, l_next, d_next);
      MT_TRACE(1000 + i);
      mbar_wait(bar_s_full, i & 1);
      tc_fence_after();
      MT_TRACE(1100 + i);
      float sv[32], dp[32];
      tmem_ld32(tm_s + t_lane + quarter * 32, sv);
      tmem_ld32(tm_dp + t_lane + quarter * 32, dp);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_s_free);
      MT_TRACE(1200 + i);
      uint32_t pk[16], dk[16];
      const float nde = -de * sc;  // dS = p * (dp - delta) * scale = p * fma(dp, scale, -delta * scale)
      if (kvalid >= BT) {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float p0 = ex2(fmaf(sv[c], scale_log2, -l2));
          const float p1 = ex2(fmaf(sv[c + 1], scale_log2, -l2));
          pk[c >> 1] = pack_bf16(p0, p1);
          dk[c >> 1] = pack_bf16(p0 * fmaf(dp[c], sc, nde), p1 * fmaf(dp[c + 1], sc, nde));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const bool v0 = quarter * 32 + c < kvalid, v1 = quarter * 32 + c + 1 < kvalid;
          const float p0 = v0 ? ex2(fmaf(sv[c], scale_log2, -l2)) : 0.f;
          const float p1 = v1 ? ex2(fmaf(sv[c + 1], scale_log2, -l2)) : 0.f;
          pk[c >> 1] = pack_bf16(p0, p1);
          dk[c >> 1] = pack_bf16(p0 * fmaf(dp[c], sc, nde), p1 * fmaf(dp[c + 1], sc, nde));
        }
      }
      // P / dS buffers are free once the MMAs of pair i-1 have completed (that is what dq_full(i-1) tracks)
      MT_TRACE(1300 + i);
      if (i > 0) mbar_wait(bar_dq_full + 8 * ((i - 1) & 1), ((i - 1) >> 1) & 1);
      MT_TRACE(1400 + i);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = (((quarter & 1) * 4 + q) ^ sw) << 4;
        *reinterpret_cast<uint4*>(p_row + chunk) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        *reinterpret_cast<uint4*>(ds_row + chunk) = make_uint4(dk[4 * q], dk[4 * q + 1], dk[4 * q + 2], dk[4 * q + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_pds_full);
      MT_TRACE(1500 + i);
      if (i > 0) drain_dq(i - 1);
      MT_TRACE(1600 + i);
    }
    drain_dq(n_q - 1);
    if (cw == 0 && lane == 0) { MT_TRACE_DUMP("cmp"); }
    if (issuer) bulk_wait_group_all();
    // ---- dK / dV of this key tile: the last dq_full also covers the last dV / dK MMAs ---------------------------------
    {
      float a[3][4], c2[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        tmem_ld4(tm_dk + t_lane + quarter * 12 + c * 4, a[c]);
        tmem_ld4(tm_dv + t_lane + quarter * 12 + c * 4, c2[c]);
      }
      tmem_ld_wait();
      const int slot = k0 + row;
      const int pos = s * bg.g + off + slot * bg.r;
      if (slot < bg.m && pos < seg_end) {
        float* dst = dqkv + (int64_t)pos * (3 * E) + E + h * DH + quarter * 12;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          red_add_v4(dst + c * 4, a[c][0], a[c][1], a[c][2], a[c][3]);
          red_add_v4(dst + E + c * 4, c2[c][0], c2[c][1], c2[c][2], c2[c][3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, BWD_TMEM_COLS);
}

int dilated_attn_bwd_sm100(const mt_dilated_geometry* geom, const void* qkv, int64_t qkv_ld, int64_t n_alloc,
                           const void* dattn, const float* lse, const float* delta_br, float* dqkv, cudaStream_t st) {
  Sm100Params P;
  int rc = make_sm100_params(geom, &P);
  if (rc) return rc;
  MT_REQUIRE(n_alloc >= P.geo.N && n_alloc % 128 == 0, "dilated_attn_bwd: n_alloc must be a multiple of 128 >= n_tokens");
  MT_REQUIRE(qkv_ld % 8 == 0 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)dattn & 15) == 0 && ((uintptr_t)dqkv & 15) == 0,
             "dilated_attn_bwd: buffers must be 16-byte aligned");
  TensorMaps maps, do_maps, dq_maps;
  memset(&maps, 0, sizeof(maps));
  memset(&do_maps, 0, sizeof(do_maps));
  memset(&dq_maps, 0, sizeof(dq_maps));
  const int64_t E = (int64_t)P.geo.H * DH;
  for (int b = 0; b < P.geo.nb; ++b) {
    rc = encode_branch_map(&maps.m[b], qkv, qkv_ld, n_alloc, P.geo.b[b].r);
    if (rc) return rc;
    rc = encode_branch_map(&do_maps.m[b], dattn, E, n_alloc, P.geo.b[b].r);  // dattn: [n_alloc, 768], zero tail rows
    if (rc) return rc;
    rc = encode_branch_map_f32(&dq_maps.m[b], dqkv, 3 * E, n_alloc, P.geo.b[b].r);  // dqkv: [n_alloc, 2304] fp32
    if (rc) return rc;
  }
  int* flag = error_flag();
  MT_REQUIRE(flag != nullptr, "dilated_attn_bwd: cannot allocate the error flag");
  MT_CUDA(cudaFuncSetAttribute(dilated_bwd_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::TOTAL));
  dilated_bwd_sm100_kernel<<<P.item_prefix[P.geo.nb], BWD_THREADS, BwdSmem::TOTAL, st>>>(maps, do_maps, dq_maps, P, lse,
                                                                                       delta_br, dqkv, flag);
  return check_launch("dilated_bwd_sm100_kernel");
}

// =====================================================================================================================
// backward, second version: transposed formulation with the A operands in TMEM
// =====================================================================================================================
// The first version is bound by SHARED-MEMORY bandwidth (P and dS written to smem, re-read as A operands of three N = 48
// MMAs, dQ staged through smem: ~320 KB of smem traffic per tile pair against 128 B/clk).  Here the CTA (one key tile)
// computes the TRANSPOSED tiles, keys on the TMEM lanes, in half tiles of 64 queries:
//     S^T  = K  Q_h^T      dP^T = V dO_h^T        A = K / V resident in TMEM (copied once), B = Q / dO half tile
//     P^T  = exp(S^T scale - lse[q])              dS^T = P^T (dP^T - delta[q]) scale       (per-query stats from smem)
//     dV  += P^T dO_h      dK  += dS^T Q_h        A = P^T / dS^T as packed bf16 written back to TMEM over the very
//                                                 columns each thread has just read (tcgen05.st), B in place (MN-major)
//     dQ   = dS K  (once per 128 queries)         A = dS^T tile in smem read as an MN-major operand, B = K in place
// S^T / dP^T are double buffered in TMEM (2 x 128 columns), so the MMAs of half g+1 overlap the exponentials of half g.
// Shared-memory traffic per 128-query tile drops to ~150 KB; dQ goes out with red.global.add.v4.f32 from registers.
static constexpr int BWD2_THREADS = 64 + 512 + 128;  // TMA, MMA, 16 compute warps, 4 statistics / dQ-drain warps
struct Bwd2Smem {
  static constexpr int K = 0;
  static constexpr int V = K + TILE_BYTES;
  static constexpr int NQ = 3;                             // Q / dO stages (the TMA latency of a tile is ~2 tile times)
  static constexpr int Q = V + TILE_BYTES;                 // [NQ]
  static constexpr int DO = Q + NQ * TILE_BYTES;           // [NQ]
  static constexpr int DS = DO + NQ * TILE_BYTES;          // [2] dS^T tiles: [128 key rows][2 blocks of 64 queries]
  static constexpr int NSTAT = 8;                          // per-query stats ring (written 3 tiles ahead)
  static constexpr int STATS = DS + 4 * TILE_BYTES;        // [NSTAT][2][128] floats: lse * log2e, -delta * scale
  static constexpr int BAR = STATS + NSTAT * 2 * BT * 4;
  // kv_full, st_full[2], pt_full[2], dq_full, dq_free, kvt_full, all_done, qdo_full[NQ], qdo_empty[NQ], stats_full[NSTAT]
  static constexpr int NBAR = 9 + 2 * NQ + NSTAT;
  static constexpr int TMEM_PTR = BAR + NBAR * 8;
  static constexpr int TOTAL = TMEM_PTR + 16;
};

__global__ void __launch_bounds__(BWD2_THREADS, 1)
dilated_bwd2_sm100_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ TensorMaps do_maps,
                          const Sm100Params P, const float* __restrict__ lse, const float* __restrict__ delta_br,
                          float* __restrict__ dqkv, int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  // roles by warp id: compute 0-15, statistics / dQ drain 16-19, TMA 20, MMA 21 (the scheduler favours high warp ids:
  // the single-thread MMA issuer must never starve behind the math warps of its sub-partition)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_TMA = 20, W_MMA = 21, W_EPI = 16;
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  int oi = 0;
  while (oi + 1 < P.geo.nb && (int)blockIdx.x >= P.item_prefix[oi + 1]) ++oi;
  const int b = P.order[oi];
  const BranchGeom bg = P.geo.b[b];
  int local = blockIdx.x - P.item_prefix[oi];
  const int kt = local % P.tiles[b];
  local /= P.tiles[b];
  const int h = local % P.geo.H;
  const int s = local / P.geo.H;
  const int H = P.geo.H, N = P.geo.N, E = H * DH;
  const int off = (h * bg.r) / H;
  const int jseg = (s * bg.g) / bg.r;
  const int n_q = P.tiles[b];
  const int k0 = kt * BT;
  const int seg_end = min(N, (s + 1) * bg.g);
  const int slot_h = h - off * bg.hpb;

  constexpr int NQ = Bwd2Smem::NQ;
  const uint32_t bar_kv_full = sbase + Bwd2Smem::BAR + 0;
  const uint32_t bar_st_full = sbase + Bwd2Smem::BAR + 8;      // [2]
  const uint32_t bar_pt_full = sbase + Bwd2Smem::BAR + 24;     // [2]
  const uint32_t bar_dq_full = sbase + Bwd2Smem::BAR + 40;
  const uint32_t bar_dq_free = sbase + Bwd2Smem::BAR + 48;
  const uint32_t bar_kvt_full = sbase + Bwd2Smem::BAR + 56;
  const uint32_t bar_done = sbase + Bwd2Smem::BAR + 64;        // every MMA of the CTA has completed
  const uint32_t bar_qdo_full = sbase + Bwd2Smem::BAR + 72;    // [NQ]
  const uint32_t bar_qdo_empty = bar_qdo_full + 8 * NQ;        // [NQ]
  const uint32_t bar_stats_full = bar_qdo_empty + 8 * NQ;      // [NSTAT]
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Bwd2Smem::TMEM_PTR);
  constexpr int NCOMP = 512;

  if (threadIdx.x == 0) {
    mbar_init(bar_kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_st_full + 8 * i, 1);
      mbar_init(bar_pt_full + 8 * i, NCOMP / 2);   // one compute group per half-tile parity
    }
    for (int i = 0; i < NQ; ++i) {
      mbar_init(bar_qdo_full + 8 * i, 1);
      mbar_init(bar_qdo_empty + 8 * i, 1);
    }
    for (int i = 0; i < Bwd2Smem::NSTAT; ++i) mbar_init(bar_stats_full + 8 * i, 128);
    mbar_init(bar_dq_full, 1);
    mbar_init(bar_dq_free, 128);
    mbar_init(bar_kvt_full, 256);
    mbar_init(bar_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&maps.m[b]);
    tma_prefetch_desc(&do_maps.m[b]);
  }
  if (warp == W_MMA) {
    tmem_alloc(smem_u32((const void*)tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // columns: [0,64) S^T buf0 | [64,128) dP^T buf0 | [128,192) S^T buf1 | [192,256) dP^T buf1 | dV 256 | dK 320 | dQ 384
  //          K (bf16 pairs) 448..471 | V 480..503
  const uint32_t tm_dv = tmem + 256, tm_dk = tmem + 320, tm_dq = tmem + 384, tm_k = tmem + 448, tm_v = tmem + 480;

  if (warp == W_TMA) {
    // ===== TMA producer ===============================================================================================
    if (lane == 0) {
      const void* map = &maps.m[b];
      const void* dmap = &do_maps.m[b];
      mbar_expect_tx(bar_kv_full, 2 * TILE_BYTES);
      tma_load_3d(sbase + Bwd2Smem::K, map, bar_kv_full, E + h * DH, off, jseg + k0);
      tma_load_3d(sbase + Bwd2Smem::V, map, bar_kv_full, 2 * E + h * DH, off, jseg + k0);
      for (int i = 0; i < n_q; ++i) {
        const int st = i % NQ, use = i / NQ;
        mbar_wait(bar_qdo_empty + 8 * st, (use & 1) ^ 1);
        mbar_expect_tx(bar_qdo_full + 8 * st, 2 * TILE_BYTES);
        tma_load_3d(sbase + Bwd2Smem::Q + st * TILE_BYTES, map, bar_qdo_full + 8 * st, h * DH, off, jseg + i * BT);
        tma_load_3d(sbase + Bwd2Smem::DO + st * TILE_BYTES, dmap, bar_qdo_full + 8 * st, h * DH, off, jseg + i * BT);
      }
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer =================================================================================================
    constexpr uint32_t IDESC_ST = umma_idesc_bf16(BT, 64, 0, 0);    // A = K / V (TMEM), B = Q / dO half tile (K-major)
    constexpr uint32_t IDESC_TS = umma_idesc_bf16(BT, DH, 0, 1);    // A = P^T / dS^T (TMEM), B = dO / Q (MN-major)
    constexpr uint32_t IDESC_DQ = umma_idesc_bf16(BT, DH, 1, 1);    // A = dS^T tile (smem, MN-major), B = K (MN-major)
    const int n_half = 2 * n_q;
    // every descriptor is built once; inside the loops an MMA is one UTCHMMA plus a constant descriptor advance
    uint64_t qk_desc[NQ], dk_desc[NQ], qm_desc[NQ], dm_desc[NQ];   // Q / dO stages as K-major and as MN-major operands
#pragma unroll
    for (int st = 0; st < NQ; ++st) {
      qk_desc[st] = umma_smem_desc(sbase + Bwd2Smem::Q + st * TILE_BYTES, 16, 1024);
      dk_desc[st] = umma_smem_desc(sbase + Bwd2Smem::DO + st * TILE_BYTES, 16, 1024);
      qm_desc[st] = umma_smem_desc(sbase + Bwd2Smem::Q + st * TILE_BYTES, TILE_BYTES, 1024);
      dm_desc[st] = umma_smem_desc(sbase + Bwd2Smem::DO + st * TILE_BYTES, TILE_BYTES, 1024);
    }
    const uint64_t k_mn_desc = umma_smem_desc(sbase + Bwd2Smem::K, TILE_BYTES, 1024);
    const uint64_t ds_desc0 = umma_smem_desc(sbase + Bwd2Smem::DS, TILE_BYTES, 1024);
    const uint64_t ds_desc1 = umma_smem_desc(sbase + Bwd2Smem::DS + 2 * TILE_BYTES, TILE_BYTES, 1024);
    auto pick = [&](const uint64_t (&a)[NQ], int st) { return st == 0 ? a[0] : (st == 1 ? a[1] : a[2]); };
    static_assert(NQ == 3, "pick() is written for three stages");
    auto issue_st = [&](int g) {  // S^T and dP^T of half tile g into TMEM buffer g & 1
      if (elect_one()) {
        const int st = (g >> 1) % NQ, hh = g & 1;
        const uint64_t q = umma_desc_adv(pick(qk_desc, st), hh * 64 * 128), d = umma_desc_adv(pick(dk_desc, st), hh * 64 * 128);
        const uint32_t ts = tmem + (g & 1) * 128, td = ts + 64;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_ts(ts, tm_k + k * 8, umma_desc_adv(q, k * 32), IDESC_ST, k > 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_ts(td, tm_v + k * 8, umma_desc_adv(d, k * 32), IDESC_ST, k > 0);
        umma_commit(bar_st_full + 8 * (g & 1));
      }
      __syncwarp();
    };
    MT_TRACE_DECL
    mbar_wait(bar_kvt_full, 0);       // K and V are in TMEM
    mbar_wait(bar_qdo_full, 0);
    tc_fence_after();
    MT_TRACE(0);
    issue_st(0);
    issue_st(1);
    for (int g = 0; g < n_half; ++g) {
      const int i = g >> 1, hh = g & 1;
      mbar_wait(bar_pt_full + 8 * (g & 1), (g >> 1) & 1);   // P^T, dS^T of half g are in TMEM (+ dS^T half in smem)
      MT_TRACE(100 + g);
      tc_fence_after();
      if (elect_one()) {
        const int st = i % NQ;
        const uint64_t q = umma_desc_adv(pick(qm_desc, st), hh * 64 * 128), d = umma_desc_adv(pick(dm_desc, st), hh * 64 * 128);
        const uint32_t ts = tmem + (g & 1) * 128, td = ts + 64;
        // packed P^T / dS^T: query pair (2c, 2c+1) of 16-query quarter qq sits in column 16*qq + c of its buffer;
        // 64 queries = 4 k-steps of 16 (rows of the dO / Q half tile: 2048 B each)
        umma_ts(tm_dv, ts, d, IDESC_TS, g > 0);
#pragma unroll
        for (int k = 1; k < 4; ++k) umma_ts(tm_dv, ts + k * 16, umma_desc_adv(d, k * 2048), IDESC_TS, 1);
        umma_ts(tm_dk, td, q, IDESC_TS, g > 0);
#pragma unroll
        for (int k = 1; k < 4; ++k) umma_ts(tm_dk, td + k * 16, umma_desc_adv(q, k * 2048), IDESC_TS, 1);
        if (hh == 1) umma_commit(bar_qdo_empty + 8 * st);   // last readers of this Q / dO stage (dQ reads K, dS)
      }
      __syncwarp();
      if (g + 2 < n_half) {          // the buffer is free once the MMAs above have consumed it (the pipe is in order)
        if (((g + 2) & 1) == 0) mbar_wait(bar_qdo_full + 8 * (((g + 2) >> 1) % NQ), (((g + 2) >> 1) / NQ) & 1);
        tc_fence_after();
        MT_TRACE(200 + g);
        issue_st(g + 2);
      }
      if (hh == 1) {                 // both halves of query tile i are done: dQ_i = dS_i K
        if (i > 0) mbar_wait(bar_dq_free, (i - 1) & 1);
        tc_fence_after();
        MT_TRACE(300 + g);
        if (elect_one()) {
          const uint64_t ds = (i & 1) ? ds_desc1 : ds_desc0;
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)   // contraction over the 128 keys = rows of the dS^T tile and of K
            umma_ss(tm_dq, umma_desc_adv(ds, k * 2048), umma_desc_adv(k_mn_desc, k * 2048), IDESC_DQ, k > 0);
          umma_commit(bar_dq_full);
        }
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
    if (lane == 0) { MT_TRACE_DUMP("mma2"); }
  } else if (warp < W_EPI) {
    // ===== compute: two groups of 8 warps ping-pong over the half tiles (group 0: even halves, group 1: odd halves), so
    // one group's exponentials overlap the other group's TMEM / shared-memory traffic and barrier waits.
    // thread = (key row, 32 of the 64 queries of its half tile, processed as two 16-query chunks)
    const int cw = warp;
    const int lane_grp = warp & 3;               // TMEM lanes of this warp
    const int qq = cw >> 2;                      // used for the K / V copy and the final dK / dV columns
    const int grp = cw >> 3;                     // which half of every query tile this warp works on
    const int sub = (cw >> 2) & 1;               // which 32 queries of the half
    const int row = lane_grp * 32 + lane;        // key slot k0 + row
    const uint32_t t_lane = (uint32_t)(lane_grp * 32) << 16;
    const int sw = row & 7;
    const bool key_ok = (k0 + row) < bg.m;       // rows past m belong to the next segment: P = dS = 0
    const float LOG2E = 1.4426950408889634f;
    const float scale_log2 = P.scale_log2, sc = P.scale;
    float* stats = reinterpret_cast<float*>(smem + Bwd2Smem::STATS);

    // ---- K and V tiles -> TMEM (A operands of S^T / dP^T for the whole CTA): quarter 0 copies K, quarter 1 copies V
    mbar_wait(bar_kv_full, 0);
    if (qq < 2) {
      const uint8_t* src = smem + (qq == 0 ? Bwd2Smem::K : Bwd2Smem::V) + row * 128;
      uint32_t w[24];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + ((c ^ sw) << 4));
        w[4 * c] = u.x; w[4 * c + 1] = u.y; w[4 * c + 2] = u.z; w[4 * c + 3] = u.w;
      }
      const uint32_t dst = (qq == 0 ? tm_k : tm_v) + t_lane;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t r8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) r8[e] = w[8 * c + e];
        tmem_st8(dst + c * 8, r8);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_kvt_full);
    }

    MT_TRACE_DECL
    for (int i = 0; i < n_q; ++i) {
      mbar_wait(bar_stats_full + 8 * (i % Bwd2Smem::NSTAT), (i / Bwd2Smem::NSTAT) & 1);
      const float* st_l = stats + (i % Bwd2Smem::NSTAT) * 2 * BT;
      uint8_t* ds_tile = smem + Bwd2Smem::DS + (i & 1) * 2 * TILE_BYTES;
      {
        const int hh = grp;
        const int g = 2 * i + hh;
        MT_TRACE(1000 + g);
        mbar_wait(bar_st_full + 8 * (g & 1), (g >> 1) & 1);
        tc_fence_after();
        MT_TRACE(1100 + g);
        uint8_t* drow = ds_tile + hh * TILE_BYTES + row * 128;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int q16 = sub * 2 + t;           // 16-query chunk of the half tile
          const uint32_t ts = tmem + (g & 1) * 128 + t_lane + q16 * 16, td = ts + 64;
          float sv[16], dp[16];
          tmem_ld16(ts, sv);
          tmem_ld16(td, dp);
          float l2[16], nde[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 a = *reinterpret_cast<const float4*>(st_l + hh * 64 + q16 * 16 + c * 4);
            const float4 e = *reinterpret_cast<const float4*>(st_l + BT + hh * 64 + q16 * 16 + c * 4);
            l2[4 * c] = a.x; l2[4 * c + 1] = a.y; l2[4 * c + 2] = a.z; l2[4 * c + 3] = a.w;
            nde[4 * c] = e.x; nde[4 * c + 1] = e.y; nde[4 * c + 2] = e.z; nde[4 * c + 3] = e.w;
          }
          tmem_ld_wait();
          uint32_t pk[8], dk[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float p0 = ex2(fmaf(sv[c], scale_log2, -l2[c]));
            float p1 = ex2(fmaf(sv[c + 1], scale_log2, -l2[c + 1]));
            if (!key_ok) p0 = p1 = 0.f;
            pk[c >> 1] = pack_bf16(p0, p1);
            dk[c >> 1] = pack_bf16(p0 * fmaf(dp[c], sc, nde[c]), p1 * fmaf(dp[c + 1], sc, nde[c + 1]));
          }
          // packed results over the first 8 of the 16 columns this thread has just read (nobody else touches them)
          tmem_st8(ts, pk);
          tmem_st8(td, dk);
          // dS^T half tile for dQ = dS K: [key row][64 queries] block hh, 16 queries = 2 swizzled 16-byte chunks
          *reinterpret_cast<uint4*>(drow + (((2 * q16) ^ sw) << 4)) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
          *reinterpret_cast<uint4*>(drow + (((2 * q16 + 1) ^ sw) << 4)) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
        }
        MT_TRACE(1300 + g);
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(bar_pt_full + 8 * (g & 1));
        MT_TRACE(1400 + g);
      }
    }
#ifdef MT_DEBUG_TRACE
    if (lane == 0 && tron_) for (int q_ = 0; q_ + 1 < trn_; q_ += 2) printf("cmp%d %lld %lld\n", cw, tr_[q_], tr_[q_ + 1] & 0xffffffffll);
#endif
    mbar_wait(bar_done, 0);                      // (a parity wait on dq_full would alias: these warps run 2 tiles ahead)
    tc_fence_after();
    // ---- dK / dV of this key tile (the last dq_full commit also covers the last dV / dK MMAs) --------------------------
    {
      float a[3][4], c2[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        tmem_ld4(tm_dk + t_lane + qq * 12 + c * 4, a[c]);
        tmem_ld4(tm_dv + t_lane + qq * 12 + c * 4, c2[c]);
      }
      tmem_ld_wait();
      const int slot = k0 + row;
      const int pos = s * bg.g + off + slot * bg.r;
      if (slot < bg.m && pos < seg_end) {
        float* dst = dqkv + (int64_t)pos * (3 * E) + E + h * DH + qq * 12;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          red_add_v4(dst + c * 4, a[c][0], a[c][1], a[c][2], a[c][3]);
          red_add_v4(dst + E + c * 4, c2[c][0], c2[c][1], c2[c][2], c2[c][3]);
        }
      }
    }
  } else {
    // ===== statistics + dQ drain (4 warps, one query row per thread) ==================================================
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;
    const uint32_t t_lane = (uint32_t)(lane_grp * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const float sc = P.scale;
    float* stats = reinterpret_cast<float*>(smem + Bwd2Smem::STATS);
    // per-query statistics of tile i -> smem ring (lse * log2e, -delta * scale): strided 4-byte global loads, issued
    // two tiles ahead of their use by the compute warps
    auto load_stats = [&](int i, float& l, float& d) {
      const int slot = i * BT + row;
      const int pos = s * bg.g + off + slot * bg.r;
      const bool qok = i < n_q && slot < bg.m && pos < seg_end;
      l = qok ? lse[(int64_t)pos * H + h] * LOG2E : INFINITY;   // +inf -> P = 0 for rows that do not exist
      d = qok ? -delta_br[bg.lse_off + (int64_t)pos * bg.hpb + slot_h] * sc : 0.f;
    };
    auto store_stats = [&](int i, float l, float d) {
      if (i < n_q) {
        float* dst = stats + (i % Bwd2Smem::NSTAT) * 2 * BT;
        dst[row] = l;
        dst[BT + row] = d;
        mbar_arrive(bar_stats_full + 8 * (i % Bwd2Smem::NSTAT));
      }
    };
    float l_n, d_n;
    for (int i = 0; i < 3; ++i) {
      load_stats(i, l_n, d_n);
      store_stats(i, l_n, d_n);
    }
    for (int i = 0; i < n_q; ++i) {
      load_stats(i + 3, l_n, d_n);   // requested now, consumed after the drain below (hides the L2 latency)
      // dQ of query tile i: TMEM -> registers -> red.global.add.v4.f32 (48 columns of this thread's query row)
      mbar_wait(bar_dq_full, i & 1);
      tc_fence_after();
      float v[3][16];
#pragma unroll
      for (int c = 0; c < 3; ++c) tmem_ld16(tm_dq + t_lane + c * 16, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_dq_free);
      const int slot = i * BT + row;
      const int pos = s * bg.g + off + slot * bg.r;
      if (slot < bg.m && pos < seg_end) {
        float* dst = dqkv + (int64_t)pos * (3 * E) + h * DH;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
          for (int e = 0; e < 16; e += 4) red_add_v4(dst + c * 16 + e, v[c][e], v[c][e + 1], v[c][e + 2], v[c][e + 3]);
      }
      store_stats(i + 3, l_n, d_n);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

int dilated_attn_bwd2_sm100(const mt_dilated_geometry* geom, const void* qkv, int64_t qkv_ld, int64_t n_alloc,
                            const void* dattn, const float* lse, const float* delta_br, float* dqkv, cudaStream_t st) {
  Sm100Params P;
  int rc = make_sm100_params(geom, &P);
  if (rc) return rc;
  MT_REQUIRE(n_alloc >= P.geo.N && n_alloc % 128 == 0, "dilated_attn_bwd: n_alloc must be a multiple of 128 >= n_tokens");
  MT_REQUIRE(qkv_ld % 8 == 0 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)dattn & 15) == 0 && ((uintptr_t)dqkv & 15) == 0,
             "dilated_attn_bwd: buffers must be 16-byte aligned");
  TensorMaps maps, do_maps;
  memset(&maps, 0, sizeof(maps));
  memset(&do_maps, 0, sizeof(do_maps));
  const int64_t E = (int64_t)P.geo.H * DH;
  for (int b = 0; b < P.geo.nb; ++b) {
    rc = encode_branch_map(&maps.m[b], qkv, qkv_ld, n_alloc, P.geo.b[b].r);
    if (rc) return rc;
    rc = encode_branch_map(&do_maps.m[b], dattn, E, n_alloc, P.geo.b[b].r);
    if (rc) return rc;
  }
  int* flag = error_flag();
  MT_REQUIRE(flag != nullptr, "dilated_attn_bwd: cannot allocate the error flag");
  MT_CUDA(cudaFuncSetAttribute(dilated_bwd2_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd2Smem::TOTAL));
  dilated_bwd2_sm100_kernel<<<P.item_prefix[P.geo.nb], BWD2_THREADS, Bwd2Smem::TOTAL, st>>>(maps, do_maps, P, lse,
                                                                                          delta_br, dqkv, flag);
  return check_launch("dilated_bwd2_sm100_kernel");
}

// =====================================================================================================================
// backward, third version: the per-query statistics ride on the tensor cores
// =====================================================================================================================
// Same transposed formulation as version 2 (keys on the TMEM lanes, K / V resident in TMEM, P^T / dS^T written back to
// TMEM as the A operands of dV / dK).  What changed, each item measured on B200 (tools/ubench, MT_DEBUG_TRACE):
//   * the 64-column (128-byte) tile rows only carry 48 head columns, so columns 48..51 are used as an AUGMENTED
//     contraction: a dedicated warp overwrites them in every Q tile with the 3-way bf16 split of -lse/scale (and a
//     -32768 mask column) and in every dO tile with the split of -delta; K' / V' in TMEM carry ones there.  The MMAs
//     then deliver S^T - lse/scale and dP^T - delta directly (fp32 accumulation of exact 1 x bf16 products): no
//     per-query statistics ring in shared memory, no broadcast loads, no subtraction per element.  Key rows past the
//     segment's m get a one in the mask column (S = -32768 -> P = 0), queries that do not exist get -32768 as their lse.
//   * all 16 compute warps work on the SAME 64-query half tile (16 scores per thread) while the MMAs of the other
//     half run, instead of two groups of 8 warps ping-ponging over 32 scores per thread.
//   * the MMA issue loop is unrolled over stages x halves: every descriptor is a base plus a compile-time constant
//     (the single issuing thread shares its scheduler with busy warps; its instructions are critical-path latency).
//   * dQ, dK and dV leave through swizzled fp32 staging tiles and TMA reduce-adds (cp.reduce.async.bulk.tensor)
//     instead of red.global from registers, which costs one L1 request per row and instruction (~1500 LSU cycles per
//     dQ tile).  fp32 reduce-adds saturate at ~5.3 TB/s chip-wide (tools/ubench/red_rate.cu): ~1300 cycles per dQ tile
//     and SM, so the staging tile is a buffer of its own and nobody but the drain warps ever waits for it.
//   * separate warps for the TMA issue, the statistics columns and the dQ drain; the first loads are issued before
//     the CTA-wide setup barrier.
// scale is applied to dQ and dK when they leave TMEM (dS is kept as P (dP - delta)).
static constexpr int BWD3_THREADS = 32 * 23;   // 16 compute, 4 dQ drain, TMA, statistics, MMA
struct Bwd3Smem {
  static constexpr int NQ = 3;
  static constexpr int K = 0;
  static constexpr int V = K + TILE_BYTES;
  static constexpr int Q = V + TILE_BYTES;                 // [NQ]
  static constexpr int DO = Q + NQ * TILE_BYTES;           // [NQ]
  static constexpr int DS = DO + NQ * TILE_BYTES;          // [2] dS^T tiles: [128 key rows][2 blocks of 64 queries]
  static constexpr int STG = DS + 4 * TILE_BYTES;          // fp32 staging tile of the dQ reduce-add (24 KB)
  static constexpr int BAR = STG + BT * DH * 4;
  // kv_full, kvt_full, st_full[2], pt_full[2], dq_full, dq_free, done, qdo_full[NQ], qdo_empty[NQ], aug_full[NQ]
  static constexpr int NBAR = 9 + 3 * NQ;
  static constexpr int TMEM_PTR = BAR + NBAR * 8;
  static constexpr int TOTAL = TMEM_PTR + 16;
  // fp32 staging of a [128][48] gradient tile: columns 0..31 as [128][32] with the 128 B
  // swizzle, columns 32..47 as [128][16] with the 64 B swizzle (row pitch = swizzle span: one-row-per-thread stores are
  // bank-conflict free); each part is the box of one TMA reduce-add
  static constexpr int STG16 = BT * 128;
};
static_assert(Bwd3Smem::TOTAL <= 232448, "shared memory of the backward kernel");

__device__ __forceinline__ void split3_bf16(float a, uint32_t& w0, uint32_t& w1, float fourth) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(a);
  const float r1 = a - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);
  w0 = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(mid) << 16);
  w1 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(r2)) |
       ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(fourth)) << 16);
}

// 16-byte chunk c (4 floats, columns 4c..4c+3 of 48) of row `row` of a staged fp32 gradient tile
__device__ __forceinline__ void stage_chunk(uint8_t* stg, int row, int c, float4 v) {
  uint8_t* dst = c < 8 ? stg + row * 128 + ((c ^ (row & 7)) << 4)
                       : stg + Bwd3Smem::STG16 + row * 64 + (((c - 8) ^ ((row >> 1) & 3)) << 4);
  *reinterpret_cast<float4*>(dst) = v;
}

__global__ void __launch_bounds__(BWD3_THREADS, 1)
dilated_bwd3_sm100_kernel(const __grid_constant__ TensorMaps maps, const __grid_constant__ TensorMaps do_maps,
                          const __grid_constant__ TensorMaps dq32_maps, const __grid_constant__ TensorMaps dq16_maps,
                          const Sm100Params P, const float* __restrict__ lse, const float* __restrict__ delta_br,
                          int* __restrict__ err_flag) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  // roles by warp id: compute 0-15, dQ drain 16-19, TMA 20, statistics columns 21, MMA 22 (last: the scheduler
  // favours high warp ids and the single-thread issuer must never starve)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int W_DRAIN = 16, W_TMA = 20, W_AUG = 21, W_MMA = 22;
  if ((sbase & 1023u) != 0) {
    if (threadIdx.x == 0) atomicExch(err_flag, 1);
    return;
  }
  int oi = 0;
  while (oi + 1 < P.geo.nb && (int)blockIdx.x >= P.item_prefix[oi + 1]) ++oi;
  const int b = P.order[oi];
  const BranchGeom bg = P.geo.b[b];
  int local = blockIdx.x - P.item_prefix[oi];
  const int kt = local % P.tiles[b];
  local /= P.tiles[b];
  const int h = local % P.geo.H;
  const int s = local / P.geo.H;
  const int H = P.geo.H, N = P.geo.N, E = H * DH;
  const int off = (h * bg.r) / H;
  const int jseg = (s * bg.g) / bg.r;
  const int k0 = kt * BT;
  const int seg_end = min(N, (s + 1) * bg.g);
  const int slot_h = h - off * bg.hpb;
  // Whole tiles of zero padding (positions >= N in the last segment) are skipped: a zero key tile has K = V = 0, so it
  // adds nothing to dQ and its own dK / dV rows are scratch; a padding query tile has P = 0.
  const int seg_lo = s * bg.g + off;
  const int c_real = seg_end > seg_lo ? (seg_end - seg_lo + bg.r - 1) / bg.r : 0;   // real slots of this (segment, head)
  if (k0 >= c_real) return;
  const int n_q = min(P.tiles[b], (c_real + BT - 1) / BT);

  constexpr int NQ = Bwd3Smem::NQ;
  const uint32_t bar_kv_full = sbase + Bwd3Smem::BAR + 0;
  const uint32_t bar_kvt_full = sbase + Bwd3Smem::BAR + 8;
  const uint32_t bar_st_full = sbase + Bwd3Smem::BAR + 16;     // [2]
  const uint32_t bar_pt_full = sbase + Bwd3Smem::BAR + 32;     // [2]
  const uint32_t bar_dq_full = sbase + Bwd3Smem::BAR + 48;
  const uint32_t bar_dq_free = sbase + Bwd3Smem::BAR + 56;
  const uint32_t bar_done = sbase + Bwd3Smem::BAR + 64;        // every MMA of the CTA has completed
  const uint32_t bar_qdo_full = sbase + Bwd3Smem::BAR + 72;    // [NQ] TMA landed
  const uint32_t bar_qdo_empty = bar_qdo_full + 8 * NQ;        // [NQ]
  const uint32_t bar_aug_full = bar_qdo_empty + 8 * NQ;        // [NQ] statistics columns written into the stage
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Bwd3Smem::TMEM_PTR);
  constexpr int NCOMP = 512;

  if (warp == W_TMA && lane == 0) {
    // barrier setup and the first loads by the producer itself, ahead of the CTA-wide sync: the first tiles are on
    // the critical path of every CTA (one CTA per SM, nothing overlaps its prologue)
    mbar_init(bar_kv_full, 1);
    mbar_init(bar_kvt_full, 256);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_st_full + 8 * i, 1);
      mbar_init(bar_pt_full + 8 * i, NCOMP);
    }
    for (int i = 0; i < NQ; ++i) {
      mbar_init(bar_qdo_full + 8 * i, 1);
      mbar_init(bar_qdo_empty + 8 * i, 1);
      mbar_init(bar_aug_full + 8 * i, 32);
    }
    mbar_init(bar_dq_full, 1);
    mbar_init(bar_dq_free, 128);
    mbar_init(bar_done, 1);
    fence_barrier_init();
    const void* map = &maps.m[b];
    const void* dmap = &do_maps.m[b];
    mbar_expect_tx(bar_kv_full, 2 * TILE_BYTES);
    tma_load_3d(sbase + Bwd3Smem::K, map, bar_kv_full, E + h * DH, off, jseg + k0);
    tma_load_3d(sbase + Bwd3Smem::V, map, bar_kv_full, 2 * E + h * DH, off, jseg + k0);
    for (int i = 0; i < NQ && i < n_q; ++i) {
      mbar_expect_tx(bar_qdo_full + 8 * i, 2 * TILE_BYTES);
      tma_load_3d(sbase + Bwd3Smem::Q + i * TILE_BYTES, map, bar_qdo_full + 8 * i, h * DH, off, jseg + i * BT);
      tma_load_3d(sbase + Bwd3Smem::DO + i * TILE_BYTES, dmap, bar_qdo_full + 8 * i, h * DH, off, jseg + i * BT);
    }
    tma_prefetch_desc(&dq32_maps.m[b]);
    tma_prefetch_desc(&dq16_maps.m[b]);
  }
  if (warp == W_MMA) {
    tmem_alloc(smem_u32((const void*)tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // columns: [0,64) S^T buf0 | [64,128) dP^T buf0 | [128,192) S^T buf1 | [192,256) dP^T buf1 | dV 256 | dK 320 | dQ 384
  //          K' (64 bf16 = 32 columns) 448 | V' 480
  const uint32_t tm_dv = tmem + 256, tm_dk = tmem + 320, tm_dq = tmem + 384, tm_k = tmem + 448, tm_v = tmem + 480;

  if (warp == W_TMA) {
    // ===== TMA producer (tiles >= NQ; the first NQ were issued above) ================================================
    if (lane == 0) {
      const void* map = &maps.m[b];
      const void* dmap = &do_maps.m[b];
      for (int i = NQ; i < n_q; ++i) {
        const int st = i % NQ, use = i / NQ;
        mbar_wait(bar_qdo_empty + 8 * st, (use & 1) ^ 1);
        mbar_expect_tx(bar_qdo_full + 8 * st, 2 * TILE_BYTES);
        tma_load_3d(sbase + Bwd3Smem::Q + st * TILE_BYTES, map, bar_qdo_full + 8 * st, h * DH, off, jseg + i * BT);
        tma_load_3d(sbase + Bwd3Smem::DO + st * TILE_BYTES, dmap, bar_qdo_full + 8 * st, h * DH, off, jseg + i * BT);
      }
    }
  } else if (warp == W_AUG) {
    // ===== statistics columns: lane owns rows lane + 32 j of every Q / dO tile ==========================================
    const float inv_sc = 1.f / P.scale;
    float lr[4], dr[4], ln[4], dn[4];
    // raw per-query statistics of tile i: strided 4-byte loads issued one tile ahead; consumed (and masked) at use
    auto load_raw = [&](int i, float (&l)[4], float (&d)[4]) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int slot = i * BT + lane + 32 * j;
        const int pos = s * bg.g + off + slot * bg.r;
        const bool qok = i < n_q && slot < bg.m && pos < seg_end;
        const int pc = qok ? pos : 0;
        l[j] = lse[(int64_t)pc * H + h];
        d[j] = delta_br[bg.lse_off + (int64_t)pc * bg.hpb + slot_h];
      }
    };
    load_raw(0, lr, dr);
    for (int i = 0; i < n_q; ++i) {
      load_raw(i + 1, ln, dn);
      const int st = i % NQ;
      mbar_wait(bar_qdo_full + 8 * st, (i / NQ) & 1);      // the TMA has written the stage
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = lane + 32 * j;
        const int slot = i * BT + row;
        const int pos = s * bg.g + off + slot * bg.r;
        const bool qok = slot < bg.m && pos < seg_end;
        uint32_t w0, w1;
        // columns 48..55 (16-byte chunk 6) of the row: [hi, mid, lo, mask, 0, 0, 0, 0]
        split3_bf16(qok ? -lr[j] * inv_sc : -32768.f, w0, w1, -32768.f);   // -32768 -> P = 0 for rows that do not exist
        *reinterpret_cast<uint4*>(smem + Bwd3Smem::Q + st * TILE_BYTES + row * 128 + ((6 ^ (row & 7)) << 4)) =
            make_uint4(w0, w1, 0u, 0u);
        split3_bf16(qok ? -dr[j] : 0.f, w0, w1, 0.f);
        *reinterpret_cast<uint4*>(smem + Bwd3Smem::DO + st * TILE_BYTES + row * 128 + ((6 ^ (row & 7)) << 4)) =
            make_uint4(w0, w1, 0u, 0u);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_aug_full + 8 * st);
#pragma unroll
      for (int j = 0; j < 4; ++j) { lr[j] = ln[j]; dr[j] = dn[j]; }
    }
  } else if (warp == W_MMA) {
    // ===== MMA issuer =================================================================================================
    constexpr uint32_t IDESC_ST = umma_idesc_bf16(BT, 64, 0, 0);    // A = K' / V' (TMEM), B = Q' / dO' half tile (K-major)
    constexpr uint32_t IDESC_TS = umma_idesc_bf16(BT, DH, 0, 1);    // A = P^T / dS^T (TMEM), B = dO / Q (MN-major)
    constexpr uint32_t IDESC_DQ = umma_idesc_bf16(BT, DH, 1, 1);    // A = dS^T tile (smem, MN-major), B = K (MN-major)
    const int n_half = 2 * n_q;
    const uint64_t qk_desc = umma_smem_desc(sbase + Bwd3Smem::Q, 16, 1024);          // K-major views (stage 0)
    const uint64_t dk_desc = umma_smem_desc(sbase + Bwd3Smem::DO, 16, 1024);
    const uint64_t qm_desc = umma_smem_desc(sbase + Bwd3Smem::Q, TILE_BYTES, 1024);  // MN-major views (stage 0)
    const uint64_t dm_desc = umma_smem_desc(sbase + Bwd3Smem::DO, TILE_BYTES, 1024);
    const uint64_t k_mn_desc = umma_smem_desc(sbase + Bwd3Smem::K, TILE_BYTES, 1024);
    const uint64_t ds_desc0 = umma_smem_desc(sbase + Bwd3Smem::DS, TILE_BYTES, 1024);
    const uint64_t ds_desc1 = umma_smem_desc(sbase + Bwd3Smem::DS + 2 * TILE_BYTES, TILE_BYTES, 1024);
    auto issue_st = [&](auto U) {  // S'^T and dP'^T of a half tile (stage U / 2, half U & 1) into TMEM buffer U & 1
      constexpr int u = decltype(U)::value;
      if (elect_one()) {
        constexpr uint32_t so = (uint32_t)((u >> 1) * TILE_BYTES + (u & 1) * 64 * 128);
        const uint32_t ts = tmem + (u & 1) * 128, td = ts + 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(ts, tm_k + k * 8, umma_desc_adv(qk_desc, so + k * 32), IDESC_ST, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(td, tm_v + k * 8, umma_desc_adv(dk_desc, so + k * 32), IDESC_ST, k > 0);
        umma_commit(bar_st_full + 8 * (u & 1));
      }
      __syncwarp();
    };
    MT_TRACE_DECL
    mbar_wait(bar_kvt_full, 0);       // K' and V' are in TMEM
    mbar_wait(bar_qdo_full, 0);
    mbar_wait(bar_aug_full, 0);
    tc_fence_after();
    MT_TRACE(0);
    issue_st(std::integral_constant<int, 0>{});
    issue_st(std::integral_constant<int, 1>{});
    // one half tile: t = trip of the unrolled loop (tile i = NQ * t + U / 2)
    auto half = [&](int g, int t, auto U) {
      constexpr int u = decltype(U)::value;
      constexpr int st = u >> 1, hh = u & 1;
      const int i = g >> 1;
      mbar_wait(bar_pt_full + 8 * hh, i & 1);   // P^T, dS^T of half g are in TMEM (+ dS^T half in smem)
      tc_fence_after();
      MT_TRACE(100 + g);
      if (elect_one()) {
        constexpr uint32_t so = (uint32_t)(st * TILE_BYTES + hh * 64 * 128);
        const uint32_t ts = tmem + hh * 128, td = ts + 64;
        // packed P^T / dS^T: query pair (2c, 2c+1) of 16-query quarter qq sits in column 16*qq + c of its buffer;
        // 64 queries = 4 k-steps of 16 (rows of the dO / Q half tile: 2048 B each)
        umma_ts(tm_dv, ts, umma_desc_adv(dm_desc, so), IDESC_TS, g > 0);
#pragma unroll
        for (int k = 1; k < 4; ++k) umma_ts(tm_dv, ts + k * 16, umma_desc_adv(dm_desc, so + k * 2048), IDESC_TS, 1);
        umma_ts(tm_dk, td, umma_desc_adv(qm_desc, so), IDESC_TS, g > 0);
#pragma unroll
        for (int k = 1; k < 4; ++k) umma_ts(tm_dk, td + k * 16, umma_desc_adv(qm_desc, so + k * 2048), IDESC_TS, 1);
        if (hh == 1) umma_commit(bar_qdo_empty + 8 * st);   // last readers of this Q / dO stage (dQ reads K, dS)
      }
      __syncwarp();
      if (g + 2 < n_half) {          // the buffer is free once the MMAs above have consumed it (the pipe is in order)
        constexpr int st2 = (st + 1) % NQ;                  // stage of tile i + 1
        if (hh == 0) {
          const uint32_t par = (uint32_t)(t + (st2 == 0 ? 1 : 0)) & 1u;   // (i + 1) / NQ
          mbar_wait(bar_qdo_full + 8 * st2, par);
          mbar_wait(bar_aug_full + 8 * st2, par);
        }
        tc_fence_after();
        MT_TRACE(200 + g);
        issue_st(std::integral_constant<int, 2 * st2 + hh>{});
      }
      if (hh == 1) {                 // both halves of query tile i are done: dQ_i = dS_i K
        if (i > 0) mbar_wait(bar_dq_free, (i - 1) & 1);
        tc_fence_after();
        MT_TRACE(300 + g);
        if (elect_one()) {
          const uint64_t ds = (i & 1) ? ds_desc1 : ds_desc0;
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)   // contraction over the 128 keys = rows of the dS^T tile and of K
            umma_ss(tm_dq, umma_desc_adv(ds, k * 2048), umma_desc_adv(k_mn_desc, k * 2048), IDESC_DQ, k > 0);
          umma_commit(bar_dq_full);
        }
        __syncwarp();
      }
    };
    auto trip = [&](int g0, int t, auto... Us) {
      ((g0 + decltype(Us)::value < n_half ? (half(g0 + decltype(Us)::value, t, Us), 0) : 0), ...);
    };
    static_assert(NQ == 3, "the issue loop is unrolled for three stages");
    for (int g = 0, t = 0; g < n_half; g += 2 * NQ, ++t)
      trip(g, t, std::integral_constant<int, 0>{}, std::integral_constant<int, 1>{}, std::integral_constant<int, 2>{},
           std::integral_constant<int, 3>{}, std::integral_constant<int, 4>{}, std::integral_constant<int, 5>{});
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
    if (lane == 0) { MT_TRACE_DUMP("mma3"); }
  } else if (warp < W_DRAIN) {
    // ===== compute: 16 warps, thread = (key row, 16 queries of the current 64-query half tile) =======================
    const int lane_grp = warp & 3;               // TMEM lanes of this warp
    const int qq = warp >> 2;                    // 16-query quarter of every half tile; also the dK / dV column group
    const int row = lane_grp * 32 + lane;        // key slot k0 + row
    const uint32_t t_lane = (uint32_t)(lane_grp * 32) << 16;
    const int sw = row & 7;
    const bool key_ok = (k0 + row) < bg.m;       // rows past m belong to the next segment: mask column -> P = dS = 0
    const float scale_log2 = P.scale_log2, sc = P.scale;
    MT_TRACE_DECL
    MT_TRACE(1);

    // ---- K' and V' -> TMEM (A operands of S^T / dP^T for the whole CTA): quarter 0 copies K, quarter 1 copies V.
    // columns 48..50 = 1 (the three statistics columns), column 51 of K' = 1 for masked key rows, the rest 0
    mbar_wait(bar_kv_full, 0);
    MT_TRACE(2);
    if (qq < 2) {
      const uint8_t* src = smem + (qq == 0 ? Bwd3Smem::K : Bwd3Smem::V) + row * 128;
      uint32_t w[32];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const uint4 u = *reinterpret_cast<const uint4*>(src + ((c ^ sw) << 4));
        w[4 * c] = u.x; w[4 * c + 1] = u.y; w[4 * c + 2] = u.z; w[4 * c + 3] = u.w;
      }
      w[24] = 0x3f803f80u;
      w[25] = (qq == 0 && !key_ok) ? 0x3f803f80u : 0x00003f80u;
#pragma unroll
      for (int c = 26; c < 32; ++c) w[c] = 0u;
      const uint32_t dst = (qq == 0 ? tm_k : tm_v) + t_lane;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) r8[e] = w[8 * c + e];
        tmem_st8(dst + c * 8, r8);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_kvt_full);
    }

    const int n_half = 2 * n_q;
    MT_TRACE(3);
    for (int g = 0; g < n_half; ++g) {
      const int i = g >> 1, hh = g & 1;
      uint8_t* drow = smem + Bwd3Smem::DS + (i & 1) * 2 * TILE_BYTES + hh * TILE_BYTES + row * 128;
      const uint32_t ts = tmem + hh * 128 + t_lane + qq * 16, td = ts + 64;
      mbar_wait(bar_st_full + 8 * hh, i & 1);
      tc_fence_after();
      float sv[16], dp[16];
      tmem_ld16(ts, sv);
      tmem_ld16(td, dp);
      tmem_ld_wait();
      uint32_t pk[8], dk[8];
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const float p0 = ex2(sv[c] * scale_log2);
        const float p1 = ex2(sv[c + 1] * scale_log2);
        pk[c >> 1] = pack_bf16(p0, p1);
        dk[c >> 1] = pack_bf16(p0 * dp[c], p1 * dp[c + 1]);
      }
      // packed results over the first 8 of the 16 columns this thread has just read (nobody else touches them)
      tmem_st8(ts, pk);
      tmem_st8(td, dk);
      // dS^T half tile for dQ = dS K: [key row][64 queries] block hh, 16 queries = 2 swizzled 16-byte chunks
      *reinterpret_cast<uint4*>(drow + (((2 * qq) ^ sw) << 4)) = make_uint4(dk[0], dk[1], dk[2], dk[3]);
      *reinterpret_cast<uint4*>(drow + (((2 * qq + 1) ^ sw) << 4)) = make_uint4(dk[4], dk[5], dk[6], dk[7]);
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(bar_pt_full + 8 * hh);
    }
    MT_TRACE(4);
    mbar_wait(bar_done, 0);
    tc_fence_after();
    MT_TRACE(5);
    // ---- dK / dV of this key tile: TMEM -> fp32 staging (dK in dS^T buffer 0, dV in buffer 1) -> TMA reduce-adds.
    // Rows past the segment's m carry zeros (mask column); zero-key rows (position >= N) land in the scratch rows of dqkv.
    {
      float a[3][4], c2[3][4];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        tmem_ld4(tm_dk + t_lane + qq * 12 + c * 4, a[c]);
        tmem_ld4(tm_dv + t_lane + qq * 12 + c * 4, c2[c]);
      }
      tmem_ld_wait();
      uint8_t* stg_k = smem + Bwd3Smem::DS;
      uint8_t* stg_v = smem + Bwd3Smem::DS + 2 * TILE_BYTES;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        stage_chunk(stg_k, row, qq * 3 + c, make_float4(a[c][0] * sc, a[c][1] * sc, a[c][2] * sc, a[c][3] * sc));
        stage_chunk(stg_v, row, qq * 3 + c, make_float4(c2[c][0], c2[c][1], c2[c][2], c2[c][3]));
      }
      fence_proxy_async_smem();
      named_bar_sync(2, NCOMP);
      if (threadIdx.x == 0) {
        const uint32_t sk = sbase + Bwd3Smem::DS, sv_ = sbase + Bwd3Smem::DS + 2 * TILE_BYTES;
        tma_reduce_add_3d(&dq32_maps.m[b], sk, E + h * DH, off, jseg + k0);
        tma_reduce_add_3d(&dq16_maps.m[b], sk + Bwd3Smem::STG16, E + h * DH + 32, off, jseg + k0);
        tma_reduce_add_3d(&dq32_maps.m[b], sv_, 2 * E + h * DH, off, jseg + k0);
        tma_reduce_add_3d(&dq16_maps.m[b], sv_ + Bwd3Smem::STG16, 2 * E + h * DH + 32, off, jseg + k0);
        bulk_commit_group();
        bulk_wait_group_read<0>();          // the staging must outlive the reads; the adds complete with the kernel
      }
    }
    MT_TRACE(6);
    if (warp == 0 && lane == 0) { MT_TRACE_DUMP("cmp3"); }
  } else {
    // ===== dQ drain (4 warps, one query row per thread) ================================================================
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;
    const uint32_t t_lane = (uint32_t)(lane_grp * 32) << 16;
    const float sc = P.scale;
    const bool leader = threadIdx.x == W_DRAIN * 32;
    MT_TRACE_DECL
    for (int i = 0; i < n_q; ++i) {
      // dQ of query tile i: TMEM -> registers -> fp32 staging tile -> one TMA reduce-add per box
      MT_TRACE(2000 + i);
      mbar_wait(bar_dq_full, i & 1);
      tc_fence_after();
      MT_TRACE(2100 + i);
      float v[3][16];
#pragma unroll
      for (int c = 0; c < 3; ++c) tmem_ld16(tm_dq + t_lane + c * 16, v[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_dq_free);
      if (leader) bulk_wait_group_read<0>();   // the previous reduce-add has read the staging tile
      named_bar_sync(1, 128);
      uint8_t* stg = smem + Bwd3Smem::STG;
#pragma unroll
      for (int c = 0; c < 12; ++c)
        stage_chunk(stg, row, c, make_float4(v[c >> 2][(c & 3) * 4] * sc, v[c >> 2][(c & 3) * 4 + 1] * sc,
                                             v[c >> 2][(c & 3) * 4 + 2] * sc, v[c >> 2][(c & 3) * 4 + 3] * sc));
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (leader) {
        // rows past the segment's m (next segment / padding) carry dS = 0 -> they add zeros; rows >= n_alloc are clipped
        tma_reduce_add_3d(&dq32_maps.m[b], sbase + Bwd3Smem::STG, h * DH, off, jseg + i * BT);
        tma_reduce_add_3d(&dq16_maps.m[b], sbase + Bwd3Smem::STG + Bwd3Smem::STG16, h * DH + 32, off, jseg + i * BT);
        bulk_commit_group();
      }
      MT_TRACE(2200 + i);
    }
    if (leader) bulk_wait_group_read<0>();
    if (leader) { MT_TRACE_DUMP("drn3"); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

int dilated_attn_bwd3_sm100(const mt_dilated_geometry* geom, const void* qkv, int64_t qkv_ld, int64_t n_alloc,
                            const void* dattn, const float* lse, const float* delta_br, float* dqkv, cudaStream_t st) {
  Sm100Params P;
  int rc = make_sm100_params(geom, &P);
  if (rc) return rc;
  MT_REQUIRE(n_alloc >= P.geo.N && n_alloc % 128 == 0, "dilated_attn_bwd: n_alloc must be a multiple of 128 >= n_tokens");
  MT_REQUIRE(qkv_ld % 8 == 0 && ((uintptr_t)qkv & 15) == 0 && ((uintptr_t)dattn & 15) == 0 && ((uintptr_t)dqkv & 15) == 0,
             "dilated_attn_bwd: buffers must be 16-byte aligned");
  TensorMaps maps, do_maps, dq32_maps, dq16_maps;
  memset(&maps, 0, sizeof(maps));
  memset(&do_maps, 0, sizeof(do_maps));
  memset(&dq32_maps, 0, sizeof(dq32_maps));
  memset(&dq16_maps, 0, sizeof(dq16_maps));
  const int64_t E = (int64_t)P.geo.H * DH;
  for (int b = 0; b < P.geo.nb; ++b) {
    rc = encode_branch_map(&maps.m[b], qkv, qkv_ld, n_alloc, P.geo.b[b].r);
    if (rc) return rc;
    rc = encode_branch_map(&do_maps.m[b], dattn, E, n_alloc, P.geo.b[b].r);
    if (rc) return rc;
    rc = encode_branch_map_f32(&dq32_maps.m[b], dqkv, 3 * E, n_alloc, P.geo.b[b].r, 32);
    if (rc) return rc;
    rc = encode_branch_map_f32(&dq16_maps.m[b], dqkv, 3 * E, n_alloc, P.geo.b[b].r, 16);
    if (rc) return rc;
  }
  int* flag = error_flag();
  MT_REQUIRE(flag != nullptr, "dilated_attn_bwd: cannot allocate the error flag");
  MT_CUDA(cudaFuncSetAttribute(dilated_bwd3_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd3Smem::TOTAL));
  dilated_bwd3_sm100_kernel<<<P.item_prefix[P.geo.nb], BWD3_THREADS, Bwd3Smem::TOTAL, st>>>(
      maps, do_maps, dq32_maps, dq16_maps, P, lse, delta_br, flag);
  return check_launch("dilated_bwd3_sm100_kernel");
}

}  // namespace mt
