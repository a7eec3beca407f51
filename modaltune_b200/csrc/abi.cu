// Error reporting, version and geometry validation shared by every entry point of the C ABI.
#include <cstdarg>
#include <cstring>

#include "mt_common.cuh"

namespace mt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int make_dilated_geom(const mt_dilated_geometry* g, DilatedGeom* out) {
  if (g == nullptr) {
    set_error("geometry is NULL");
    return MT_E_BADARG;
  }
  if (g->n_tokens <= 0 || g->n_heads <= 0 || g->head_dim <= 0 || g->n_branches <= 0 ||
      g->n_branches > MT_MAX_BRANCHES) {
    set_error("bad geometry: N=%d H=%d D=%d branches=%d", g->n_tokens, g->n_heads, g->head_dim, g->n_branches);
    return MT_E_BADARG;
  }
  out->N = g->n_tokens;
  out->H = g->n_heads;
  out->D = g->head_dim;
  out->nb = g->n_branches;
  int64_t o_off = 0, l_off = 0;
  for (int b = 0; b < g->n_branches; ++b) {
    BranchGeom& bg = out->b[b];
    bg.r = g->ratio[b];
    if (bg.r <= 0 || g->n_heads % bg.r != 0) {
      set_error("dilation %d must divide the head count %d", bg.r, g->n_heads);
      return MT_E_UNSUPPORTED;
    }
    if (g->seg_len[b] <= 0) {
      set_error("segment length must be positive");
      return MT_E_BADARG;
    }
    bg.g = g->seg_len[b] < g->n_tokens ? g->seg_len[b] : g->n_tokens;
    bg.n_seg = (g->n_tokens + bg.g - 1) / bg.g;
    bg.m = (bg.g + bg.r - 1) / bg.r;
    if (bg.n_seg > 1 && bg.g % bg.r != 0) {
      // the reference itself scatters to the wrong positions here (dilated_attention.py:124); unreachable for
      // N <= 185363 with the GigaPath segment lengths
      set_error("segment length %d not divisible by dilation %d with %d segments", bg.g, bg.r, bg.n_seg);
      return MT_E_UNSUPPORTED;
    }
    bg.hpb = g->n_heads / bg.r;
    bg.pow2 = ((bg.r & (bg.r - 1)) == 0 && (bg.hpb & (bg.hpb - 1)) == 0 && g->n_tokens < (1 << 24)) ? 1 : 0;
    bg.log2_hpb = 0;
    while ((1 << bg.log2_hpb) < bg.hpb) ++bg.log2_hpb;
    bg.inv_g = 1.0f / (float)bg.g;
    bg.o_off = o_off;
    bg.lse_off = l_off;
    o_off += (int64_t)g->n_tokens * bg.hpb * g->head_dim;
    l_off += (int64_t)g->n_tokens * bg.hpb;
  }
  return 0;
}

}  // namespace mt

extern "C" const char* mt_last_error(void) { return mt::g_err; }
extern "C" int mt_version(void) { return 100; }
extern "C" int mt_device_is_sm100(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    mt::set_error("no CUDA device");
    return MT_E_BADARG;
  }
  return p.major == 10 ? 1 : 0;
}
