"""GPU, experiment build (-DMT_DEBUG_TIMELINE): per-CTA (SM, loop length, entry / exit clock) records of the tcgen05
attention kernels -> how much of each SM's time is spent INSIDE CTAs, how long a CTA lives as a function of its loop
length (in-CTA fixed cost = prologue + epilogue), and how long the SM idles between two CTAs.

    MT_BUILD_OUT=build_exp/libmt_timeline.so MT_EXTRA_NVCC_FLAGS=-DMT_DEBUG_TIMELINE python -m modaltune_b200.build
    MODALTUNE_B200_LIB=build_exp/libmt_timeline.so python tools/attn_timeline.py [n_tokens]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from modaltune_b200 import _lib, ops  # noqa: E402
from modaltune_b200.slide_encoder import DILATED_RATIO, optimal_segment_lengths  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
FWD_IMPL = int(sys.argv[2]) if len(sys.argv) > 2 else 1
BWD_IMPL = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = "cuda"
lib = _lib.load()
lib.mt_debug_timeline.restype = ctypes.c_int
lib.mt_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]


def timeline(n=32768):
    buf = np.zeros((n, 4), dtype=np.int64)
    assert lib.mt_debug_timeline(buf.ctypes.data, n) == 0
    return buf[buf[:, 3] != 0]


def report(name, tl, per_sm_slots):
    sm, n, t0, t1 = tl[:, 0], tl[:, 1], tl[:, 2], tl[:, 3]
    dur = t1 - t0
    span = busy = 0
    gaps = []
    for s in np.unique(sm):
        m = sm == s
        a, b = t0[m], t1[m]
        order = np.argsort(a)
        a, b = a[order], b[order]
        span += b.max() - a.min()
        # union coverage of the CTA intervals on this SM
        cur_a, cur_b, cov = a[0], b[0], 0
        for x, y in zip(a[1:], b[1:]):
            if x > cur_b:
                cov += cur_b - cur_a
                gaps.append(x - cur_b)
                cur_a, cur_b = x, y
            else:
                cur_b = max(cur_b, y)
        cov += cur_b - cur_a
        busy += cov
    A = np.stack([n.astype(np.float64), np.ones(len(n))], 1)
    (a_fit, b_fit), *_ = np.linalg.lstsq(A, dur.astype(np.float64), rcond=None)
    kernel_span = max((t1[sm == s].max() - t0[sm == s].min()) for s in np.unique(sm))
    print(f"{name}: {len(tl)} CTAs on {len(np.unique(sm))} SMs, longest SM span {kernel_span} cycles")
    print(f"   CTA lifetime = {a_fit:.0f} cycles x loop length + {b_fit:.0f} cycles (in-CTA prologue + epilogue), "
          f"{per_sm_slots} CTA(s) resident per SM")
    print(f"   fraction of SM time covered by at least one CTA: {busy / span:.3f}; "
          f"idle gaps between CTAs: n = {len(gaps)}, median {np.median(gaps) if gaps else 0:.0f}, mean {np.mean(gaps) if gaps else 0:.0f} cycles")
    for L in np.unique(n):
        d = dur[n == L]
        print(f"      loop length {L:3d}: {len(d):5d} CTAs, lifetime median {np.median(d):8.0f}  p10 {np.percentile(d, 10):8.0f}  p90 {np.percentile(d, 90):8.0f}  "
              f"-> {np.median(d) / L:6.0f} cycles per tile")
    ends = np.array([t1[sm == s].max() - t0.min() for s in np.unique(sm)])   # clock64 is per SM: only a rough alignment
    return a_fit, b_fit


geom = ops.Geometry.get(N, optimal_segment_lengths(), DILATED_RATIO)
g = torch.Generator().manual_seed(0)
qkv = torch.zeros(geom.n_alloc, 2304)
qkv[:N] = torch.randn(N, 2304, generator=g)
qkv = qkv.to(torch.bfloat16).to(dev)
gamma, beta = torch.ones(768, device=dev), torch.zeros(768, device=dev)
dy = torch.randn(N, 768, generator=g).to(dev)
for _ in range(2):
    o, l = ops.dilated_attn_fwd(geom, qkv, FWD_IMPL)
    y, _, lse, m, r = ops.dilated_merge_ln_fwd(geom, o, l, gamma, beta)
    dattn, delta = ops.dilated_merge_ln_bwd(geom, dy, o, l, gamma, m, r)
    dq = ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, BWD_IMPL)
torch.cuda.synchronize()
timeline()
o, l = ops.dilated_attn_fwd(geom, qkv, FWD_IMPL)
report(f"forward impl {FWD_IMPL} N={N}", timeline(), 2)
dq = ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, BWD_IMPL)
report(f"backward impl {BWD_IMPL} N={N}", timeline(), 1)

# ---- per-item records of the persistent backward (impl 2): time inside an item vs the bubble between two items ---------
if BWD_IMPL == 2:
    lib.mt_debug_item_timeline.restype = ctypes.c_int
    lib.mt_debug_item_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
    buf = np.zeros((65536, 4), dtype=np.int64)
    cnt = lib.mt_debug_item_timeline(buf.ctypes.data, 65536)   # includes the warm-up launches: keep the last launch
    dq = ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, BWD_IMPL)
    cnt = lib.mt_debug_item_timeline(buf.ctypes.data, 65536)
    rec = buf[:cnt]
    inside = rec[:, 3] - rec[:, 2]
    A = np.stack([rec[:, 1].astype(np.float64), np.ones(cnt)], 1)
    (a_in, b_in), *_ = np.linalg.lstsq(A, inside.astype(np.float64), rcond=None)
    bubbles, after_len = [], []
    for c in np.unique(rec[:, 0]):
        r = rec[rec[:, 0] == c]
        r = r[np.argsort(r[:, 2])]
        bubbles += list(r[1:, 2] - r[:-1, 3])
        after_len += list(r[:-1, 1])
    bubbles, after_len = np.array(bubbles), np.array(after_len)
    print(f"persistent backward, {cnt} items: first score tile -> last P^T of an item = {a_in:.0f} cycles x loop length + {b_in:.0f}")
    print(f"   bubble between the last half tile of an item and the first score tile of the next: median {np.median(bubbles):.0f}, "
          f"mean {bubbles.mean():.0f}, p10 {np.percentile(bubbles, 10):.0f}, p90 {np.percentile(bubbles, 90):.0f} cycles")
    for L in np.unique(rec[:, 1]):
        m = rec[:, 1] == L
        print(f"      loop length {L:3d}: {m.sum():5d} items, inside median {np.median(inside[m]):8.0f} -> {np.median(inside[m]) / L:6.0f} per tile; "
              f"bubble after such an item: median {np.median(bubbles[after_len == L]) if (after_len == L).any() else 0:.0f}")
