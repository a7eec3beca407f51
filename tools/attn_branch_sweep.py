"""GPU: where does the time of the tcgen05 dilated-attention kernels go -- steady state per 128 x 128 tile pair, or fixed
cost per CTA?  Times the forward and backward kernels on SINGLE-branch geometries (one (segment length, dilation) pair
each) so that every CTA of a launch has the same loop length, and fits  t = a * tile_pairs + b * CTAs.

    python tools/attn_branch_sweep.py [n_tokens]         # default 10001
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modaltune_b200 import ops  # noqa: E402
from modaltune_b200.slide_encoder import DILATED_RATIO, optimal_segment_lengths  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
FWD_IMPL = int(sys.argv[2]) if len(sys.argv) > 2 else 1
BWD_IMPL = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = "cuda"
SMS = 148


def med(fn, reps=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def count(n_tokens, sl, r):
    g = min(sl, n_tokens)
    n_seg = -(-n_tokens // g)
    ctas = pairs = 0
    for s in range(n_seg):
        lo, hi = s * g, min(n_tokens, (s + 1) * g)
        for h in range(16):
            off = (h * r) // 16
            c = max(0, -(-(hi - lo - off) // r))
            t = -(-c // 128)
            ctas += t
            pairs += t * t
    return ctas, pairs


g = torch.Generator().manual_seed(0)
n_alloc = -(-N // 128) * 128
qkv = torch.zeros(n_alloc, 2304)
qkv[:N] = torch.randn(N, 2304, generator=g)
qkv = qkv.to(torch.bfloat16).to(dev)
dattn = torch.zeros(n_alloc, 768)
dattn[:N] = torch.randn(N, 768, generator=g)
dattn = dattn.to(torch.bfloat16).to(dev)
lse = torch.full((N, 16), 6.0, device=dev)

cases = [(sl, r) for sl, r in zip(optimal_segment_lengths(), DILATED_RATIO)]
cases += [(256, 1), (512, 1), (2048, 1), (4096, 1), (16384, 1), (2048, 2), (4096, 4)]
rows = []
print(f"N = {N}; fwd impl {FWD_IMPL}, bwd impl {BWD_IMPL}; cycles at 1.9 GHz; tile pair = 128 queries x 128 keys of one head")
print(f"{'branch':>14} {'CTAs':>6} {'pairs':>7} {'len':>5} | {'fwd ms':>8} {'cyc/pair/SM':>12} | {'bwd ms':>8} {'cyc/pair/SM':>12}")
for sl, r in cases:
    geom = ops.Geometry(N, [sl], [r])
    ctas, pairs = count(N, sl, r)
    delta = torch.zeros(geom.lse_elems, device=dev)
    tf = med(lambda: ops.dilated_attn_fwd(geom, qkv, FWD_IMPL))
    tb = med(lambda: ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, BWD_IMPL))
    cf, cb = tf * 1e-3 * 1.9e9 * SMS / pairs, tb * 1e-3 * 1.9e9 * SMS / pairs
    rows.append((ctas, pairs, tf, tb))
    print(f"{sl:>9}/r{r:<3} {ctas:>6} {pairs:>7} {pairs / ctas:>5.1f} | {tf:>8.4f} {cf:>12.0f} | {tb:>8.4f} {cb:>12.0f}")

# least squares t = a * pairs + b * ctas
A = torch.tensor([[p, c] for c, p, _, _ in rows], dtype=torch.float64)
for name, col in (("fwd", 2), ("bwd", 3)):
    y = torch.tensor([r_[col] for r_ in rows], dtype=torch.float64)
    sol = torch.linalg.lstsq(A, y[:, None]).solution.flatten()
    a, b = float(sol[0]), float(sol[1])
    print(f"{name}: a = {a * 1e-3 * 1.9e9 * SMS:.0f} cycles per tile pair and SM,  b = {b * 1e-3 * 1.9e9 * SMS:.0f} cycles per CTA "
          f"(per SM; the forward runs 2 CTAs per SM)")
geom = ops.Geometry.get(N, optimal_segment_lengths(), DILATED_RATIO)
delta = torch.zeros(geom.lse_elems, device=dev)
tf = med(lambda: ops.dilated_attn_fwd(geom, qkv, FWD_IMPL))
tb = med(lambda: ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, BWD_IMPL))
f, b = ops.attention_flops(geom)
print(f"all five branches in one launch: fwd {tf:.4f} ms = {f / tf / 1e9:.0f} TFLOP/s, bwd {tb:.4f} ms = {b / tb / 1e9:.0f} TFLOP/s; "
      f"sum of the five single-branch launches: fwd {sum(r_[2] for r_ in rows[:5]):.4f} bwd {sum(r_[3] for r_ in rows[:5]):.4f}")
