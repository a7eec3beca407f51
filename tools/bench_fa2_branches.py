"""Same-box comparator: the kernel the REFERENCE runs for its dilated attention, timed on the B200 it shares with ours.

The reference calls ``flash_attn_func(q, k, v, ..., return_attn_probs=True)`` once per (segment length, dilation)
branch, five times per encoder layer, on zero-padded sparse copies of q / k / v
(``torchscale/component/flash_attention.py:11-28``, ``multihead_attention.py:109-119``,
``dilated_attention.py:82-144``).  The image ships flash-attn 2.8 (FA2 ``mma.sync`` kernels built for sm_100): this
tool times

* ``kernels``: the five ``flash_attn_func`` calls, forward and backward, on pre-gathered [n_seg, m, 16, 48] bf16 tensors
  (what the reference's attention kernels alone cost), and
* ``with_gather_scatter``: the same plus a plain-torch gather (pad, split into segments, keep every r-th position per
  head group) and scatter + log-sum-exp merge around them (the data movement the reference performs per layer),

next to ``mt_dilated_attn_fwd`` / ``mt_dilated_attn_bwd`` (+ the merge kernels) of this repo on the SAME token count, with
CUDA events.  flash-attn is library code used here ONLY as the comparator; nothing in the product imports it.

    python tools/bench_fa2_branches.py [n_tokens ...]          # default 10001 32769; prints one JSON object
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

HEADS, HEAD_DIM = 16, 48


def branch_shapes(n_tokens, segment_lengths, ratios):
    """(n_seg, m, g, r) per branch: batch and sequence length of the reference's flash_attn_func call."""
    out = []
    for sl, r in zip(segment_lengths, ratios):
        g = min(int(sl), n_tokens)
        n_seg = -(-n_tokens // g)
        m = -(-g // r)
        out.append((n_seg, m, g, int(r)))
    return out


def gather(x, n_seg, m, g, r):
    """x [N, 16, 48] -> [n_seg, m, 16, 48]: zero-pad to whole segments and to a multiple of r, head h keeps the local
    positions floor(h*r/16) + j*r.  Plain torch (pad + strided slices + one stack), as a framework would do it."""
    N = x.shape[0]
    x = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, n_seg * g - N))
    x = x.view(n_seg, g, HEADS, HEAD_DIM)
    if m * r > g:
        x = torch.nn.functional.pad(x, (0, 0, 0, 0, 0, m * r - g))
    hpb = HEADS // r
    x = x.view(n_seg, m, r, r, hpb, HEAD_DIM)            # [seg, slot, residue, head group, head in group, d]
    return torch.diagonal(x, dim1=2, dim2=3).permute(0, 1, 4, 2, 3).reshape(n_seg, m, HEADS, HEAD_DIM).contiguous()


def scatter(o, lse, n_seg, m, g, r, N):
    """inverse of ``gather`` for an output [n_seg, m, 16, 48] and its lse [n_seg, 16, m]: dense [N, 16, 48] with zeros /
    -1e8 where a head does not own a position."""
    hpb = HEADS // r
    dense = o.new_zeros(n_seg, m, r, r, hpb, HEAD_DIM)
    torch.diagonal(dense, dim1=2, dim2=3).copy_(o.view(n_seg, m, r, hpb, HEAD_DIM).permute(0, 1, 3, 4, 2))
    dense = dense.view(n_seg, m * r, HEADS, HEAD_DIM)[:, :g].reshape(n_seg * g, HEADS, HEAD_DIM)[:N]
    l = lse.new_full((n_seg, m, r, r, hpb), -1e8)
    torch.diagonal(l, dim1=2, dim2=3).copy_(lse.view(n_seg, r, hpb, m).permute(0, 3, 2, 1))
    l = l.view(n_seg, m * r, HEADS)[:, :g].reshape(n_seg * g, HEADS)[:N]
    return dense, l


def merge(outs, lses):
    with torch.no_grad():
        L = torch.stack(lses, 0)
        w = torch.softmax(L, 0)
    out = 0
    for wb, ob in zip(w, outs):
        out = out + ob * wb.unsqueeze(-1).to(ob.dtype)
    return out


def _time(fn, reps, warm=3):
    for _ in range(warm):
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
    ts.sort()
    return ts[len(ts) // 2]


def run(n_tokens, reps=7, dev="cuda"):
    from flash_attn.flash_attn_interface import flash_attn_func

    from modaltune_b200 import config, ops
    from modaltune_b200.slide_encoder import DILATED_RATIO, optimal_segment_lengths

    seg = optimal_segment_lengths()
    shapes = branch_shapes(n_tokens, seg, DILATED_RATIO)
    geom = ops.Geometry.get(n_tokens, seg, DILATED_RATIO)
    g = torch.Generator().manual_seed(n_tokens)
    qkv = torch.zeros(geom.n_alloc, 3 * HEADS * HEAD_DIM)
    qkv[:n_tokens] = torch.randn(n_tokens, 3 * HEADS * HEAD_DIM, generator=g)
    qkv = qkv.to(torch.bfloat16).to(dev)
    q, k, v = (qkv[:n_tokens, i * 768:(i + 1) * 768].reshape(n_tokens, HEADS, HEAD_DIM).contiguous() for i in range(3))
    d_out = torch.randn(n_tokens, HEADS, HEAD_DIM, generator=g).to(torch.bfloat16).to(dev)

    # ---- FA2, kernels only: pre-gathered operands, five calls ---------------------------------------------------------
    sparse = [tuple(gather(t, *s).requires_grad_(True) for t in (q, k, v)) for s in shapes]
    d_sparse = [gather(d_out, *s) for s in shapes]

    def fa2_fwd():
        return [flash_attn_func(a, b, c, dropout_p=0.0, softmax_scale=None, causal=False, return_attn_probs=True)[:2]
                for a, b, c in sparse]

    t_fwd = _time(fa2_fwd, reps)
    outs = fa2_fwd()

    def fa2_bwd():
        for (o, _), (a, b, c), d in zip(outs, sparse, d_sparse):
            torch.autograd.grad(o, (a, b, c), d, retain_graph=True)

    t_bwd = _time(fa2_bwd, reps)

    # ---- FA2 with the gather / scatter / merge around it ----------------------------------------------------------------
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))

    def ref_core():
        o_d, l_d = [], []
        for s in shapes:
            a, b, c = gather(qr, *s), gather(kr, *s), gather(vr, *s)
            o, lse, _ = flash_attn_func(a, b, c, dropout_p=0.0, softmax_scale=None, causal=False, return_attn_probs=True)
            od, ld = scatter(o, lse[:, :, :o.shape[1]], *s, n_tokens)
            o_d.append(od)
            l_d.append(ld)
        return merge(o_d, l_d)

    t_core_fwd = _time(lambda: ref_core(), reps)
    attn_ref = ref_core()
    t_core_bwd = _time(lambda: torch.autograd.grad(attn_ref, (qr, kr, vr), d_out, retain_graph=True), reps)

    # ---- ours ---------------------------------------------------------------------------------------------------------
    impl_f, impl_b = config.AUTO_IMPL["fwd"], config.AUTO_IMPL["bwd"]
    ones, zeros = torch.ones(768, device=dev), torch.zeros(768, device=dev)
    t_ours_fwd = _time(lambda: ops.dilated_attn_fwd(geom, qkv, impl_f), reps)
    o_br, lse_br = ops.dilated_attn_fwd(geom, qkv, impl_f)
    t_ours_merge = _time(lambda: ops.dilated_merge_ln_fwd(geom, o_br, lse_br, ones, zeros), reps)
    y, attn, lse, mean, rstd = ops.dilated_merge_ln_fwd(geom, o_br, lse_br, ones, zeros, want_attn=True)
    dy = d_out.reshape(n_tokens, 768).float()
    t_ours_merge_bwd = _time(lambda: ops.dilated_merge_ln_bwd(geom, dy, o_br, lse_br, ones, mean, rstd), reps)
    dattn, delta = ops.dilated_merge_ln_bwd(geom, dy, o_br, lse_br, ones, mean, rstd)
    t_ours_bwd = _time(lambda: ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, impl_b), reps)

    # agreement of the two implementations on the same inputs (bf16 both): merged attention output
    err = float((attn.float() - attn_ref.detach().reshape(n_tokens, 768).float()).abs().max() /
                attn_ref.detach().float().abs().max())
    f_fwd, f_bwd = ops.attention_flops(geom)
    padded = sum(4.0 * HEAD_DIM * n * HEADS * m * m for n, m, _, _ in shapes)
    return {
        "n_tokens": n_tokens, "branch_shapes_batch_seqlen": [(n, m) for n, m, _, _ in shapes],
        "algorithmic_gflop_fwd": f_fwd / 1e9, "as_padded_by_the_reference_gflop_fwd": padded / 1e9,
        "fa2_kernels_ms": {"fwd": t_fwd, "bwd": t_bwd},
        "fa2_with_gather_scatter_ms": {"fwd": t_core_fwd, "bwd": t_core_bwd},
        "ours_ms": {"fwd": t_ours_fwd, "bwd": t_ours_bwd, "merge_ln_fwd": t_ours_merge, "merge_ln_bwd": t_ours_merge_bwd},
        "speedup_vs_fa2_kernels": {"fwd": t_fwd / t_ours_fwd, "bwd": t_bwd / t_ours_bwd},
        "speedup_vs_fa2_with_gather_scatter": {"fwd": t_core_fwd / (t_ours_fwd + t_ours_merge),
                                               "bwd": t_core_bwd / (t_ours_bwd + t_ours_merge_bwd)},
        "tflops_algorithmic": {"fa2_fwd": f_fwd / t_fwd / 1e9, "fa2_bwd": f_bwd / t_bwd / 1e9,
                               "ours_fwd": f_fwd / t_ours_fwd / 1e9, "ours_bwd": f_bwd / t_ours_bwd / 1e9},
        "merged_output_rel_diff_ours_vs_fa2": err,
        "flash_attn_version": __import__("flash_attn").__version__,
    }


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [10001, 32769]
    assert torch.cuda.is_available()
    res = {"comparator": "flash_attn_func (FA2 for sm_100) on the reference's five per-layer shapes, bf16, d = 48",
           "cases": [run(n) for n in sizes]}
    print(json.dumps(res))
    return res


if __name__ == "__main__":
    main()
