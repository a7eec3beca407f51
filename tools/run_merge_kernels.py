"""GPU: time the branch-merge + inner_attn_ln kernels alone at the bench size (CUDA-graph replay of 20 calls)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import ops
from modaltune_b200.slide_encoder import DILATED_RATIO, optimal_segment_lengths
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
dev = "cuda"
geom = ops.Geometry.get(N, optimal_segment_lengths(), DILATED_RATIO)
g = torch.Generator().manual_seed(0)
o = torch.randn(geom.o_elems, generator=g).to(torch.bfloat16).to(dev)
l = torch.randn(geom.lse_elems, generator=g).to(dev)
gamma, beta = torch.ones(768, device=dev), torch.zeros(768, device=dev)
dy = torch.randn(N, 768, generator=g).to(dev)
y, _, lse, m, r = ops.dilated_merge_ln_fwd(geom, o, l, gamma, beta)


def timed(fn, rep=20):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(rep):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / rep


tf = timed(lambda: ops.dilated_merge_ln_fwd(geom, o, l, gamma, beta))
tb = timed(lambda: ops.dilated_merge_ln_bwd(geom, dy, o, l, gamma, m, r))
d, de = ops.dilated_merge_ln_bwd(geom, dy, o, l, gamma, m, r)
print(f"N={N}: merge_ln fwd {tf:.1f} us, bwd {tb:.1f} us (bwd includes the zero tail fill); "
      f"checksums {float(y.float().abs().sum()):.6e} {float(d.float().abs().sum()):.6e} {float(de.abs().sum()):.6e}")
