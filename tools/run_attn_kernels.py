"""GPU: run the dilated-attention kernels (tcgen05 fwd + bwd, merge) alone at the bench size -- the ncu target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import ops
from modaltune_b200.slide_encoder import DILATED_RATIO, optimal_segment_lengths
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bwd_impl = int(sys.argv[3]) if len(sys.argv) > 3 else 1
fwd_impl = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = "cuda"
geom = ops.Geometry.get(N, optimal_segment_lengths(), DILATED_RATIO)
g = torch.Generator().manual_seed(0)
qkv = torch.zeros(geom.n_alloc, 2304)
qkv[:N] = torch.randn(N, 2304, generator=g)
qkv = qkv.to(torch.bfloat16).to(dev)
gamma, beta = torch.ones(768, device=dev), torch.zeros(768, device=dev)
dy = torch.randn(N, 768, generator=g).to(dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for i in range(reps):
    ev[0].record()
    o, l = ops.dilated_attn_fwd(geom, qkv, fwd_impl)
    ev[1].record()
    y, _, lse, m, r = ops.dilated_merge_ln_fwd(geom, o, l, gamma, beta)
    dattn, delta = ops.dilated_merge_ln_bwd(geom, dy, o, l, gamma, m, r)
    ev[2].record()
    dq = ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, bwd_impl)
    ev[3].record()
torch.cuda.synchronize()
f, b = ops.attention_flops(geom)
tf, tb = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])
print(f"N={N} fwd {tf:.3f} ms {f/tf/1e9:.1f} TFLOP/s   bwd {tb:.3f} ms {b/tb/1e9:.1f} TFLOP/s   checksum {float(dq.abs().sum()):.4e}")
