"""GPU: time the Injector / Extractor cross-attention kernels alone at the bench shapes (CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)
for name, lq, lk in (("injector", 10000, 66), ("extractor", 66, 10000), ("prompt-sa", 66, 66)):
    q, k, v = (torch.randn(n, 192, generator=g).to(dev) for n in (lq, lk, lk))
    d_o = torch.randn(lq, 192, generator=g).to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(5):
        ev[0].record()
        o, lse = ops.cross_attn_fwd(q, k, v, 12)
        ev[1].record()
        dq, dk, dv = ops.cross_attn_bwd(q, k, v, o, d_o, lse, 12)
        ev[2].record()
    torch.cuda.synchronize()
    flop = 4.0 * lq * lk * 192
    tf, tb = ev[0].elapsed_time(ev[1]) * 1e3, ev[1].elapsed_time(ev[2]) * 1e3
    print(f"{name:10s} lq={lq:6d} lk={lk:6d}  fwd {tf:7.1f} us ({flop / tf / 1e6:6.2f} TFLOP/s)   bwd {tb:7.1f} us "
          f"({2.5 * flop / tb / 1e6:6.2f} TFLOP/s)")
