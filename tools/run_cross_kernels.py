"""GPU: time the Injector / Extractor cross-attention kernels alone at the bench shapes (CUDA-graph replay of 20
back-to-back calls, CUDA events: the kernels take a few microseconds, far less than a Python launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)
shapes = (("injector", 10000, 66), ("extractor", 66, 10000), ("prompt-sa", 66, 66))
REP = 20


def timed(fn):
    fn(); torch.cuda.synchronize()
    if "--plain" in sys.argv:   # under ncu: one plain launch per call is enough
        return float("nan")
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(REP):
            fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / REP


for impl, (name, lq, lk) in ((i, s) for i in (0, 1) for s in shapes):
    q, k, v = (torch.randn(n, 192, generator=g).to(dev) for n in (lq, lk, lk))
    d_o = torch.randn(lq, 192, generator=g).to(dev)
    o, lse = ops.cross_attn_fwd(q, k, v, 12, impl=impl)
    tf = timed(lambda: ops.cross_attn_fwd(q, k, v, 12, impl=impl))
    tb = timed(lambda: ops.cross_attn_bwd(q, k, v, o, d_o, lse, 12, impl=impl))
    flop = 4.0 * lq * lk * 192
    print(f"impl {impl} ({'TF32 mma.sync' if impl else 'SIMT fp32'}) {name:10s} lq={lq:6d} lk={lk:6d}  fwd {tf:7.1f} us "
          f"({flop / tf / 1e6:6.2f} TFLOP/s)   bwd {tb:7.1f} us ({2.5 * flop / tb / 1e6:6.2f} TFLOP/s)")
