"""GPU, experiment build (-DMT_DEBUG_TRACE): run one backward launch on a single-branch geometry so that the traced CTA's
items all have the same loop length; the kernel prints its phase stamps (tools/trace_timeline.py merges them).

    MODALTUNE_B200_LIB=build_exp/libmt_trace.so python tools/attn_trace.py [n_tokens] [segment_length] [ratio] [bwd_impl]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from modaltune_b200 import ops  # noqa: E402
from modaltune_b200.slide_encoder import DILATED_RATIO, optimal_segment_lengths  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
sl = int(sys.argv[2]) if len(sys.argv) > 2 else 0
r = int(sys.argv[3]) if len(sys.argv) > 3 else 1
impl = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = "cuda"
geom = ops.Geometry(N, [sl], [r]) if sl else ops.Geometry.get(N, optimal_segment_lengths(), DILATED_RATIO)
g = torch.Generator().manual_seed(0)
qkv = torch.zeros(geom.n_alloc, 2304)
qkv[:N] = torch.randn(N, 2304, generator=g)
qkv = qkv.to(torch.bfloat16).to(dev)
dattn = torch.zeros(geom.n_alloc, 768)
dattn[:N] = torch.randn(N, 768, generator=g)
dattn = dattn.to(torch.bfloat16).to(dev)
lse = torch.full((N, 16), 6.0, device=dev)
delta = torch.zeros(geom.lse_elems, device=dev)
ops.dilated_attn_bwd(geom, qkv, dattn, lse, delta, impl)
torch.cuda.synchronize()
