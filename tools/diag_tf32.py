import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import ops
g = torch.Generator().manual_seed(0)
x = torch.randn(1100, 768, generator=g).cuda().requires_grad_(True)
w = (torch.randn(192, 768, generator=g) * 0.05).cuda().requires_grad_(True)
b = torch.randn(192, generator=g).cuda().requires_grad_(True)
dy = torch.randn(1100, 192, generator=g).cuda()
def rel(a, r): return float((a.double() - r.double()).norm() / r.double().norm())
yr = torch.nn.functional.linear(x.double(), w.double(), b.double())
gr = torch.autograd.grad(yr, [x, w, b], dy.double())
for name, fn in (("fp32", lambda: torch.nn.functional.linear(x, w, b)), ("tf32", lambda: ops.linear_tf32(x, w, b)),
                 ("bf16", lambda: torch.nn.functional.linear(x.bfloat16(), w.bfloat16(), b.bfloat16()).float())):
    y = fn()
    gs = torch.autograd.grad(y, [x, w, b], dy)
    print(name, "y", f"{rel(y, yr):.2e}", "dx", f"{rel(gs[0], gr[0]):.2e}", "dw", f"{rel(gs[1], gr[1]):.2e}", "db", f"{rel(gs[2], gr[2]):.2e}")
print("allow_tf32 after:", torch.backends.cuda.matmul.allow_tf32)
