"""GPU, -DMT_DEBUG_TRACE build: one launch of mt_linear_sm100 per epilogue flavour; the kernel prints where its epilogue
warps spend their cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import _lib, ops
dev = "cuda"; M = 10001
g = torch.Generator().manual_seed(0)
a = torch.randn(M, 768, generator=g).to(torch.bfloat16).to(dev)
w = (torch.randn(3072, 768, generator=g) * 0.05).to(torch.bfloat16).to(dev)
bias = torch.randn(3072, generator=g).to(dev)
for _ in range(2):
    ops.linear_sm100(a, w); torch.cuda.synchronize()
print("--- bf16 + bias"); ops.linear_sm100(a, w, bias=bias, want_f32=False, want_bf16=True); torch.cuda.synchronize()
stats = torch.empty(M, 24, 2, device=dev)
print("--- gelu + stats"); ops.linear_sm100(a, w, mode=_lib.MT_EPI_GELU_STATS, bias=bias, want_f32=True, want_bf16=True, stats=stats); torch.cuda.synchronize()
