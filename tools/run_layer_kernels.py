"""GPU: one frozen encoder layer, forward + backward (dX), at the bench size in bf16 mode -- the ncu target for the
non-attention kernels of the layer (LayerNorm, merge+LN, GELU+LN backward, casts, the tcgen05 GEMMs).

    python tools/run_layer_kernels.py [n_tokens] [reps] [gemm: sm100|cublas]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from modaltune_b200 import config, factory  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10001
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gemm = sys.argv[3] if len(sys.argv) > 3 else "sm100"
dev = "cuda"
model = factory.build_model(factory.SMALL_GROUPS, device=dev)
layer = model.encoder.layers[5]
g = torch.Generator().manual_seed(0)
x = torch.randn(1, N, 768, generator=g).to(dev).requires_grad_(True)
dy = torch.randn(1, N, 768, generator=g).to(dev)
mask = torch.zeros(1, N, dtype=torch.bool, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
with config.using(mode="bf16", attn_impl="auto", gemm=gemm):
    for _ in range(reps):
        ev[0].record()
        y, _ = layer(x, encoder_padding_mask=mask)
        ev[1].record()
        (gx,) = torch.autograd.grad(y, x, dy)
        ev[2].record()
torch.cuda.synchronize()
print(f"N={N} gemm={gemm}: layer forward {ev[0].elapsed_time(ev[1]):.3f} ms, backward {ev[1].elapsed_time(ev[2]):.3f} ms, "
      f"checksum {float(gx.abs().sum()):.4e}")
