"""Diagnostic (GPU): per-parameter gradient cosine of the bf16 mode vs the fp32 mode over slide sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import config, synthetic, train_step
from modaltune_b200 import factory as helpers
dev = "cuda"
model = helpers.build_model(helpers.SMALL_GROUPS, device=dev)
proj = helpers.build_projector(0, dev)
def run(mode, slide):
    model.zero_grad()
    with config.using(mode=mode):
        loss, logits = train_step.forward_backward(model, proj, slide)
    return logits.float().cpu(), {k: p.grad.detach().double().cpu().clone() for k, p in model.named_parameters() if p.requires_grad}
def cos(a, b):
    a, b = a.flatten(), b.flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))
for L in [int(a) for a in sys.argv[1:]]:
    slide = train_step.slide_to_device(synthetic.synthetic_slide(L, seed=77 + L, group_sizes=helpers.SMALL_GROUPS), dev)
    rl, ref = run("fp32", slide)
    bl, g = run("bf16", slide)
    gmax = max(float(v.norm()) for v in ref.values())
    cs = sorted((cos(g[k], ref[k]), k) for k in ref if float(ref[k].norm()) > 1e-4 * gmax)
    allc = cos(torch.cat([g[k].flatten() for k in ref]), torch.cat([ref[k].flatten() for k in ref]))
    print(f"L={L:6d} logits rel {helpers.relerr(bl, rl):.2e} global cos {allc:.6f} min cos {cs[0][0]:.5f} ({cs[0][1]}) #<0.999: {sum(c < 0.999 for c, _ in cs)}/{len(cs)}")
