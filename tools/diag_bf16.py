"""Diagnostic (GPU): which part of the bf16 path limits the gradient cosine?  Reference = our fp32 mode on the GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import config, ops, synthetic, train_step, adapter_modules, slide_encoder
from modaltune_b200 import factory as helpers

dev = "cuda"
L = int(sys.argv[1]) if len(sys.argv) > 1 else 520
model = helpers.build_model(helpers.SMALL_GROUPS, device=dev)
proj = helpers.build_projector(0, dev)
slide = train_step.slide_to_device(synthetic.synthetic_slide(L, seed=77, group_sizes=helpers.SMALL_GROUPS), dev)


def run(mode):
    model.zero_grad()
    with config.using(mode=mode, attn_impl="simt"):
        loss, logits = train_step.forward_backward(model, proj, slide)
    return logits.float().cpu(), {k: p.grad.detach().double().cpu().clone() for k, p in model.named_parameters() if p.requires_grad}


def cos(a, b):
    a, b = a.flatten(), b.flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


ref_logits, ref = run("fp32")
gmax = max(float(g.norm()) for g in ref.values())


def report(tag):
    logits, g = run("bf16")
    cs = sorted((cos(g[k], ref[k]), k) for k in ref if float(ref[k].norm()) > 1e-4 * gmax)
    print(f"{tag:28s} logits rel {helpers.relerr(logits, ref_logits):.2e}  min cos {cs[0][0]:.5f} ({cs[0][1]})  "
          f"#<0.999: {sum(c < 0.999 for c, _ in cs)}  5 worst: {[round(c, 5) for c, _ in cs[:5]]}")


report("baseline bf16")

# A: cross-attention core in fp32
orig_ca = ops.cross_attention
ops.cross_attention = lambda q, k, v, h: orig_ca(q.float(), k.float(), v.float(), h).to(q.dtype)
report("A: cross-attn core fp32")
ops.cross_attention = orig_ca

# B: encoder layers fp32, adapter bf16
orig_fel = ops.frozen_encoder_layer
orig_refresh = ops.FrozenLayerWeights.refresh
ops.frozen_encoder_layer = lambda x, W, geom, cdt, impl: orig_fel(x, W, geom, torch.float32, impl)
ops.FrozenLayerWeights.refresh = lambda self, layer, dtype: orig_refresh(self, layer, torch.float32)
report("B: encoder fp32")
ops.frozen_encoder_layer = orig_fel
ops.FrozenLayerWeights.refresh = orig_refresh

# C: adapter fp32 (LN outputs + linears), encoder bf16
orig_cd = config.compute_dtype
import modaltune_b200.adapter_modules as am
am_config_cd = am.config.compute_dtype
class _Cfg:
    def __getattr__(self, n):
        return getattr(config, n)
    def compute_dtype(self):
        return torch.float32
am.config = _Cfg()
report("C: adapter fp32")
am.config = config

# D: adapter linears exact fp32 (no TF32)
orig_lt = ops.linear_tf32
ops.linear_tf32 = lambda x, w, b: torch.nn.functional.linear(x, w, b)
report("D: adapter exact fp32")
ops.frozen_encoder_layer = lambda x, W, geom, cdt, impl: orig_fel(x, W, geom, torch.float32, impl)
ops.FrozenLayerWeights.refresh = lambda self, layer, dtype: orig_refresh(self, layer, torch.float32)
report("D+B: + encoder fp32")
orig_embed = type(model).embed
def embed32(self, x, coords):
    with config.using(mode="fp32"):
        return orig_embed(self, x, coords)
type(model).embed = embed32
report("D+B+embed fp32")
