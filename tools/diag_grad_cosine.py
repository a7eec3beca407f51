"""GPU diagnostic: per-parameter gradient cosine of the bf16 mode (SIMT and tcgen05 attention) against the CPU oracle,
for the 10-pathway toy model and the 331-pathway model.  Writes gpurun_out/grad_cosine.json; the tensors below 0.999
are the candidates for tests/golden/bf16_grad_cosine_exceptions.json (named exceptions with their measured value).

    python tools/diag_grad_cosine.py [L]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from modaltune_b200 import config, factory, synthetic, train_step  # noqa: E402
from oracle import modaltune_oracle as O  # noqa: E402

dev = "cuda"
L = int(sys.argv[1]) if len(sys.argv) > 1 else 520
torch.set_num_threads(os.cpu_count())


def cos(a, b):
    a, b = a.flatten().double().cpu(), b.flatten().double().cpu()
    return float(a @ b / (a.norm() * b.norm() + 1e-300))


out = {}
for tag, groups in (("toy10", factory.SMALL_GROUPS), ("full331", None)):
    model = factory.build_model(groups, device=dev)
    proj = factory.build_projector(0, dev)
    slide = synthetic.synthetic_slide(L, seed=77, group_sizes=groups)
    sd = {k: v.detach().cpu().clone().requires_grad_(v.requires_grad) for k, v in model.named_parameters()}
    genes = [slide["genes"][i] for i in range(len(slide["genes"]))]
    loss_o, logits_o = O.training_step(sd, synthetic.seeded_projector_state(0), slide["x"][0], slide["coords"][0], genes,
                                       slide["clinical"], slide["text"])
    loss_o.backward()
    gmax = max(float(v.grad.norm()) for v in sd.values() if v.grad is not None)
    keys = [k for k, p in model.named_parameters() if p.requires_grad and float(sd[k].grad.norm()) >= 1e-4 * gmax]
    dslide = train_step.slide_to_device(slide, dev)
    out[tag] = {"n_live": len(keys)}
    for mode, attn in (("fp32", "simt"), ("bf16", "simt"), ("bf16", "auto")):
        model.zero_grad()
        with config.using(mode=mode, attn_impl=attn):
            loss, logits = train_step.forward_backward(model, proj, dslide)
        g = dict(model.named_parameters())
        cs = {k: cos(g[k].grad, sd[k].grad) for k in keys}
        low = {k: round(c, 6) for k, c in sorted(cs.items(), key=lambda kv: kv[1]) if c < 0.9995}
        rel = float((logits.float().cpu() - logits_o.detach()).abs().max() / logits_o.detach().abs().max())
        out[tag][f"{mode}/{attn}"] = {"logits_rel": rel, "loss": float(loss), "loss_oracle": float(loss_o),
                                      "min_cos": min(cs.values()), "n_below_0.999": sum(c < 0.999 for c in cs.values()),
                                      "below_0.9995": low,
                                      "rel_norm_of_low": {k: float(sd[k].grad.norm()) / gmax for k in low}}
        print(tag, mode, attn, "logits rel %.2e" % rel, "min cos %.5f" % min(cs.values()),
              "#<0.999:", sum(c < 0.999 for c in cs.values()), "of", len(cs))
        for k, c in list(low.items())[:12]:
            print("   ", c, k, "|g|/gmax %.1e" % (float(sd[k].grad.norm()) / gmax))
    del model
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "grad_cosine.json"), "w") as f:
    json.dump(out, f, indent=1)
