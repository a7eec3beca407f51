"""Round-2 golden fixtures at the BENCH sizes, written from the UNMODIFIED reference.  Build-container only.

    python tests/golden/make_golden_r2.py            # ~15 min on 8 cores, peak ~35 GB of host memory

What it adds to ``make_golden.py`` (whose fixtures stop at N = 2049 tokens / 10 toy pathways):

* ``encoder_layer_10k.pt``  -- reference ``EncoderLayer`` (layer 5, fp32) forward + input gradient at N = 10 001 tokens
  (BASELINE config 2): sampled rows of y / gx, norms.  This is the geometry ``bench.py`` runs.
* ``encoder_layer_32k.pt``  -- the same at N = 32 769 (config 3: 33 x 1024, 6 x 5792 and the 2 x 32 768 branch whose
  second segment is all padding).  The reference forward runs under ``no_grad`` (its autograd graph over the padded
  [2, 16, 8192, 8192] score tensors does not fit this container); the input gradient comes from the oracle, which the
  same script first checks against the reference forward at this size (and which ``make_golden.py`` pins against the
  reference gradient to 1e-11 at smaller N).  ``gx_source`` in the fixture says so.
* ``training_step_331.pt`` -- one full training step (3 task passes + KL loss + backward) of the reference with the
  real 331-pathway gene encoder (sizes of ``dataset/gene_pathway_processed_v2.csv``) at 400 tiles, fp32: logits, loss
  and the compact gradient summaries of all trainable tensors.

Inputs and weights are regenerated from seeds (``modaltune_b200/synthetic.py``); nothing but outputs is stored.
"""
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shims  # noqa: E402
from make_golden import grad_summary, relerr  # noqa: E402
from modaltune_b200 import synthetic  # noqa: E402
from oracle import modaltune_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count())
LAYER = 5


def layer_fixture(model, sd32, N, seed, with_ref_grad):
    lay = model.encoder.layers[LAYER]
    seglens = [int(s) for s in model.encoder.args.segment_length]
    sd = {k: v for k, v in sd32.items() if k.startswith(f"encoder.layers.{LAYER}.")}
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(1, N, 768, generator=gen, dtype=torch.float32)
    dy = torch.randn(1, N, 768, generator=gen, dtype=torch.float32)
    mask = torch.zeros(1, N).bool()
    t0 = time.time()
    if with_ref_grad:
        xr = x.clone().requires_grad_(True)
        y_ref, _ = lay(xr, encoder_padding_mask=mask)
        (gx_ref,) = torch.autograd.grad(y_ref, xr, dy)
        y_ref, gx_ref = y_ref.detach()[0], gx_ref[0]
    else:
        with torch.no_grad():
            y_ref, _ = lay(x, encoder_padding_mask=mask)
        y_ref, gx_ref = y_ref[0], None
    t_ref = time.time() - t0
    t0 = time.time()
    xo = x[0].clone().requires_grad_(True)
    y_or = O.encoder_layer(sd, LAYER, xo, seglens, O.DILATED_RATIO)
    (gx_or,) = torch.autograd.grad(y_or, xo, dy[0])
    t_or = time.time() - t0
    rep = {"N": N, "out_rel_fp32": relerr(y_or.detach(), y_ref), "seconds_ref_vs_oracle": [round(t_ref, 1), round(t_or, 1)]}
    if gx_ref is not None:
        rep["grad_rel_fp32"] = relerr(gx_or, gx_ref)
        assert rep["grad_rel_fp32"] < 1e-4, rep
    assert rep["out_rel_fp32"] < 1e-4, rep
    gx = gx_ref if gx_ref is not None else gx_or
    rows = torch.arange(0, N, max(1, N // 400))
    fix = {"layer": LAYER, "weight_seed": 0, "N": N, "seed": seed, "rows": rows, "y_rows": y_ref[rows].clone(),
           "gx_rows": gx[rows].clone(), "y_norm": float(y_ref.double().norm()), "gx_norm": float(gx.double().norm()),
           "gx_source": "reference autograd" if gx_ref is not None else
           "oracle autograd (reference forward checked at this size; reference gradient pinned at N <= 5793)"}
    return fix, rep


def main():
    t0 = time.time()
    report = {}
    torch.manual_seed(0)
    model, cfg = ref_shims.build_reference_model(clinical=True, multi_task=3, gene_group_sizes=None)   # 331 pathways
    model.eval()
    synthetic.seeded_init_(model.named_parameters(), seed=0)
    sd32 = {k: v.detach().clone() for k, v in model.state_dict().items() if k != "pos_embed"}
    print(f"reference (331 pathways) built in {time.time() - t0:.0f}s")

    fix, rep = layer_fixture(model, sd32, 10001, 71, with_ref_grad=True)
    torch.save(fix, os.path.join(HERE, "encoder_layer_10k.pt"))
    report["encoder_layer_10k"] = rep
    print("layer 10k:", rep)

    fix, rep = layer_fixture(model, sd32, 32769, 72, with_ref_grad=False)
    torch.save(fix, os.path.join(HERE, "encoder_layer_32k.pt"))
    report["encoder_layer_32k"] = rep
    print("layer 32k:", rep)

    # ---- full step with the 331-pathway gene encoder ----------------------------------------------------------------
    from train_modaltune import Projection_layer  # the reference's own projector class
    proj_sd = synthetic.seeded_projector_state(0)
    projector = Projection_layer(512, 256)
    projector.load_state_dict(proj_sd)
    L, seed = 400, 2400
    slide = synthetic.synthetic_slide(L, seed=seed)
    model.zero_grad()
    t1 = time.time()
    logits = torch.cat([model(x=slide["x"], coords=slide["coords"], genes=slide["genes"], clinical=slide["clinical"],
                              task_token=torch.eye(3)[t]) for t in range(3)], 0)
    text = projector(slide["text"])
    text = text / text.norm(dim=-1, keepdim=True)
    z = logits / logits.norm(dim=-1, keepdim=True)
    loss = torch.nn.KLDivLoss(reduction="sum")(torch.nn.functional.log_softmax(z, dim=1),
                                               torch.nn.functional.softmax(text[[0, 1, 3], :], dim=1)) * 10
    loss.backward()
    t_ref = time.time() - t1
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.requires_grad}
    sd = {k: v.detach().clone().requires_grad_(k in grads) for k, v in sd32.items()}
    genes = [slide["genes"][i] for i in range(len(slide["genes"]))]
    t1 = time.time()
    loss_o, logits_o = O.training_step(sd, proj_sd, slide["x"][0], slide["coords"][0], genes, slide["clinical"], slide["text"])
    loss_o.backward()
    t_or = time.time() - t1
    gmax = max(float(g.norm()) for g in grads.values())
    dead = sorted(k for k, g in grads.items() if float(g.norm()) < 1e-6 * gmax)
    cos = lambda a, b: float((a.flatten().double() @ b.flatten().double()) / (a.double().norm() * b.double().norm() + 1e-300))
    coss = {k: cos(sd[k].grad, grads[k]) for k in grads if k not in dead}
    report["step_331"] = {"L": L, "logits_rel": relerr(logits_o.detach(), logits.detach()),
                          "loss_rel": abs(float(loss_o) - float(loss)) / abs(float(loss)),
                          "min_grad_cos": min(coss.values()), "n_trainable_tensors": len(grads),
                          "structurally_zero_grads": dead, "seconds_ref_vs_oracle": [round(t_ref, 1), round(t_or, 1)]}
    print("step 331:", report["step_331"])
    assert report["step_331"]["logits_rel"] < 1e-5 and report["step_331"]["min_grad_cos"] > 0.9999
    torch.save({"L": L, "seed": seed, "weight_seed": 0, "logits": logits.detach().float(), "loss": float(loss),
                "dead": dead, "grads": {k: grad_summary(k, g) for k, g in grads.items()}},
               os.path.join(HERE, "training_step_331.pt"))
    report["wall_seconds"] = round(time.time() - t0, 1)
    report["torch"] = torch.__version__
    with open(os.path.join(HERE, "ORACLE_VALIDATION_R2.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(report, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
