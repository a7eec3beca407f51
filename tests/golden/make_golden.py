"""Pin the oracle against the UNMODIFIED reference and write the golden fixtures.  Build-container only.

    python tests/golden/make_golden.py            # ~10 min on 8 cores

1. runs reference modules (``/root/reference`` + ``ref_shims``) and ``oracle/modaltune_oracle.py`` side by side in
   fp64 on identical seeded weights / inputs and records the worst disagreement per component in
   ``tests/golden/ORACLE_VALIDATION.json``;
2. stores REFERENCE outputs (fp32 model, the dtype the reference would run on CPU) as compact fixtures
   ``tests/golden/*.pt`` that ``tests/test_oracle_golden.py`` replays against the oracle everywhere (no reference
   needed), and that the ``-m gpu`` tests replay against the CUDA path.

Inputs and weights are never stored: they are regenerated from seeds by ``modaltune_b200/synthetic.py``.
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
from modaltune_b200 import synthetic  # noqa: E402
from oracle import modaltune_oracle as O  # noqa: E402

torch.set_num_threads(os.cpu_count())
SMALL_GROUPS = [3, 5, 7, 2, 9, 4, 6, 8, 1, 12]
report = {}


def relerr(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))


def grad_summary(name, g, k=64):
    """Compact, seed-reproducible description of a gradient tensor: norm, k sampled entries, 4 random projections."""
    flat = g.flatten().double()
    gen = torch.Generator().manual_seed(zlib_seed(name))
    idx = torch.randint(0, flat.numel(), (min(k, flat.numel()),), generator=gen)
    proj = torch.stack([(torch.randn(flat.numel(), generator=gen, dtype=torch.float64) @ flat) for _ in range(4)])
    return {"norm": float(flat.norm()), "idx": idx, "vals": flat[idx].float(), "proj": proj.float()}


def zlib_seed(name):
    import zlib
    return zlib.crc32(name.encode()) % (2**31)


def main():
    t0 = time.time()
    torch.manual_seed(0)
    model, cfg = ref_shims.build_reference_model(clinical=True, multi_task=3, gene_group_sizes=SMALL_GROUPS)
    model.eval()
    synthetic.seeded_init_(model.named_parameters(), seed=0)
    sd32 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    seglens = [int(s) for s in model.encoder.args.segment_length]
    assert seglens == O.optimal_segment_lengths(), seglens
    print(f"reference built in {time.time() - t0:.0f}s; segment lengths {seglens}")

    # ---- positional table ------------------------------------------------------------------------------------------
    tab = O.sincos_table()
    pe = model.pos_embed[0]
    gen = torch.Generator().manual_seed(1)
    ii, jj = torch.randint(0, 1000, (512,), generator=gen), torch.randint(0, 1000, (512,), generator=gen)
    want = pe[1 + ii * 1000 + jj]
    got = torch.cat([tab[jj], tab[ii]], -1)
    report["pos_embed_max_abs"] = float((want - got).abs().max())
    report["pos_embed_row0_abs"] = float(pe[0].abs().max())
    print("pos table:", report["pos_embed_max_abs"], report["pos_embed_row0_abs"])

    # ---- dilated attention (fp64), several geometries --------------------------------------------------------------
    layer0 = model.encoder.layers[0]
    attn = layer0.self_attn
    attn64 = __import__("copy").deepcopy(attn).double()
    sd64_l0 = {"a." + k: v.double() for k, v in attn.state_dict().items()}
    fixtures = {}
    cases = [  # (N, segment_lengths or None for the real ones)
        (75, [16, 24, 32, 64, 128]), (97, [8, 32, 64, 96, 256]), (130, [32, 64, 128, 256, 512]),
        (1025, None), (1500, None), (5793, None), (2049, [256, 512, 1024, 4096, 8192]),
    ]
    worst_o, worst_g = 0.0, 0.0
    for N, sl in cases:
        sl_use = sl or seglens
        attn64.args.segment_length = sl_use
        gen = torch.Generator().manual_seed(100 + N)
        x = torch.randn(1, N, 768, generator=gen, dtype=torch.float64).requires_grad_(True)
        dy = torch.randn(1, N, 768, generator=gen, dtype=torch.float64)
        y_ref, _ = attn64(x, x, x)
        (gx_ref,) = torch.autograd.grad(y_ref, x, dy)
        x2 = x.detach().clone().requires_grad_(True)
        y_or = O.dilated_self_attention(sd64_l0, "a", x2[0], sl_use, O.DILATED_RATIO)
        (gx_or,) = torch.autograd.grad(y_or, x2, dy[0])
        eo, eg = relerr(y_or, y_ref[0]), relerr(gx_or[0], gx_ref[0])
        worst_o, worst_g = max(worst_o, eo), max(worst_g, eg)
        print(f"dilated attn N={N} sl={sl_use}: out rel {eo:.2e} grad rel {eg:.2e}")
        if N <= 2049:
            rows = torch.arange(0, N, max(1, N // 48))
            fixtures[f"N{N}"] = {"N": N, "segment_lengths": sl_use, "seed": 100 + N, "rows": rows,
                                 "y_rows": y_ref[0][rows].float(), "gx_rows": gx_ref[0][rows].float(),
                                 "y_norm": float(y_ref.norm()), "gx_norm": float(gx_ref.norm())}
    attn64.args.segment_length = seglens
    report["dilated_attention_out_rel_fp64"] = worst_o
    report["dilated_attention_grad_rel_fp64"] = worst_g
    torch.save({"layer": "encoder.layers.0.self_attn", "weight_seed": 0, "cases": fixtures},
               os.path.join(HERE, "dilated_attention.pt"))

    # ---- encoder layer (fp64; GELU is fp32 inside the reference) ----------------------------------------------------
    lay64 = __import__("copy").deepcopy(model.encoder.layers[5]).double()
    sd64 = {k: v.double() for k, v in sd32.items() if k.startswith("encoder.layers.5.")}
    N = 1200
    gen = torch.Generator().manual_seed(7)
    x = torch.randn(1, N, 768, generator=gen, dtype=torch.float64).requires_grad_(True)
    dy = torch.randn(1, N, 768, generator=gen, dtype=torch.float64)
    y_ref, _ = lay64(x, encoder_padding_mask=torch.zeros(1, N).bool())
    (gx_ref,) = torch.autograd.grad(y_ref, x, dy)
    x2 = x.detach().clone().requires_grad_(True)
    y_or = O.encoder_layer(sd64, 5, x2[0], seglens, O.DILATED_RATIO)
    (gx_or,) = torch.autograd.grad(y_or, x2, dy[0])
    report["encoder_layer_out_rel_fp64"] = relerr(y_or, y_ref[0])
    report["encoder_layer_grad_rel_fp64"] = relerr(gx_or[0], gx_ref[0])
    print("encoder layer:", report["encoder_layer_out_rel_fp64"], report["encoder_layer_grad_rel_fp64"])
    rows = torch.arange(0, N, 25)
    torch.save({"layer": 5, "weight_seed": 0, "N": N, "seed": 7, "rows": rows, "y_rows": y_ref[0][rows].float(),
                "gx_rows": gx_ref[0][rows].float(), "y_norm": float(y_ref.norm()), "gx_norm": float(gx_ref.norm())},
               os.path.join(HERE, "encoder_layer.pt"))

    # ---- injector / extractor / prompt SA (fp64) ---------------------------------------------------------------------
    blk = __import__("copy").deepcopy(model.interactions[2]).double()
    sd64 = {k: v.double() for k, v in sd32.items()}
    L, M = 700, 13
    gen = torch.Generator().manual_seed(11)
    xs = torch.randn(1, L, 768, generator=gen, dtype=torch.float64).requires_grad_(True)
    cs = torch.randn(1, M, 768, generator=gen, dtype=torch.float64).requires_grad_(True)
    pe64 = torch.randn(M, 768, generator=gen, dtype=torch.float64) * 0.02
    dyx = torch.randn(1, L, 768, generator=gen, dtype=torch.float64)
    dyc = torch.randn(1, M, 768, generator=gen, dtype=torch.float64)
    y_ref = blk.injector(query=xs, feat=cs, pos=pe64)
    g_ref = torch.autograd.grad(y_ref, [xs, cs], dyx)
    y_or = O.injector(sd64, "interactions.2.injector", xs[0], cs[0], pe64)
    g_or = torch.autograd.grad(y_or, [xs, cs], dyx[0])
    report["injector_out_rel_fp64"] = relerr(y_or, y_ref[0])
    report["injector_grad_rel_fp64"] = max(relerr(a[0], b[0]) for a, b in zip(g_or, g_ref))
    inj_fix = {"y_rows": y_ref[0][::20].float(), "gx_rows": g_ref[0][0][::20].float(), "gc": g_ref[1][0].float()}
    y_ref = blk.extractor(query=cs, feat=xs, pos=pe64)
    g_ref = torch.autograd.grad(y_ref, [xs, cs], dyc)
    y_or = O.extractor(sd64, "interactions.2.extractor", cs[0], xs[0], pe64)
    g_or = torch.autograd.grad(y_or, [xs, cs], dyc[0])
    report["extractor_out_rel_fp64"] = relerr(y_or, y_ref[0])
    report["extractor_grad_rel_fp64"] = max(relerr(a[0], b[0]) for a, b in zip(g_or, g_ref))
    ext_fix = {"y": y_ref[0].float(), "gx_rows": g_ref[0][0][::20].float(), "gc": g_ref[1][0].float()}
    psa = __import__("copy").deepcopy(model.prompt_selfattention[1]).double()
    y_ref = psa(cs, pe64)
    y_or = O.prompt_self_attention(sd64, "prompt_selfattention.1", cs[0], pe64)
    report["prompt_sa_out_rel_fp64"] = relerr(y_or, y_ref[0])
    print({k: v for k, v in report.items() if "injector" in k or "extractor" in k or "prompt" in k})
    torch.save({"block": 2, "weight_seed": 0, "L": L, "M": M, "seed": 11, "injector": inj_fix, "extractor": ext_fix,
                "prompt_sa_y": y_ref[0].float()}, os.path.join(HERE, "adapter_blocks.pt"))

    # ---- full training step: 3 task passes + loss + grads -------------------------------------------------------------
    proj_sd = synthetic.seeded_projector_state(0)
    from train_modaltune import Projection_layer  # the reference's own projector class
    projector = Projection_layer(512, 256)
    projector.load_state_dict(proj_sd)

    def ref_step(model, slide, dtype):
        kw = dict(x=slide["x"].to(dtype), coords=slide["coords"].to(dtype),
                  genes={k: v.to(dtype) for k, v in slide["genes"].items()}, clinical=slide["clinical"].to(dtype))
        logits = torch.cat([model(**kw, task_token=torch.eye(3, dtype=dtype)[t]) for t in range(3)], 0)
        text = projector.to(dtype)(slide["text"].to(dtype))
        text = text / text.norm(dim=-1, keepdim=True)
        z = logits / logits.norm(dim=-1, keepdim=True)
        loss = torch.nn.KLDivLoss(reduction="sum")(torch.nn.functional.log_softmax(z, dim=1),
                                                   torch.nn.functional.softmax(text[[0, 1, 3], :], dim=1)) * 10
        return loss, logits

    steps = {}
    for L, dtype in [(300, torch.float64), (300, torch.float32), (1100, torch.float32)]:
        slide = synthetic.synthetic_slide(L, seed=2000 + L, group_sizes=SMALL_GROUPS)
        model.to(dtype)
        model.zero_grad()
        t1 = time.time()
        loss, logits = ref_step(model, slide, dtype)
        loss.backward()
        t_ref = time.time() - t1
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.requires_grad}
        sd = {k: v.detach().to(dtype).requires_grad_(k in grads) for k, v in model.state_dict().items()
              if k != "pos_embed"}
        genes = [slide["genes"][i].to(dtype) for i in range(len(SMALL_GROUPS))]
        t1 = time.time()
        loss_o, logits_o = O.training_step(sd, proj_sd, slide["x"][0].to(dtype), slide["coords"][0].to(dtype), genes,
                                           slide["clinical"].to(dtype), slide["text"].to(dtype))
        loss_o.backward()
        t_or = time.time() - t1
        # gradients that are structurally zero (a bias that is constant along an axis the next LayerNorm removes:
        # gene_encoder.mlp_mixer.*.0.fn.3.bias, gene_encoder.pathway_compression.bias) are rounding noise on both
        # sides; they are required to be tiny, not to be parallel.
        gmax = max(float(g.norm()) for g in grads.values())
        dead = sorted(k for k, g in grads.items() if float(g.norm()) < 1e-6 * gmax)
        for k in dead:
            assert float(sd[k].grad.norm()) < 1e-5 * gmax, k
        report[f"step_L{L}_{str(dtype).split('.')[-1]}_structurally_zero_grads"] = dead
        coss = {k: cos(sd[k].grad, grads[k]) for k in grads if k not in dead}
        tag = f"L{L}_{str(dtype).split('.')[-1]}"
        report[f"step_{tag}_logits_rel"] = relerr(logits_o.detach(), logits.detach())
        report[f"step_{tag}_loss_rel"] = abs(float(loss_o) - float(loss)) / abs(float(loss))
        report[f"step_{tag}_min_grad_cos"] = min(coss.values())
        report[f"step_{tag}_max_grad_rel"] = max(relerr(sd[k].grad, grads[k]) for k in coss)
        report[f"step_{tag}_n_trainable_tensors"] = len(grads)
        report[f"step_{tag}_seconds_ref_vs_oracle"] = [round(t_ref, 1), round(t_or, 1)]
        print(tag, {k: v for k, v in report.items() if tag in k})
        if dtype == torch.float32:
            steps[tag] = {"L": L, "seed": 2000 + L, "weight_seed": 0, "group_sizes": SMALL_GROUPS,
                          "logits": logits.detach().float(), "loss": float(loss),
                          "grads": {k: grad_summary(k, g) for k, g in grads.items()}}
    torch.save(steps, os.path.join(HERE, "training_step.pt"))
    model.float()

    report["reference_commit"] = "martellab-sri/ModalTune (mounted read-only at /root/reference)"
    report["torch"] = torch.__version__
    report["wall_seconds"] = round(time.time() - t0, 1)
    with open(os.path.join(HERE, "ORACLE_VALIDATION.json"), "w") as f:
        json.dump(report, f, indent=1, sort_keys=True)
    print(json.dumps(report, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
