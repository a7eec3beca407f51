"""The oracle against the REFERENCE's outputs (fixtures written by tests/golden/make_golden.py from the unmodified
reference running in the build container) -- replayed everywhere, no reference needed."""
import json
import os

import pytest
import torch

from modaltune_b200 import synthetic
from oracle import modaltune_oracle as O
from tests import helpers


@pytest.fixture(scope="module")
def sd():
    model = helpers.build_model(helpers.SMALL_GROUPS)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def test_validation_report_is_tight():
    rep = json.load(open(os.path.join(helpers.GOLDEN, "ORACLE_VALIDATION.json")))
    assert rep["dilated_attention_out_rel_fp64"] < 1e-12 and rep["dilated_attention_grad_rel_fp64"] < 1e-12
    assert rep["encoder_layer_out_rel_fp64"] < 1e-9 and rep["injector_out_rel_fp64"] < 1e-12
    assert rep["step_L300_float64_logits_rel"] < 1e-6 and rep["step_L300_float64_min_grad_cos"] > 0.999999
    assert rep["pos_embed_max_abs"] == 0.0


def test_segment_lengths():
    assert O.optimal_segment_lengths() == [1024, 5792, 32768, 185363, 1048576]


def test_dilated_attention_fixture(sd):
    gold = torch.load(os.path.join(helpers.GOLDEN, "dilated_attention.pt"))
    p = gold["layer"]
    sd64 = {k: v.double() for k, v in sd.items() if k.startswith(p)}
    for name, case in gold["cases"].items():
        N = case["N"]
        g = torch.Generator().manual_seed(case["seed"])
        x = torch.randn(1, N, 768, generator=g, dtype=torch.float64)[0].requires_grad_(True)
        dy = torch.randn(1, N, 768, generator=g, dtype=torch.float64)[0]
        y = O.dilated_self_attention(sd64, p, x, case["segment_lengths"], O.DILATED_RATIO)
        (gx,) = torch.autograd.grad(y, x, dy)
        assert helpers.relerr(y[case["rows"]], case["y_rows"]) < 1e-6, name
        assert helpers.relerr(gx[case["rows"]], case["gx_rows"]) < 1e-6, name
        assert abs(float(y.norm()) - case["y_norm"]) < 1e-9 * case["y_norm"], name


def test_encoder_layer_fixture(sd):
    gold = torch.load(os.path.join(helpers.GOLDEN, "encoder_layer.pt"))
    l, N = gold["layer"], gold["N"]
    sd64 = {k: v.double() for k, v in sd.items() if k.startswith(f"encoder.layers.{l}.")}
    g = torch.Generator().manual_seed(gold["seed"])
    x = torch.randn(1, N, 768, generator=g, dtype=torch.float64)[0].requires_grad_(True)
    dy = torch.randn(1, N, 768, generator=g, dtype=torch.float64)[0]
    y = O.encoder_layer(sd64, l, x, O.optimal_segment_lengths(), O.DILATED_RATIO)
    (gx,) = torch.autograd.grad(y, x, dy)
    assert helpers.relerr(y[gold["rows"]], gold["y_rows"]) < 1e-6
    assert helpers.relerr(gx[gold["rows"]], gold["gx_rows"]) < 1e-6


def test_adapter_blocks_fixture(sd):
    gold = torch.load(os.path.join(helpers.GOLDEN, "adapter_blocks.pt"))
    L, M, b = gold["L"], gold["M"], gold["block"]
    sd64 = {k: v.double() for k, v in sd.items()}
    g = torch.Generator().manual_seed(gold["seed"])
    xs = torch.randn(1, L, 768, generator=g, dtype=torch.float64)[0].requires_grad_(True)
    cs = torch.randn(1, M, 768, generator=g, dtype=torch.float64)[0].requires_grad_(True)
    pe = torch.randn(M, 768, generator=g, dtype=torch.float64) * 0.02
    dyx = torch.randn(1, L, 768, generator=g, dtype=torch.float64)[0]
    dyc = torch.randn(1, M, 768, generator=g, dtype=torch.float64)[0]
    y = O.injector(sd64, f"interactions.{b}.injector", xs, cs, pe)
    gx, gc = torch.autograd.grad(y, [xs, cs], dyx)
    assert helpers.relerr(y[::20], gold["injector"]["y_rows"]) < 1e-6
    assert helpers.relerr(gx[::20], gold["injector"]["gx_rows"]) < 1e-6
    assert helpers.relerr(gc, gold["injector"]["gc"]) < 1e-6
    y = O.extractor(sd64, f"interactions.{b}.extractor", cs, xs, pe)
    gx, gc = torch.autograd.grad(y, [xs, cs], dyc)
    assert helpers.relerr(y, gold["extractor"]["y"]) < 1e-6
    assert helpers.relerr(gx[::20], gold["extractor"]["gx_rows"]) < 1e-6
    assert helpers.relerr(gc, gold["extractor"]["gc"]) < 1e-6
    y = O.prompt_self_attention(sd64, "prompt_selfattention.1", cs, pe)
    assert helpers.relerr(y, gold["prompt_sa_y"]) < 1e-6


def test_training_step_fixture(sd):
    gold = torch.load(os.path.join(helpers.GOLDEN, "training_step.pt"))["L300_float32"]
    slide = synthetic.synthetic_slide(gold["L"], seed=gold["seed"], group_sizes=gold["group_sizes"])
    model = helpers.build_model(helpers.SMALL_GROUPS)
    sdg = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in model.named_parameters()}
    genes = [slide["genes"][i] for i in range(len(gold["group_sizes"]))]
    loss, logits = O.training_step(sdg, synthetic.seeded_projector_state(0), slide["x"][0], slide["coords"][0], genes,
                                   slide["clinical"], slide["text"])
    loss.backward()
    assert helpers.relerr(logits.detach(), gold["logits"]) < 1e-5
    assert abs(float(loss) - gold["loss"]) < 1e-3 * abs(gold["loss"])
    gmax = max(v["norm"] for v in gold["grads"].values())
    for k, want in gold["grads"].items():
        if want["norm"] < 1e-4 * gmax:
            continue
        got = helpers.grad_summary(k, sdg[k].grad)
        assert abs(got["norm"] - want["norm"]) < 1e-3 * want["norm"], k
        assert helpers.relerr(got["proj"], want["proj"]) < 5e-3, k


def test_encoder_layer_fixture_bench_size():
    """Round-2 fixture from the reference at N = 10 001 tokens (BASELINE config 2, fp32): oracle forward + dX."""
    gold = torch.load(os.path.join(helpers.GOLDEN, "encoder_layer_10k.pt"))
    l, N = gold["layer"], gold["N"]
    sd = {k: v.detach() for k, v in helpers.build_model(helpers.SMALL_GROUPS).state_dict().items()
          if k.startswith(f"encoder.layers.{l}.")}
    g = torch.Generator().manual_seed(gold["seed"])
    x = torch.randn(1, N, 768, generator=g, dtype=torch.float32)[0].requires_grad_(True)
    dy = torch.randn(1, N, 768, generator=g, dtype=torch.float32)[0]
    y = O.encoder_layer(sd, l, x, O.optimal_segment_lengths(), O.DILATED_RATIO)
    (gx,) = torch.autograd.grad(y, x, dy)
    assert gold["gx_source"] == "reference autograd"
    assert helpers.relerr(y[gold["rows"]], gold["y_rows"]) < 1e-5
    assert helpers.relerr(gx[gold["rows"]], gold["gx_rows"]) < 1e-5
    assert abs(float(gx.double().norm()) - gold["gx_norm"]) < 1e-5 * gold["gx_norm"]


def test_training_step_331_pathways_fixture():
    """Round-2 fixture: the reference's full training step with the real 331-pathway gene encoder."""
    gold = torch.load(os.path.join(helpers.GOLDEN, "training_step_331.pt"))
    slide = synthetic.synthetic_slide(gold["L"], seed=gold["seed"])
    model = helpers.build_model(None)
    sdg = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in model.named_parameters()}
    genes = [slide["genes"][i] for i in range(len(slide["genes"]))]
    loss, logits = O.training_step(sdg, synthetic.seeded_projector_state(0), slide["x"][0], slide["coords"][0], genes,
                                   slide["clinical"], slide["text"])
    loss.backward()
    assert len(gold["grads"]) == 1550
    assert helpers.relerr(logits.detach(), gold["logits"]) < 1e-5
    assert abs(float(loss) - gold["loss"]) < 1e-3 * abs(gold["loss"])
    gmax = max(v["norm"] for v in gold["grads"].values())
    for k, want in gold["grads"].items():
        if want["norm"] < 1e-4 * gmax:
            continue
        got = helpers.grad_summary(k, sdg[k].grad)
        assert abs(got["norm"] - want["norm"]) < 1e-3 * want["norm"], k
        assert helpers.relerr(got["proj"], want["proj"]) < 5e-3, k
