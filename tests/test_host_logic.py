"""Host-side logic of the product (module plumbing + autograd wiring of the fused nodes) on CPU.

The CUDA kernels are replaced by the torch stand-ins of ``tests/cpu_kernels.py`` (test infrastructure); what is checked
is everything AROUND the kernels: parameter naming, the dX-only backward chain of the frozen encoder layer, the
Injector / Extractor plumbing, and the training step against the reference's golden fixture.
"""
import os

import pytest
import torch

from modaltune_b200 import config, synthetic, train_step
from tests import cpu_kernels, helpers


@pytest.fixture(scope="module")
def model():
    return helpers.build_model(helpers.SMALL_GROUPS)


def test_training_step_matches_reference_golden(model):
    gold = torch.load(os.path.join(helpers.GOLDEN, "training_step.pt"))["L300_float32"]
    slide = synthetic.synthetic_slide(gold["L"], seed=gold["seed"], group_sizes=gold["group_sizes"])
    proj = helpers.build_projector(0)
    model.zero_grad()
    with cpu_kernels.installed(), config.using(mode="fp32"):
        loss, logits = train_step.forward_backward(model, proj, slide)
    assert helpers.relerr(logits, gold["logits"]) < 1e-4
    assert abs(float(loss) - gold["loss"]) / abs(gold["loss"]) < 1e-3
    grads = {k: p.grad for k, p in model.named_parameters() if p.requires_grad}
    assert set(grads) == set(gold["grads"])
    gmax = max(v["norm"] for v in gold["grads"].values())
    for k, want in gold["grads"].items():
        got = helpers.grad_summary(k, grads[k])
        if want["norm"] < 1e-6 * gmax:  # structurally zero gradients (see make_golden.py)
            assert got["norm"] < 1e-4 * gmax, k
            continue
        assert abs(got["norm"] - want["norm"]) <= 2e-3 * want["norm"] + 1e-7 * gmax, k
        cos = torch.dot(got["proj"].double(), want["proj"].double()) / (got["proj"].norm() * want["proj"].norm() + 1e-30)
        assert cos > 0.999 or want["norm"] < 1e-4 * gmax, (k, float(cos))


def _step_grads(model, slide, proj):
    for p in model.parameters():
        p.grad = None
    loss, logits = train_step.forward_backward(model, proj, slide)
    return float(loss), logits.clone(), {n: p.grad.clone() for n, p in model.named_parameters()
                                         if p.requires_grad and p.grad is not None}


@pytest.mark.parametrize("flag", ["injector_fused", "shared_extractor_kv", "split_param_grads"])
def test_round2_restructurings_do_not_change_the_step(model, flag):
    """The host-side restructurings of round 2 -- the Injector as one autograd node with composed projections, one shared
    k | v projection for the three extractors of the last block (LayerNorm affine folded into the stacked weights), one
    gradient per (parameter, task pass) summed by multi-tensor adds -- against the path before each of them (its switch in
    ``modaltune_b200.config`` turned off): same logits, same gradient for every live parameter (fp32, kernel stand-ins)."""
    slide = synthetic.synthetic_slide(160, seed=21, group_sizes=helpers.SMALL_GROUPS)
    proj = helpers.build_projector(0)
    with cpu_kernels.installed(), config.using(mode="fp32"):
        on = _step_grads(model, slide, proj)
        old = config._state[flag]
        config._state[flag] = False
        try:
            off = _step_grads(model, slide, proj)
        finally:
            config._state[flag] = old
    assert helpers.relerr(on[1], off[1]) < 1e-5
    assert set(on[2]) == set(off[2])
    gmax = max(float(v.abs().max()) for v in off[2].values())
    for n, g in off[2].items():
        if float(g.abs().max()) < 1e-6 * gmax:      # structurally zero gradients: rounding noise on both sides
            continue
        assert helpers.relerr(on[2][n], g) < 2e-4, (flag, n)


def test_densify_and_gather_buffers_keep_every_gradient(model):
    """``FlatGradAllReduce.densify`` (single rank) and ``gather_buffers`` (what is all-reduced): the transposed slices of
    the stacked SNN gradients become dense parameter-layout tensors through ONE gather, values unchanged."""
    slide = synthetic.synthetic_slide(90, seed=22, group_sizes=helpers.SMALL_GROUPS)
    proj = helpers.build_projector(0)
    flat = train_step.FlatGradAllReduce([p for p in model.parameters() if p.requires_grad])
    with cpu_kernels.installed(), config.using(mode="fp32"):
        flat.zero()
        train_step.forward_backward(model, proj, slide)
    want = {id(p): p.grad.clone() for p in flat.params if p.grad is not None}
    assert any(not p.grad.is_contiguous() for p in flat.params if p.grad is not None)   # the cat-ed first SNN layer
    flat.densify()
    assert all(p.grad.is_contiguous() for p in flat.params if p.grad is not None)
    assert all(torch.equal(p.grad, want[id(p)]) for p in flat.params if id(p) in want)
    bufs = flat.gather_buffers()
    assert sum(b.numel() for b in bufs) == flat.numel
    assert all(torch.equal(p.grad, want[id(p)]) for p in flat.params if id(p) in want)
