"""Train-mode stochastic ops of the frozen encoder on the B200: Dropout(0.25) + DropPath on the two residual branches
(torchscale/architecture/encoder.py:149-152,169-170; feedforward_network.py:142) and on the embedded tokens (:339).

Masks cannot be bit-matched with the reference's torch generator, so the checks are:
* statistics of the counter-based masks, their independence across seeds / call sites, and that the backward kernel
  regenerates exactly the mask of the forward kernels;
* a train-mode layer forward + backward against the oracle with the SAME masks (read back from the kernels) injected.
"""
import pytest
import torch

from modaltune_b200 import config, ops
from modaltune_b200.slide_encoder import DILATED_RATIO, optimal_segment_lengths
from oracle import modaltune_oracle as O
from tests import helpers

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _factors(drop, rows=512, cols=768):
    """keep / (1 - p) * path_scale of every element, read back through the forward kernel: 0 + D(1)."""
    zeros = torch.zeros(rows, cols, device=DEV)
    ones = torch.ones(rows, cols, device=DEV)
    return ops.residual_bias_add(zeros, ones, None, drop=drop)


def test_dropout_masks_statistics_and_consistency():
    seed = torch.tensor([1234567891011], device=DEV, dtype=torch.int64)
    d = ops.DropSpec(0.25, seed, 6)
    m = _factors(d)
    kept = (m > 0).float().mean().item()
    assert abs(kept - 0.75) < 0.01, kept
    assert torch.equal(m[m > 0], torch.full_like(m[m > 0], 1.0 / 0.75))
    assert abs(m.mean().item() - 1.0) < 0.02                                   # unbiased
    assert abs((m[:, ::2] * m[:, 1::2]).mean().item() - 1.0) < 0.03            # neighbours are uncorrelated
    # the other forward kernel and the backward kernel see the same mask
    x = torch.randn(512, 768, device=DEV)
    a = torch.randn(512, 768, device=DEV)
    g, b = torch.ones(768, device=DEV), torch.zeros(768, device=DEV)
    x_out, _, _, _ = ops.add_layernorm_fwd(x, a, g, b, torch.float32, drop=d)
    assert torch.allclose(x_out, x + a * m, atol=1e-6)
    for dt in (torch.float32, torch.bfloat16):
        gm = ops.dropout_bwd_cast(a, dt, d)
        assert torch.equal(gm, (a * m).to(dt))
    # independent masks per call site and per seed; DropPath scales / kills the whole branch
    assert not torch.equal(_factors(ops.DropSpec(0.25, seed, 7)), m)
    assert not torch.equal(_factors(ops.DropSpec(0.25, seed + 1, 6)), m)
    ps = torch.tensor([1.0 / 0.9], device=DEV)
    assert torch.allclose(_factors(ops.DropSpec(0.25, seed, 6, ps)), m / 0.9)
    assert float(_factors(ops.DropSpec(0.25, seed, 6, torch.zeros(1, device=DEV))).abs().max()) == 0.0
    assert torch.equal(_factors(ops.DropSpec(0.0, seed, 6)), torch.ones_like(m))


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_encoder_layer_train_mode_against_oracle_with_the_same_masks(mode, tol, monkeypatch):
    model = helpers.build_model(helpers.SMALL_GROUPS, device=DEV)
    l, N = 7, 701
    layer = model.encoder.layers[l]
    assert layer.dropout == 0.25 and abs(layer.drop_path_prob - 0.1 * l / 11) < 1e-9   # config of the reference JSON
    captured = []
    real = ops.train_rng

    def spy(*a, **k):
        captured.append(real(*a, **k))
        return captured[-1]

    monkeypatch.setattr(ops, "train_rng", spy)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, N, 768, generator=g).to(DEV).requires_grad_(True)
    dy = torch.randn(1, N, 768, generator=g).to(DEV)
    torch.manual_seed(11)
    layer.train()
    try:
        with config.using(mode=mode, attn_impl="simt" if mode == "fp32" else "auto"):
            y, _ = layer(x)
            (gx,) = torch.autograd.grad(y, x, dy)
            y_eval_like = None
    finally:
        layer.eval()
    assert len(captured) == 1
    masks = [_factors(d, rows=N).cpu().double() for d in captured[0]]
    assert 0.70 < (masks[0] > 0).double().mean() < 0.80 or float(masks[0].abs().max()) == 0.0   # DropPath may kill a branch
    sd = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
    xo = x.detach().cpu().double()[0].requires_grad_(True)
    yo = O.encoder_layer(sd, l, xo, optimal_segment_lengths(), DILATED_RATIO, masks=masks)
    (gxo,) = torch.autograd.grad(yo, xo, dy.cpu().double()[0])
    assert helpers.relerr(y[0].cpu(), yo) < tol, helpers.relerr(y[0].cpu(), yo)
    assert helpers.relerr(gx[0].cpu(), gxo) < 2 * tol, helpers.relerr(gx[0].cpu(), gxo)
    # and the masks did something: the eval-mode output differs
    with config.using(mode=mode, attn_impl="simt" if mode == "fp32" else "auto"):
        y_eval, _ = layer(x)
    assert helpers.relerr(y_eval[0].cpu(), yo) > 10 * tol


def test_train_mode_step_is_stochastic_and_eval_is_not():
    """Whole adapter in train(): embedding dropout + per-layer masks make two forwards differ, gradients stay finite;
    eval() stays deterministic."""
    from modaltune_b200 import synthetic, train_step
    model = helpers.build_model(helpers.SMALL_GROUPS, device=DEV)
    slide = train_step.slide_to_device(synthetic.synthetic_slide(600, seed=3, group_sizes=helpers.SMALL_GROUPS), DEV)
    proj = helpers.build_projector(0, DEV)
    model.train()
    try:
        torch.manual_seed(0)
        a = train_step.multitask_forward(model, slide)
        b = train_step.multitask_forward(model, slide)
        assert helpers.relerr(a, b) > 1e-3
        loss = train_step.distill_loss(b, train_step.text_targets(proj, slide["text"]))
        loss.backward()
        grads = [p.grad for p in model.parameters() if p.requires_grad and p.grad is not None]
        assert len(grads) > 100 and all(torch.isfinite(gr).all() for gr in grads)
    finally:
        model.eval()
    with torch.no_grad():
        assert torch.equal(train_step.multitask_forward(model, slide), train_step.multitask_forward(model, slide))
