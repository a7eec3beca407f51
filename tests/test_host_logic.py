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
