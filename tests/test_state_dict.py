"""Checkpoint compatibility with the reference: identical parameter names / shapes / trainable set (fixture written by
tests/golden/make_state_dict_keys.py from the unmodified reference), and load_state_dict round trips."""
import json
import os

import pytest
import torch

from modaltune_b200 import synthetic
from tests import helpers

KEYS = json.load(open(os.path.join(helpers.GOLDEN, "state_dict_keys.json")))


@pytest.mark.parametrize("clinical", [True, False])
def test_state_dict_matches_reference(clinical):
    name = "longnetvit_gene_clinical_adapter" if clinical else "longnetvit_gene_adapter"
    model = helpers.build_model(None, clinical=clinical)
    want = KEYS[name]
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == want["state_dict"]
    assert sorted(k for k, p in model.named_parameters() if p.requires_grad) == want["trainable"]
    assert sum(p.numel() for p in model.parameters() if not p.requires_grad) == want["n_frozen"]
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == want["n_trainable"]


def test_reference_style_checkpoint_round_trip(tmp_path):
    a = helpers.build_model(helpers.SMALL_GROUPS, seed=3)
    path = os.path.join(tmp_path, "best_model_weights.pt")
    torch.save(a.state_dict(), path)                       # utils/base_trainer.py:320-340 saves exactly this
    b = helpers.build_model(helpers.SMALL_GROUPS, seed=4)
    sd = torch.load(path)
    sd["pos_embed"] = torch.zeros(1, 8, 768)               # a reference checkpoint never has it (persistent=False);
    missing, unexpected = b.load_state_dict(sd, strict=False)  # a stray one must not break loading
    assert missing == [] and unexpected == ["pos_embed"]
    for (k, p), (_, q) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(p, q), k


def test_pos_embed_rows_match_reference_formula():
    model = helpers.build_model(helpers.SMALL_GROUPS)
    pos = torch.tensor([0, 1, 2, 1000, 1001, 54321, 1000000])
    rows = model.pos_embed_rows(pos)
    assert float(rows[0].abs().max()) == 0.0
    i, j = (pos - 1) // 1000, (pos - 1) % 1000
    from oracle import modaltune_oracle as O
    tab = O.sincos_table()
    want = torch.cat([tab[j], tab[i]], -1)
    assert torch.equal(rows[1:], want[1:])
    c = torch.tensor([[[256.0 * 3 + 17, 256.0 * 998 + 255.9]]])
    assert int(model.coords_to_pos(c)) == 3 * 1000 + 998 + 1
