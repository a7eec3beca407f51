"""Seeded construction of the ModalTune-GigaPath model and of the frozen text projector (random init).

What ``train_modaltune.py:118-125`` does with ``Aggregator.create(name, **json_config, multi_task=3)`` followed by the
name-keyed seeded re-draw of ``synthetic.seeded_init_`` (identical weights in the reference, the oracle and here,
whatever the construction order).  Used by ``bench.py``, ``__graft_entry__.smoke()`` and the tests.
"""
from __future__ import annotations

from typing import List, Optional

from . import synthetic
from .longvit_adapter import GIGAPATH_CONFIG, Aggregator
from .train_step import Projection_layer

SMALL_GROUPS = [3, 5, 7, 2, 9, 4, 6, 8, 1, 12]   # toy pathway sizes of the small fixtures


def build_model(group_sizes: Optional[List[int]] = None, clinical: bool = True, multi_task: int = 3, seed: int = 0,
                device="cpu"):
    """``group_sizes`` None = the 331 pathways of the reference's table; eval mode."""
    sizes = group_sizes if group_sizes is not None else synthetic.pathway_sizes()
    groups = {i: ["g"] * n for i, n in enumerate(sizes)}
    cfg = dict(GIGAPATH_CONFIG)
    name = "longnetvit_gene_clinical_adapter" if clinical else "longnetvit_gene_adapter"
    if not clinical:
        cfg.pop("clinfeat_dim")
    model = Aggregator.create(name, gene_group_defination=groups, **cfg, multi_task=multi_task)
    model.eval()
    synthetic.seeded_init_(model.named_parameters(), seed=seed)
    return model.to(device)


def build_projector(seed: int = 0, device="cpu"):
    proj = Projection_layer(512, 256)
    proj.load_state_dict(synthetic.seeded_projector_state(seed))
    return proj.to(device).eval()
