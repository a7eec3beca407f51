"""Shared builders for the tests (seeded model / slide / golden comparison)."""
import os
import zlib

import torch

from modaltune_b200 import synthetic
from modaltune_b200.longvit_adapter import GIGAPATH_CONFIG, Aggregator
from modaltune_b200.train_step import Projection_layer

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL_GROUPS = [3, 5, 7, 2, 9, 4, 6, 8, 1, 12]


def build_model(group_sizes=None, clinical=True, multi_task=3, seed=0, device="cpu"):
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


def build_projector(seed=0, device="cpu"):
    proj = Projection_layer(512, 256)
    proj.load_state_dict(synthetic.seeded_projector_state(seed))
    return proj.to(device).eval()


def grad_summary(name, g, k=64):
    """Must mirror tests/golden/make_golden.py:grad_summary."""
    flat = g.flatten().double().cpu()
    gen = torch.Generator().manual_seed(zlib.crc32(name.encode()) % (2**31))
    idx = torch.randint(0, flat.numel(), (min(k, flat.numel()),), generator=gen)
    proj = torch.stack([(torch.randn(flat.numel(), generator=gen, dtype=torch.float64) @ flat) for _ in range(4)])
    return {"norm": float(flat.norm()), "idx": idx, "vals": flat[idx].float(), "proj": proj.float()}


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))
