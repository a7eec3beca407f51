"""Shared builders for the tests (seeded model / slide / golden comparison)."""
import os
import zlib

import torch

from modaltune_b200.factory import SMALL_GROUPS, build_model, build_projector  # noqa: F401

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def grad_summary(name, g, k=64):
    """Must mirror tests/golden/make_golden.py:grad_summary."""
    flat = g.flatten().double().cpu()
    gen = torch.Generator().manual_seed(zlib.crc32(name.encode()) % (2**31))
    idx = torch.randint(0, flat.numel(), (min(k, flat.numel()),), generator=gen)
    proj = torch.stack([(torch.randn(flat.numel(), generator=gen, dtype=torch.float64) @ flat) for _ in range(4)])
    return {"norm": float(flat.norm()), "idx": idx, "vals": flat[idx].float(), "proj": proj.float()}


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))
