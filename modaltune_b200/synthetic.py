"""Seeded synthetic slides and seeded weights.

There is no network for datasets or checkpoints, so every test, golden fixture and bench line is driven by the
generators below.  They are deterministic CPU generators (``torch.Generator``), so the build container (where the
reference is available) and the GPU box (where it is not) see bit-identical inputs and weights.

Input layout follows ``data_utils/datasets.py:180-285`` of the reference: tile features ``[1, L, 1536]``, tile
coordinates in pixels ``[1, L, 2]`` (multiples of 256), one tensor ``[1, n_i]`` per pathway (331 pathways, sizes from
``dataset/gene_pathway_processed_v2.csv``), clinical features ``[1, 5]`` and CONCH text embeddings ``[4, 512]``.
"""
from __future__ import annotations

import json
import math
import os
import zlib
from typing import Dict, List, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def pathway_sizes() -> List[int]:
    """The 331 pathway group sizes (1..199, sum 9731) of the reference's pathway table."""
    with open(os.path.join(_HERE, "data", "pathway_sizes.json")) as f:
        return json.load(f)


def synthetic_slide(n_tiles: int, seed: int = 0, group_sizes: Optional[List[int]] = None, in_chans: int = 1536,
                    dtype=torch.float32) -> Dict[str, object]:
    """One synthetic case.  Coordinates: ``n_tiles`` distinct cells of a G x G grid (G = ceil(sqrt(L/0.6)) <= 999),
    row-major sorted, times 256 px."""
    g = torch.Generator().manual_seed(int(seed))
    if group_sizes is None:
        group_sizes = pathway_sizes()
    feats = torch.randn(1, n_tiles, in_chans, generator=g, dtype=torch.float32).to(dtype)
    G = min(999, max(2, math.ceil(math.sqrt(n_tiles / 0.6))))
    assert G * G >= n_tiles, "too many tiles for a 999 x 999 grid"
    cells = torch.randperm(G * G, generator=g)[:n_tiles].sort().values
    coords = torch.stack([(cells // G).float() * 256.0, (cells % G).float() * 256.0], dim=-1)[None]
    genes = {i: torch.randn(1, n, generator=g, dtype=torch.float32).to(dtype) for i, n in enumerate(group_sizes)}
    clinical = torch.randn(1, 5, generator=g, dtype=torch.float32).to(dtype)
    text = torch.randn(4, 512, generator=g, dtype=torch.float32).to(dtype)
    return {"x": feats, "coords": coords, "genes": genes, "clinical": clinical, "text": text}


_SUBLN_SCALED = ("fc1.weight", "fc2.weight", "out_proj.weight", "v_proj.weight")


def _is_norm_weight(name: str, p: torch.Tensor) -> bool:
    if p.dim() != 1 or not name.endswith("weight"):
        return False
    return True  # every 1-D ".weight" on this path is a LayerNorm scale


@torch.no_grad()
def seeded_init_(named_params, seed: int = 0) -> None:
    """Overwrite parameters in place with name-keyed seeded draws.

    The draw for a parameter depends only on ``(seed, name, shape)``, so the reference model, the oracle and the
    CUDA modules end up with identical weights regardless of construction order.  Distributions mimic the
    reference's initialisers (xavier for matrices, sub-LN scaling of fc1/fc2/out_proj/v_proj in the encoder,
    ``TS/architecture/encoder.py:269-285``) but keep every LayerNorm scale/bias, Linear bias and the Injector gate
    gamma away from their degenerate init (1 / 0 / 0) so that every gradient is live (SURVEY.md §8c gotchas).
    """
    for name, p in named_params:
        g = torch.Generator().manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2**63 - 1))
        r = torch.randn(p.shape, generator=g, dtype=torch.float32)
        if name.endswith("gamma"):
            v = 0.05 * r
        elif name in ("cls_token", "gene_pe", "gene_cls"):
            v = 0.02 * r
        elif p.dim() >= 2:
            fan_out, fan_in = p.shape[0], p[0].numel()
            std = math.sqrt(2.0 / (fan_in + fan_out))
            if name.startswith("encoder.") and name.endswith(_SUBLN_SCALED):
                std *= math.sqrt(math.log(24.0))
            v = std * r
        elif _is_norm_weight(name, p):
            v = 1.0 + 0.1 * r
        else:
            v = 0.02 * r
        p.copy_(v.to(p.dtype))


def seeded_projector_state(seed: int = 0, in_dim: int = 512, out_dim: int = 256) -> Dict[str, torch.Tensor]:
    """Weights of the frozen random text projector (``train_modaltune.py:44-59``), reference parameter names."""
    shapes = {
        "conv1.0.weight": (out_dim, in_dim, 1, 1), "conv1.0.bias": (out_dim,),
        "conv1.1.weight": (out_dim, 1, 1), "conv1.1.bias": (out_dim, 1, 1),
        "conv1.3.weight": (out_dim, out_dim, 1, 1), "conv1.3.bias": (out_dim,),
    }
    sd = {k: torch.empty(s) for k, s in shapes.items()}
    g = torch.Generator().manual_seed(int(seed) + 7919)
    for k, t in sd.items():
        r = torch.randn(t.shape, generator=g)
        if k == "conv1.1.weight":
            t.copy_(1.0 + 0.1 * r)
        elif k.endswith("bias"):
            t.copy_(0.02 * r)
        else:
            t.copy_(r * math.sqrt(2.0 / (t.shape[0] + t[0].numel())))
    return sd
