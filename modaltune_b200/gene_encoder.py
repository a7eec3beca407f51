"""Genomic pathway encoder producing the modal tokens the Injector / Extractor consume.

Same parameters (names, shapes) and arithmetic as the reference's ``GeneEncoder_Group``
(``models/genomic_utils/gene_encoder.py:98-223``): one 2-layer SNN (Linear + ELU + AlphaDropout) per pathway, a
3-deep MLP-Mixer over the 331 pathway tokens, LayerNorm + Linear(256 -> 768), and ``pathway_compression``
Linear(331 -> 64) over the token axis.  It is ~0.1 GFLOP and not a kernel target; what matters
at B200 speeds is launch count, so the 331 sequential tiny Linears of ``gene_encode`` (:201-203) are evaluated as two
grouped GEMMs: layer 1 as a block-indicator matmul over the concatenated gene vector, layer 2 as one ``bmm``.
"""
from __future__ import annotations

from functools import partial
from typing import Dict, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


class PreNormResidual(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.fn(self.norm(x)) + x


def _snn_block(dim1, dim2, dropout):
    return nn.Sequential(nn.Linear(dim1, dim2), nn.ELU(), nn.AlphaDropout(dropout))


def _feed_forward(dim, expansion_factor, dropout, dense):
    inner = int(dim * expansion_factor)
    return nn.Sequential(dense(dim, inner), nn.GELU(), nn.Dropout(dropout), dense(inner, dim), nn.Dropout(dropout))


class GeneEncoder_Group(nn.Module):
    def __init__(self, output_dim: int, latent_dim: int, group_sizes: Dict, n_groups: int = 64, depth: int = 5,
                 cls_token: bool = False, expansion_groups=4, expansion_dim=0.5, dropout: float = 0.25,
                 n_classes: int = 2, mode: str = "classifier", final_groups: int = 8, **kwargs):
        super().__init__()
        assert mode == "feature" and not cls_token, "ModalTune builds the gene encoder in feature mode without cls token"
        self.mode = mode
        self.sizes = [len(v) for v in group_sizes.values()]
        self.gene_networks = nn.ModuleList([
            nn.Sequential(_snn_block(n, latent_dim, dropout), _snn_block(latent_dim, latent_dim, dropout))
            for n in self.sizes])
        chan_first, chan_last = partial(nn.Conv1d, kernel_size=1), nn.Linear
        self.mlp_mixer = nn.Sequential(
            *[nn.Sequential(PreNormResidual(latent_dim, _feed_forward(n_groups, expansion_groups, dropout, chan_first)),
                            PreNormResidual(latent_dim, _feed_forward(latent_dim, expansion_dim, dropout, chan_last)))
              for _ in range(depth)],
            nn.LayerNorm(latent_dim), nn.Linear(latent_dim, output_dim))
        self.n_groups = final_groups
        self.pathway_compression = nn.Linear(n_groups, final_groups)
        seg = torch.repeat_interleave(torch.arange(len(self.sizes)), torch.tensor(self.sizes))
        ind = torch.zeros(len(self.sizes), int(sum(self.sizes)))
        ind[seg, torch.arange(seg.numel())] = 1.0
        self.register_buffer("_group_indicator", ind, persistent=False)  # [G, sum n_i] 0/1

    def _pathway_tokens(self, x: Sequence[torch.Tensor]) -> torch.Tensor:
        """The per-pathway SNNs (gene_encode :201-203) as grouped GEMMs -> [G, latent]."""
        G = len(self.sizes)
        nets = self.gene_networks
        # train mode: the AlphaDropout of every SNN block is one draw over the batched [G, latent] tensor -- element-wise
        # i.i.d. masks, the same distribution as 331 modules drawing theirs one after the other (the reference's loop costs
        # ~6 000 tiny kernels per training step at B200 speeds)
        p1, p2 = nets[0][0][2].p, nets[0][1][2].p
        xs = torch.cat([x[i].reshape(-1) for i in range(G)]).float()                 # [sum n_i]
        w1 = torch.cat([nets[i][0][0].weight.t() for i in range(G)], 0)               # [sum n_i, latent]
        b1 = torch.stack([nets[i][0][0].bias for i in range(G)], 0)
        h = F.alpha_dropout(F.elu(self._group_indicator @ (xs.unsqueeze(1) * w1) + b1), p1, self.training)   # [G, latent]
        w2 = torch.stack([nets[i][1][0].weight for i in range(G)], 0)                 # [G, latent, latent]
        b2 = torch.stack([nets[i][1][0].bias for i in range(G)], 0)
        return F.alpha_dropout(F.elu(torch.bmm(w2, h.unsqueeze(2)).squeeze(2) + b2), p2, self.training)

    def gene_encode(self, x):
        h = self._pathway_tokens(x).unsqueeze(0)                                       # [1, G, latent]
        h = self.mlp_mixer(h)
        return self.pathway_compression(h.permute(0, 2, 1)).permute(0, 2, 1)           # [1, final_groups, out]

    def forward(self, x):
        return self.gene_encode(x)
