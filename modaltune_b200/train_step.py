"""One ModalTune training step on the hot path: three task passes forward, the KL-distillation loss, one backward.

Host-side mirror of ``MILTextGeneTrainer_multitask.multitask_forward`` + the loss of ``train_one_epoch``
(``train_modaltune.py:156-179, 211-235``) and of the frozen random text projector (``:44-59``); the data-parallel
gradient exchange replaces DDP's bucketed all-reduce (``utils/base_trainer.py:205-211``) with ONE NCCL all-reduce over
a flat buffer of the trainable gradients.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

NUM_TASKS = 3          # utils/constants.py:45-49: 0 General, 1 Diagnosis, 2 Survival
TEXT_ROWS = (0, 1, 3)  # train_modaltune.py:230-233


class Projection_layer(nn.Module):
    """Frozen random text projector (train_modaltune.py:44-59): Conv1x1(512->256), LN([256,1,1]), ReLU, Conv1x1."""

    def __init__(self, in_dim: int = 512, out_dim: int = 256):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(in_dim, out_dim, 1), nn.LayerNorm([out_dim, 1, 1]), nn.ReLU(),
                                   nn.Conv2d(out_dim, out_dim, 1))
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, x):
        return self.conv1(x.unsqueeze(-1).unsqueeze(-1)).squeeze(-1).squeeze(-1)


def text_targets(projector: nn.Module, text: torch.Tensor) -> torch.Tensor:
    t = projector(text.float())
    return t / t.norm(dim=-1, keepdim=True)


def multitask_forward(model, slide: Dict, task_ids: Sequence[int] = (0, 1, 2)) -> torch.Tensor:
    """[len(task_ids), 256] task-conditioned embeddings (train_modaltune.py:156-179)."""
    eye = torch.eye(NUM_TASKS, device=slide["x"].device)
    outs = []
    for t in task_ids:
        if "clinical" in slide and getattr(model, "_HAS_CLINICAL", False):
            outs.append(model(x=slide["x"], coords=slide["coords"], genes=slide["genes"], clinical=slide["clinical"],
                              task_token=eye[t]))
        else:
            outs.append(model(x=slide["x"], coords=slide["coords"], genes=slide["genes"], task_token=eye[t]))
    return torch.cat(outs, 0)


def distill_loss(logits: torch.Tensor, text_proj: torch.Tensor) -> torch.Tensor:
    """10 * KLDiv_sum(log_softmax(z/|z|), softmax(t[[0,1,3]]))   (train_modaltune.py:225-233)."""
    z = logits.float()
    z = z / z.norm(dim=-1, keepdim=True)
    tgt = F.softmax(text_proj[list(TEXT_ROWS)], dim=1)
    return F.kl_div(F.log_softmax(z, dim=1), tgt, reduction="sum") * 10.0


def forward_backward(model, projector, slide: Dict):
    """One slide step: returns (loss, logits).  Gradients are left in ``p.grad`` of the trainable parameters."""
    logits = multitask_forward(model, slide)
    loss = distill_loss(logits, text_targets(projector, slide["text"]))
    loss.backward()
    return loss.detach(), logits.detach()


def slide_to_device(slide: Dict, device, non_blocking: bool = True) -> Dict:
    out = {}
    for k, v in slide.items():
        if isinstance(v, dict):
            out[k] = {kk: vv.to(device, non_blocking=non_blocking) for kk, vv in v.items()}
        else:
            out[k] = v.to(device, non_blocking=non_blocking)
    return out


class FlatGradAllReduce:
    """Data-parallel gradient exchange: one all-reduce (sum, then / world) over a single contiguous fp32 buffer that the
    ``.grad`` of every trainable parameter aliases.  Replaces DDP's 25 MB-bucket all-reduces."""

    def __init__(self, params: Sequence[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            assert p.dtype == torch.float32
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()
        off = 0
        for p in self.params:  # re-alias in case an optimizer / zero_grad(set_to_none) dropped the views
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def all_reduce(self):
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())
