"""Slide-sharded data parallelism with world_size 2 on CPU (gloo): every rank runs its own slide, ONE
all-reduce over the flat gradient buffer, result == mean of the per-slide gradients computed serially."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from modaltune_b200 import config, synthetic, train_step
from tests import cpu_kernels, helpers

L = 70


def _grads_for(seed):
    model = helpers.build_model(helpers.SMALL_GROUPS)
    proj = helpers.build_projector(0)
    flat = train_step.FlatGradAllReduce([p for p in model.parameters() if p.requires_grad])
    flat.zero()
    slide = synthetic.synthetic_slide(L, seed=seed, group_sizes=helpers.SMALL_GROUPS)
    with cpu_kernels.installed(), config.using(mode="fp32"):
        train_step.forward_backward(model, proj, slide)
    return model, flat


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    model, flat = _grads_for(500 + rank)
    flat.all_reduce()
    # .grad of every parameter aliases the reduced flat buffer
    off = 0
    for p in flat.params:
        assert p.grad.data_ptr() == flat.flat.data_ptr() + 4 * off
        off += p.numel()
    assert off == flat.numel
    torch.save(flat.flat.clone(), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_flat_allreduce_world2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"rank{r}.pt")) for r in (0, 1))
    assert torch.equal(r0, r1)
    torch.set_num_threads(4)
    serial = sum(_grads_for(500 + r)[1].gather() for r in (0, 1)) / 2
    assert helpers.relerr(r0, serial) < 1e-5


def test_zero_drops_grads_and_gather_realiases():
    model = helpers.build_model(helpers.SMALL_GROUPS)
    flat = train_step.FlatGradAllReduce([p for p in model.parameters() if p.requires_grad])
    for p in flat.params:
        p.grad = torch.ones_like(p)
    buf = flat.gather()
    assert buf.numel() == sum(p.numel() for p in model.parameters() if p.requires_grad) and float(buf.min()) == 1.0
    buf.mul_(2.0)
    assert all(float(p.grad.max()) == 2.0 for p in flat.params)   # views of the flat buffer
    flat.zero()
    assert all(p.grad is None for p in flat.params)
