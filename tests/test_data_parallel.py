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


def test_gather_buffers_exchanges_stacked_gradients_in_place():
    """``gather_buffers`` (what a captured step all-reduces with more than one rank): the per-pathway SNN gradients stay
    the slices of the stacked gradient tensors autograd delivers (no copies), everything else goes into one small flat
    buffer, every gradient is covered exactly once, and ``p.grad`` aliases the buffers (a scaling of the buffers -- the
    all-reduce -- shows up in every ``p.grad``)."""
    model, flat = _grads_for(511)
    want = {id(p): p.grad.clone() for p in flat.params if p.grad is not None}
    bufs = flat.gather_buffers()
    assert len(bufs) > 1 and all(b.is_contiguous() and b.dtype == torch.float32 for b in bufs)
    assert sum(b.numel() for b in bufs) == flat.numel                      # every gradient exactly once
    n_views = sum(1 for p in flat.params if p.grad is not None and p.grad._base is not None
                  and any(p.grad._base is b for b in bufs[1:]))
    assert n_views >= 4 * len(helpers.SMALL_GROUPS)                        # the SNN weights and biases of every pathway
    for b in bufs:
        b.mul_(0.5)
    for p in flat.params:
        if id(p) in want:
            assert torch.equal(p.grad, want[id(p)] * 0.5)
