"""Import shims that let the UNMODIFIED reference (``/root/reference``) run on CPU in the build container.

Only used by ``tests/golden/make_golden.py`` and ``tests/golden/validate_oracle.py`` (the scripts that pin the oracle
against the reference).  Nothing on the GPU box imports this file: ``/root/reference`` does not exist there.

What is shimmed -- none of it touches arithmetic except ``flash_attn_func``:

* ``timm`` / ``fairscale`` / ``lifelines`` / ``warmup_scheduler`` / TITAN snapshot modules: absent packages, stubbed.
* NumPy 2: ``eval("[np.int64(1024), ...]")`` in ``torchscale/architecture/config.py:76`` needs ``np`` in that
  module's globals.
* ``flash_attn_func``: the reference has no CPU implementation (``torchscale/component/flash_attention.py:143-146``
  sets it to ``None``).  The restatement below follows the call contract of ``flash_attention.py:11-28`` and its
  consumer ``multihead_attention.py:112-119``: ``q,k,v [b,l,h,d] -> out [b,l,h,d], lse [b,h,l]``, softmax scale
  ``1/sqrt(d)``, no mask, dtype preserving; ``lse`` carries no gradient (flash-attn returns it detached).
"""
import math
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("MODALTUNE_REFERENCE", "/root/reference")


def plain_flash_attn_func(q, k, v, dropout=0.0, bias=None, softmax_scale=None, is_causal=False):
    assert bias is None and not is_causal and dropout == 0.0
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(q.shape[-1])
    s = torch.einsum("blhd,bmhd->bhlm", q, k) * scale
    lse = torch.logsumexp(s, dim=-1)
    p = torch.exp(s - lse.unsqueeze(-1))
    out = torch.einsum("bhlm,bmhd->blhd", p, v)
    return out, lse.detach()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _timm_drop_path(x, drop_prob: float = 0.0, training: bool = False, scale_by_keep: bool = True):
    if drop_prob == 0.0 or not training:
        return x
    keep = 1 - drop_prob
    mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
    if keep > 0.0 and scale_by_keep:
        mask.div_(keep)
    return x * mask


def install():
    """Install the shims and put the reference on ``sys.path``.  Idempotent."""
    if getattr(install, "_done", False):
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}; golden generation only runs in the build container")
    # -- absent third-party packages -------------------------------------------------------------------------
    if "timm" not in sys.modules:
        timm = _stub("timm", create_model=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("timm stub")))
        _stub("timm.models")
        _stub("timm.models.registry", register_model=lambda f: f)
        _stub("timm.models.layers", drop_path=_timm_drop_path)
        timm.models = sys.modules["timm.models"]
    if "fairscale" not in sys.modules:
        _stub("fairscale")
        _stub("fairscale.nn", checkpoint_wrapper=lambda m, *a, **k: m, wrap=lambda m, *a, **k: m)
    if "lifelines" not in sys.modules:
        _stub("lifelines", CoxPHFitter=object)
        _stub("lifelines.utils", concordance_index=lambda *a, **k: 0.5)
    if "warmup_scheduler" not in sys.modules:
        _stub("warmup_scheduler", GradualWarmupScheduler=object)
    if "safetensors" not in sys.modules:
        try:
            import safetensors  # noqa: F401
        except Exception:
            _stub("safetensors", safe_open=None)
    # -- TITAN snapshot (un-vendored; models/aggregators/__init__.py star-imports titan_adapter) -------------
    snap = "b2fb4f475256eb67c6e9ccbf2d6c9c3f25f20791"
    if snap not in sys.modules:
        _stub(snap)
        _stub(snap + ".vision_transformer", VisionTransformer=type("VisionTransformer", (torch.nn.Module,), {}))
        _stub(snap + ".configuration_titan", TitanConfig=type("TitanConfig", (), {}))
    # -- reference on the path -------------------------------------------------------------------------------
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # LongNet.py:7-8 appends .../gigapath to sys.path and imports ``torchscale`` as a top-level package.
    giga = os.path.join(REFERENCE_ROOT, "models", "prov_gigapath", "gigapath")
    if giga not in sys.path:
        sys.path.append(giga)
    # -- NumPy 2 fix: the segment-length string is eval()'ed inside torchscale.architecture.config -----------
    import torchscale.architecture.config as ts_config

    ts_config.np = np
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    # -- flash_attn_func restatement; the live module copy is the top-level ``torchscale`` one ---------------
    import torchscale.component.flash_attention as ts_fa
    import torchscale.component.multihead_attention as ts_mha

    ts_fa.flash_attn_func = plain_flash_attn_func
    ts_mha.flash_attn_func = plain_flash_attn_func
    install._done = True


def build_reference_model(clinical=True, multi_task=3, gene_group_sizes=None, config_overrides=None, quiet=True):
    """Build the reference ``LongNetGene[SimpleClinical]Adapter`` exactly as ``train_modaltune.py:118-125`` does."""
    import contextlib
    import io
    import json

    install()
    from models.aggregators import Aggregator

    with open(os.path.join(REFERENCE_ROOT, "model_configs", "modaltune_gigapath_config.json")) as f:
        cfg = json.load(f)
    if config_overrides:
        cfg.update(config_overrides)
    if gene_group_sizes is None:
        here = os.path.dirname(os.path.abspath(__file__))
        with open(os.path.join(here, "..", "..", "modaltune_b200", "data", "pathway_sizes.json")) as f:
            gene_group_sizes = json.load(f)
    groups = {i: ["g"] * n for i, n in enumerate(gene_group_sizes)}
    name = "longnetvit_gene_clinical_adapter" if clinical else "longnetvit_gene_adapter"
    if not clinical:
        cfg.pop("clinfeat_dim", None)
    ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
    with ctx:
        model = Aggregator.create(name, gene_group_defination=groups, **cfg, multi_task=multi_task)
    return model, cfg
