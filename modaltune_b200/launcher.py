"""Run the reference's own training scripts, byte-unchanged, on the B200 module classes.

    python -m modaltune_b200.launcher [launcher options] train_modaltune.py [the script's own arguments ...]

    launcher options
      --reference DIR        the reference checkout (default: $MODALTUNE_REFERENCE, then /root/reference)
      --classes ours|reference   whose ``LongNetGene*Adapter`` classes ``Aggregator.create`` hands to the script
      --synthetic DIR        write a synthetic dataset (json splits, feature bags, text / clinical dicts, genomics csv) to
                             DIR and append the matching ``--train_json ... --clinical_location ...`` script arguments
      --cases N --tiles L    size of that dataset (cases per split, tiles per slide)
      --seeded-init S        overwrite every parameter with ``synthetic.seeded_init_(seed=S)`` right after construction
                             (identical weights whichever classes run: what the parity tests compare)
      --mode bf16|fp32       compute mode of the B200 classes (``config.set_mode``)
      --deterministic        parity runs: ``model.train()`` leaves the modules in eval mode (Dropout / DropPath masks
                             cannot be bit-matched between two implementations; the training loop itself is untouched)
      --device cpu|N         the scripts declare ``--device`` as an int (a CUDA ordinal); this widens the declaration so
                             that ``cpu`` can be passed too, and forwards the value

What the launcher does NOT do is touch the script: ``train_modaltune.py`` / ``train_modaltune_pancancer.py`` are executed
with ``runpy`` exactly as ``python train_modaltune.py ...`` would (``__name__ == "__main__"``, their own argparse,
``Trainer.__init__``, DataLoader, AdamW, GradScaler, probes, checkpoints).  Around them it

1. provides the third-party modules the scripts import but this image lacks -- ``timm`` (``register_model``,
   ``drop_path``), ``fairscale`` (identity wrappers), ``lifelines`` (a small Cox-score stand-in for the probe evaluation),
   ``warmup_scheduler`` (a working ``GradualWarmupScheduler``), the un-vendored TITAN snapshot modules -- and the two
   NumPy-2 fixes the reference needs (``np.Inf``; ``np`` in ``torchscale.architecture.config``'s globals);
2. swaps the model classes through the reference's own registry (``models/aggregators/aggregators.py:23-41``): the
   scripts call ``Aggregator.create(args.mil_name, ...)`` (``train_modaltune.py:123-125``) and get
   ``modaltune_b200.longvit_adapter`` classes, whose constructor keywords, forward signature, ``is_multi`` attribute and
   parameter names are the reference's;
3. optionally generates the dataset files the reference's ``FeaturesGeneTextDataset`` reads
   (``data_utils/datasets.py:144-285``) from ``synthetic.synthetic_slide``.
"""
from __future__ import annotations

import argparse
import json
import os
import runpy
import sys
import types
from typing import List, Optional

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TITAN_SNAPSHOT = "b2fb4f475256eb67c6e9ccbf2d6c9c3f25f20791"   # utils/constants.py:22-23 of the reference


def find_reference(explicit: Optional[str] = None) -> str:
    for cand in (explicit, os.environ.get("MODALTUNE_REFERENCE"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "train_modaltune.py")):
            return os.path.abspath(cand)
    raise FileNotFoundError("no reference checkout found: pass --reference DIR or set MODALTUNE_REFERENCE "
                            "(a checkout of martellab-sri/ModalTune)")


# ---------------------------------------------------------------------------------------------------------------------
# absent third-party packages
# ---------------------------------------------------------------------------------------------------------------------
def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _drop_path(x, drop_prob: float = 0.0, training: bool = False, scale_by_keep: bool = True):
    """timm.models.layers.drop_path (stochastic depth per sample)."""
    if drop_prob == 0.0 or not training:
        return x
    keep = 1 - drop_prob
    mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
    if keep > 0.0 and scale_by_keep:
        mask.div_(keep)
    return x * mask


class GradualWarmupScheduler:
    """warmup_scheduler.GradualWarmupScheduler as the scripts use it (``train_modaltune.py:150-153``): the learning rate
    climbs linearly from lr to multiplier * lr over ``total_epoch`` epochs, then ``after_scheduler`` takes over."""

    def __init__(self, optimizer, multiplier, total_epoch, after_scheduler=None):
        self.optimizer, self.multiplier, self.total_epoch, self.after = optimizer, multiplier, total_epoch, after_scheduler
        self.epoch = 0
        self.base = [g["lr"] for g in optimizer.param_groups]

    def step(self):
        self.epoch += 1
        if self.epoch <= self.total_epoch:
            for g, b in zip(self.optimizer.param_groups, self.base):
                g["lr"] = b * ((self.multiplier - 1.0) * self.epoch / self.total_epoch + 1.0)
        elif self.after is not None:
            if self.epoch == self.total_epoch + 1:
                self.after.base_lrs = [b * self.multiplier for b in self.base]
            self.after.step()


class _CoxStandIn:
    """lifelines.fitters.coxph_fitter.CoxPHFitter for the probe evaluation of the training loop
    (``train_modaltune.py:372-380``): a ridge-regularised linear risk fitted by least squares on the log durations and
    scored by Harrell's concordance index.  Evaluation-probe plumbing, not part of the accelerated path."""

    def __init__(self, penalizer: float = 0.0, **kw):
        self.penalizer, self.w = penalizer, None

    def fit(self, df, duration_col, event_col, **kw):
        x = df.drop(columns=[duration_col, event_col]).to_numpy(dtype=np.float64)
        y = -np.log(np.maximum(df[duration_col].to_numpy(dtype=np.float64), 1e-6))
        xc = x - x.mean(0, keepdims=True)
        self.w = np.linalg.solve(xc.T @ xc + (self.penalizer + 1e-3) * len(x) * np.eye(x.shape[1]), xc.T @ (y - y.mean()))
        return self

    def score(self, df, scoring_method="concordance_index"):
        dcol, ecol = "durations", "vital_status"
        x = df.drop(columns=[dcol, ecol]).to_numpy(dtype=np.float64)
        risk = x @ self.w
        t, e = df[dcol].to_numpy(dtype=np.float64), df[ecol].to_numpy(dtype=np.float64)
        num = den = 0.0
        for i in range(len(t)):
            if e[i] != 1:
                continue
            later = t > t[i]
            den += later.sum()
            num += (risk[i] > risk[later]).sum() + 0.5 * (risk[i] == risk[later]).sum()
        return float(num / den) if den > 0 else 0.5


def install_shims(reference_root: str) -> None:
    """Idempotent.  Puts the reference on ``sys.path`` and fills in what this image lacks (module docstring, item 1)."""
    if getattr(install_shims, "_done", None) == reference_root:
        return
    if "timm" not in sys.modules:
        try:
            import timm  # noqa: F401
        except Exception:
            timm = _stub("timm", create_model=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("timm is not installed")))
            _stub("timm.models")
            _stub("timm.models.registry", register_model=lambda f: f)
            _stub("timm.models.layers", drop_path=_drop_path)
            timm.models = sys.modules["timm.models"]
    if "fairscale" not in sys.modules:
        try:
            import fairscale  # noqa: F401
        except Exception:
            _stub("fairscale")
            _stub("fairscale.nn", checkpoint_wrapper=lambda m, *a, **k: m, wrap=lambda m, *a, **k: m)
    if "lifelines" not in sys.modules:
        try:
            import lifelines  # noqa: F401
        except Exception:
            ll = _stub("lifelines", CoxPHFitter=_CoxStandIn)
            _stub("lifelines.utils", concordance_index=lambda *a, **k: 0.5)
            fit = _stub("lifelines.fitters")
            cox = _stub("lifelines.fitters.coxph_fitter", CoxPHFitter=_CoxStandIn)
            ll.fitters, fit.coxph_fitter, ll.utils = fit, cox, sys.modules["lifelines.utils"]
    if "warmup_scheduler" not in sys.modules:
        try:
            import warmup_scheduler  # noqa: F401
        except Exception:
            _stub("warmup_scheduler", GradualWarmupScheduler=GradualWarmupScheduler)
    if "wandb" not in sys.modules:
        try:
            import wandb  # noqa: F401
        except Exception:
            run = types.SimpleNamespace(finish=lambda: None, name="offline", config=types.SimpleNamespace())
            _stub("wandb", init=lambda *a, **k: run, log=lambda *a, **k: None, define_metric=lambda *a, **k: None,
                  config=types.SimpleNamespace(),
                  plot=types.SimpleNamespace(confusion_matrix=lambda **k: None, roc_curve=lambda **k: None))
    if TITAN_SNAPSHOT not in sys.modules:   # models/aggregators/__init__.py star-imports titan_adapter
        _stub(TITAN_SNAPSHOT)
        _stub(TITAN_SNAPSHOT + ".vision_transformer", VisionTransformer=type("VisionTransformer", (torch.nn.Module,), {}))
        _stub(TITAN_SNAPSHOT + ".configuration_titan", TitanConfig=type("TitanConfig", (), {}))
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    giga = os.path.join(reference_root, "models", "prov_gigapath", "gigapath")   # LongNet.py:7-8 imports `torchscale` top-level
    if giga not in sys.path:
        sys.path.append(giga)
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                                   # utils/base_trainer.py:476
    import torchscale.architecture.config as ts_config  # eval("[np.int64(1024), ...]") at :76 needs `np` in its globals

    ts_config.np = np
    install_shims._done = reference_root


def swap_classes(which: str, seeded_init: Optional[int] = None, deterministic: bool = False) -> None:
    """Point the reference's registry at the B200 classes (``which == "ours"``) or back at its own (``"reference"``),
    and optionally post-process every constructed model (seeded weights, eval-mode ``train()``).  The scripts only ever
    reach the models through ``Aggregator.create`` (``train_modaltune.py:123-125``), so nothing else needs to change."""
    from models.aggregators import Aggregator as RefAggregator  # the reference's registry

    if not hasattr(RefAggregator, "_mt_original"):
        RefAggregator._mt_original = (dict(RefAggregator.subclasses), RefAggregator.create.__func__)
    originals, orig_create = RefAggregator._mt_original
    RefAggregator.subclasses.clear()
    RefAggregator.subclasses.update(originals)
    if which == "ours":
        from . import longvit_adapter as ours

        for name in ("longnetvit_gene_adapter", "longnetvit_gene_clinical_adapter"):
            RefAggregator.subclasses[name] = ours.Aggregator.subclasses[name]

    def create(cls, subclass_name, **params):
        model = orig_create(cls, subclass_name, **params)
        if seeded_init is not None:
            from . import synthetic

            synthetic.seeded_init_(model.named_parameters(), seed=seeded_init)
        if deterministic:
            model.eval()
            model.train = lambda mode=True: model
        return model

    RefAggregator.create = classmethod(create)


# ---------------------------------------------------------------------------------------------------------------------
# synthetic dataset in the reference's file formats (data_utils/datasets.py:144-285)
# ---------------------------------------------------------------------------------------------------------------------
def write_synthetic_dataset(root: str, reference_root: str, cases: int = 4, tiles: int = 300, seed: int = 0,
                            tiles_range: Optional[tuple] = None) -> List[str]:
    """json splits (train / val / test, ``cases`` each), one feature bag ``{"features": [n, 1536], "coords": [n, 2]}`` per
    slide, ``{case_id: [4, 512]}`` text embeddings, ``{case_id: [5]}`` clinical features and a genomics csv with the
    pathway genes of ``dataset/gene_pathway_processed_v2.csv``.  Returns the script arguments that point at them."""
    import pandas as pd

    from . import synthetic

    os.makedirs(os.path.join(root, "features"), exist_ok=True)
    genes = pd.read_csv(os.path.join(reference_root, "dataset", "gene_pathway_processed_v2.csv"), usecols=["gene"])["gene"].tolist()
    rng = np.random.default_rng(seed)
    text, clinical, rows, splits = {}, {}, [], {"train": [], "val": [], "test": []}
    n = 0
    for split in splits:
        for _ in range(cases):
            case = f"SYN-{n:04d}"
            L = tiles if tiles_range is None else int(np.exp(rng.uniform(np.log(tiles_range[0]), np.log(tiles_range[1]))))
            s = synthetic.synthetic_slide(L, seed=seed * 1000 + n, group_sizes=[1])
            path = os.path.join(root, "features", f"{case}.pt")
            torch.save({"features": s["x"][0].clone(), "coords": s["coords"][0].clone()}, path)
            text[case] = s["text"].clone()
            clinical[case] = s["clinical"][0].clone()
            g = torch.Generator().manual_seed(seed * 1000 + n + 7)
            rows.append([case] + torch.randn(len(genes), generator=g).tolist())
            splits[split].append({"case_id": case, "case_submitter_id": case, "features_path": path,
                                  "primary_class": int(n % 2), "vital_status": int(rng.integers(0, 2)),
                                  "durations": float(rng.uniform(30, 3000)), "project_id": "SYN"})
            n += 1
    for split, data in splits.items():
        with open(os.path.join(root, f"{split}.json"), "w") as f:
            json.dump({"data": data}, f)
    torch.save(text, os.path.join(root, "text.pt"))
    torch.save(clinical, os.path.join(root, "clinical.pt"))
    pd.DataFrame(rows, columns=["case_id"] + genes).to_csv(os.path.join(root, "genomics.csv"), index=False)
    return ["--train_json", os.path.join(root, "train.json"), "--val_json", os.path.join(root, "val.json"),
            "--test_json", os.path.join(root, "test.json"), "--text_location", os.path.join(root, "text.pt"),
            "--clinical_location", os.path.join(root, "clinical.pt"), "--genomics_csv_path", os.path.join(root, "genomics.csv"),
            "--output_path", os.path.join(root, "results")]


def main(argv: Optional[List[str]] = None) -> None:
    ap = argparse.ArgumentParser(prog="python -m modaltune_b200.launcher", description=__doc__.split("\n\n")[0],
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default=None)
    ap.add_argument("--classes", default="ours", choices=["ours", "reference"])
    ap.add_argument("--synthetic", default=None, metavar="DIR")
    ap.add_argument("--cases", type=int, default=4)
    ap.add_argument("--tiles", type=int, default=300)
    ap.add_argument("--seeded-init", type=int, default=None)
    ap.add_argument("--mode", default=None, choices=["bf16", "fp32"])
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--device", default=None)
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)

    ref = find_reference(args.reference)
    script = args.script if os.path.isabs(args.script) else os.path.join(ref, args.script)
    if not os.path.isfile(script):
        raise FileNotFoundError(script)
    install_shims(ref)
    sys.modules.pop("utils.defaut_args", None)   # a fresh argparse parser if a script already ran in this process
    if args.mode:
        from . import config

        config.set_mode(args.mode)
    swap_classes(args.classes, args.seeded_init, args.deterministic)
    extra = write_synthetic_dataset(args.synthetic, ref, args.cases, args.tiles) if args.synthetic else []
    if args.device is not None:
        from utils.defaut_args import parser as ref_parser   # the very parser object the script extends and parses with

        for action in ref_parser._actions:
            if action.dest == "device":
                action.type = lambda v: int(v) if str(v).lstrip("-").isdigit() else str(v)
        extra += ["--device", str(args.device)]
    sys.argv = [script] + list(args.script_args) + extra
    cwd = os.getcwd()
    os.chdir(ref)   # the scripts resolve model_configs / dataset relative to their own location, results relative to cwd
    try:
        runpy.run_path(script, run_name="__main__")
    finally:
        os.chdir(cwd)


if __name__ == "__main__":
    main()
