"""Same-box END-TO-END comparator: the reference's own GPU path on the B200 it shares with ours.

The UNMODIFIED reference modules (staged under oracle/_ref/reference by ``__graft_entry__.build()``; /root/reference in the
build container) run one ModalTune training step -- three task passes through ``LongNetGeneSimpleClinicalAdapter``,
the KL-distillation loss, ``loss.backward()`` -- under ``torch.autocast(float16)`` with the installed flash-attn
(``train_modaltune.py:156-179, 211-235``; FA2 kernels built for sm_100), on the same seeded weights and synthetic slide
the B200 path is benchmarked on.  CUDA events, a few steps.  This is the "secondary comparator" of BASELINE.md section 4.
Library / reference code, used ONLY as a measured baseline: nothing in the product imports it.

    python tools/bench_reference_gpu.py [tiles] [steps]      # prints one JSON object
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def run(tiles: int = 10000, steps: int = 3, warmup: int = 2, dev="cuda"):
    from modaltune_b200 import factory, launcher, synthetic

    staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "reference")
    ref = launcher.find_reference(staged if os.path.isdir(staged) else None)   # measurement tooling: the staged copy
    launcher.install_shims(ref)
    launcher.swap_classes("reference")
    from models.aggregators import Aggregator  # the reference's registry and classes

    with open(os.path.join(ref, "model_configs", "modaltune_gigapath_config.json")) as f:
        cfg = json.load(f)
    groups = {i: ["g"] * n for i, n in enumerate(synthetic.pathway_sizes())}
    model = Aggregator.create("longnetvit_gene_clinical_adapter", gene_group_defination=groups, **cfg, multi_task=3)
    model.eval()                       # eval mode with autograd on, as the B200 arm
    synthetic.seeded_init_(model.named_parameters(), seed=0)
    model = model.to(dev)
    proj = factory.build_projector(0, dev)
    slide = synthetic.synthetic_slide(tiles, seed=1000)
    x, coords = slide["x"].to(dev), slide["coords"].to(dev)
    genes = {k: v.to(dev) for k, v in slide["genes"].items()}
    clinical, text = slide["clinical"].to(dev), slide["text"].to(dev)
    eye = torch.eye(3, device=dev)

    def step():
        for p in model.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            logits = torch.cat([model(x=x, coords=coords, genes=genes, clinical=clinical, task_token=eye[t])
                                for t in range(3)], 0)
            t = proj(text)
            t = t / t.norm(dim=-1, keepdim=True)
            z = logits / logits.norm(dim=-1, keepdim=True)
            loss = torch.nn.functional.kl_div(torch.nn.functional.log_softmax(z, dim=1),
                                              torch.nn.functional.softmax(t[[0, 1, 3], :], dim=1), reduction="sum") * 10
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"what": "unmodified reference modules, fp16 autocast + flash-attn (FA2 for sm_100), 3 task passes fwd + KL loss + bwd, "
                    "eval mode, same seeded weights and synthetic slide generator",
            "tiles": tiles, "steps": steps, "warmup": warmup, "ms_per_step": ms, "slides_per_s": 1e3 / ms,
            "loss": float(loss), "flash_attn_version": __import__("flash_attn").__version__,
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}


if __name__ == "__main__":
    tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    print(json.dumps(run(tiles, steps)))
