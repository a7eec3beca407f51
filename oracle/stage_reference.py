"""TEST / MEASUREMENT INFRASTRUCTURE: stage the reference tree for the GPU box.

``/root/reference`` exists only in the build container.  The reference is pure Python (nothing to compile), so the
"build" of ``oracle/_ref`` is a file copy: the ``.py`` sources, the model configs and the pathway table the trainer reads
(``dataset/gene_pathway_processed_v2.csv``) go to ``oracle/_ref/reference/`` -- a git-ignored directory that travels to
the GPU box with the gpurun snapshot exactly like the built ``.so`` does, and is never committed.  What uses it there:

* ``modaltune_b200.launcher``: runs the reference's byte-unchanged ``train_modaltune.py`` /
  ``train_modaltune_pancancer.py`` on the B200 with the B200 module classes swapped in;
* the ``-m gpu`` launcher test and ``bench.py``'s ``reference_gpu`` comparator: the reference's own GPU path (its
  modules under autocast with the installed flash-attn) on the same box.

    python oracle/stage_reference.py [--src /root/reference]
"""
import argparse
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "reference")
KEEP_DIRS = ("models", "utils", "data_utils", "model_configs")
KEEP_FILES = ("train_modaltune.py", "train_modaltune_pancancer.py", "dataset/gene_pathway_processed_v2.csv")


def stage(src: str = "/root/reference") -> str:
    if not os.path.isdir(src):
        raise FileNotFoundError(f"reference tree not found at {src}")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for d in KEEP_DIRS:
        shutil.copytree(os.path.join(src, d), os.path.join(DST, d),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pt", "*.pth", "*.bin"))
    for f in KEEP_FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
        shutil.copyfile(os.path.join(src, f), os.path.join(DST, f))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as fh:
        fh.write(f"{src}\nstaged by oracle/stage_reference.py; git-ignored; do not edit\n")
    return DST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    print(stage(ap.parse_args().src))
