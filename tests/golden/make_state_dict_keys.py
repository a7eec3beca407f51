"""Build-container only: dump the parameter/buffer names and shapes of the REFERENCE models (full 331-pathway config)
to tests/golden/state_dict_keys.json, so tests/test_state_dict.py can check checkpoint compatibility without the
reference tree."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

out = {}
for clinical in (True, False):
    model, _ = ref_shims.build_reference_model(clinical=clinical, multi_task=3)
    name = "longnetvit_gene_clinical_adapter" if clinical else "longnetvit_gene_adapter"
    out[name] = {
        "state_dict": {k: list(v.shape) for k, v in model.state_dict().items()},
        "trainable": sorted(k for k, p in model.named_parameters() if p.requires_grad),
        "n_frozen": sum(p.numel() for p in model.parameters() if not p.requires_grad),
        "n_trainable": sum(p.numel() for p in model.parameters() if p.requires_grad),
    }
    print(name, len(out[name]["state_dict"]), out[name]["n_frozen"], out[name]["n_trainable"])
with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
    json.dump(out, f)
