"""GPU: per-parameter comparison of the gradients of a CUDA-graph replay and of an eager step (pass streams on)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import config, synthetic, train_step
from modaltune_b200 import factory as helpers
DEV = "cuda"
config.set_pass_streams(os.environ.get("MODALTUNE_B200_PASS_STREAMS", "1") != "0")
model = helpers.build_model(helpers.SMALL_GROUPS, device=DEV)
proj = helpers.build_projector(0, DEV)
params = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
flat = train_step.FlatGradAllReduce([p for _, p in params])
slides = [synthetic.synthetic_slide(700, seed=s, group_sizes=helpers.SMALL_GROUPS) for s in (31, 32)]
packed = [train_step.pack_host_slide(s) for s in slides]
sizes = packed[0][1]
if os.environ.get("DIAG_EAGER_FIRST", "0") != "0":
    pre = train_step.unpack_slide({k: v.to(DEV) for k, v in packed[0][0].items()}, sizes)
    if os.environ.get("DIAG_EAGER_FIRST") == "3":
        with config.using(mode="fp32", attn_impl="simt"):
            train_step.forward_backward(model, proj, pre)
    else:
        train_step.forward_backward(model, proj, pre)
    flat.zero()
    if os.environ.get("DIAG_EAGER_FIRST") == "2":
        import gc; del pre; gc.collect(); torch.cuda.synchronize()
graphed = train_step.GraphedStep(model, proj, packed[0][0], sizes, flat)
res = []
for rep in range(3):
    graphed(packed[1][0]); torch.cuda.synchronize()
    res.append(graphed.grads.clone())
flat.zero()
dev_slide = train_step.unpack_slide({k: v.to(DEV) for k, v in packed[1][0].items()}, sizes)
train_step.forward_backward(model, proj, dev_slide)
g_eager = flat.gather().clone()
flat.zero()
train_step.forward_backward(model, proj, dev_slide)
g_eager2 = flat.gather().clone()
def cos(a, b): return float((a.double() @ b.double()) / (a.double().norm() * b.double().norm() + 1e-300))
print("graph vs graph", cos(res[0], res[1]), cos(res[1], res[2]), " eager vs eager", cos(g_eager, g_eager2), " graph vs eager", cos(res[0], g_eager))
off = 0; rows = []
for n, p in params:
    k = p.numel()
    rows.append((cos(res[0][off:off + k], g_eager[off:off + k]), cos(g_eager2[off:off + k], g_eager[off:off + k]), float(g_eager[off:off+k].norm()), n))
    off += k
rows.sort()
for r in rows[:25]: print("%.6f  (eager/eager %.6f)  |g|=%.3e  %s" % r)
