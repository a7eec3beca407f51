"""GPU: which aten ops (by input shape) account for copy / add / fill / reduce kernel time in one step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from modaltune_b200 import synthetic, train_step
from modaltune_b200 import factory as helpers
dev = "cuda"
model = helpers.build_model(None, device=dev)
proj = helpers.build_projector(0, dev)
flat = train_step.FlatGradAllReduce([p for p in model.parameters() if p.requires_grad])
slide = train_step.slide_to_device(synthetic.synthetic_slide(10000, seed=1), dev)
def step():
    flat.zero()
    return train_step.forward_backward(model, proj, slide)
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
ev = prof.key_averages(group_by_input_shape=True)
rows = sorted(((e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:90]) for e in ev if e.self_device_time_total > 0), reverse=True)
for t, c, k, sh in rows[:45]:
    print(f"{t/1e3:8.3f} ms {c:5d}x {k[:38]:38s} {sh}")
