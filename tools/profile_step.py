"""GPU: where does a slide step spend its time?  torch.profiler kernel table + GPU-busy vs wall."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from modaltune_b200 import config, synthetic, train_step, ops
from modaltune_b200 import factory as helpers
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = "cuda"
model = helpers.build_model(None, device=dev)
proj = helpers.build_projector(0, dev)
flat = train_step.FlatGradAllReduce([p for p in model.parameters() if p.requires_grad])
slide = train_step.slide_to_device(synthetic.synthetic_slide(L, seed=1), dev)
def step():
    flat.zero()
    return train_step.forward_backward(model, proj, slide)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 3 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = prof.key_averages()
rows = sorted(((e.device_time_total, e.count, e.key) for e in ev if e.device_time_total > 0 and e.device_type.name == "CUDA"), reverse=True)
tot = sum(r[0] for r in rows)
print(f"GPU kernel time total {tot/1e3:.2f} ms over {sum(r[1] for r in rows)} kernels")
for t, c, k in rows[:70]:
    print(f"{t/1e3:9.3f} ms {c:6d}x {100*t/tot:5.1f}%  {k[:110]}")
if "--ops" in sys.argv:   # ATen operators by device time, grouped by input shape: where the autograd glue comes from
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof2:
        step(); torch.cuda.synchronize()
    ops_ = [e for e in prof2.key_averages(group_by_input_shape=True) if e.device_type.name == "CPU" and e.self_device_time_total > 0]
    ops_.sort(key=lambda e: -e.self_device_time_total)
    print("\nATen operators by SELF device time (grouped by input shapes)")
    for e in ops_[:40]:
        print(f"{e.self_device_time_total/1e3:9.3f} ms {e.count:5d}x  {e.key[:40]:40s} {str(e.input_shapes)[:110]}")
