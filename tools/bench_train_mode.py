"""GPU: the train-mode step (model.train(), Dropout / DropPath in the frozen encoder, AdamW) as a captured CUDA graph,
timed alone; environment switches of modaltune_b200.config select the A / B variants."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from modaltune_b200 import factory, synthetic, train_step
L = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = "cuda"
model = factory.build_model(None, device=dev)
proj = factory.build_projector(0, dev)
flat = train_step.FlatGradAllReduce([p for p in model.parameters() if p.requires_grad])
host = train_step.pack_host_slide(synthetic.synthetic_slide(L, seed=4000))
opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-6)
model.train()
g = train_step.GraphedStep(model, proj, {k: v.to(dev) for k, v in host[0].items()}, host[1], flat)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
def it(timed=False):
    opt.zero_grad()
    if timed: ev[0].record()
    g()
    if timed: ev[1].record()
    opt.step()
    if timed: ev[2].record()
for _ in range(3): it()
torch.cuda.synchronize()
tg = to = 0.0
for _ in range(8):
    it(True); torch.cuda.synchronize()
    tg += ev[0].elapsed_time(ev[1]); to += ev[1].elapsed_time(ev[2])
sw = {k: v for k, v in os.environ.items() if k.startswith("MODALTUNE_B200_")}
print(f"train-mode graph step {tg / 8:.2f} ms + AdamW {to / 8:.2f} ms  {sw}")
