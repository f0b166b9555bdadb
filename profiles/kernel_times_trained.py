"""Per-kernel device time (torch.profiler) at the initial point and after N SVI iterations: what a long fit spends its time in.
Usage (GPU box): python profiles/kernel_times_trained.py [workload] [train_iters]"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from tapqir_b200.models.cosmos import cosmos  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
dev = torch.device("cuda", 0)
ds, nb, fb, desc = bench.make_shard(workload, 0, 1, dev)
model = cosmos(device="cuda:0", dtype="float")
model.data = ds
model.init(nbatch_size=nb, fbatch_size=fb)


def report(tag, steps=10):
    for _ in range(3):
        model.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            model.step()
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total / steps) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    print(f"--- {workload} {tag}: " + "; ".join(f"{n.split('(')[0].replace('void tq::', '').replace('tq::', '')[:28]} {us:.0f}" for n, us in rows[:6])
          + f"; sum {sum(r[1] for r in rows):.0f} us/step", flush=True)


report("initial point")
done = 0
for target in (1000, iters):
    for _ in range(target - done):
        model.step()
    done = target
    report(f"after {done} iterations")
eng = model.engine
u = eng.named_unconstrained()
hc = (u["h_loc"] + u["h_beta"]).exp()
print(f"height concentrations < 10: {(hc < 10).float().mean().item():.2f}; size (x, y guide) median {(2 + u['size'].exp()).median().item():.0f}; "
      f"w_size median {(2 + u['w_size'].exp()).median().item():.0f}; m_probs < 0.5: {(torch.sigmoid(u['m_probs']) < 0.5).float().mean().item():.2f}")
