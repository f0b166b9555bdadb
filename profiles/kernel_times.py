"""Per-kernel device time of the default bench workload via torch.profiler (CUPTI), warm caches.
Usage (GPU box): python profiles/kernel_times.py [workload] [steps]"""
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from tapqir_b200.models.cosmos import cosmos  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
ds, nb, fb, desc = bench.make_shard(workload, 0, 1, dev)
model = cosmos(device="cuda:0", dtype="float")
model.data = ds
model.init(nbatch_size=nb, fbatch_size=fb)
for _ in range(5):
    model.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        model.step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / steps, e.count / steps) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"workload {workload}: {desc}")
for name, us, n in rows:
    print(f"{name[:70]:70s} {us:9.1f} us/step  x{n:.1f}  {100 * us / tot:5.1f}%")
print(f"{'sum of kernels':70s} {tot:9.1f} us/step")
