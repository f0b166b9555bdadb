"""Parameters of a TRAINED C2 model (bench.py's workload: 100 AOIs x 1000 frames, seed 0) for offline study of the
guide-site regimes: python profiles/dump_trained.py [iters] -> gpurun_out/trained_c2.pt (unconstrained, fp32),
plus the worklist length (sites redone in double) and the step time at that point."""
import sys; sys.path.insert(0, '.')
import os, torch
from tapqir_b200.models import models
from tapqir_b200.utils.simulate import simulate

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ds = simulate(100, 1000, C=1, P=14, seed=0, device="cuda", aoi_chunk=50)
m = models["cosmos"](device="cuda", dtype="float")
m.data = ds
m.init(lr=0.005, nbatch_size=100, fbatch_size=1000)
for i in range(iters):
    m.step()
torch.cuda.synchronize()
eng = m.engine
counts = []
for _ in range(5):
    m.step(); torch.cuda.synchronize()
    counts.append(int(eng.work_count[0].item()))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(20):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.step(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"iters {iters}: worklist {counts} of {9 * eng.nb * eng.fb} sites; step {sum(ts) / len(ts) * 1e3:.1f} us")
os.makedirs("gpurun_out", exist_ok=True)
torch.save({k: v.detach().cpu().clone() for k, v in eng.named_unconstrained().items()}, "gpurun_out/trained_c2.pt")
