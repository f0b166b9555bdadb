"""Is a fit run-to-run deterministic?  Two models on the same data and seed, stepped in lockstep; parameters compared bit
for bit every `every` iterations (first divergence reported), then per-kernel time of the site kernel of both.
Usage (GPU box): python profiles/r2s2_determinism.py [workload] [iters] [every] [models]
(models = 1: one fit, printing an exact checksum of the parameters -- to compare builds / switches such as TQ_SITE_SPLIT)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from tapqir_b200.models.cosmos import cosmos  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c3s8"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
every = int(sys.argv[3]) if len(sys.argv) > 3 else 250
n_models = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda", 0)
ds, nb, fb, desc = bench.make_shard(workload, 0, 1, dev)
models = []
for _ in range(n_models):
    m = cosmos(device="cuda:0", dtype="float")
    m.data = ds
    m.init(nbatch_size=nb, fbatch_size=fb)
    models.append(m)
first = None
for it in range(1, iters + 1):
    losses = [m.step() for m in models]
    if it % every == 0 or it == iters:
        a, b = models[0].engine, models[-1].engine
        checksum = int(a.lparams.view(torch.int32).to(torch.int64).sum().item()) ^ int(a.gparams.view(torch.int64).sum().item())
        same_l = torch.equal(a.lparams, b.lparams)
        same_g = torch.equal(a.gparams, b.gparams)
        dl = (a.lparams - b.lparams).abs().max().item()
        wc = [int(e.work_count[0].item()) for e in (a, b)]
        print(f"iter {it}: local identical {same_l} (max diff {dl:.3e}), global identical {same_g}, loss {losses[0].item():.10e} / {losses[-1].item():.10e}, "
              f"worklist {wc}, parameter checksum {checksum}", flush=True)
        if first is None and not (same_l and same_g):
            first = it
print("first divergence at or before iteration", first)
