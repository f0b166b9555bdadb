"""Run under torchrun (N >= 2): a few SVI steps on per-rank simulated shards; prints the loss sequence of rank 0 and whether
all ranks hold bit-identical global parameters.  Compare TQ_ALLREDUCE=p2p (NVLink peer-memory push) with =nccl."""
import os, sys; sys.path.insert(0, '.')
import torch
from tapqir_b200.models import models
from tapqir_b200.utils.simulate import simulate

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.distributed.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "cosmos"
ds = simulate(6, 40, C=1, P=14, seed=rank, device=dev, params={"kon": 0.2, "koff": 0.2} if name == "cosmos+hmm" else None)
m = models[name](device=str(dev), dtype="float")
m.data = ds
m.init(lr=0.005, nbatch_size=6, fbatch_size=40, rank=rank, world_size=world, presharded=True)
losses = [float(m.step().item()) for _ in range(12)]
g = m.engine.gparams.clone()
gathered = [torch.empty_like(g) for _ in range(world)]
torch.distributed.all_gather(gathered, g)
same = all(torch.equal(gathered[0], t) for t in gathered)
if rank == 0:
    mode = "p2p" if m.engine.p2p is not None else "nccl"
    print(f"{name} allreduce={mode} world={world} identical_globals={same} losses={[round(l, 3) for l in losses[:3]]}...{round(losses[-1], 3)}"
          + (f" timeout_marker={m.engine.p2p.timed_out()}" if m.engine.p2p is not None else ""), flush=True)
torch.cuda.synchronize()
m.engine.release_graph()
os._exit(0)
