"""Run under torchrun (N >= 2): the multi-rank fit path end to end on real GPUs.

  1. a dataset of 7 AOIs (NOT divisible by the ranks) saved by rank 0, loaded by every rank; Model.init shards it into
     balanced AOI blocks; 201 iterations of Model.run (checkpoints at iterations 0 and 200, one `.rank<r>` file each);
  2. every rank holds bit-identical global parameters; the peer-memory all-reduce reports no timeout;
  3. run() consolidated the rank files into ONE reference-layout cosmos_model.tpqr: rank 0 resumes it on a single GPU
     (world_size 1) and finds the concatenation of every rank's AOI-local parameters + Adam moments;
  4. a second 2-rank model resumes from the consolidated file alone (rank files removed): each rank gets its block back;
  5. NaN injected into ONE rank's parameters: every rank raises together at the next checkpoint, rank 0's new seed is
     used by all, the fit restarts from the checkpoint and continues (no deadlock).
"""
import logging
import os
import shutil
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from tapqir_b200.models import models  # noqa: E402
from tapqir_b200.utils.dataset import save  # noqa: E402
from tapqir_b200.utils.simulate import simulate  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.distributed.init_process_group("nccl", device_id=dev)
logging.basicConfig(level=logging.WARNING)
box = [tempfile.mkdtemp(prefix="tq_mg_") if rank == 0 else None]
torch.distributed.broadcast_object_list(box, src=0)
path = Path(box[0])
if rank == 0:
    save(simulate(7, 60, C=1, P=14, seed=0), path)
torch.distributed.barrier()
quiet = lambda it: it

m = models["cosmos"](device=str(dev), dtype="float")
m.load(path)
m.init(lr=0.005, nbatch_size=2, fbatch_size=30, rank=rank, world_size=world)
sizes = m._shard_sizes()
assert sum(sizes) == 7 and max(sizes) - min(sizes) <= 1 and m.engine.Nt == sizes[rank]
m.run(201, progress_bar=quiet)
g = m.engine.gparams.clone()
gathered = [torch.empty_like(g) for _ in range(world)]
torch.distributed.all_gather(gathered, g)
assert all(torch.equal(gathered[0], t) for t in gathered), "ranks hold different global parameters"
assert m.engine.p2p is None or m.engine.p2p.timed_out() == 0
mode = "p2p" if m.engine.p2p is not None else "nccl"
files = sorted(p.name for p in (path / ".tapqir").iterdir() if p.name.startswith("cosmos_model"))
assert files == ["cosmos_model.tpqr"] + [f"cosmos_model.tpqr.rank{r}" for r in range(world)], files

# 3. single-GPU resume of the consolidated file (rank 0 only; the others keep their own blocks for the comparison)
ckpt_rank = torch.load(path / ".tapqir" / f"cosmos_model.tpqr.rank{rank}", map_location="cpu", weights_only=False)
if rank == 0:
    one = models["cosmos"](device=str(dev), dtype="float")
    one.load(path)
    one.init(lr=0.005, nbatch_size=2, fbatch_size=30)
    assert one.iter == 200 and one.engine.Nt == 7
    whole = {k: v.cpu() for k, v in one.engine.named_unconstrained().items()}
    one.engine.close()
else:
    whole = None
lo = sum(sizes[:rank])
for k, v in ckpt_rank["params"]["params"].items():
    box = [whole[k] if rank == 0 else None]
    torch.distributed.broadcast_object_list(box, src=0)
    ref = box[0]
    axis = m._aoi_axis(k, v.dim())
    blk = ref if axis is None else ref.narrow(axis, lo, sizes[rank])
    assert torch.equal(blk, v.cpu()), k

# 4. multi-rank resume from the consolidated file alone
torch.distributed.barrier()
(path / ".tapqir" / f"cosmos_model.tpqr.rank{rank}").unlink()
torch.distributed.barrier()
again = models["cosmos"](device=str(dev), dtype="float")
again.load(path)
again.init(lr=0.005, nbatch_size=2, fbatch_size=30, rank=rank, world_size=world)
assert again.iter == 200
for k, v in ckpt_rank["params"]["params"].items():
    assert torch.equal(again.engine.named_unconstrained()[k].cpu(), v.cpu()), k
assert torch.equal(again.engine.lm.cpu(), m.engine.ll.pack({k: ckpt_rank["optimizer"][k]["state"][0]["exp_avg"] for k in m.engine.ll.shapes}).to(again.engine.lm.dtype))

# 5. NaN on one rank -> collective restart
again.iter = 399          # the next iteration checkpoints
if rank == world - 1:
    again.engine.named_unconstrained()["h_loc"][0, 0, 0, 0] = float("nan")
again.run(3, progress_bar=quiet)
assert bool(torch.isfinite(again.engine.lparams).all()) and again.iter >= 200
seeds = [None] * world
torch.distributed.all_gather_object(seeds, again.seed)
assert len(set(seeds)) == 1, seeds
if rank == 0:
    print(f"multi-rank fit OK: world={world} allreduce={mode} shards={sizes} files={files} restart_seed={seeds[0]} iter={again.iter}", flush=True)
    shutil.rmtree(path, ignore_errors=True)
torch.distributed.barrier()
torch.cuda.synchronize()
m.engine.close()
again.engine.close()
torch.distributed.destroy_process_group()
