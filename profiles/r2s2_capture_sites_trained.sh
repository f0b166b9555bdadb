# One 8-GPU rank's C3 shard (625 k units) after 3000 SVI iterations -- the state a long fit spends its time in:
#  (1) parameters of the first 8 AOIs dumped for offline regime analysis, (2) ncu --set full of site_fast_kernel with the
#  source-correlated page exported on the box
mkdir -p gpurun_out
python - <<'P'
import sys; sys.path.insert(0, '.')
import torch, bench
from tapqir_b200.models.cosmos import cosmos
dev = torch.device("cuda", 0)
ds, nb, fb, desc = bench.make_shard("c3s8", 0, 1, dev)
m = cosmos(device="cuda:0", dtype="float"); m.data = ds; m.init(nbatch_size=nb, fbatch_size=fb)
for i in range(3000): m.step()
torch.cuda.synchronize()
u = m.engine.named_unconstrained()
sub = {k: (v[:, :8] if v.dim() == 4 else v[:8] if v.dim() == 3 else v).detach().cpu().clone() for k, v in u.items()}
torch.save(sub, "gpurun_out/trained_c3s8_sub.pt")
print({k: tuple(v.shape) for k, v in sub.items()})
P
CMD="python bench.py --workload c3s8 --steps 3 --warmup 3 --no-cpu-baseline --no-subs --trained-iters 0 --train-iters 3000"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:site_fast_kernel -s 3004 -c 1 \
    -o gpurun_out/sites_trained_s2 $CMD > gpurun_out/ncu_sites_trained_s2.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_sites_trained_s2.log
ncu -i gpurun_out/sites_trained_s2.ncu-rep --page raw --csv > gpurun_out/sites_trained_s2_raw.csv 2>/dev/null
ncu -i gpurun_out/sites_trained_s2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/sites_trained_s2_src.csv 2>/dev/null
rm -f gpurun_out/sites_trained_s2.ncu-rep
ls -la gpurun_out/sites_trained_s2* gpurun_out/trained_c3s8_sub.pt
