"""Exploratory fits on simulated data (validation aid, not a test): classification quality vs the simulated labels and a
NaN hunt.  python profiles/fit_probe.py [cosmos|cosmos+hmm] [iters]"""
import sys; sys.path.insert(0, '.')
import torch, numpy as np
from tapqir_b200.models import models
from tapqir_b200.utils.simulate import simulate

def mcc(pred, true):
    tp = ((pred == 1) & (true == 1)).sum(); tn = ((pred == 0) & (true == 0)).sum()
    fp = ((pred == 1) & (true == 0)).sum(); fn = ((pred == 0) & (true == 1)).sum()
    d = np.sqrt(float(tp + fp) * float(tp + fn) * float(tn + fp) * float(tn + fn))
    return (float(tp) * float(tn) - float(fp) * float(fn)) / d if d else 0.0

name = sys.argv[1] if len(sys.argv) > 1 else "cosmos+hmm"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
prm = {"kon": 0.2, "koff": 0.2} if name == "cosmos+hmm" else None
ds = simulate(20, 300, C=1, P=14, seed=3, params=prm, device="cuda")
m = models[name](device="cuda", dtype="float")
m.data = ds
m.init(lr=0.005, nbatch_size=20, fbatch_size=300)
m.engine.use_graph = False
eng = m.engine
for i in range(iters):
    l = m.step()
    if i % 100 == 0 or i == iters - 1:
        lv = float(l.item())
        if lv != lv:
            print("NaN loss at iteration", i)
            for nm in ("acc", "hacc", "rec", "Lm", "samples", "gs", "g_rate", "chain_v", "chain_a", "chain_rows", "qm", "lgrads", "ggrads", "lparams", "gparams", "gstate"):
                t = getattr(eng, nm, None)
                if t is not None:
                    bad = ~torch.isfinite(t)
                    print(f"  {nm:10s} nonfinite {int(bad.sum())} / {t.numel()}", (bad.nonzero()[:4].flatten().tolist() if bad.any() else ""))
            if hasattr(eng, "acc"):
                print("  acc", eng.acc.tolist())
            rec = eng.rec
            bad = (~torch.isfinite(rec)).nonzero()
            if len(bad):
                r, u = bad[0].tolist()
                print("  first bad rec row", r, "unit", u, "site", r // 6, "entry", r % 6)
                print("  samples of that unit", eng.samples[:, u].tolist())
                names = eng.named_unconstrained()
                n_, f_ = u // eng.F, u % eng.F
                for k in ("b_loc", "b_beta", "h_loc", "h_beta", "w_mean", "w_size", "x_mean", "y_mean", "size"):
                    v = names[k]
                    print("   ", k, (v[n_, f_, 0].item() if v.dim() == 3 else v[:, n_, f_, 0].tolist()))
            break
        if i % 1000 == 0:
            print(i, lv, flush=True)
torch.cuda.synchronize()
zp = m.z_probs
pred = (zp[:10, :, 0, 1] > 0.5).numpy().astype(int)
true = ds.labels["z"][:, :, 0]
print(name, "iters", i + 1, "MCC", round(mcc(pred, true), 4), "mean z true/pred", true.mean().round(3), pred.mean().round(3))
