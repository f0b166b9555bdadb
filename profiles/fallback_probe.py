"""How often does a TRAINED model leave the fp32 site forms?  Trains on simulated data, then evaluates the production
site function (host build, tests/hostcheck) on the trained parameters with fresh draws and counts the statuses.
python profiles/fallback_probe.py [cosmos|cosmos+hmm] [iters]"""
import ctypes, sys; sys.path.insert(0, '.')
import numpy as np, torch
from tapqir_b200.models import models, layout as L
from tapqir_b200.utils.simulate import simulate
from tests import hostcheck
from oracle import cosmos_oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "cosmos"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
prm = {"kon": 0.2, "koff": 0.2} if name == "cosmos+hmm" else None
ds = simulate(20, 300, C=1, P=14, seed=3, params=prm, device="cuda")
m = models[name](device="cuda", dtype="float")
m.data = ds
m.init(lr=0.005, nbatch_size=20, fbatch_size=300)
for i in range(iters):
    l = m.step()
print("loss", float(l.item()))
p = {k: v.detach().cpu() for k, v in m.engine.named_unconstrained().items()}
hc = hostcheck.load()
mc = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float64)
hc.hc_site_status_rng.restype = ctypes.c_int
rng = np.random.default_rng(0)
sites = {"b": (0, "b_loc", "b_beta"), "h": (1, "h_loc", "h_beta"), "w": (3, "w_mean", "w_size"), "x": (5, "x_mean", "size"), "y": (7, "y_mean", "size")}
for tag, (s, n0, n1) in sites.items():
    a, b = p[n0].flatten().numpy(), p[n1].flatten().numpy()
    idx = rng.choice(len(a), size=min(4000, len(a)), replace=False)
    st = np.array([hc.hc_site_status_rng(s, ctypes.c_float(a[i]), ctypes.c_float(b[i]), ctypes.c_float(5.0), ctypes.c_float(3.0),
                                         ctypes.byref(mc), ctypes.c_uint64(i)) for i in idx])
    print(f"site {tag}: done {np.mean(st == 0):.4f}  fallback {np.mean(st == 1):.4f}  fallback-before-draw {np.mean(st == 2):.4f}")
print("size quantiles", np.quantile(2 + np.exp(p["size"].flatten().numpy()), [0.05, 0.25, 0.5, 0.75, 0.95]).round(1))
print("w_size quantiles", np.quantile(2 + np.exp(p["w_size"].flatten().numpy()), [0.05, 0.25, 0.5, 0.75, 0.95]).round(1))
print("h conc quantiles", np.quantile(np.exp(p["h_loc"].flatten().numpy() + p["h_beta"].flatten().numpy()), [0.05, 0.25, 0.5, 0.75, 0.95]).round(2))

# step time with the TRAINED parameters (the benchmark's default is the initial point, where every site is in the fp32 regime)
eng = m.engine
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for _ in range(3):
    m.step()
ts = []
for _ in range(20):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.step(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"trained-state step: {np.mean(ts)*1e3:.1f} us for {eng.nb * eng.fb} units")
