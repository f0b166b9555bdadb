"""Which guide sites of a TRAINED model leave the fp32 forms, and why.  Offline (CPU): reads gpurun_out/trained_c2.pt
(profiles/dump_trained.py) and runs the production site function (host build, tests/hostcheck) with fresh draws.
python profiles/site_regimes.py [file]"""
import ctypes, sys; sys.path.insert(0, '.')
import numpy as np, torch
from tapqir_b200.models import layout as L
from tests import hostcheck
from oracle import cosmos_oracle as O

p = torch.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trained_c2.pt")
hc = hostcheck.load()
mc = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float64)
f32 = lambda t: np.ascontiguousarray(t.flatten().float().numpy())
ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
sites = {"b": (0, "b_loc", "b_beta", None), "h0": (1, "h_loc", "h_beta", 0), "h1": (2, "h_loc", "h_beta", 1),
         "w0": (3, "w_mean", "w_size", 0), "w1": (4, "w_mean", "w_size", 1), "x0": (5, "x_mean", "size", 0), "x1": (6, "x_mean", "size", 1),
         "y0": (7, "y_mean", "size", 0), "y1": (8, "y_mean", "size", 1)}
tot_fb = 0
for tag, (s, n0, n1, k) in sites.items():
    a, b = (p[n0], p[n1]) if k is None else (p[n0][k], p[n1][k])
    a, b = f32(a), f32(b)
    n = len(a)
    if s == 0:
        bm = f32(p["background_mean_loc"].expand_as(p[n0])); bs = f32(p["background_std_loc"].expand_as(p[n0]))
    else:
        bm = bs = np.zeros(n, np.float32)
    st = np.zeros(n, np.int32); var = np.zeros(n, np.float64)
    hc.hc_site_status_batch(s, n, ptr(a), ptr(b), ptr(bm), ptr(bs), ctypes.byref(mc), ctypes.c_uint64(5), ptr(st), ptr(var))
    fb = st != 0
    tot_fb += fb.sum()
    line = f"site {tag}: fallback {fb.mean():.4f} (before draw {np.mean(st == 2):.4f})"
    if s >= 3:
        m1 = 1 / (1 + np.exp(-a.astype(np.float64))); S = 2 + np.exp(b.astype(np.float64)); c1, c0 = S * m1, S * (1 - m1)
        x = var; y = 1 - x
        bound = S * x * y
        tierA = (c1 > 6) & (c0 > 6) & (bound >= 2.5)
        cause = {
            "S>64 outside A": fb & ~tierA & (S > 64),
            "series bx>=2": fb & ~tierA & (S <= 64) & (((x <= 0.5) & (bound < 2.5) & (c0 * x >= 2)) | ((y <= 0.5) & (bound < 2.5) & (c1 * y >= 2))),
            "mirror ay>=2": fb & ~tierA & (S <= 64) & (((x >= 0.5) & (bound < 0.75) & (c1 * y >= 2)) | ((y >= 0.5) & (bound < 0.75) & (c0 * x >= 2))),
            "rice": fb & ~tierA & (S <= 64) & (c1 > 6) & (c0 > 6),
        }
        line += "  S quantiles " + str(np.quantile(S, [0.05, 0.5, 0.95]).round(1)) + "  tierA " + f"{tierA.mean():.3f}  "
        line += "  ".join(f"{k}: {v.mean():.4f}" for k, v in cause.items())
    else:
        conc = np.exp(a.astype(np.float64) + b.astype(np.float64))
        line += "  conc quantiles " + str(np.quantile(conc, [0.05, 0.5, 0.95]).round(2)) + f"  x<0.8&conc>30: {np.mean(fb & (var < 0.8) & (conc > 30)):.4f}"
    print(line)
print("total fallback sites per step:", tot_fb)
