"""Regime classes of the guide sites of a TRAINED model, offline: reads gpurun_out/trained_c3s8_sub.pt (parameters of the first
8 AOIs of one 8-GPU rank's C3 shard after 3000 iterations, written by profiles/r2s2_capture_sites_trained.sh), draws every site once
with the host build of the production site function and counts the deferred sites per class (csrc/cosmos_sites_fast.cuh).
python profiles/r2s2_site_classes.py"""
import ctypes, sys; sys.path.insert(0, '.')
import numpy as np, torch, collections
from tapqir_b200.models import layout as L
from tests import hostcheck
from oracle import cosmos_oracle as O
p = torch.load("gpurun_out/trained_c3s8_sub.pt")
hc = hostcheck.load()
mc = L.ModelConst.make(O.DEFAULT_PRIORS, 14, torch.float64)
f32 = lambda t: np.ascontiguousarray(t.flatten().float().numpy())
ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
sites = {"b": (0, "b_loc", "b_beta", None), "h0": (1, "h_loc", "h_beta", 0), "h1": (2, "h_loc", "h_beta", 1),
         "w0": (3, "w_mean", "w_size", 0), "w1": (4, "w_mean", "w_size", 1), "x0": (5, "x_mean", "size", 0), "x1": (6, "x_mean", "size", 1),
         "y0": (7, "y_mean", "size", 0), "y1": (8, "y_mean", "size", 1)}
def branch(x, boundary, a, b):
    r = np.where((x <= 0.5) & (boundary < 2.5), 0, np.where((x >= 0.5) & (boundary < 0.75), 1, np.where((a > 6) & (b > 6), 2, 3)))
    return r
tot = 0; totdef = 0
allcls = collections.Counter()
for tag, (s, n0, n1, k) in sites.items():
    a, b = (p[n0], p[n1]) if k is None else (p[n0][k], p[n1][k])
    a, b = f32(a), f32(b); n = len(a)
    if s == 0:
        bm = f32(p["background_mean_loc"].expand_as(p[n0])); bs = f32(p["background_std_loc"].expand_as(p[n0]))
    else:
        bm = bs = np.zeros(n, np.float32)
    st = np.zeros(n, np.int32); var = np.zeros(n, np.float64)
    hc.hc_site_status_batch(s, n, ptr(a), ptr(b), ptr(bm), ptr(bs), ctypes.byref(mc), ctypes.c_uint64(5), ptr(st), ptr(var))
    if s < 3:
        conc = np.exp(a.astype(np.float64) + b.astype(np.float64)); x = var
        bulk = (conc > 10) & (x >= 0.8)
        cls = (x < 0.8) * 1 + (conc > 8) * 2 + (conc > 10) * 4
        extra = f"conc q05/50/95 {np.quantile(conc,[.05,.5,.95]).round(2)} alpha<1: {(conc<1).mean():.3f}"
    else:
        m1 = 1 / (1 + np.exp(-a.astype(np.float64))); S = 2 + np.exp(b.astype(np.float64)); c1, c0 = S * m1, S * (1 - m1)
        x = var; y = 1 - x; bound = S * x * y
        ua = (x - m1) / m1; ub = -(x - m1) / (1 - m1)
        bulk = (c1 > 6) & (c0 > 6) & (bound >= 2.5) & (ua > -1) & (ub > -1)
        dens = np.where(S <= 16, 0, np.where(c1 <= 6, 1, np.where(c0 <= 6, 2, 3)))
        cls = dens * 16 + branch(x, bound, c1, c0) * 4 + branch(y, bound, c0, c1)
        extra = f"S q05/50/95 {np.quantile(S,[.05,.5,.95]).round(1)} c<1: {((c1<1)|(c0<1)).mean():.3f}"
    d = ~bulk
    cc = collections.Counter(cls[d].tolist())
    for c_, v in cc.items(): allcls[(0 if s < 3 else 1, c_)] += v
    tot += n; totdef += d.sum()
    print(f"site {tag}: deferred {d.mean():.3f}  {extra}  top classes {[(c_, round(v/n,3)) for c_, v in cc.most_common(6)]}")
print("overall deferred", totdef / tot)
print("class mix (type, cls): share of deferred", [(k, round(v / totdef, 3)) for k, v in allcls.most_common(16)])
