"""Row N2 at BASELINE config 3 size (1000 AOIs x 5000 frames, K = 2): device time of the credible intervals over the whole
(K, Nt, F, Q) arrays and of SNR / chi2 over the resident pixels, next to scipy on the host for a sample of the same
elements.   Usage (GPU box): python profiles/r2_stats_timing.py [Nt F]"""
import ctypes
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tapqir_b200 import _lib  # noqa: E402

Nt, F = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1000, 5000)
lib = _lib.load()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
U = Nt * F
r = lambda n: torch.rand(n, generator=g, device=dev, dtype=torch.float64)
# guides as a fit leaves them: background Gamma(150 * beta, beta), beta ~ 0.5-5; height of present / absent spots; width,
# x, y AffineBeta with sizes 10-2000
cases = {
    "background (Gamma, U)": ("gamma", 150 * (0.5 + 4.5 * r(U)), 0.5 + 4.5 * r(U)),
    "height (Gamma, 2U)": ("gamma", torch.exp(np.log(0.5) + np.log(6000) * r(2 * U)), 1e-3 + r(2 * U)),
    "width (Beta, 2U)": ("beta", 5 + 500 * r(2 * U), 5 + 500 * r(2 * U)),
    "x (Beta, 2U)": ("beta", 5 + 1000 * r(2 * U), 5 + 1000 * r(2 * U)),
    "y (Beta, 2U)": ("beta", 5 + 1000 * r(2 * U), 5 + 1000 * r(2 * U)),
}
total = 0.0
for name, (fam, a, b) in cases.items():
    lo, hi = torch.empty_like(a), torch.empty_like(a)
    fn = lib.tq_gamma_interval if fam == "gamma" else lib.tq_beta_interval
    fn(1024, _lib.ptr(a), _lib.ptr(b), ctypes.c_double(0.95), _lib.ptr(lo), _lib.ptr(hi), _lib.stream_ptr())   # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(fn(a.numel(), _lib.ptr(a), _lib.ptr(b), ctypes.c_double(0.95), _lib.ptr(lo), _lib.ptr(hi), _lib.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    total += ms
    # scipy on a sample of the same elements
    import scipy.stats as st

    k = 20000
    an, bn = a[:k].cpu().numpy(), b[:k].cpu().numpy()
    t0 = time.perf_counter()
    L, Uq = (st.gamma(an, scale=1 / bn) if fam == "gamma" else st.beta(an, bn)).interval(0.95)
    dt = time.perf_counter() - t0
    err = max(np.max(np.abs(lo[:k].cpu().numpy() - L) / L), np.max(np.abs(hi[:k].cpu().numpy() - Uq) / Uq))
    print(f"{name:24s} {a.numel() / 1e6:6.1f} M elements: {ms:8.1f} ms on the device; scipy {dt / k * 1e6:6.2f} us/element "
          f"-> {dt / k * a.numel():7.1f} s for the array on one host core; max rel. difference on the sample {err:.1e}")
print(f"all intervals of a {Nt} x {F} fit: {total:.0f} ms on the device")

P = 14
pix = torch.randint(200, 400, (U, P, P), generator=g, device=dev, dtype=torch.int32).to(torch.uint16)
xy = torch.full((U, 2), 6.5, device=dev)
f = lambda n, lo_, hi_: (lo_ + (hi_ - lo_) * torch.rand(n, generator=g, device=dev)).float()
h, w, x, y, b = f(2 * U, 500, 3000), f(2 * U, 1.0, 2.0), f(2 * U, -3, 3), f(2 * U, -3, 3), f(U, 100, 200)
snr, chi2 = torch.empty(2 * U, device=dev), torch.empty(U, device=dev)
for it in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.tq_snr_chi2(U, P, _lib.TQ_PIX_U16, _lib.ptr(pix), _lib.ptr(xy), _lib.ptr(h), _lib.ptr(w), _lib.ptr(x), _lib.ptr(y),
                               _lib.ptr(b), 7.0, 90.0, 0.5, _lib.ptr(snr), _lib.ptr(chi2), _lib.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
print(f"SNR + chi2 over {U / 1e6:.1f} M resident patches: {e0.elapsed_time(e1):.1f} ms "
      f"({U * P * P * 2 / e0.elapsed_time(e1) / 1e6:.0f} GB/s of pixels)")
