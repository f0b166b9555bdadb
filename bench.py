"""
Benchmark of the cosmos SVI hot path (ELBO forward + backward + dense Adam), AOI-frames/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload c2|c3|c2mb]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full SVI step over one minibatch of synthetic (tapqir-simulated) data.
Default workload = BASELINE.json configs[1]: N=100 AOIs x F=1000 frames, P=14, K=2, C=1, O=3 offset
bins, full local batch, per GPU (weak scaling: every rank holds its own 100-AOI shard; only the
(C, 18) accumulator vector is all-reduced).  Prints ONE JSON line (rank 0).

Timing: CUDA events on the launching stream around every step, L2 flushed (256 MiB write) before
each timed step outside the event pair, max over ranks; clocks sampled with nvidia-smi during the
timed region.  `--impl reference` times the reference's CPU path restated by oracle/ (Pyro cannot be
installed here) on a bounded sample of the same workload.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (AOIs per GPU, frames, nb, fb, description)
    "c2": (100, 1000, 100, 1000, "cosmos C2: simulated N=100 AOIs x F=1000 frames per GPU, P=14, K=2, C=1, O=3, full local batch"),
    "c2mb": (100, 1000, 10, 512, "cosmos C2 with the reference-default minibatch 10 AOIs x 512 frames (main.py:1428-1431)"),
    "c3": (1000, 5000, 1000, 5000, "cosmos C3: simulated N=1000 AOIs x F=5000 frames on this GPU, full batch"),
    "c1": (5, 100, 5, 100, "cosmos C1: simulated N=5 AOIs x F=100 frames, full batch"),
    # BASELINE configs 4 and 5 (parity-test cases; here for per-config timings, not the headline line)
    "c4": (500, 2000, 500, 2000, "cosmos C4: two-channel (C=2) simulated N=500 AOIs x F=2000 frames on this GPU, full batch"),
    "c5": (200, 2000, 200, 2000, "cosmos+hmm C5: simulated N=200 AOIs x F=2000 frames on this GPU, all frames per step"),
}
WORKLOAD_CHANNELS = {"c4": 2}
WORKLOAD_MODEL = {"c5": "cosmos+hmm"}
O_BINS = 3   # offset bins of the simulated data (simulate.py:92,103): three IDENTICAL bins, merged to one on upload


def algorithmic_work(o_exec):
    """Work per unit (one 14x14 patch, K=2, 4 configurations, forward + backward) of the likelihood kernel for
    ``o_exec`` distinct offset bins -- DESIGN.md section 4.1.  SURVEY.md 8d estimates 980*O + 3276 MUFU ops and
    8232*O + 63220 flops; the figures here are the TIGHTER counts of this repository's formulation (shared
    log(D - delta_j); lgamma and digamma from one lg2 + one rcp; one rcp for 1/a and 1/sum; no exp / log-sum
    at all for a single bin), so that `frac` cannot be flattered by work the kernel does not need:
      MUFU:  P^2 * O [lg2(D - delta)] + M P^2 * (O + 3) [ex2 per bin, rcp, lg2 a, lg2 sum]  (O > 1)
             P^2     + M P^2 * 2      [rcp a, lg2 a]                                          (O = 1)
             + 2 K P [separable render]
      FP32:  lane-operations (an FMA counts once) of the packed two-pixel sweep, counted in its SASS:
             98 pairs * (2 * packed + scalar) + ~600 per-patch prologue/epilogue."""
    P2, M, K, P = 196, 4, 2, 14
    if o_exec == 1:
        mufu = P2 + M * P2 * 2 + 2 * K * P
        fp32_ops = 98 * (2 * 85 + 14) + 600               # 47 FFMA2 + 24 FADD2 + 14 FMUL2, 14 scalar per pair
    else:
        mufu = P2 * o_exec + M * P2 * (o_exec + 3) + 2 * K * P
        fp32_ops = 98 * (2 * (35 + 53 * o_exec) - 2 * 1 + 13) + 600   # 193 packed at O = 3
    return mufu, fp32_ops


HBM_BYTES_PER_UNIT = 604 + 504
KSMOGN_HBM_BYTES_PER_UNIT = 392 + 8 + 4 * 9 + 4 * 4 + 4 * 4 + 4 * 10  # pixels, xy, samples, W in; L, grads out


def dbg(msg):
    if os.environ.get("BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', 0)} +{time.perf_counter():.1f}s] {msg}", file=sys.stderr, flush=True)


def dist_env():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


class ClockSampler:
    """SM clock and throttle reasons of one GPU, polled through NVML (nvidia_ml_py) every ~2 ms on a
    background thread while the timed region runs (nvidia-smi -lms is too coarse for a 10 ms region)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            names = {
                "hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap,
            }

            def poll():
                while not self._stop.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        self.reasons.update(n for n, bit in names.items() if mask & bit)
                    except Exception:
                        pass
                    time.sleep(0.002)

            self._thread = threading.Thread(target=poll, daemon=True)
            self._thread.start()
        except Exception as err:  # NVML missing: report it, do not fail the bench
            self.error = repr(err)
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measure_peaks(lib, _lib, device):
    """FP32 FMA and MUFU issue peaks of this GPU (ops/s), best of 5 launches."""
    import ctypes

    scratch = torch.zeros(16, device=device)
    out = {}
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    for name, fn in (("fma", lib.tq_peak_fma), ("mufu", lib.tq_peak_mufu)):
        ops = ctypes.c_double()
        best = 0.0
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(fn(sms * 8, 4096, _lib.ptr(scratch), ctypes.byref(ops), _lib.stream_ptr(device)))
            e1.record()
            torch.cuda.synchronize(device)
            if it:
                best = max(best, ops.value / (e0.elapsed_time(e1) * 1e-3))
        out[name] = best
    return out


def make_shard(workload, rank, device, offset_hist=0):
    from tapqir_b200.utils.simulate import simulate

    n_aoi, n_frames, nb, fb, desc = WORKLOADS[workload]
    kinetic = {"kon": 0.2, "koff": 0.2} if WORKLOAD_MODEL.get(workload) == "cosmos+hmm" else None   # test_tapqir.py:31-33
    kw = {}
    if offset_hist:     # SURVEY 8d "secondary realism run": a non-degenerate offset histogram of `offset_hist` distinct bins
        s = torch.arange(90 - offset_hist // 2, 90 - offset_hist // 2 + offset_hist, dtype=torch.float64)
        w = torch.exp(-0.5 * ((s - 90) / max(offset_hist / 8.0, 1.0)) ** 2) + 1e-4
        kw = dict(offset_samples=s, offset_weights=w / w.sum())
        desc += f"; offsets: Gaussian-shaped histogram of {offset_hist} distinct integer bins around 90 instead of O=3"
    ds = simulate(n_aoi, n_frames, C=WORKLOAD_CHANNELS.get(workload, 1), P=14, seed=rank, device=device, aoi_chunk=50,
                  params=kinetic, **kw)
    return ds, nb, fb, desc


def run_native(args):
    from oracle import cosmos_oracle as O
    from tapqir_b200 import _lib
    from tapqir_b200.models import models as model_registry

    rank, world, local = dist_env()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    lib = _lib.load()  # raises if the sm_100a library is missing: no fallback

    ds, nb, fb, desc = make_shard(args.workload, rank, device, args.offset_hist)
    model = model_registry[WORKLOAD_MODEL.get(args.workload, "cosmos")](device=str(device), dtype="float")
    model.data = ds
    model.merge_offsets = not args.keep_offset_bins
    model.init(lr=0.005, nbatch_size=nb, fbatch_size=fb, rank=rank, world_size=world, presharded=True)
    eng = model.engine
    units_per_step = eng.nb * eng.fb  # AOI-frames per rank per step
    patches_per_step = units_per_step * eng.C   # units of the kernels: one per (AOI, frame, channel)
    launches_per_step = model.launches_per_step

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    dbg('model ready')
    for _ in range(args.train_iters):
        model.step()
    for _ in range(max(args.warmup, 3)):
        model.step()
    barrier()
    dbg('warm-up done')

    # ---- device-resident timing ("value") ----------------------------------------------------------
    evs = []
    with ClockSampler(local) as clocks:
        barrier()
        t_wall = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)  # evict the 126 MB L2 (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.step()
            e1.record()
            evs.append((e0, e1))
        barrier()
        t_wall = time.perf_counter() - t_wall
    dbg('timed region done')
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(total_ms, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = total_ms.item() / args.steps
    value = units_per_step * world / (ms_per_step * 1e-3)

    # ---- dominant kernel alone (roofline) -------------------------------------------------------------
    k_ms = []
    for _ in range(max(3, min(args.steps, 10))):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        model.step(time_likelihood=(e0, e1))
        torch.cuda.synchronize(device)
        k_ms.append(e0.elapsed_time(e1))
    k_ms_avg = sum(k_ms) / len(k_ms)
    dbg('kernel timing done')
    peaks = measure_peaks(lib, _lib, device)
    measured = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm_peak = measured.get("hbm_gbs", 6650.0)
    hbm_src = "measured" if "hbm_gbs" in measured else "fallback"
    o_exec = int(eng.store.offset_samples.numel())   # distinct offset bins the kernels loop over
    MUFU_PER_UNIT, FP32_OPS_PER_UNIT = algorithmic_work(o_exec)
    FLOP_PER_UNIT = 2 * FP32_OPS_PER_UNIT                # FMA = 2 flops, like the measured peak
    mufu_ach = MUFU_PER_UNIT * patches_per_step / (k_ms_avg * 1e-3)
    flop_ach = FLOP_PER_UNIT * patches_per_step / (k_ms_avg * 1e-3)
    hbm_ach = KSMOGN_HBM_BYTES_PER_UNIT * patches_per_step / (k_ms_avg * 1e-3) / 1e9
    t_mufu, t_fp32 = MUFU_PER_UNIT / peaks["mufu"], FLOP_PER_UNIT / (2 * peaks["fma"])
    t_hbm = KSMOGN_HBM_BYTES_PER_UNIT / (hbm_peak * 1e9)
    bound = max((t_mufu, "mufu"), (t_fp32, "fp32"), (t_hbm, "hbm"))
    roof_units_per_s = 1.0 / bound[0]
    traffic = None
    tj = ROOT / "profiles" / "traffic.json"
    if tj.exists():
        t = json.loads(tj.read_text()).get({1: "ksmogn_stream_kernel", 3: "ksmogn_stream_kernel_o3"}.get(o_exec, "none"), {})
        if t.get("workload") == args.workload and t.get("units_per_launch") == patches_per_step:
            traffic = {"dram_bytes_per_launch": t["dram_bytes_read"] + t["dram_bytes_write"],
                       "algorithmic_bytes_per_launch": KSMOGN_HBM_BYTES_PER_UNIT * patches_per_step, "source": t["source"]}
    roofline = {
        "kernel": f"ksmogn_stream_kernel<uint16,{o_exec if o_exec <= 4 else 0},true,true> (fused render + offset-marginalised "
                  f"likelihood fwd+bwd" + ("; O > 4 runs the two-pass form, the work counts are the cached-offset form's" if o_exec > 4 else "") + ")",
        "bound": bound[1],
        "achieved": (mufu_ach / 1e12) if bound[1] == "mufu" else (flop_ach / 1e12 if bound[1] == "fp32" else hbm_ach),
        "peak": (peaks["mufu"] / 1e12) if bound[1] == "mufu" else (2 * peaks["fma"] / 1e12 if bound[1] == "fp32" else hbm_peak),
        "unit": "Top/s (MUFU)" if bound[1] == "mufu" else ("TFLOP/s" if bound[1] == "fp32" else "GB/s"),
        "frac": (patches_per_step / (k_ms_avg * 1e-3)) / roof_units_per_s,
        "traffic": traffic,
        "kernel_ms": k_ms_avg,
        "kernel_share_of_step": k_ms_avg / ms_per_step,
        "algorithmic_per_unit": {"mufu_ops": MUFU_PER_UNIT, "fp32_flop": FLOP_PER_UNIT, "hbm_bytes": KSMOGN_HBM_BYTES_PER_UNIT,
                                 "offset_bins_executed": o_exec,
                                 "survey_8d_estimate": {"mufu_ops": 980 * o_exec + 3276, "fp32_flop": 8232 * o_exec + 63220}},
        "peaks_measured_here": {"mufu_Tops": peaks["mufu"] / 1e12, "fp32_TFLOPs": 2 * peaks["fma"] / 1e12,
                                "hbm_GBs": hbm_peak, "hbm_source": hbm_src + " (MEASURED_PEAKS.json)"},
        "hbm_view": {"achieved_GBs": hbm_ach, "peak_GBs": hbm_peak, "frac": hbm_ach / hbm_peak},
        "step_roofline_frac": value * eng.C / world / roof_units_per_s,
    }

    # ---- end to end through the public API with host buffers --------------------------------------------
    host_pix = eng.store.pixels.cpu().pin_memory()
    host_xy = eng.store.xy.cpu().pin_memory()
    loss_host = torch.zeros(1, dtype=torch.float64).pin_memory()
    dbg('e2e buffers ready')
    for _ in range(2):
        model.step_from_host(host_pix, host_xy, loss_host, prefetch_next=(host_pix, host_xy))
    barrier()
    dbg('e2e warm-up done')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        # this step's inputs come from pinned host memory (uploaded during the previous step's compute)
        model.step_from_host(host_pix, host_xy, loss_host, prefetch_next=(host_pix, host_xy))
    e1.record()
    barrier()
    dbg('e2e done')
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(e2e_ms, op=torch.distributed.ReduceOp.MAX)
    e2e_value = units_per_step * world * args.steps / (e2e_ms.item() * 1e-3)
    e2e = {"value": e2e_value, "unit": "AOI-frames/s",
           "h2d_bytes_per_step": host_pix.numel() * host_pix.element_size() + host_xy.numel() * host_xy.element_size(),
           "d2h_bytes_per_step": 8}

    # ---- the same measurement after training ----------------------------------------------------------------
    # The timed steps above start from the initial variational parameters (the contract: W warm-up steps, then K timed
    # ones).  As SVI proceeds the guides of absent spots relax to small concentrations, a few percent of the guide sites
    # leave the fp32 forms and neighbouring units take different branches; the step slows down.  Reported beside the
    # headline so that it is not mistaken for the steady state of a long fit.
    trained = None
    if args.trained_iters > 0 and args.train_iters == 0:
        for _ in range(args.trained_iters):
            model.step()
        barrier()
        tev = []
        for _ in range(args.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.step()
            e1.record()
            tev.append((e0, e1))
        barrier()
        t_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in tev)], dtype=torch.float64, device=device)
        if world > 1:
            torch.distributed.all_reduce(t_ms, op=torch.distributed.ReduceOp.MAX)
        trained = {"extra_svi_iterations": args.trained_iters,
                   "ms_per_step": t_ms.item() / args.steps, "value": units_per_step * world / (t_ms.item() / args.steps * 1e-3),
                   "unit": "AOI-frames/s"}
    dbg('trained-state timing done')

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ---------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_sample(args, budget_s=20.0)

    if rank == 0:
        line = {
            "metric": "cosmos SVI AOI-frames/sec (ELBO fwd+bwd+Adam)", "value": value, "unit": "AOI-frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "model": model.name, "channels": eng.C, "nb_per_gpu": eng.nb, "fb": eng.fb, "offset_bins": args.offset_hist or O_BINS,
                       "offset_bins_distinct": o_exec, "train_iters_before_timing": args.train_iters,
                       "parallelism": f"aoi-shard x{world}", "l2": "flushed (256 MiB write) before every timed step",
                       "local_terms_dtype": "f32 (double fallback outside the fp32 regimes)", "likelihood_dtype": "f32"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "trained_state": trained,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks.summary(),
            "wall_s_timed_region": t_wall, "final_loss": float(eng.loss.item()),
        }
        print(json.dumps(line))
    if world > 1:
        # Tear-down: the captured CUDA graphs hold NCCL work; drop them, drain the device, and leave
        # without destroy_process_group() (it can block behind graph-captured collectives).
        barrier()
        model.engine.release_graph()
        torch.cuda.synchronize(device)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_cpu_sample(args, budget_s=20.0):
    """The oracle (CPU, fp64, all host threads) on the reference-default 10 x 512 minibatch of the
    same kind of data; returns the cpu_baseline object."""
    from oracle import cosmos_oracle as O
    from tapqir_b200.utils.simulate import simulate

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_aoi, n_frames = 20, 600
    nb, fb = 10, 512
    ds = simulate(n_aoi, n_frames, seed=0)
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    svi = O.OracleSVI(data, nbatch_size=nb, fbatch_size=fb, seed=0)
    svi.step()
    times, t_start = [], time.perf_counter()
    while len(times) < 3 or (time.perf_counter() - t_start < budget_s and len(times) < 40):
        t0 = time.perf_counter()
        svi.step()
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {"value": nb * fb / med, "unit": "AOI-frames/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} steps of the fp64 torch oracle, minibatch {nb} AOIs x {fb} frames drawn from a "
                      f"simulated {n_aoi} x {n_frames} dataset (O=3), median step {med * 1e3:.0f} ms"}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation cannot be imported (pyro-ppl, funsor,
    pykeops are not installable here), so this times its restatement in oracle/ on the host cores."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    n_aoi, n_frames, _, _, desc = WORKLOADS[args.workload]
    from oracle import cosmos_oracle as O
    from tapqir_b200.utils.simulate import simulate

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    nb, fb = 10, 512
    ds = simulate(20, 600, seed=0)
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    svi = O.OracleSVI(data, nbatch_size=nb, fbatch_size=fb, seed=0)
    for _ in range(max(args.warmup, 1)):
        svi.step()
    steps = min(args.steps, 30)
    t0 = time.perf_counter()
    for _ in range(steps):
        svi.step()
    dt = (time.perf_counter() - t0) / steps
    value = nb * fb / dt
    sample = (f"fp64 torch oracle (port of the reference step, Pyro not installable), {steps} steps of the "
              f"reference-default minibatch {nb} AOIs x {fb} frames from a simulated 20 x 600 dataset, O=3")
    print(json.dumps({
        "impl": "reference", "metric": "cosmos SVI AOI-frames/sec (ELBO fwd+bwd+Adam)", "value": value,
        "unit": "AOI-frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": max(args.warmup, 1),
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "AOI-frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "AOI-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--train-iters", type=int, default=0,
                    help="untimed SVI iterations before the warm-up: times the TRAINED state (absent spots' guides relax to "
                         "small concentrations, a few percent of the sites leave the fp32 forms) instead of the initial point")
    ap.add_argument("--trained-iters", type=int, default=2000,
                    help="after the timed steps, run this many more SVI iterations and time K steps again (reported as "
                         "`trained_state`); 0 to skip")
    ap.add_argument("--offset-hist", type=int, default=0, metavar="O",
                    help="simulate with a histogram of O distinct offset bins (SURVEY 8d's secondary realism run, e.g. 64) "
                         "instead of the reference simulator's three identical bins; not the headline configuration")
    ap.add_argument("--keep-offset-bins", action="store_true",
                    help="do not merge the simulator's three identical offset bins (exercises the O = 3 kernels)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
