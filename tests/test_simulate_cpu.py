"""
CPU: the simulator that feeds bench.py and the parity tests (tapqir_b200/utils/simulate.py) against summary
statistics of a movie produced by the reference's OWN ``utils/simulate.py::simulate`` (Predictive over the unconditioned
cosmos model, pixels from ``KSMOGN.rsample``), run verbatim by tests/golden/make_golden_step.py::run_simulate_case with the
constants of the reference's test-suite (test/test_tapqir.py:22-50).  Random streams differ, so the comparison is in
distribution: tolerances are a few standard errors of each statistic at this size (40 AOIs x 100 frames).
"""

from pathlib import Path

import torch

from tapqir_b200.utils.simulate import simulate


def _stats(d, prm, P):
    img = d.images.double()
    corners = torch.stack([img[..., 0, 0], img[..., 0, P - 1], img[..., P - 1, 0], img[..., P - 1, P - 1]], -1)
    patch_sum = img.sum((-1, -2)) - (prm["background"] + prm["offset"] - 0.5) * P * P
    return img, corners, patch_sum


def test_simulated_movie_matches_the_reference_simulator_in_distribution():
    ref = torch.load(Path(__file__).resolve().parent / "golden" / "ref_simulate_stats.pt", weights_only=False)
    prm, N, F, C, P = ref["params"], ref["N"], ref["F"], ref["C"], ref["P"]
    d = simulate(N, F, C=C, P=P, seed=0)                      # the defaults ARE the reference test-suite's constants
    img, corners, patch_sum = _stats(d, prm, P)
    # structure: exactly what the reference writes
    assert tuple(img.shape) == ref["shape"] and bool((img == img.floor()).all()) == ref["integral"]
    assert torch.equal(d.is_ontarget, ref["is_ontarget"]) and torch.equal(torch.unique(d.xy).double(), ref["xy_unique"])
    assert torch.equal(d.offset.samples.double(), ref["offset_samples"])
    torch.testing.assert_close(d.offset.weights.double(), ref["offset_weights"])
    assert d.labels.shape == ref["labels_shape"] and d.labels.dtype.names == ref["labels_fields"]
    assert d.labels["aoi"][3, 7, 0] == 3 and d.labels["frame"][3, 7, 0] == 7
    # specific binding: Bernoulli(pi) per on-target AOI-frame
    n = d.labels["z"].size
    zf = float(d.labels["z"].mean())
    se = (prm["pi"] * (1 - prm["pi"]) / n) ** 0.5
    assert abs(zf - prm["pi"]) < 4 * se and abs(zf - ref["z_fraction"]) < 4 * 2 ** 0.5 * se
    # noise model far from the target: floor(Gamma(b / g, 1 / g) + offset) plus the tails of non-specific spots
    m, v = corners.mean().item(), corners.var().item()
    se_m = (v / corners.numel()) ** 0.5
    assert abs(m - ref["corner_mean"]) < 4 * 2 ** 0.5 * se_m
    assert m > prm["background"] + prm["offset"] - 0.5 and abs(m - (prm["background"] + prm["offset"] - 0.5)) < 2.0
    assert abs(v / ref["corner_var"] - 1) < 0.08 and v > prm["background"] * prm["gain"]      # Var = b g + 1/12 + spots
    # photons above background per patch: on-target (specific + non-specific spots) and off-target (non-specific only)
    on, off = patch_sum[: N // 2].mean().item(), patch_sum[N // 2:].mean().item()
    assert abs(on / ref["patch_sum_on"] - 1) < 0.15 and abs(off / ref["patch_sum_off"] - 1) < 0.15
    assert on > 1.5 * off
    q = torch.quantile(patch_sum.flatten(), torch.tensor([0.9, 0.99], dtype=torch.float64))
    assert abs(q[0].item() / ref["patch_sum_q"][1].item() - 1) < 0.05 and abs(q[1].item() / ref["patch_sum_q"][2].item() - 1) < 0.10
    assert abs(img.mean().item() - ref["pixel_mean"]) < 0.6 and img.min().item() > prm["offset"]


def test_simulator_is_reproducible_and_seed_dependent():
    a, b, c = simulate(4, 6, seed=3), simulate(4, 6, seed=3), simulate(4, 6, seed=4)
    assert torch.equal(a.images, b.images) and not torch.equal(a.images, c.images)


def test_kinetic_recipe_matches_the_reference_simulator_in_distribution():
    """kon / koff simulation (the data of BASELINE config 5): the reference's hmm model in its sequential form, simulated by
    the reference's own ``simulate`` (test/test_tapqir.py:30-33), against ours: stationary occupancy, transition
    frequencies, photons per on- / off-target patch."""
    ref = torch.load(Path(__file__).resolve().parent / "golden" / "ref_simulate_stats.pt", weights_only=False)["hmm"]
    prm, N, F, P = ref["params"], ref["N"], ref["F"], 14
    d = simulate(N, F, C=1, P=P, seed=1, params={"kon": prm["kon"], "koff": prm["koff"]})
    assert tuple(d.images.shape) == ref["shape"] and d.labels.shape == ref["labels_shape"]
    z = torch.as_tensor(d.labels["z"])[..., 0]
    prev, cur = z[:, :-1], z[:, 1:]
    p01 = ((prev == 0) & (cur == 1)).sum().item() / (prev == 0).sum().item()
    p10 = ((prev == 1) & (cur == 0)).sum().item() / (prev == 1).sum().item()
    se = (0.2 * 0.8 / (prev.numel() / 2)) ** 0.5
    assert abs(p01 - prm["kon"]) < 4 * se and abs(p10 - prm["koff"]) < 4 * se
    assert abs(p01 - ref["p01"]) < 6 * se and abs(p10 - ref["p10"]) < 6 * se
    # occupancy: the chain mixes slowly (correlation time ~ 1 / (kon + koff)), so ~ N/2 * F * (kon + koff) / 2 effective draws
    occ_se = (0.25 / (z.numel() * (prm["kon"] + prm["koff"]) / 2)) ** 0.5
    assert abs(z.double().mean().item() - 0.5) < 4 * occ_se and abs(z.double().mean().item() - ref["z_fraction"]) < 6 * occ_se
    _, _, patch_sum = _stats(d, {"background": 150, "offset": 90.0}, P)
    on, off = patch_sum[: N // 2].mean().item(), patch_sum[N // 2:].mean().item()
    # 5000 patches per class; the off-target sum rides on ~750 non-specific spots of 3000 photons: +-4 % on either side
    assert abs(on / ref["patch_sum_on"] - 1) < 0.10 and abs(off / ref["patch_sum_off"] - 1) < 0.20
