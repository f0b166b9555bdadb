"""CPU: dataset container helpers that do not need a device."""

import torch

from tapqir_b200.utils.dataset import OffsetData, merge_offset_support


def test_merge_offset_support_is_an_identity_for_the_marginal():
    s = torch.tensor([90.0, 88.0, 90.0, 91.0, 88.0, 90.0], dtype=torch.float64)
    w = torch.tensor([0.1, 0.2, 0.25, 0.15, 0.05, 0.25], dtype=torch.float64)
    logits = OffsetData(s, w).logits
    ms, ml = merge_offset_support(s, logits)
    assert ms.tolist() == [88.0, 90.0, 91.0]
    torch.testing.assert_close(ml.exp(), torch.tensor([0.25, 0.6, 0.15], dtype=torch.float64))
    # sum_j w_j f(D - delta_j) for an arbitrary f
    D = torch.linspace(95, 300, 7, dtype=torch.float64)[:, None]
    f = lambda y: torch.lgamma(y / 7.0) - 0.3 * y
    full = torch.logsumexp(logits + f(D - s), -1)
    merged = torch.logsumexp(ml + f(D - ms), -1)
    torch.testing.assert_close(full, merged, rtol=1e-13, atol=1e-13)


def test_merge_offset_support_leaves_distinct_bins_alone():
    s = torch.arange(80.0, 96.0)
    w = torch.softmax(-0.5 * ((s - 90) / 3) ** 2, 0)
    logits = OffsetData(s, w).logits
    ms, ml = merge_offset_support(s, logits)
    assert ms is s and ml is logits


# ---- the data.tpqr contract, against a file written by the reference's own utils/dataset.py ---------------------------
def test_reads_a_data_file_written_by_the_reference_and_reports_what_the_reference_reports(tmp_path):
    """tests/golden/ref_data/data.tpqr was written by the reference's ``save`` (dataset.py:195-213) and facts.pt holds what
    its ``CosmosDataset`` / ``OffsetData`` / ``fetch`` report for it (tests/golden/make_golden_step.py::run_data_case)."""
    from pathlib import Path

    import numpy as np

    from tapqir_b200.utils.dataset import load, save

    ref_dir = Path(__file__).resolve().parent / "golden" / "ref_data"
    facts = torch.load(ref_dir / "facts.pt", weights_only=False)
    ds = load(ref_dir)
    assert (ds.N, ds.Nc, ds.Nt, ds.F, ds.C, ds.P) == tuple(facts[k] for k in ("N", "Nc", "Nt", "F", "C", "P"))
    assert ds.channels == facts["channels"] and ds.name == facts["name"]
    assert torch.equal(ds.median.double(), facts["median"]) and torch.equal(ds.x, facts["x"]) and torch.equal(ds.y, facts["y"])
    off = ds.offset
    assert (off.min, off.max) == (facts["offset_min"], facts["offset_max"])
    assert abs(off.mean - facts["offset_mean"]) <= 1e-12 * facts["offset_mean"] and abs(off.var - facts["offset_var"]) <= 1e-9
    torch.testing.assert_close(off.logits, facts["offset_logits"], rtol=1e-14, atol=0)
    obs, target, ont = ds.fetch(facts["fetch_ndx"], facts["fetch_fdx"], torch.arange(ds.C))       # dataset.py:140-151
    assert torch.equal(obs, facts["fetch_obs"]) and torch.equal(target, facts["fetch_target"]) and torch.equal(ont, facts["fetch_ontarget"])
    assert ds.labels.dtype.names == ("aoi", "frame", "z") and ds.labels.shape == (ds.N, ds.F, ds.C)
    assert not bool(ds.mask[2]) and ds.mask.sum().item() == 3
    # and back: our save() writes the same key set with identical contents
    save(ds, tmp_path)
    ours = torch.load(tmp_path / "data.tpqr", weights_only=False)
    theirs = torch.load(ref_dir / "data.tpqr", weights_only=False)
    assert list(ours.keys()) == list(theirs.keys())
    for k, v in theirs.items():
        if isinstance(v, torch.Tensor):
            assert torch.equal(ours[k], v) and ours[k].dtype == v.dtype, k
        elif isinstance(v, np.ndarray):
            assert (ours[k] == v).all() and ours[k].dtype == v.dtype, k
        else:
            assert ours[k] == v, k


def test_missing_data_file_raises_the_reference_exception(tmp_path):
    from tapqir_b200.exceptions import TapqirFileNotFoundError
    from tapqir_b200.utils.dataset import load

    try:
        load(tmp_path)
    except TapqirFileNotFoundError as err:                       # exceptions.py:19-30: (name, path) and the message
        assert err.name == "data" and str(err.path).endswith("data.tpqr") and "Unable to find data file" in str(err)
    else:
        raise AssertionError("no exception")
