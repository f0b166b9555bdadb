"""
CPU: the glimpse-ingestion oracle and the host-side helpers of tapqir_b200.imscroll against the golden
vectors made from the reference's own ``bin_hist`` (tests/golden/make_golden_glimpse.py), and against each
other for the offset post-processing.
"""

from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import glimpse_oracle as GO
from tapqir_b200.imscroll import glimpse_reader as GR

GOLDEN = Path(__file__).parent / "golden" / "ref_glimpse.pt"


@pytest.fixture(scope="module")
def cases():
    return torch.load(GOLDEN, weights_only=False)


@pytest.mark.parametrize("impl", [GO.bin_hist, GR.bin_hist], ids=["oracle", "product"])
def test_bin_hist_matches_reference_bit_for_bit(cases, impl):
    old = torch.get_default_dtype()
    try:
        for case in cases:
            torch.set_default_dtype(torch.float64 if "64" in case["default"] else torch.float32)
            out_s, out_w = impl(case["samples"], case["weights"], case["s"])
            assert out_s.dtype == case["out_samples"].dtype and out_w.dtype == case["out_weights"].dtype
            assert torch.equal(out_s, case["out_samples"]), (len(case["samples"]), case["s"])
            assert torch.equal(out_w, case["out_weights"]), (len(case["samples"]), case["s"])
    finally:
        torch.set_default_dtype(old)


def test_offset_distribution_matches_oracle():
    rng = np.random.default_rng(3)
    for min_data, bin_size in [(40, 3), (95, 1), (200, 5)]:
        counts = np.zeros(65536, dtype=np.int64)
        vals = np.clip(rng.normal(90, 6, 20000).round().astype(int), 60, 140)
        np.add.at(counts, vals, 1)
        counts[300] = 3   # far tail, removed with the top 0.5 %
        offsets = {int(v): int(c) for v, c in enumerate(counts) if c}
        ref_s, ref_w = GO.offset_distribution(offsets, min_data, bin_size)
        out_s, out_w = GR.offset_distribution(counts, min_data, bin_size)
        assert torch.equal(out_s, ref_s) and torch.equal(out_w, ref_w)
        assert abs(out_w.sum().item() - 1.0) < 1e-6
        if min_data <= vals.min():
            assert out_s[0].item() == min_data - 1


def test_decode_frame_is_big_endian_plus_offset():
    raw = np.array([[-32768, -1], [0, 32767]], dtype=">i2").tobytes()
    assert GO.decode_frame(raw, 2, 2).tolist() == [[0, 32767], [32768, 65535]]


def test_crop_loop_rounds_half_to_even():
    frames = np.arange(2 * 20 * 20).reshape(2, 20, 20)
    # raw - (P-1)/2 = 4.5 and 5.5 -> shifts 4 and 6 (ties to even), like Python's round()
    data, xy = GO.crop_loop(frames, np.array([[7.0, 8.0]]), np.zeros((2, 2)), P=6)
    assert data[0, 0, 0, 0] == frames[0, 6, 4] and xy[0, 0].tolist() == [3.0, 2.0]


# ---- header / AOI table / drift / labels parsing against the reference's own GlimpseDataset constructor ------------------
import pytest


@pytest.mark.parametrize("case", ["mat_full", "text_range", "mat_nolabels"])
def test_glimpse_dataset_parsing_matches_reference_constructor(case):
    """tests/golden/ref_glimpse_folder/ holds a small synthetic glimpse folder (header.mat, driftlist.mat, AOI tables in the
    three accepted layouts, spot-picker intervals) and facts.pt what the reference's ``GlimpseDataset.__init__``
    (glimpse_reader.py:55-159, run verbatim by tests/golden/make_golden_step.py) makes of it."""
    from pathlib import Path

    import numpy as np
    import torch

    from tapqir_b200.imscroll.glimpse_reader import GlimpseDataset

    folder = Path(__file__).resolve().parent / "golden" / "ref_glimpse_folder"
    ref = torch.load(folder / "facts.pt", weights_only=False)[case]
    kw = dict(ref["kwargs"])
    kw["glimpse-folder"] = str(folder / "glimpse")
    for k in ("driftlist", "ontarget-aoiinfo", "offtarget-aoiinfo", "ontarget-labels"):
        if kw[k] is not None:
            kw[k] = str(folder / kw[k])
    g = GlimpseDataset(**kw)
    assert (g.height, g.width, list(g.dtypes), (g.offset_x, g.offset_y), g.name) == \
        (ref["height"], ref["width"], ref["dtypes"], ref["offset"], ref["name"])
    assert float(g.header["time1"]) == ref["time1"] and np.array_equal(np.asarray(g.header["filenumber"]), ref["filenumber"])
    for d in ref["dtypes"]:
        assert np.array_equal(g.aoiinfo[d].index.values, ref["aoiinfo"][d]["index"])
        assert np.array_equal(g.aoiinfo[d][["frame", "ave", "y", "x", "pixnum"]].values, ref["aoiinfo"][d]["values"])
        if ref["labels"][d] is None:
            assert g.labels[d] is None
        else:
            assert g.labels[d].dtype == ref["labels"][d].dtype and np.array_equal(g.labels[d], ref["labels"][d])
    assert np.array_equal(g.cumdrift.index.values, ref["cumdrift"]["index"])
    assert np.array_equal(g.cumdrift[["dy", "dx", "ttb"]].values, ref["cumdrift"]["values"])       # bit for bit
    assert g.N == 4 and g.Nc == 3 and g.F == len(ref["cumdrift"]["index"])


def test_read_glimpse_of_the_reference_pins_the_oracle_and_the_host_side_histogram():
    """tests/golden/ref_glimpse_folder/out/data.tpqr was written by the reference's whole ``read_glimpse``
    (glimpse_reader.py:304-472: frame decoding, drift-corrected AOI cropping, offset histogram, 0.5 % tail fold, ``bin_hist``)
    run verbatim on the synthetic movie of that folder under the float32 default ``tapqir glimpse`` runs with.  Bit for
    bit: the oracle's frame loop and offset post-processing (what the CUDA path is tested against on the GPU,
    tests/test_glimpse_gpu.py), and the host-side ``offset_distribution`` of the product fed the same counts."""
    from pathlib import Path

    import numpy as np
    import torch

    from tapqir_b200.imscroll.glimpse_reader import offset_distribution
    from tapqir_b200.utils.dataset import load

    folder = Path(__file__).resolve().parent / "golden" / "ref_glimpse_folder"
    facts = torch.load(folder / "facts.pt", weights_only=False)
    rg, hdr = facts["read_glimpse"], facts["mat_nolabels"]                  # mat_nolabels: the same tables and frame range
    ref = load(folder / "out")
    P, f1, f2 = rg["P"], rg["frame_start"], rg["frame_end"]
    frames = rg["decoded"].numpy().astype(np.int64)[f1 - 1:f2]              # frame numbers are 1-based
    assert list(hdr["cumdrift"]["index"]) == list(range(f1, f2 + 1))
    xy = np.concatenate([hdr["aoiinfo"][d]["values"][:, [3, 2]] for d in ("ontarget", "offtarget")], 0)     # columns x, y
    cum = hdr["cumdrift"]["values"][:, [1, 0]]                              # (dy, dx, ttb) -> (dx, dy)
    data, target = GO.crop_loop(frames, xy, cum, P)
    assert ref.images.dtype == torch.int64 and np.array_equal(ref.images[:, :, 0].numpy(), data)
    assert np.array_equal(ref.xy[:, :, 0].numpy(), target)
    assert ref.is_ontarget.tolist() == [True] * 4 + [False] * 3 and ref.name == "golden-movie" and ref.channels == ("green",)
    assert np.array_equal(ref.ttb[:, 0].numpy(), hdr["cumdrift"]["values"][:, 2]) and abs(ref.time1.item() - hdr["time1"]) < 1e-3
    labels = facts["mat_full"]["labels"]["ontarget"][:, f1 - 1:f2]
    assert ref.labels.shape == (4, f2 - f1 + 1, 1) and np.array_equal(ref.labels[..., 0], labels)
    # offsets: pooled counts of the dark corner, guard bin, tail fold, thinning -- in float32 like the reference's run
    ox, oy = hdr["offset"]
    counts = GO.offset_counts(frames, ox, oy, rg["offset_P"])
    assert torch.get_default_dtype() == torch.float32
    s, w = GO.offset_distribution(counts, int(data.min()), rg["bin_size"])
    assert torch.equal(s, ref.offset.samples) and torch.equal(w, ref.offset.weights) and w.dtype == torch.float32
    dense = np.zeros(65536, dtype=np.int64)                                 # the product's form of the same counts
    for value, count in counts.items():
        dense[value] = count
    ps, pw = offset_distribution(dense, int(data.min()), rg["bin_size"])
    assert torch.equal(ps, ref.offset.samples) and torch.equal(pw, ref.offset.weights)
