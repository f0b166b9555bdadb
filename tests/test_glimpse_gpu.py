"""
GPU: glimpse ingestion kernels (csrc/glimpse.cu through the C ABI) against oracle/glimpse_oracle.py --
bit-exact, as index / integer work must be -- and ``read_glimpse`` end to end on a synthetic movie written
in the glimpse layout (header.mat, N.glimpse files of big-endian int16 frames, driftlist.mat, aoiinfo text).
"""

import numpy as np
import pytest
import torch

from oracle import glimpse_oracle as GO

pytestmark = pytest.mark.gpu


def synthetic_frames(F, H, W, seed):
    rng = np.random.default_rng(seed)
    decoded = rng.integers(60, 4000, size=(F, H, W)).astype(np.int64)
    decoded[:, :12, :12] = rng.integers(85, 97, size=(F, 12, 12))          # a dark corner for the offset region
    raw = (decoded - 2**15).astype(">i2")                                  # what the file holds
    return decoded, raw


def run_kernels(raw, aoi_xy, drift, P, chunks, off):
    from tapqir_b200 import _lib

    lib, p = _lib.load(), _lib.ptr
    F, H, W = raw.shape
    N = len(aoi_xy)
    dev = torch.device("cuda")
    as_u16 = torch.from_numpy(np.ascontiguousarray(raw).view(np.uint16).copy()).to(dev)   # bytes untouched
    xy_d, drift_d = torch.from_numpy(aoi_xy.copy()).to(dev), torch.from_numpy(drift.copy()).to(dev)
    patches = torch.zeros(N, F, P, P, dtype=torch.uint16, device=dev)
    target = torch.zeros(N, F, 2, dtype=torch.float64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    counts = torch.zeros(65536, dtype=torch.int64, device=dev)
    st = _lib.stream_ptr(dev)
    f0 = 0
    for fc in chunks:
        chunk = as_u16[f0:f0 + fc].contiguous()
        _lib.check(lib.tq_crop_aois(p(chunk), H, W, fc, f0, p(xy_d), p(drift_d), N, F, P, p(patches), p(target), p(status), st))
        _lib.check(lib.tq_offset_hist(p(chunk), H, W, fc, off[0], off[1], off[2], p(counts), st))
        f0 += fc
    torch.cuda.synchronize()
    return patches.cpu(), target.cpu(), status.item(), counts.cpu().numpy()


@pytest.mark.parametrize("P,chunks", [(6, [5]), (14, [2, 2, 1]), (7, [1, 4])])
def test_crop_and_histogram_are_bit_exact(P, chunks):
    F, H, W = 5, 40, 48
    decoded, raw = synthetic_frames(F, H, W, seed=P)
    rng = np.random.default_rng(P + 100)
    N = 9
    aoi_xy = rng.uniform(P, min(H, W) - P - 2, size=(N, 2))
    # windows whose corner lands exactly on a half: round-half-even must agree with Python's round()
    aoi_xy[0] = [10.0 + 0.5 * (P - 1) + 0.5, 12.0 + 0.5 * (P - 1) + 1.5]
    aoi_xy[1] = [15.5 + 0.5 * (P - 1), 16.5 + 0.5 * (P - 1)]
    drift = np.cumsum(rng.normal(0, 0.3, size=(F, 2)), 0)
    drift[0] = 0.0
    off = (1, 2, 9)
    patches, target, status, counts = run_kernels(raw, aoi_xy, drift, P, chunks, off)
    ref_data, ref_xy = GO.crop_loop(decoded, aoi_xy, drift, P)
    assert status == 0
    assert np.array_equal(patches.numpy().astype(np.int64), ref_data)
    assert np.array_equal(target.numpy(), ref_xy)                       # same double arithmetic: identical bits
    ref_counts = GO.offset_counts(decoded, *off)
    assert {int(v): int(c) for v, c in enumerate(counts) if c} == dict(ref_counts)


def test_window_outside_the_frame_is_flagged():
    decoded, raw = synthetic_frames(2, 30, 30, seed=1)
    aoi_xy = np.array([[15.0, 15.0], [1.0, 15.0]])
    patches, target, status, _ = run_kernels(raw, aoi_xy, np.zeros((2, 2)), 8, [2], (0, 0, 4))
    assert status == 1
    assert np.array_equal(patches[0].numpy().astype(np.int64), GO.crop_loop(decoded, aoi_xy[:1], np.zeros((2, 2)), 8)[0][0])
    assert patches[1].abs().sum() == 0 if patches.dtype.is_signed else int(patches[1].to(torch.int64).sum()) == 0


def write_movie(folder, decoded, raw, aoi_on, aoi_off, drift_steps, frames_per_file=3):
    """header.mat + N.glimpse + driftlist.mat + aoiinfo text files in the layout glimpse_reader.py parses."""
    from scipy.io import savemat

    F, H, W = raw.shape
    gdir = folder / "glimpse"
    gdir.mkdir()
    filenumber, offset = [], []
    for f in range(F):
        number, pos = f // frames_per_file, f % frames_per_file
        filenumber.append(number)
        offset.append(pos * H * W * 2)
        with open(gdir / f"{number}.glimpse", "ab") as fid:
            fid.write(raw[f].tobytes())
    vid = dict(height=float(H), width=float(W), filenumber=np.array(filenumber, dtype=np.int32), offset=np.array(offset, dtype=np.int64),
               ttb=np.arange(F, dtype=float) * 50.0, time1=1234.0)
    savemat(gdir / "header.mat", {"vid": vid})
    # driftlist: frame, dy, dx (per-frame increments; frame numbers start at 1)
    dl = np.zeros((F, 3))
    dl[:, 0] = np.arange(1, F + 1)
    dl[:, 1:] = drift_steps
    savemat(folder / "driftlist.mat", {"driftlist": dl})

    def aoi_file(name, xy):
        # frame, ave, y, x, pixnum, aoi -- MATLAB indexing (+1)
        rows = np.stack([np.full(len(xy), 1.0), np.full(len(xy), 10.0), xy[:, 1] + 1, xy[:, 0] + 1, np.full(len(xy), 5.0),
                         np.arange(1, len(xy) + 1)], 1)
        np.savetxt(folder / name, rows)
        return folder / name

    return dict(gdir=gdir, driftlist=folder / "driftlist.mat", on=aoi_file("ontarget.dat", aoi_on), off=aoi_file("offtarget.dat", aoi_off))


def test_read_glimpse_end_to_end(tmp_path):
    from tapqir_b200.imscroll import read_glimpse
    from tapqir_b200.utils.dataset import load

    F, H, W, P = 8, 50, 60, 14
    decoded, raw = synthetic_frames(F, H, W, seed=9)
    rng = np.random.default_rng(9)
    aoi_on = rng.uniform(P, 30, size=(4, 2))
    aoi_off = rng.uniform(P, 30, size=(3, 2))
    steps = rng.normal(0, 0.2, size=(F, 2))          # dy, dx increments
    steps[0] = 0.0                                   # the frame the AOIs were picked in
    files = write_movie(tmp_path, decoded, raw, aoi_on, aoi_off, steps)
    kwargs = {"P": P, "num-channels": 1, "dataset": "synthetic", "offset-P": 10, "bin-size": 3, "offset-x": 0, "offset-y": 0,
              "use-offtarget": True, "frame-range": False, "frame-start": None, "frame-end": None, "labels": False,
              "channels": [{"name": "green", "glimpse-folder": str(files["gdir"]), "driftlist": str(files["driftlist"]),
                            "ontarget-aoiinfo": str(files["on"]), "offtarget-aoiinfo": str(files["off"]),
                            "ontarget-labels": None, "offtarget-labels": None}]}
    ds = read_glimpse(tmp_path, None, **kwargs)

    # the same movie through the oracle: cumulative drift relative to frame 1 (where the AOIs were picked)
    cum = np.zeros((F, 2))
    cum[1:] = np.cumsum(steps[1:], 0)
    cum = cum[:, ::-1]                                # (dy, dx) -> (dx, dy)
    xy = np.concatenate([aoi_on, aoi_off], 0)
    ref_data, ref_xy = GO.crop_loop(decoded, xy, cum, P)
    assert np.array_equal(ds.images[:, :, 0].numpy(), ref_data)
    np.testing.assert_allclose(ds.xy[:, :, 0].numpy(), ref_xy, rtol=0, atol=1e-12)   # text aoiinfo round trip
    assert ds.is_ontarget.tolist() == [True] * 4 + [False] * 3
    ref_s, ref_w = GO.offset_distribution(GO.offset_counts(decoded, 0, 0, 10), int(ref_data.min()), 3)
    assert torch.equal(ds.offset.samples.cpu(), ref_s) and torch.equal(ds.offset.weights.cpu(), ref_w)
    back = load(tmp_path)
    assert torch.equal(back.images, ds.images) and back.name == "synthetic"
