"""
Glimpse movie ingestion with the reference's interface (tapqir/imscroll/glimpse_reader.py):
``read_glimpse(path, progress_bar, **kwargs)`` with the same option names, ``GlimpseDataset`` with the
same attributes (``header, aoiinfo, cumdrift, labels, dtypes, N, Nc, F, height, width``) and
``bin_hist``; the result is the same ``data.tpqr``.

What moved to the GPU: the reference decodes every frame with numpy and then
runs a Python loop over frames x AOIs to cut the PxP windows and a ``np.unique`` per frame for the offset
histogram (glimpse_reader.py:354-381).  Here the raw bytes of a chunk of frames are uploaded as they are
in the file (big-endian int16) and two kernels (csrc/glimpse.cu, ``tq_crop_aois`` / ``tq_offset_hist``) do
the decoding, the window arithmetic (double, round-half-even like Python's ``round``) and the counting.
Index arithmetic is bit-exact with the reference loop (tests/test_glimpse_gpu.py against
oracle/glimpse_oracle.py).  Parsing of header / aoiinfo / driftlist / interval files stays on the host
(scipy.io + pandas, as in the reference); the diagnostic PNG plots (matplotlib) are not produced.
"""

import logging
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

from tapqir_b200.utils.dataset import CosmosDataset, save

logger = logging.getLogger(__name__)

AOI_COLUMNS = ["frame", "ave", "y", "x", "pixnum", "aoi"]


def bin_hist(samples, weights, s):
    """
    Thin an offset histogram: the first bin is kept, the remaining ones are merged in groups of ``s`` (sample =
    middle bin of the group, weight = group total), a shorter last group likewise (glimpse_reader.py:22-37).
    Accumulates in torch's default dtype, column by column, like the reference.
    """
    rest = len(samples) - 1
    full, tail = divmod(rest, s)
    out_s = torch.zeros(1 + full + (1 if tail else 0), dtype=torch.int)
    out_w = torch.zeros(len(out_s))
    out_s[0], out_w[0] = samples[0], weights[0]
    if full:
        groups_s = samples[1:1 + full * s].reshape(full, s)
        groups_w = weights[1:1 + full * s].reshape(full, s)
        out_s[1:1 + full] = groups_s[:, s // 2]
        for col in range(s):
            out_w[1:1 + full] += groups_w[:, col]
    if tail:
        start = 1 + full * s
        out_s[-1] = samples[start + tail // 2]
        out_w[-1] = weights[start:].sum()
    return out_s, out_w


def _aoi_table(source):
    """aoiinfo2 table from a .mat (two layouts) or a whitespace text file (glimpse_reader.py:76-97)."""
    import pandas as pd
    from scipy.io import loadmat

    try:
        mat = loadmat(source)
    except ValueError:
        table = np.loadtxt(source)
    else:
        table = mat["aoiinfo2"] if "aoiinfo2" in mat else mat["aoifits"]["aoiinfo2"][0, 0]
    df = pd.DataFrame(table, columns=AOI_COLUMNS).astype({"aoi": int}).set_index("aoi")
    df["x"] -= 1  # MATLAB -> python indexing
    df["y"] -= 1
    return df


def _cumulative_drift(drift_df, ref_frame):
    """dx, dy relative to the frame the AOIs were picked in (glimpse_reader.py:99-108)."""
    cols = ["dx", "dy"]
    after = drift_df.loc[ref_frame + 1:, cols]
    drift_df.loc[ref_frame + 1:, cols] = after.cumsum(axis=0).values
    first = drift_df.index[1]
    before = -drift_df.loc[ref_frame:first:-1, cols]
    drift_df.loc[ref_frame - 1::-1, cols] = before.cumsum(axis=0).values
    return drift_df


def _interval_labels(path, aoi_index, frame_index):
    """Spot-picker intervals -> (aoi, frame, z, spotpicker) record array (glimpse_reader.py:115-149)."""
    from scipy.io import loadmat

    rec = np.zeros((len(aoi_index), len(frame_index)),
                   dtype=[("aoi", int), ("frame", int), ("z", bool), ("spotpicker", float)])
    rec["aoi"] = np.asarray(aoi_index).reshape(-1, 1)
    rec["frame"] = np.asarray(frame_index)
    for row in loadmat(path)["Intervals"]["CumulativeIntervalArray"][0, 0]:
        kind, start, end, aoi = row[0], int(row[1]), int(row[2]), int(row[-1])
        if kind in (-2.0, 0.0, 2.0):
            value = 0
        elif kind in (-3.0, 1.0, 3.0):
            value = 1
        else:
            continue
        rec["spotpicker"][(rec["aoi"] == aoi) & (rec["frame"] >= start) & (rec["frame"] <= end)] = value
    rec["z"] = rec["spotpicker"]
    return rec


class GlimpseDataset:
    """Header, AOI locations, cumulative drift and (optional) labels of one channel of a glimpse movie."""

    def __init__(self, c=0, **kwargs):
        import pandas as pd
        from scipy.io import loadmat

        self.config = kwargs
        self.folder = Path(kwargs["glimpse-folder"])
        self.dtypes = ["ontarget"] + (["offtarget"] if kwargs["use-offtarget"] else [])
        vid = loadmat(self.folder / "header.mat")["vid"]
        self.header = {name: np.squeeze(vid[0, 0][i]) for i, name in enumerate(vid.dtype.names)}
        self.height, self.width = int(self.header["height"]), int(self.header["width"])

        drift = pd.DataFrame(loadmat(kwargs["driftlist"])["driftlist"][:, :3], columns=["frame", "dy", "dx"])
        drift = drift.astype({"frame": int}).set_index("frame")
        drift["ttb"] = self.header["ttb"]
        self.aoiinfo = {dtype: _aoi_table(kwargs[f"{dtype}-aoiinfo"]) for dtype in self.dtypes}
        drift = _cumulative_drift(drift, int(self.aoiinfo["ontarget"].at[1, "frame"]))
        if kwargs["frame-range"]:
            drift = drift.loc[int(kwargs["frame-start"]):int(kwargs["frame-end"])]
        self.cumdrift = drift

        self.labels = defaultdict(lambda: None)
        for dtype in self.dtypes:
            if kwargs["labels"] and kwargs[f"{dtype}-labels"] is not None:
                self.labels[dtype] = _interval_labels(kwargs[f"{dtype}-labels"], self.aoiinfo[dtype].index.values,
                                                      self.cumdrift.index.values)
        self.name = kwargs["name"]
        self.c = c
        self.offset_x, self.offset_y = kwargs["offset-x"], kwargs["offset-y"]

    # ---- frames ---------------------------------------------------------------------------------------------
    def raw_frame(self, frame, out=None):
        """The frame's bytes as stored (big-endian int16), viewed as native uint16 WITHOUT swapping."""
        number = int(self.header["filenumber"][frame - 1])
        with open(self.folder / f"{number}.glimpse", "rb") as fid:
            fid.seek(int(self.header["offset"][frame - 1]))
            raw = np.fromfile(fid, dtype=np.uint16, count=self.height * self.width)
        raw = raw.reshape(self.height, self.width)
        if out is not None:
            out[...] = raw
            return out
        return raw

    def __getitem__(self, key):
        """Decoded frame(s) on the host (compatibility with the reference; the GPU path uses raw_frame)."""
        if isinstance(key, slice):
            return np.stack([self[f] for f in range(key.start, key.stop, key.step or 1)], 0)
        return self.raw_frame(key).view(">i2").astype(np.int64) + 2**15

    def __len__(self):
        return self.F

    @property
    def N(self):
        return len(self.aoiinfo["ontarget"])

    @property
    def Nc(self):
        return len(self.aoiinfo["offtarget"]) if "offtarget" in self.dtypes else 0

    @property
    def F(self):
        return len(self.cumdrift)

    def __repr__(self):
        return f"{type(self).__name__}(N={self.N}, Nc={self.Nc}, F={self.F})"


def offset_distribution(counts, min_data, bin_size):
    """
    Offset samples / weights from the pooled pixel counts of the offset regions (glimpse_reader.py:411-433):
    a guard bin below the darkest data pixel, normalisation, the top 0.5 % folded into the last kept bin,
    thinning with :func:`bin_hist`.  ``counts``: (65536,) integer array.
    """
    counts = np.asarray(counts)
    values = np.nonzero(counts)[0]
    samples, weights = values.astype(np.int64), counts[values].astype(np.int64)
    if min_data <= samples[0]:
        samples = np.insert(samples, 0, min_data - 1)
        weights = np.insert(weights, 0, 1)
    weights = weights / weights.sum()
    drop = weights.cumsum() > 0.995
    dropped = weights[drop].sum()
    samples, weights = samples[~drop], weights[~drop]
    weights[-1] += dropped
    return bin_hist(torch.tensor(samples, dtype=torch.int), torch.tensor(weights), bin_size)


def crop_movie(movie, P, offset_P, device, counts, progress_bar=None, chunk_frames=64):
    """
    All AOIs x all frames of one channel on the GPU.  Returns patches (N_total, F, P, P) uint16 and target_xy
    (N_total, F, 2) float64 (on-target AOIs first), and adds the offset-region pixel counts to ``counts``
    ((65536,) int64 CUDA tensor).
    """
    import ctypes

    from tapqir_b200 import _lib

    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("tapqir_b200 has no CPU execution path: read_glimpse needs device='cuda'")
    H, W = movie.height, movie.width
    xy = np.concatenate([movie.aoiinfo[d][["x", "y"]].values for d in movie.dtypes], 0).astype(np.float64)
    drift = movie.cumdrift[["dx", "dy"]].values.astype(np.float64)
    N, F = len(xy), len(drift)
    frames = list(movie.cumdrift.index)
    p = _lib.ptr
    with torch.cuda.device(dev):
        xy_d = torch.from_numpy(np.ascontiguousarray(xy)).to(dev)
        drift_d = torch.from_numpy(np.ascontiguousarray(drift)).to(dev)
        patches = torch.zeros(N, F, P, P, dtype=torch.uint16, device=dev)
        target = torch.zeros(N, F, 2, dtype=torch.float64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        staging = [torch.empty(chunk_frames, H, W, dtype=torch.uint16).pin_memory() for _ in range(2)]
        chunk_d = [torch.empty(chunk_frames, H, W, dtype=torch.uint16, device=dev) for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        st = _lib.stream_ptr(dev)
        it = range(0, F, chunk_frames)
        for i, f0 in enumerate(progress_bar(it) if progress_bar is not None else it):
            slot = i % 2
            fc = min(chunk_frames, F - f0)
            done[slot].synchronize()   # the kernels that read this slot two chunks ago have finished
            host = staging[slot].numpy()
            for j in range(fc):
                movie.raw_frame(frames[f0 + j], out=host[j])
            chunk_d[slot][:fc].copy_(staging[slot][:fc], non_blocking=True)
            _lib.check(lib.tq_crop_aois(p(chunk_d[slot]), H, W, fc, f0, p(xy_d), p(drift_d), N, F, P, p(patches), p(target),
                                        p(status), st), "tq_crop_aois")
            _lib.check(lib.tq_offset_hist(p(chunk_d[slot]), H, W, fc, int(movie.offset_x), int(movie.offset_y), int(offset_P),
                                          p(counts), st), "tq_offset_hist")
            done[slot].record()
        torch.cuda.synchronize(dev)
        if status.item() != 0:
            raise ValueError("an AOI window leaves the field of view (check aoiinfo / driftlist against the frame size)")
    return patches, target


def read_glimpse(path, progress_bar, **kwargs):
    """Preprocess glimpse files into ``<path>/data.tpqr`` (same options as the reference's read_glimpse)."""
    kwargs = dict(kwargs)
    P, C = kwargs.pop("P"), kwargs.pop("num-channels")
    name, channels = kwargs.pop("dataset"), kwargs.pop("channels")
    offset_P, bin_size = kwargs.pop("offset-P"), kwargs.pop("bin-size")
    device = kwargs.pop("device", "cuda")
    path = Path(path)

    counts = torch.zeros(65536, dtype=torch.int64, device=device)
    per_channel, time1, ttb = [], [], []
    for c in range(C):
        logger.info(f"Channel #{c} ({channels[c]['name']})")
        movie = GlimpseDataset(**kwargs, **channels[c], c=c)
        time1.append(float(movie.header["time1"]))
        ttb.append(movie.cumdrift["ttb"].values)
        patches, target = crop_movie(movie, P, offset_P, device, counts, progress_bar)
        # target positions must sit in the central pixel (glimpse_reader.py:383-386)
        assert (target > 0.5 * P - 1).all() and (target < 0.5 * P).all()
        per_channel.append((movie, patches.cpu(), target.cpu()))

    logger.info("Processing extracted AOIs ...")
    movie = per_channel[0][0]
    data = torch.stack([pc[1].to(torch.int64) for pc in per_channel], 2)       # (N, F, C, P, P)
    target_xy = torch.stack([pc[2] for pc in per_channel], 2)                 # (N, F, C, 2)
    sizes = [len(movie.aoiinfo[d]) for d in movie.dtypes]
    is_ontarget = torch.cat([torch.full((n,), d == "ontarget", dtype=torch.bool) for n, d in zip(sizes, movie.dtypes)])
    labels = []
    for n, d in zip(sizes, movie.dtypes):
        per = [pc[0].labels[d] for pc in per_channel]
        if all(l is not None for l in per):
            labels.append(np.stack(per, -1))
    labels = np.concatenate(labels, 0) if labels else None
    offset_samples, offset_weights = offset_distribution(counts.cpu().numpy(), int(data.min().item()), bin_size)

    dataset = CosmosDataset(data, target_xy, is_ontarget, labels=labels, offset_samples=offset_samples,
                            offset_weights=offset_weights, time1=torch.as_tensor(time1),
                            ttb=torch.as_tensor(np.array(ttb)).T, name=name,
                            channels=tuple(ch["name"] for ch in channels))
    logger.info(f"Dataset: N={dataset.N} on-target AOIs, Nc={dataset.Nc} off-target AOIs, F={dataset.F} frames, "
                f"C={dataset.C} channels, Px={dataset.P} pixels, Py={dataset.P} pixels")
    save(dataset, path)
    return dataset
