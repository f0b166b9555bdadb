"""
Dataset container with the reference's interface (tapqir/utils/dataset.py:18-222) plus a
device-resident pixel store for the CUDA path.

Same attribute / property names (``images, xy, is_ontarget, mask, labels, offset, N, Nc, Nt, F, C,
P, x, y, median, vmin, vmax, fetch``) and the same ``data.tpqr`` key set (dataset.py:195-213), so
files written by either implementation load in the other.

What is different, by design: ``fetch`` in the reference gathers the minibatch on the CPU and
copies it to the device every step (dataset.py:140-151).  Here :meth:`CosmosDataset.device_store`
uploads the whole pixel array once (``uint16`` when the pixels are integers in [0, 65535], as
produced by glimpse_reader.py:183-186 and simulate.py:122; ``float32`` otherwise) and the step
kernel gathers directly from HBM with the minibatch indices.
"""

import logging
from collections import namedtuple
from pathlib import Path

import torch

from tapqir_b200.exceptions import TapqirFileNotFoundError

logger = logging.getLogger(__name__)


class OffsetData(namedtuple("OffsetData", ["samples", "weights"])):
    """Empirical camera-offset distribution (dataset.py:18-37)."""

    @property
    def min(self):
        return torch.min(self.samples).item()

    @property
    def max(self):
        return torch.max(self.samples).item()

    @property
    def logits(self):
        # probs_to_logits: log of the eps-clamped weights (dataset.py:27-29)
        eps = torch.finfo(self.weights.dtype).eps
        return torch.log(self.weights.clamp(min=eps, max=1 - eps))

    @property
    def mean(self):
        return torch.sum(self.samples * self.weights).item()

    @property
    def var(self):
        return torch.sum(self.samples**2 * self.weights).item() - self.mean**2


DeviceStore = namedtuple("DeviceStore", ["pixels", "xy", "is_ontarget", "mask", "offset_samples", "offset_logits"])


def merge_offset_support(samples, logits):
    """
    Distinct support points of the empirical offset distribution: bins with the same sample value are
    merged and their weights added.  The likelihood marginalises the offset as ``sum_j w_j f(D - delta_j)``
    (ksmogn.py:222-238), so this is an identity, and the kernels' work per pixel is proportional to the
    number of bins.  The reference's simulator writes three identical bins (simulate.py:92,103); histograms
    from real movies (glimpse_reader.py:421) have distinct bins and pass through unchanged.
    """
    uniq, inverse = torch.unique(samples, sorted=True, return_inverse=True)
    if uniq.numel() == samples.numel():
        return samples, logits
    weights = torch.zeros_like(uniq, dtype=torch.float64).index_add_(0, inverse, logits.double().exp())
    return uniq, weights.log().to(logits.dtype)


class CosmosDataset:
    """AOI x frame x channel stack of PxP patches with target positions and labels."""

    def __init__(self, images, xy, is_ontarget, mask=None, labels=None, offset_samples=None,
                 offset_weights=None, device=torch.device("cpu"), time1=None, ttb=None, name=None,
                 channels=None):
        self.images = images
        self.xy = xy
        self.is_ontarget = is_ontarget
        self.mask = torch.ones_like(is_ontarget, dtype=torch.bool) if mask is None else mask
        self.labels = labels
        self.device = torch.device(device)
        self.offset = OffsetData(offset_samples.to(self.device), offset_weights.to(self.device))
        self.time1 = time1
        self.ttb = ttb
        self.name = name
        self.channels = tuple(f"channel{c}" for c in range(self.C)) if channels is None else channels
        self._store = None
        self._median = None

    # ---- sizes -------------------------------------------------------------------------------
    @property
    def N(self) -> int:
        """On-target AOIs."""
        return int(self.is_ontarget.sum().item())

    @property
    def Nc(self) -> int:
        """Off-target (control) AOIs."""
        return int((~self.is_ontarget).sum().item())

    @property
    def Nt(self) -> int:
        return self.N + self.Nc

    @property
    def F(self) -> int:
        return self.images.shape[1]

    @property
    def C(self) -> int:
        return self.images.shape[2]

    @property
    def P(self) -> int:
        assert self.images.shape[3] == self.images.shape[4]
        return self.images.shape[3]

    @property
    def x(self) -> torch.Tensor:
        return self.xy[..., 0]

    @property
    def y(self) -> torch.Tensor:
        return self.xy[..., 1]

    @property
    def median(self) -> torch.Tensor:
        """Per-channel median pixel (dataset.py:134-138); seeds the background parameters."""
        if self._median is None:
            self._median = torch.stack([torch.median(self.images[..., c, :, :]) for c in range(self.C)])
        return self._median

    def _quantile(self, q):
        out = []
        for c in range(self.C):
            flat = self.images[..., c, :, :].flatten().float()
            # kthvalue-based: torch.quantile refuses inputs above 16M elements
            k = min(max(int(round(q * (flat.numel() - 1))) + 1, 1), flat.numel())
            out.append(torch.kthvalue(flat, k).values)
        return torch.stack(out)

    @property
    def vmin(self) -> torch.Tensor:
        return self._quantile(0.05)

    @property
    def vmax(self) -> torch.Tensor:
        return self._quantile(0.99)

    # ---- minibatch access --------------------------------------------------------------------
    def fetch(self, ndx, fdx, cdx):
        """Reference-compatible host gather (dataset.py:140-151); the CUDA step does not use it."""
        cpu = lambda i: i.cpu() if isinstance(i, torch.Tensor) else i
        ndx, fdx, cdx = cpu(ndx), cpu(fdx), cpu(cdx)
        return (
            self.images[ndx, fdx, cdx].to(self.device),
            self.xy[ndx, fdx, cdx].to(self.device),
            self.is_ontarget[ndx].to(self.device),
        )

    def device_store(self, device=None, dtype=torch.float32, aoi_slice=slice(None), merge_offsets=True) -> DeviceStore:
        """
        One-time upload of (a contiguous AOI shard of) the dataset in the layout the kernels read:
        pixels (Nt,F,C,P,P) uint16|float32, xy (Nt,F,C,2) ``dtype``, is_ontarget / mask (Nt,) uint8,
        offset samples / log-weights (O,) ``dtype``; identical offset bins merged
        (:func:`merge_offset_support`) unless ``merge_offsets=False``.
        """
        device = torch.device(device or self.device)
        key = (str(device), dtype, aoi_slice.start, aoi_slice.stop, merge_offsets)
        if self._store is not None and self._store[0] == key:
            return self._store[1]
        img = self.images[aoi_slice]
        integral = (not img.dtype.is_floating_point) or bool((img == img.floor()).all())
        if integral and img.numel() and 0 <= img.min().item() and img.max().item() <= 65535:
            pixels = img.to(torch.int32).to(torch.uint16).contiguous().to(device)
        else:
            pixels = img.to(torch.float32).contiguous().to(device)
        off_s, off_l = self.offset.samples, self.offset.logits
        if merge_offsets:
            off_s, off_l = merge_offset_support(off_s, off_l)
        store = DeviceStore(
            pixels=pixels,
            xy=self.xy[aoi_slice].to(device=device, dtype=dtype).contiguous(),
            is_ontarget=self.is_ontarget[aoi_slice].to(device=device, dtype=torch.uint8).contiguous(),
            mask=self.mask[aoi_slice].to(device=device, dtype=torch.uint8).contiguous(),
            offset_samples=off_s.to(device=device, dtype=dtype).contiguous(),
            offset_logits=off_l.to(device=device, dtype=dtype).contiguous(),
        )
        self._store = (key, store)
        return store

    def __repr__(self):
        return (
            f"{self.__class__.__name__}: {self.name}\n"
            f"  images  (N={self.N} on-target, Nc={self.Nc} off-target, F={self.F}, C={self.C}, "
            f"P={self.P}, P={self.P})\n"
            f"  offset.samples {tuple(self.offset.samples.shape)}  offset.weights {tuple(self.offset.weights.shape)}"
        )


def save(obj: CosmosDataset, path):
    """Write ``<path>/data.tpqr`` with the reference's key set (dataset.py:195-213)."""
    path = Path(path)
    payload = {
        "images": obj.images,
        "xy": obj.xy,
        "is_ontarget": obj.is_ontarget,
        "mask": obj.mask,
        "labels": obj.labels,
        "offset_samples": obj.offset.samples.cpu(),
        "offset_weights": obj.offset.weights.cpu(),
        "name": obj.name,
        "time1": obj.time1,
        "ttb": obj.ttb,
        "channels": obj.channels,
    }
    torch.save(payload, path / "data.tpqr")
    logger.info(f"Data is saved in {path / 'data.tpqr'}")


def load(path, device=torch.device("cpu")) -> CosmosDataset:
    """Read ``<path>/data.tpqr`` (dataset.py:216-222)."""
    path = Path(path)
    try:
        payload = torch.load(path / "data.tpqr", weights_only=False)
    except FileNotFoundError:
        raise TapqirFileNotFoundError("data", path / "data.tpqr")
    return CosmosDataset(**payload, device=device)
