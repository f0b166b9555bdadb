"""
CPU restatement (numpy, plain Python loops) of the ingestion arithmetic of the reference's
``tapqir/imscroll/glimpse_reader.py`` -- TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing under
tapqir_b200/).  Pinned bit for bit (tests/test_glimpse_cpu.py): ``bin_hist`` against the reference's own function
(tests/golden/ref_glimpse.pt, made by tests/golden/make_golden_glimpse.py from the reference source); the frame loop
(glimpse_reader.py:168-186, :354-381) and the offset post-processing (:411-433) against ``data.tpqr`` as written by the
reference's whole ``read_glimpse`` run verbatim on a synthetic movie (tests/golden/ref_glimpse_folder/, made by
tests/golden/make_golden_step.py).
"""

from collections import OrderedDict, defaultdict

import numpy as np
import torch


def decode_frame(raw_bytes, height, width):
    """glimpse_reader.py:181-186: big-endian int16 + 2**15 (the reference relies on numpy-1 value-based promotion
    of ``int16 + 32768`` to a wider integer; spelled out here because numpy 2 raises instead)."""
    return np.frombuffer(raw_bytes, dtype=">i2", count=height * width).reshape(height, width).astype(np.int64) + 2**15


def crop_loop(frames, aoi_xy, cumdrift, P):
    """
    glimpse_reader.py:338-381.  frames: (F, H, W) decoded ints; aoi_xy: (N, 2) [x, y]; cumdrift: (F, 2) [dx, dy].
    Returns data (N, F, P, P) int64 and target_xy (N, F, 2) float64.
    """
    N, F = len(aoi_xy), len(cumdrift)
    raw_target_xy = np.expand_dims(aoi_xy, axis=1) + cumdrift          # :338-341
    data = np.zeros((N, F, P, P), dtype="int")
    target_xy = np.zeros((N, F, 2))
    for f in range(F):
        img = frames[f]
        for n in range(N):
            shiftx = round(raw_target_xy[n, f, 0] - 0.5 * (P - 1))      # :365  (Python round: ties to even)
            shifty = round(raw_target_xy[n, f, 1] - 0.5 * (P - 1))      # :366
            data[n, f, :, :] += img[shifty:shifty + P, shiftx:shiftx + P]
            target_xy[n, f, 0] = raw_target_xy[n, f, 0] - shiftx
            target_xy[n, f, 1] = raw_target_xy[n, f, 1] - shifty
    return data, target_xy


def offset_counts(frames, offset_x, offset_y, offset_P, offsets=None):
    """glimpse_reader.py:356-362: pooled value -> count dictionary of the offset region of every frame."""
    offsets = defaultdict(int) if offsets is None else offsets
    for img in frames:
        region = img[offset_y:offset_y + offset_P, offset_x:offset_x + offset_P]
        values, counts = np.unique(region, return_counts=True)
        for value, count in zip(values, counts):
            offsets[int(value)] += int(count)
    return offsets


def bin_hist(samples, weights, s):
    """
    glimpse_reader.py:22-37 as explicit loops: bin 0 is kept; the following bins are merged s at a time (sample:
    the group's element s // 2, weight: the group's weights added one after the other in torch's DEFAULT dtype, as
    the reference's in-place ``+=`` on ``torch.zeros(n)`` does); a shorter last group takes element r // 2 and the
    (input-dtype) sum of what is left.
    """
    dt = torch.get_default_dtype()
    count = len(samples) - 1
    groups, left = count // s, count % s
    out_s = [int(samples[0])]
    out_w = [weights[0].to(dt)]
    for g in range(groups):
        first = 1 + g * s
        out_s.append(int(samples[first + s // 2]))
        total = torch.zeros((), dtype=dt)
        wide = torch.promote_types(dt, weights.dtype)   # an in-place add computes in the promoted type, then casts back
        for i in range(s):
            total = (total.to(wide) + weights[first + i].to(wide)).to(dt)
        out_w.append(total)
    if left:
        first = 1 + groups * s
        out_s.append(int(samples[first + left // 2]))
        out_w.append(weights[first:].sum().to(dt))
    return torch.tensor(out_s, dtype=torch.int), torch.stack(out_w)


def offset_distribution(offsets, min_data, bin_size):
    """glimpse_reader.py:411-433 on the dictionary of :func:`offset_counts`."""
    offsets = OrderedDict(sorted(offsets.items()))
    offset_samples = np.array(list(offsets.keys()))
    offset_weights = np.array(list(offsets.values()))
    if min_data <= offset_samples[0]:
        offset_samples = np.insert(offset_samples, 0, min_data - 1)
        offset_weights = np.insert(offset_weights, 0, 1)
    offset_weights = offset_weights / offset_weights.sum()
    high_mask = offset_weights.cumsum() > 0.995
    high_weights = offset_weights[high_mask].sum()
    offset_samples = offset_samples[~high_mask]
    offset_weights = offset_weights[~high_mask]
    offset_weights[-1] += high_weights
    offset_samples = torch.tensor(offset_samples, dtype=torch.int)
    offset_weights = torch.tensor(offset_weights)
    return bin_hist(offset_samples, offset_weights, bin_size)
