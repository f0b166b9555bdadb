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
