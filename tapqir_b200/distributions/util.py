"""
Distribution utilities with the reference's names and semantics (tapqir/distributions/util.py).

``gaussian_spots`` runs the sm_100a kernel (``tq_gaussian_spots``); the prior tables
(``truncated_poisson_probs``, ``probs_m``, ``probs_theta``, ``expand_offtarget``) are a few scalars
and stay as host-side torch code -- inside the SVI step the same tables are rebuilt in-kernel from
the sampled ``lamda`` / ``pi`` (csrc/cosmos_globals.cuh).
"""

import torch

from tapqir_b200 import _lib


def gaussian_spots(height, width, x, y, target_locs, P, m=None):
    r"""
    Ideal 2-D Gaussian spot shapes (reference: distributions/util.py:15-64)

    .. math:: \mu^S_{i,j} = \frac{m\,h}{2\pi w^2}
              \exp\left(-\frac{(i-x-x^{target})^2 + (j-y-y^{target})^2}{2w^2}\right)

    Same broadcasting as the reference: ``height, width, x, y`` (and ``m``) broadcast to a common
    ``batch`` shape, ``target_locs`` to ``batch + (2,)``; returns ``batch + (P, P)`` with x along the
    last axis (in the model ``batch = (N, F, C, K)`` with ``target_locs (N, F, C, 1, 2)``; the post-fit
    statistics call it with ``batch = (K, F, Q)``, stats.py:65-72).  Forward only (the differentiable
    route is :class:`tapqir_b200.distributions.KSMOGN`); CUDA tensors only.
    """
    tensors = [height, width, x, y] + ([m] if m is not None else [])
    if any(t.requires_grad for t in tensors if isinstance(t, torch.Tensor)) and torch.is_grad_enabled():
        raise NotImplementedError("gaussian_spots is forward-only; differentiate through KSMOGN.log_prob")
    dtype = height.dtype
    shape = torch.broadcast_shapes(*[t.shape for t in tensors], target_locs.shape[:-1])
    U = 1
    for s in shape:
        U *= s
    flat = lambda t: t.to(dtype).expand(shape).reshape(1, U).contiguous()
    h, w, xx, yy = flat(height), flat(width), flat(x), flat(y)
    mm = flat(m) if m is not None else None
    tgt = target_locs.to(dtype).expand(shape + (2,)).reshape(U, 2).contiguous()
    out = torch.empty((U, 1, P, P), dtype=dtype, device=height.device)
    lib = _lib.load()
    with torch.cuda.device(height.device):
        # every broadcast element is an independent "spot": K = 1 in the kernel's (K, U) layout
        _lib.check(lib.tq_gaussian_spots(_lib.dtype_code(dtype), U, 1, P, _lib.ptr(h), _lib.ptr(w), _lib.ptr(xx),
                                         _lib.ptr(yy), _lib.ptr(tgt), _lib.ptr(mm), _lib.ptr(out),
                                         _lib.stream_ptr(height.device)), "tq_gaussian_spots")
    return out.reshape(shape + (P, P))


def truncated_poisson_probs(lamda: torch.Tensor, K: int) -> torch.Tensor:
    """Poisson pmf on 0..K-1 with the tail mass collected at K (reference: util.py:67-91)."""
    k = torch.arange(K, dtype=lamda.dtype, device=lamda.device)
    lam = lamda.unsqueeze(-1)
    head = torch.exp(torch.xlogy(k, lam) - lam - torch.lgamma(k + 1))
    return torch.cat([head, 1 - head.sum(-1, keepdim=True)], -1)


def probs_m(lamda: torch.Tensor, K: int) -> torch.Tensor:
    """p(m_k = 1 | theta, lamda), shape ``lamda.shape + (1+K, K)`` (reference: util.py:94-130)."""
    occupancy = lambda n: (
        (torch.arange(1, n + 1, dtype=lamda.dtype, device=lamda.device) * truncated_poisson_probs(lamda, n)[..., 1:]).sum(-1) / n
    )
    out = occupancy(K - 1)[..., None, None].expand(lamda.shape + (1 + K, K)).clone()
    out[..., 0, :] = occupancy(K)[..., None]
    idx = torch.arange(K)
    out[..., idx + 1, idx] = 1
    return out


def expand_offtarget(probs: torch.Tensor) -> torch.Tensor:
    """Append the off-target row e_0 (reference: util.py:133-151); result ``probs.shape + (2,)``."""
    off = torch.zeros_like(probs)
    off[..., 0] = 1
    return torch.stack([off, probs], dim=-1)


def probs_theta(K: int, device=None) -> torch.Tensor:
    """p(theta | z): row 0 = e_0, row 1 = uniform over 1..K (reference: util.py:154-173)."""
    out = torch.zeros(2, 1 + K, device=device)
    out[0, 0] = 1
    out[1, 1:] = 1 / K
    return out
