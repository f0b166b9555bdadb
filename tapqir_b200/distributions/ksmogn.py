"""
KSMOGN -- K-Spots Marginalized Offset Gamma Noise image distribution, operator-level seam.

Same constructor and ``log_prob`` semantics as the reference class
(tapqir/distributions/ksmogn.py:21-238) so ``tapqir/models/cosmos.py:310-327`` could call it
unchanged; the arithmetic is the fused sm_100a kernel ``tq_ksmogn_fwd`` / ``tq_ksmogn_fwd_bwd``
(csrc/ksmogn.cu) instead of KeOps ``Genred`` / the materialised torch branch.

.. math::
    \\mu^I = b + \\sum_k m_k \\mu^S_k, \\qquad
    p(D \\mid \\mu^I, g) = \\sum_\\delta p(\\delta)\\,\\mathrm{Gamma}(D - \\delta \\mid \\mu^I / g,\\ 1/g)

Inside the SVI step the model does not go through this class: ``cosmos.run`` calls the fused step
kernels directly.  This seam exists for drop-in use and for the operator-level parity tests.
"""

import torch
from torch.distributions import constraints
from torch.distributions.distribution import Distribution

from tapqir_b200 import _lib
from tapqir_b200.distributions.util import gaussian_spots


def _enumerated_table(dtype, device):
    return torch.tensor([[(m >> k) & 1 for k in range(_lib.K)] for m in range(_lib.M)], dtype=dtype, device=device)


class _KsmognLogProb(torch.autograd.Function):
    """log p for NM configurations of U patches; inputs already flattened to kernel layout."""

    @staticmethod
    def forward(ctx, height, width, x, y, background, gain, target, value, off_s, off_w, mcfg, P):
        U, NM = background.numel(), mcfg.shape[0]
        dtype, dev = background.dtype, background.device
        view = _lib.make_view(value, target, off_s, off_w, nb=U, fb=1, C=1, F=1, P=P)
        logp = torch.empty((NM, U), dtype=dtype, device=dev)
        lib = _lib.load()
        # the enumerated {0,1}^K table selects the specialised kernel (mcfg = NULL in the C ABI)
        ctx.enumerated = NM == _lib.M and bool(torch.equal(mcfg, _enumerated_table(dtype, dev)))
        with torch.cuda.device(dev):
            _lib.check(lib.tq_ksmogn_fwd(_lib.dtype_code(dtype), view, _lib.ptr(height), _lib.ptr(width),
                                         _lib.ptr(x), _lib.ptr(y), _lib.ptr(background), _lib.ptr(gain),
                                         None if ctx.enumerated else _lib.ptr(mcfg), NM, _lib.ptr(logp),
                                         _lib.stream_ptr(dev)), "tq_ksmogn_fwd")
        ctx.save_for_backward(height, width, x, y, background, gain, target, value, off_s, off_w, mcfg)
        ctx.P = P
        return logp

    @staticmethod
    def backward(ctx, grad_logp):
        height, width, x, y, background, gain, target, value, off_s, off_w, mcfg = ctx.saved_tensors
        U, NM = background.numel(), mcfg.shape[0]
        dtype, dev = background.dtype, background.device
        view = _lib.make_view(value, target, off_s, off_w, nb=U, fb=1, C=1, F=1, P=ctx.P)
        W = grad_logp.to(dtype).contiguous()
        g_h, g_w, g_x, g_y = (torch.empty_like(height) for _ in range(4))
        g_b, g_rate = torch.empty_like(background), torch.empty_like(background)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.tq_ksmogn_fwd_bwd(_lib.dtype_code(dtype), view, _lib.ptr(height), _lib.ptr(width),
                                             _lib.ptr(x), _lib.ptr(y), _lib.ptr(background), _lib.ptr(gain),
                                             None if ctx.enumerated else _lib.ptr(mcfg), NM, _lib.ptr(W), None, _lib.ptr(g_h), _lib.ptr(g_w),
                                             _lib.ptr(g_x), _lib.ptr(g_y), _lib.ptr(g_b), _lib.ptr(g_rate),
                                             _lib.stream_ptr(dev)), "tq_ksmogn_fwd_bwd")
        g_gain = (-(g_rate.sum()) / (gain * gain)).reshape(gain.shape)
        return g_h, g_w, g_x, g_y, g_b, g_gain, None, None, None, None, None, None


class KSMOGN(Distribution):
    """
    :param height, width, x, y: spot parameters, broadcastable to ``batch_shape + (K,)``.
    :param target_locs: target location, broadcastable to ``batch_shape + (2,)`` (x, y).
    :param background: background intensity, broadcastable to ``batch_shape``.
    :param gain: camera gain (scalar tensor).
    :param offset_samples, offset_logits: empirical offset distribution ``(O,)``.
    :param P: patch edge in pixels.
    :param m: spot presence indicator, broadcastable to ``batch_shape + (K,)``; the enumerated layout
        ``(2, 2, 1, 1, 1, K)`` of the cosmos model is evaluated in one fused pass over the 4 configs.
    :param alpha: crosstalk matrix -- not supported (crosstalk model is out of scope).
    :param use_pykeops: accepted for signature compatibility; ignored.
    """

    arg_constraints = {}
    support = constraints.positive
    has_rsample = True

    def __init__(self, height, width, x, y, target_locs, background, gain, offset_samples, offset_logits,
                 P, m=None, alpha=None, use_pykeops=True, validate_args=None):
        if alpha is not None:
            raise NotImplementedError("crosstalk (alpha) variant is outside the cosmos hot path")
        self.height, self.width, self.x, self.y = height, width, x, y
        self.target_locs, self.background, self.gain = target_locs, background, gain
        self.offset_samples, self.offset_logits = offset_samples, offset_logits
        self.P, self.m = P, m
        shapes = [height.shape, width.shape, x.shape, y.shape] + ([m.shape] if m is not None else [])
        batch = torch.broadcast_shapes(*shapes)[:-1]
        batch = torch.broadcast_shapes(batch, background.shape, target_locs.shape[:-1])
        super().__init__(batch, torch.Size([P, P]), validate_args=False)

    # ---- reference attributes ------------------------------------------------------------------
    @property
    def gaussians(self):
        return gaussian_spots(self.height, self.width, self.x, self.y, self.target_locs.unsqueeze(-2), self.P, self.m)

    @property
    def image(self):
        return self.background[..., None, None] + self.gaussians.sum(-3)

    @property
    def rate(self):
        return 1 / self.gain

    @property
    def concentration(self):
        return self.image / self.gain

    def rsample(self, sample_shape=torch.Size()):
        """Simulation helper (reference: ksmogn.py:171-185): Gamma(image/gain, 1/gain) + offset draw."""
        with torch.no_grad():
            shape = self._extended_shape(sample_shape)
            probs = torch.softmax(self.offset_logits, 0)
            odx = torch.multinomial(probs, int(torch.Size(shape).numel()), replacement=True).reshape(shape)
            conc = self.concentration.expand(shape).contiguous()
            value = torch._standard_gamma(conc) * self.gain
            value.clamp_(min=torch.finfo(value.dtype).tiny)
            return value + self.offset_samples[odx]

    # ---- the operator ---------------------------------------------------------------------------
    def _split_config_dims(self, core_shape):
        """If ``m`` varies only over leading dims where everything else is size 1, return the
        (NM, K) configuration table and the shape of those dims; else None."""
        m = self.m
        nd = max(len(core_shape), m.dim() - 1)
        core = (1,) * (nd - len(core_shape)) + tuple(core_shape)
        ms = (1,) * (nd - (m.dim() - 1)) + tuple(m.shape[:-1])
        split = 0
        for i in range(nd):
            if ms[i] > 1:
                split = i + 1
        if any(c > 1 for c in core[:split]) or any(s > 1 for s in ms[split:]):
            return None
        cfg_shape = ms[:split]
        table = m.reshape(-1, m.shape[-1])
        return table, cfg_shape, core[split:]

    def log_prob(self, value):
        dtype, dev, Kk = self.background.dtype, self.background.device, _lib.K
        core = torch.broadcast_shapes(self.height.shape[:-1], self.width.shape[:-1], self.x.shape[:-1],
                                      self.y.shape[:-1], self.background.shape, self.target_locs.shape[:-1],
                                      value.shape[:-2])
        height = self.height
        cfg_shape, table = (), None
        if self.m is not None:
            split = self._split_config_dims(core)
            if split is not None and split[0].shape[0] in (1, _lib.M):
                table, cfg_shape, core = split
                table = table.to(dtype).contiguous()
            else:
                # general broadcast: fold m into the height exactly as the reference does
                # (util.py:62-63 ``height = m * height``)
                height = self.m * height
                core = torch.broadcast_shapes(core, self.m.shape[:-1])
        if table is None:
            table = torch.ones((1, Kk), dtype=dtype, device=dev)
        if height.shape[-1] != Kk:
            raise ValueError(f"kernels are built for K={Kk} spots")
        U = 1
        for s in core:
            U *= s
        drop = lambda t, nd: t.reshape(t.shape[max(t.dim() - nd, 0):]) if t.dim() > nd else t
        nd = len(core)
        spot = lambda t: drop(t.to(dtype), nd + 1).expand(tuple(core) + (Kk,)).reshape(U, Kk).t().contiguous()
        unit = lambda t: drop(t.to(dtype), nd).expand(tuple(core)).reshape(U).contiguous()
        tgt = drop(self.target_locs.to(dtype), nd + 1).expand(tuple(core) + (2,)).reshape(U, 2).contiguous()
        pix_dtype = dtype if value.dtype.is_floating_point and value.dtype == torch.float64 else torch.float32
        val = drop(value, nd + 2).to(pix_dtype).expand(tuple(core) + (self.P, self.P)).reshape(U, self.P, self.P).contiguous()
        gain = self.gain.to(dtype).reshape(1)
        logp = _KsmognLogProb.apply(spot(height), spot(self.width), spot(self.x), spot(self.y),
                                    unit(self.background), gain, tgt, val,
                                    self.offset_samples.to(dtype).contiguous(),
                                    self.offset_logits.to(dtype).contiguous(), table, self.P)
        return logp.reshape(tuple(cfg_shape) + tuple(core))
