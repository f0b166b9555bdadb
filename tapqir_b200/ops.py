"""
The likelihood operator as ``torch.library`` custom ops over the C ABI (include/tapqir_b200.h):

    torch.ops.tapqir_b200.ksmogn_log_prob(height, width, x, y, background, gain, target, value,
                                          offset_samples, offset_logits, mcfg, P) -> (NM, U)
    torch.ops.tapqir_b200.ksmogn_log_prob_backward(W, <the same arguments>) -> 6 gradients

with the second registered as the autograd formula of the first, i.e. one explicit forward kernel
(``tq_ksmogn_fwd``) and one explicit backward kernel (``tq_ksmogn_fwd_bwd``) replacing
``KSMOGN.log_prob`` + autograd of the reference (distributions/ksmogn.py:187-238).  CUDA only: there is no
CPU implementation, calling the op on CPU tensors raises.  ``register_fake`` gives the shapes, so the ops can
be traced (FakeTensor / ``torch.export``) without a device.

Kernel layout (what ``tapqir_b200.distributions.KSMOGN.log_prob`` flattens to): ``U`` patches;
``height, width, x, y`` ``(K, U)``; ``background`` ``(U,)``; ``gain`` ``(1,)``; ``target`` ``(U, 2)`` (x, y);
``value`` ``(U, P, P)`` fp32 (or fp64 for the double kernels); offsets ``(O,)``; ``mcfg`` ``(NM, K)`` the
spot-presence table (the enumerated ``{0,1}^K`` table selects the specialised production kernel).
The SVI step itself does not go through these ops (it calls the fused step kernels, models/engine.py).
"""

from typing import Tuple

import torch
from torch import Tensor

from tapqir_b200 import _lib


def _is_enumerated(mcfg: Tensor) -> bool:
    if mcfg.shape[0] != _lib.M:
        return False
    table = torch.tensor([[(m >> k) & 1 for k in range(_lib.K)] for m in range(_lib.M)], dtype=mcfg.dtype, device=mcfg.device)
    return bool(torch.equal(mcfg, table))


@torch.library.custom_op("tapqir_b200::ksmogn_log_prob", mutates_args=(), device_types="cuda")
def ksmogn_log_prob(height: Tensor, width: Tensor, x: Tensor, y: Tensor, background: Tensor, gain: Tensor,
                    target: Tensor, value: Tensor, offset_samples: Tensor, offset_logits: Tensor, mcfg: Tensor,
                    P: int) -> Tensor:
    U, NM = background.numel(), mcfg.shape[0]
    dtype, dev = background.dtype, background.device
    # contiguous copies (if any) stay referenced until the launch below has been enqueued on this stream
    args = [t.contiguous() for t in (height, width, x, y, background, gain)]
    data = [t.contiguous() for t in (value, target, offset_samples, offset_logits)]
    view = _lib.make_view(*data, nb=U, fb=1, C=1, F=1, P=P)
    logp = torch.empty((NM, U), dtype=dtype, device=dev)
    mtable = mcfg.contiguous()
    table = None if _is_enumerated(mtable) else _lib.ptr(mtable)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.tq_ksmogn_fwd(_lib.dtype_code(dtype), view, *[_lib.ptr(t) for t in args], table, NM, _lib.ptr(logp),
                                     _lib.stream_ptr(dev)), "tq_ksmogn_fwd")
    return logp


@ksmogn_log_prob.register_fake
def _(height, width, x, y, background, gain, target, value, offset_samples, offset_logits, mcfg, P):
    return background.new_empty((mcfg.shape[0], background.numel()))


@torch.library.custom_op("tapqir_b200::ksmogn_log_prob_backward", mutates_args=(), device_types="cuda")
def ksmogn_log_prob_backward(W: Tensor, height: Tensor, width: Tensor, x: Tensor, y: Tensor, background: Tensor,
                             gain: Tensor, target: Tensor, value: Tensor, offset_samples: Tensor, offset_logits: Tensor,
                             mcfg: Tensor, P: int) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Gradients of ``sum(W * log_prob)`` w.r.t. height, width, x, y, background and gain."""
    U, NM = background.numel(), mcfg.shape[0]
    dtype, dev = background.dtype, background.device
    # contiguous copies (if any) stay referenced until the launch below has been enqueued on this stream
    args = [t.contiguous() for t in (height, width, x, y, background, gain)]
    data = [t.contiguous() for t in (value, target, offset_samples, offset_logits)]
    view = _lib.make_view(*data, nb=U, fb=1, C=1, F=1, P=P)
    Wc = W.to(dtype).contiguous()
    g_h, g_w, g_x, g_y = (torch.empty_like(args[0]) for _ in range(4))
    g_b, g_rate = torch.empty_like(args[4]), torch.empty_like(args[4])
    mtable = mcfg.contiguous()
    table = None if _is_enumerated(mtable) else _lib.ptr(mtable)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.tq_ksmogn_fwd_bwd(_lib.dtype_code(dtype), view, *[_lib.ptr(t) for t in args], table, NM, _lib.ptr(Wc),
                                         None, _lib.ptr(g_h), _lib.ptr(g_w), _lib.ptr(g_x), _lib.ptr(g_y), _lib.ptr(g_b),
                                         _lib.ptr(g_rate), _lib.stream_ptr(dev)), "tq_ksmogn_fwd_bwd")
    # the kernel differentiates w.r.t. the rate 1/gain, summed per patch
    g_gain = (-(g_rate.sum()) / (gain * gain)).reshape(gain.shape)
    return g_h, g_w, g_x, g_y, g_b, g_gain


@ksmogn_log_prob_backward.register_fake
def _(W, height, width, x, y, background, gain, target, value, offset_samples, offset_logits, mcfg, P):
    return (torch.empty_like(height), torch.empty_like(width), torch.empty_like(x), torch.empty_like(y),
            torch.empty_like(background), torch.empty_like(gain))


def _setup_context(ctx, inputs, output):
    *tensors, P = inputs
    ctx.save_for_backward(*tensors)
    ctx.P = P


def _backward(ctx, grad_logp):
    grads = ksmogn_log_prob_backward(grad_logp, *ctx.saved_tensors, ctx.P)
    return (*grads, None, None, None, None, None, None)


ksmogn_log_prob.register_autograd(_backward, setup_context=_setup_context)
