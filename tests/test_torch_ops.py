"""
The ``torch.library`` registration of the likelihood operator (tapqir_b200/ops.py): schema, shape propagation
without a device, loud failure on CPU tensors (CPU part); on the GPU the op and its registered backward against
the golden vectors of the reference's KSMOGN.log_prob (tests/golden/ref_distributions.pt).
"""

import pytest
import torch

import tapqir_b200.ops  # noqa: F401  (registers torch.ops.tapqir_b200.*)

K, M = 2, 4


def _flat_case(case, dtype, device):
    """Golden KSMOGN case -> the kernel layout of the op (U patches)."""
    c = case["inputs"]
    core = tuple(c["background"].shape)
    U = int(torch.Size(core).numel())
    spot = lambda t: t.to(dtype).reshape(U, K).t().contiguous().to(device)
    args = dict(height=spot(c["height"]), width=spot(c["width"]), x=spot(c["x"]), y=spot(c["y"]),
                background=c["background"].to(dtype).reshape(U).contiguous().to(device),
                gain=c["gain"].to(dtype).reshape(1).to(device),
                target=c["target_locs"].to(dtype).reshape(U, 2).contiguous().to(device),
                value=c["value"].to(torch.float32 if dtype == torch.float32 else dtype).reshape(U, c["P"], c["P"]).contiguous().to(device),
                offset_samples=c["offset_samples"].to(dtype).to(device),
                offset_logits=torch.distributions.utils.probs_to_logits(c["offset_weights"]).to(dtype).to(device),
                mcfg=case["m"].to(dtype).to(device))
    return args, c["P"], core, U


def test_ops_are_registered_with_the_expected_schema():
    fwd = torch.ops.tapqir_b200.ksmogn_log_prob.default
    bwd = torch.ops.tapqir_b200.ksmogn_log_prob_backward.default
    assert str(fwd._schema).startswith("tapqir_b200::ksmogn_log_prob(Tensor height, Tensor width, Tensor x, Tensor y, "
                                       "Tensor background, Tensor gain, Tensor target, Tensor value, Tensor offset_samples, "
                                       "Tensor offset_logits, Tensor mcfg, SymInt P) -> Tensor")
    assert len(bwd._schema.returns) == 6 and len(bwd._schema.arguments) == 13
    assert not any(a.alias_info is not None for a in fwd._schema.arguments)      # functional: nothing mutated


def test_shapes_propagate_without_a_device(golden):
    from torch._subclasses.fake_tensor import FakeTensorMode

    args, P, core, U = _flat_case(golden["ksmogn"]["sim_O3"], torch.float32, "cpu")
    with FakeTensorMode() as mode:
        fake = {k: mode.from_tensor(v) for k, v in args.items()}
        out = torch.ops.tapqir_b200.ksmogn_log_prob(*fake.values(), P)
        assert tuple(out.shape) == (M, U) and out.dtype == torch.float32
        grads = torch.ops.tapqir_b200.ksmogn_log_prob_backward(out, *fake.values(), P)
        assert [tuple(g.shape) for g in grads] == [(K, U)] * 4 + [(U,), (1,)]


def test_cpu_tensors_fail_loudly(golden):
    """No CPU implementation is registered: the dispatcher refuses CPU tensors instead of falling back."""
    args, P, _, _ = _flat_case(golden["ksmogn"]["sim_O3"], torch.float32, "cpu")
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.tapqir_b200.ksmogn_log_prob(*args.values(), P)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sim_O3", "hist_O16_C2"])
@pytest.mark.parametrize("dtype,tol,gtol", [(torch.float64, 1e-12, 1e-10), (torch.float32, 1e-5, 1e-5)])
def test_op_and_registered_backward_match_reference(golden, name, dtype, tol, gtol):
    case = golden["ksmogn"][name]
    args, P, core, U = _flat_case(case, dtype, "cuda")
    leaves = ("height", "width", "x", "y", "background", "gain")
    for k in leaves:
        args[k].requires_grad_(True)
    logp = torch.ops.tapqir_b200.ksmogn_log_prob(*args.values(), P)
    ref = case["log_prob"].reshape(M, U)
    assert (logp.detach().double().cpu() - ref).abs().max().item() <= tol * ref.abs().max().item()
    W = case["W"].reshape(M, U).to(dtype).cuda()
    (W * logp).sum().backward()
    for k in leaves:
        g = args[k].grad.double().cpu()
        r = case["grads"][k]
        r = r.reshape(U, K).t() if k in ("height", "width", "x", "y") else r.reshape(g.shape)
        assert (g - r).abs().max().item() <= gtol * r.abs().max().item(), k
