"""
CPU: the hmm oracle's chain algebra against brute-force enumeration of every z path (tiny F), and the host-side
layouts of the hmm variant.  (The oracle's ELBO assembly itself is "parity unpinned": pyro / funsor are absent.)
"""

import itertools

import torch

from oracle import cosmos_oracle as O
from oracle import hmm_oracle as H
from tapqir_b200.models import layout as L
from tapqir_b200.utils.simulate import simulate


def small_problem(N=3, F=3, C=1, seed=0):
    ds = simulate(N, F, C=C, seed=seed)
    data = O.OracleData(ds.images, ds.xy, ds.is_ontarget, ds.mask, ds.offset.samples, ds.offset.weights)
    g = torch.Generator().manual_seed(seed + 5)
    params = H.to_unconstrained(H.init_constrained(data), data.P, data.dtype)
    for v in params.values():
        v.add_(0.4 * torch.randn(v.shape, generator=g, dtype=v.dtype))
    ndx = torch.arange(N)
    noise = H.draw_noise(params, data, ndx, g)
    return ds, data, params, ndx, noise


def test_forward_recursion_equals_path_enumeration():
    """sum over all 2^F paths of q(path) [log p(path, emissions) - log q(path)] == the forward-recursion ELBO."""
    ds, data, params, ndx, noise = small_problem(N=2, F=3)
    total, parts = H.elbo(params, data, ndx, noise, return_parts=True)
    p = H.to_constrained(params, data.P, data.dtype)
    zt = p["z_trans"][ndx]                                                    # (nb,F,C,z',z)
    logq = torch.distributions.Categorical(probs=zt, validate_args=False).logits
    ont = data.is_ontarget[ndx].long()
    logp_init = torch.distributions.Categorical(probs=O.expand_offtarget(parts["init"])[:, :, ont].permute(2, 0, 1),
                                                validate_args=False).logits
    logp_trans = torch.distributions.Categorical(probs=O.expand_offtarget(parts["trans"])[:, :, :, ont].permute(3, 0, 1, 2),
                                                 validate_args=False).logits
    em = parts["emission"]                                                    # (z,nb,F,C)
    nb, F, C = len(ndx), data.F, data.C
    brute = torch.zeros(nb, C, dtype=data.dtype)
    for path in itertools.product(range(2), repeat=F):
        lq = torch.zeros(nb, C, dtype=data.dtype)
        lp = torch.zeros(nb, C, dtype=data.dtype)
        prev = 0
        for f, z in enumerate(path):
            lq = lq + logq[:, f, :, prev, z]
            lp = lp + (logp_init[:, :, z] if f == 0 else logp_trans[:, :, prev, z]) + em[z, :, f, :]
            prev = z
        brute = brute + lq.exp() * (lp - lq)
    assert torch.allclose(brute, parts["e_chain"] + parts["e_emit"], rtol=1e-12, atol=1e-9)


def test_z_probs_are_the_forward_marginals():
    ds, data, params, ndx, noise = small_problem(N=2, F=4)
    zp = H.z_probs(params, data)
    assert zp.shape == (2, 4, 1, 2)
    assert torch.allclose(zp.sum(-1), torch.ones(2, 4, 1, dtype=zp.dtype))


def test_hmm_layouts_round_trip():
    ll, gl = L.HmmLocalLayout(3, 4, 2), L.HmmGlobalLayout(2)
    assert gl.numel == 4 + 11 * 2 and gl.noise_numel == 2 + 7 * 2
    g = torch.Generator().manual_seed(0)
    named = {"m_probs": torch.randn(2, 2, 3, 4, 2, generator=g), "z_trans": torch.randn(3, 4, 2, 2, 2, generator=g)}
    flat = torch.zeros(ll.numel)
    for n, s in ll.shapes.items():
        if n not in ("m_probs_z0", "m_probs_z1", "z_trans"):
            named[n] = torch.randn(s, generator=g)
    ll.load_named(flat, named)
    back = ll.named(flat)
    for n, t in named.items():
        assert torch.equal(back[n], t), n
    # z_trans sits behind the cosmos layout and m_probs[z = 1]
    assert ll.offsets["z_trans"] == ll.std_numel + 2 * 3 * 4 * 2
