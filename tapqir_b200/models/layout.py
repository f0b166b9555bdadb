"""
Memory layout of the variational parameters in HBM and the small host structs the kernels take.

The reference keeps 20 separate ``pyro.param`` tensors (models/cosmos.py:471-598).  Here the 12
AOI-local tensors live back to back in ONE flat buffer (same element order as the reference shapes,
so each name is a plain view) and the 8 global tensors in a second, tiny one; gradients and the two
Adam moments use identical buffers.  One dense Adam launch then covers everything, and a shard of
AOIs is a self-contained buffer that never leaves its GPU.

    local flat buffer (Nt = AOIs held by this rank):
        background_mean_loc (Nt,1,C) | background_std_loc (Nt,1,C) | b_loc (Nt,F,C) | b_beta (Nt,F,C)
        | m_probs | h_loc | h_beta | w_mean | w_size | x_mean | y_mean | size     each (K,Nt,F,Q)
    global flat buffer:
        gain_loc | gain_beta | proximity_loc | proximity_size | pi_mean (Q,2) | pi_size (Q,1)
        | lamda_loc (Q,) | lamda_beta (Q,)

Must stay in sync with csrc/cosmos_local.cuh (LP_* enum), csrc/cosmos_globals.cuh (GlobalLayout)
and csrc/cosmos_step.cu (LocalOffsets).
"""

import ctypes
import math
from collections import OrderedDict

import torch

K = 2
S = 1
NSAMP = 1 + 4 * K       # guide samples per unit: background + (height, width, x, y) per spot
NACC = 18               # per-channel accumulators (csrc/cosmos_local.cuh ACC_*)
SAMPLE_SITES = ["background", "height", "width", "x", "y"]

LOCAL_NAMES = ["background_mean_loc", "background_std_loc", "b_loc", "b_beta", "m_probs", "h_loc", "h_beta",
               "w_mean", "w_size", "x_mean", "y_mean", "size"]
GLOBAL_NAMES = ["gain_loc", "gain_beta", "proximity_loc", "proximity_size", "pi_mean", "pi_size", "lamda_loc",
                "lamda_beta"]
PARAM_NAMES = ["pi_mean", "pi_size", "m_probs", "proximity_loc", "proximity_size", "lamda_loc", "lamda_beta",
               "gain_loc", "gain_beta", "background_mean_loc", "background_std_loc", "b_loc", "b_beta", "h_loc",
               "h_beta", "w_mean", "w_size", "x_mean", "y_mean", "size"]  # reference creation order


class ModelConst(ctypes.Structure):
    """``tq::ModelConst``: prior hyper-parameters (cosmos.py:55-64) + eps/tiny of the reference dtype."""

    _fields_ = [(n, ctypes.c_double) for n in
                ("bg_mean_std", "bg_std_std", "lamda_rate", "height_std", "width_min", "width_max",
                 "proximity_rate", "gain_std", "eps", "tiny", "logit_lim")] + [("P", ctypes.c_int)]

    @classmethod
    def make(cls, priors, P, ref_dtype=torch.float64):
        fi = torch.finfo(ref_dtype)
        return cls(priors["background_mean_std"], priors["background_std_std"], priors["lamda_rate"],
                   priors["height_std"], priors["width_min"], priors["width_max"], priors["proximity_rate"],
                   priors["gain_std"], fi.eps, fi.tiny, math.log((1.0 - fi.eps) / fi.eps), int(P))


class LocalLayout:
    def __init__(self, Nt, F, C, Kk=K):
        assert Kk == K
        self.Nt, self.F, self.C = int(Nt), int(F), int(C)
        aoi, unit = self.Nt * self.C, self.Nt * self.F * self.C
        self.shapes = OrderedDict()
        self.offsets = OrderedDict()
        off = 0
        for name in LOCAL_NAMES:
            if name in ("background_mean_loc", "background_std_loc"):
                shape = (self.Nt, 1, self.C)
            elif name in ("b_loc", "b_beta"):
                shape = (self.Nt, self.F, self.C)
            else:
                shape = (K, self.Nt, self.F, self.C)
            self.shapes[name] = shape
            self.offsets[name] = off
            off += math.prod(shape)
        self.numel = off
        assert off == 2 * aoi + 2 * unit + 8 * K * unit

    def views(self, flat):
        """name -> view of ``flat`` with the reference shape."""
        assert flat.numel() == self.numel and flat.is_contiguous()
        return OrderedDict((n, flat[self.offsets[n]: self.offsets[n] + math.prod(s)].view(s))
                           for n, s in self.shapes.items())

    def pack(self, tensors, dtype=None, device=None):
        first = next(iter(tensors.values()))
        flat = torch.empty(self.numel, dtype=dtype or first.dtype, device=device or first.device)
        for n, v in self.views(flat).items():
            v.copy_(tensors[n].reshape(v.shape))
        return flat


class GlobalLayout:
    def __init__(self, Q):
        self.Q = int(Q)
        self.shapes = OrderedDict([
            ("gain_loc", ()), ("gain_beta", ()), ("proximity_loc", ()), ("proximity_size", ()),
            ("pi_mean", (self.Q, S + 1)), ("pi_size", (self.Q, 1)), ("lamda_loc", (self.Q,)), ("lamda_beta", (self.Q,)),
        ])
        self.offsets = OrderedDict()
        off = 0
        for n, s in self.shapes.items():
            self.offsets[n] = off
            off += math.prod(s)
        self.numel = off  # 4 + 5Q
        # base variates / samples of the global sites: gain, proximity, pi (Q,2), lamda (Q,)
        self.noise_shapes = OrderedDict([("gain", ()), ("proximity", ()), ("pi", (self.Q, S + 1)), ("lamda", (self.Q,))])
        self.noise_offsets = OrderedDict()
        off = 0
        for n, s in self.noise_shapes.items():
            self.noise_offsets[n] = off
            off += math.prod(s)
        self.noise_numel = off

    def views(self, flat):
        assert flat.numel() == self.numel
        return OrderedDict((n, flat[self.offsets[n]: self.offsets[n] + math.prod(s)].view(s))
                           for n, s in self.shapes.items())

    def pack(self, tensors, dtype=None, device=None):
        first = next(iter(tensors.values()))
        flat = torch.empty(self.numel, dtype=dtype or first.dtype, device=device or first.device)
        for n, v in self.views(flat).items():
            v.copy_(tensors[n].reshape(v.shape))
        return flat

    def noise_views(self, flat):
        return OrderedDict((n, flat[self.noise_offsets[n]: self.noise_offsets[n] + math.prod(s)].view(s))
                           for n, s in self.noise_shapes.items())

    def pack_noise(self, noise, dtype=torch.float64, device="cpu"):
        flat = torch.empty(self.noise_numel, dtype=dtype, device=device)
        for n, v in self.noise_views(flat).items():
            v.copy_(noise[n].reshape(v.shape))
        return flat


def pack_local_noise(noise, dtype, device):
    """Guide base variates dict (oracle.draw_noise keys) -> (NSAMP, U) record in kernel order:
    background, height_k.., width_k.., x_k.., y_k..; each (nb, fb, C) flattened."""
    rows = [noise["background"].reshape(1, -1)]
    for site in ("height", "width", "x", "y"):
        rows.append(noise[site].reshape(K, -1))
    return torch.cat(rows, 0).to(dtype=dtype, device=device).contiguous()


# ---- hmm variant (models/hmm.py) -------------------------------------------------------------------------------------
class HmmGlobalLayout(GlobalLayout):
    """cosmos global layout with ``pi_*`` read as ``init_*``, then ``trans_mean (Q,2,2)``, ``trans_size (Q,2,1)``."""

    def __init__(self, Q):
        super().__init__(Q)
        shapes = OrderedDict()
        for n, s in self.shapes.items():
            shapes[{"pi_mean": "init_mean", "pi_size": "init_size"}.get(n, n)] = s
        shapes["trans_mean"] = (self.Q, S + 1, S + 1)
        shapes["trans_size"] = (self.Q, S + 1, 1)
        self.shapes = shapes
        self.offsets, off = OrderedDict(), 0
        for n, s in shapes.items():
            self.offsets[n] = off
            off += math.prod(s)
        self.numel = off  # 4 + 11Q
        self.noise_shapes = OrderedDict([("gain", ()), ("proximity", ()), ("init", (self.Q, S + 1)), ("lamda", (self.Q,)),
                                         ("trans", (self.Q, S + 1, S + 1))])
        self.noise_offsets, off = OrderedDict(), 0
        for n, s in self.noise_shapes.items():
            self.noise_offsets[n] = off
            off += math.prod(s)
        self.noise_numel = off


class HmmLocalLayout(LocalLayout):
    """
    Flat local buffer of the hmm variant: the cosmos layout (whose ``m_probs`` slabs hold ``m_probs[z = 0]``), then
    ``m_probs[z = 1]`` (K,Nt,F,C), then ``z_trans`` (Nt,F,C,2,2).  ``views`` exposes the raw pieces
    (``m_probs_z0``, ``m_probs_z1``); :meth:`named` / :meth:`load_named` convert from / to the reference's
    ``m_probs (1+S,K,Nt,F,C)``.
    """

    def __init__(self, Nt, F, C):
        super().__init__(Nt, F, C)
        self.std_numel = self.numel
        unit = self.Nt * self.F * self.C
        shapes, offsets = OrderedDict(), OrderedDict()
        for n in self.shapes:
            shapes["m_probs_z0" if n == "m_probs" else n] = self.shapes[n]
            offsets["m_probs_z0" if n == "m_probs" else n] = self.offsets[n]
        shapes["m_probs_z1"], offsets["m_probs_z1"] = (K, self.Nt, self.F, self.C), self.std_numel
        shapes["z_trans"], offsets["z_trans"] = (self.Nt, self.F, self.C, S + 1, S + 1), self.std_numel + K * unit
        self.shapes, self.offsets = shapes, offsets
        self.numel = self.std_numel + K * unit + unit * (S + 1) ** 2

    def named(self, flat):
        """name -> tensor with the reference's names and shapes (``m_probs`` is a stacked copy)."""
        v = self.views(flat)
        out = OrderedDict((n, t) for n, t in v.items() if not n.startswith("m_probs_z"))
        out["m_probs"] = torch.stack([v["m_probs_z0"], v["m_probs_z1"]], 0)
        return out

    def load_named(self, flat, tensors):
        v = self.views(flat)
        for n, t in v.items():
            if n.startswith("m_probs_z"):
                t.copy_(tensors["m_probs"][int(n[-1])].to(device=flat.device, dtype=flat.dtype).reshape(t.shape))
            else:
                t.copy_(tensors[n].to(device=flat.device, dtype=flat.dtype).reshape(t.shape))
