"""
Beta distribution on ``[low, high]`` in the mean / sample-size parameterisation of the reference
(tapqir/distributions/affine_beta.py:10-59, built there on pyro.distributions.AffineBeta).

Host-side helper for post-fit statistics and tests; inside the SVI step the same density, sampler
and reparameterisation gradient are evaluated in-kernel (csrc/cosmos_local.cuh).
"""

import math

import torch
from torch.distributions import Beta, constraints
from torch.distributions.transformed_distribution import TransformedDistribution
from torch.distributions.transforms import AffineTransform


class AffineBeta(TransformedDistribution):
    arg_constraints = {
        "mean": constraints.dependent,
        "sample_size": constraints.real,
        "low": constraints.real,
        "high": constraints.dependent,
    }

    def __init__(self, mean, sample_size, low, high, validate_args=None):
        mean, sample_size = torch.as_tensor(mean), torch.as_tensor(sample_size)
        width = high - low
        c1 = sample_size * (mean - low) / width
        c0 = sample_size * (high - mean) / width
        self.low_, self.high_ = low, high
        super().__init__(Beta(c1, c0, validate_args=validate_args), AffineTransform(loc=low, scale=width),
                         validate_args=validate_args)

    @property
    def concentration1(self):
        return self.base_dist.concentration1

    @property
    def concentration0(self):
        return self.base_dist.concentration0

    @property
    def low(self):
        return self.low_

    @property
    def high(self):
        return self.high_

    @property
    def scale(self):
        return self.high_ - self.low_

    @property
    def mean(self):
        return self.low_ + self.scale * self.base_dist.mean

    @property
    def variance(self):
        return self.scale**2 * self.base_dist.variance

    def _clamp(self, x):
        eps = torch.finfo(x.dtype).eps * self.scale
        return x.clamp(min=self.low_ + eps, max=self.high_ - eps)

    def sample(self, sample_shape=torch.Size()):
        with torch.no_grad():
            return self._clamp(super().sample(sample_shape))

    def rsample(self, sample_shape=torch.Size()):
        return self._clamp(super().rsample(sample_shape))
