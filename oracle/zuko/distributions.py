import math
import torch


class DiagNormal:
    """Independent(Normal(loc, scale), 1)."""

    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def log_prob(self, z):
        v = (z - self.loc) / self.scale
        lp = -0.5 * v ** 2 - torch.log(self.scale) - 0.5 * math.log(2 * math.pi)
        return lp.sum(dim=-1)

    def rsample(self, shape=()):
        shape = tuple(shape) + tuple(self.loc.shape)
        eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        return self.loc + eps * self.scale


class NormalizingFlow:
    def __init__(self, transform, base):
        self.transform, self.base = transform, base

    def log_prob(self, x):
        z, ladj = self.transform.call_and_ladj(x)
        return self.base.log_prob(z) + ladj

    def rsample(self, shape=()):
        return self.transform.inv(self.base.rsample(shape))

    def rsample_and_log_prob(self, shape=()):
        z = self.base.rsample(shape)
        x, ladj = self.transform.inv.call_and_ladj(z)
        return x, self.base.log_prob(z) - ladj
