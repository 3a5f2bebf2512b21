"""Univariate monotone transforms + autoregressive / composed wrappers."""
import math
import torch
import torch.nn.functional as F


class Transform:
    """Minimal bijection protocol: __call__, inv, log_abs_det_jacobian, call_and_ladj."""

    def __call__(self, x):
        return self._call(x)

    @property
    def inv(self):
        return _Inverse(self)

    def call_and_ladj(self, x):
        y = self._call(x)
        return y, self.log_abs_det_jacobian(x, y)


class _Inverse(Transform):
    def __init__(self, t):
        self._t = t

    def _call(self, y):
        return self._t._inverse(y)

    def _inverse(self, x):
        return self._t._call(x)

    @property
    def inv(self):
        return self._t

    def log_abs_det_jacobian(self, y, x):
        return -self._t.log_abs_det_jacobian(x, y)

    def call_and_ladj(self, y):
        # inverse first (D hyper-network passes for an autoregressive transform),
        # then ONE more pass for the log-determinant.
        x = self._t._inverse(y)
        return x, -self._t.log_abs_det_jacobian(x, y)


class MonotonicAffineTransform(Transform):
    """y = x * exp(a) + b with a soft-clipped to (-|log slope|, |log slope|)."""

    def __init__(self, shift, scale, slope=1e-3):
        self.shift = shift
        self.log_scale = scale / (1 + abs(scale / math.log(slope)))
        self.scale = self.log_scale.exp()

    def _call(self, x):
        return x * self.scale + self.shift

    def _inverse(self, y):
        return (y - self.shift) / self.scale

    def log_abs_det_jacobian(self, x, y):
        return self.log_scale.expand(x.shape)


class MonotonicRQSTransform(Transform):
    """Monotone rational-quadratic spline on [-bound, bound], identity outside."""

    def __init__(self, widths, heights, derivatives, bound=5.0, slope=1e-3):
        ls = math.log(slope)
        widths = widths / (1 + abs(2 * widths / ls))
        heights = heights / (1 + abs(2 * heights / ls))
        derivatives = derivatives / (1 + abs(derivatives / ls))
        widths = F.pad(F.softmax(widths, dim=-1), (1, 0), value=0)
        heights = F.pad(F.softmax(heights, dim=-1), (1, 0), value=0)
        derivatives = F.pad(derivatives, (1, 1), value=0)
        self.horizontal = bound * (2 * torch.cumsum(widths, dim=-1) - 1)
        self.vertical = bound * (2 * torch.cumsum(heights, dim=-1) - 1)
        self.derivatives = torch.exp(derivatives)
        self.bins = self.derivatives.shape[-1] - 1

    def _bin(self, k):
        mask = torch.logical_and(0 <= k, k < self.bins)
        k = k % self.bins
        k01 = torch.stack((k, k + 1), dim=-1)
        x0, x1 = torch.gather(self.horizontal, -1, k01).unbind(-1)
        y0, y1 = torch.gather(self.vertical, -1, k01).unbind(-1)
        d0, d1 = torch.gather(self.derivatives, -1, k01).unbind(-1)
        s = (y1 - y0) / (x1 - x0)
        return mask, x0, x1, y0, y1, d0, d1, s

    @staticmethod
    def _searchsorted(seq, value):
        return torch.searchsorted(seq, value[..., None]).squeeze(-1)

    def _call(self, x):
        k = self._searchsorted(self.horizontal, x) - 1
        mask, x0, x1, y0, y1, d0, d1, s = self._bin(k)
        z = mask * (x - x0) / (x1 - x0)
        y = y0 + (y1 - y0) * (s * z ** 2 + d0 * z * (1 - z)) / (s + (d0 + d1 - 2 * s) * z * (1 - z))
        return torch.where(mask, y, x)

    def _inverse(self, y):
        k = self._searchsorted(self.vertical, y) - 1
        mask, x0, x1, y0, y1, d0, d1, s = self._bin(k)
        y_ = mask * (y - y0)
        a = (y1 - y0) * (s - d0) + y_ * (d0 + d1 - 2 * s)
        b = (y1 - y0) * d0 - y_ * (d0 + d1 - 2 * s)
        c = -s * y_
        z = 2 * c / (-b - torch.sqrt(b ** 2 - 4 * a * c))
        x = x0 + z * (x1 - x0)
        return torch.where(mask, x, y)

    def log_abs_det_jacobian(self, x, y):
        k = self._searchsorted(self.horizontal, x) - 1
        mask, x0, x1, y0, y1, d0, d1, s = self._bin(k)
        z = mask * (x - x0) / (x1 - x0)
        jac = (s ** 2 * (2 * s * z * (1 - z) + d0 * (1 - z) ** 2 + d1 * z ** 2)
               / (s + (d0 + d1 - 2 * s) * z * (1 - z)) ** 2)
        return torch.log(jac) * mask


class DependentTransform(Transform):
    """Sum the elementwise ladj over the last ``reinterpreted`` dims."""

    def __init__(self, base, reinterpreted=1):
        self.base = base
        self.reinterpreted = reinterpreted

    def _call(self, x):
        return self.base._call(x)

    def _inverse(self, y):
        return self.base._inverse(y)

    def log_abs_det_jacobian(self, x, y):
        ladj = self.base.log_abs_det_jacobian(x, y)
        for _ in range(self.reinterpreted):
            ladj = ladj.sum(dim=-1)
        return ladj


class AutoregressiveTransform(Transform):
    """y = meta(x)(x); inverse by ``passes`` fixed-point sweeps starting from zeros."""

    def __init__(self, meta, passes):
        self.meta = meta
        self.passes = passes

    def _call(self, x):
        return self.meta(x)._call(x)

    def _inverse(self, y):
        x = torch.zeros_like(y)
        for _ in range(self.passes):
            x = self.meta(x)._inverse(y)
        return x

    def log_abs_det_jacobian(self, x, y):
        return self.meta(x).log_abs_det_jacobian(x, y)

    def call_and_ladj(self, x):
        t = self.meta(x)
        y = t._call(x)
        return y, t.log_abs_det_jacobian(x, y)


class ComposedTransform(Transform):
    def __init__(self, *transforms):
        self.transforms = transforms

    def _call(self, x):
        for t in self.transforms:
            x = t._call(x)
        return x

    def _inverse(self, y):
        for t in reversed(self.transforms):
            y = t._inverse(y)
        return y

    def log_abs_det_jacobian(self, x, y):
        _, ladj = self.call_and_ladj(x)
        return ladj

    def call_and_ladj(self, x):
        total = 0
        for t in self.transforms:
            x, ladj = t.call_and_ladj(x)
            total = total + ladj
        return x, total

    @property
    def inv(self):
        return ComposedTransform(*[t.inv for t in reversed(self.transforms)])
