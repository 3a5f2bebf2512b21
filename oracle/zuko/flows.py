"""zuko.flows.{Flow, MAF, NSF} restated (constructor signatures used at pocomc/flow.py:55-86)."""
import math
import torch
import torch.nn as nn

from .nn import MaskedMLP
from .transforms import (AutoregressiveTransform, ComposedTransform, DependentTransform,
                         MonotonicAffineTransform, MonotonicRQSTransform)
from .distributions import DiagNormal, NormalizingFlow


class MaskedAutoregressiveTransform(nn.Module):
    def __init__(self, features, context=0, passes=None, order=None,
                 univariate=MonotonicAffineTransform, shapes=((), ()), **kwargs):
        super().__init__()
        if context:
            raise NotImplementedError("pocoMC never conditions the flow (context=0)")
        self.univariate = univariate
        self.shapes = [tuple(s) for s in shapes]
        self.sizes = [int(math.prod(s)) for s in self.shapes]
        self.total = sum(self.sizes)
        self.passes = features if passes is None else min(max(passes, 1), features)
        if order is None:
            order = torch.arange(features)
        order = torch.as_tensor(order)
        self.register_buffer("order", torch.div(order, math.ceil(features / self.passes),
                                                rounding_mode="floor"))
        adjacency = self.order[:, None] > self.order
        adjacency = torch.repeat_interleave(adjacency, repeats=self.total, dim=0)
        self.hyper = MaskedMLP(adjacency, **kwargs)

    def meta(self, x):
        phi = self.hyper(x)
        phi = phi.unflatten(-1, (-1, self.total))
        phi = phi.split(self.sizes, -1)
        phi = [p.unflatten(-1, s + (1,)).squeeze(-1) for p, s in zip(phi, self.shapes)]
        return DependentTransform(self.univariate(*phi), 1)

    def forward(self, c=None):
        return AutoregressiveTransform(self.meta, self.passes)


class Flow(nn.Module):
    """Lazy flow: calling the module yields a NormalizingFlow(transform, base)."""

    def __init__(self, transform, base):
        super().__init__()
        self.transform = nn.ModuleList(transform) if isinstance(transform, (list, tuple)) else transform
        self.base = base

    def forward(self, c=None):
        ts = self.transform if isinstance(self.transform, nn.ModuleList) else [self.transform]
        transform = ComposedTransform(*[t(c) for t in ts])
        return NormalizingFlow(transform, self.base(c))


class _UnconditionalDiagNormal(nn.Module):
    def __init__(self, features):
        super().__init__()
        self.register_buffer("loc", torch.zeros(features))
        self.register_buffer("scale", torch.ones(features))

    def forward(self, c=None):
        return DiagNormal(self.loc, self.scale)


class MAF(Flow):
    def __init__(self, features, context=0, transforms=3, randperm=False, **kwargs):
        orders = [torch.arange(features), torch.flipud(torch.arange(features))]
        ts = [MaskedAutoregressiveTransform(features=features, context=context,
                                            order=torch.randperm(features) if randperm else orders[i % 2],
                                            **kwargs)
              for i in range(transforms)]
        super().__init__(ts, _UnconditionalDiagNormal(features))


class NSF(MAF):
    def __init__(self, features, context=0, bins=8, **kwargs):
        super().__init__(features=features, context=context, univariate=MonotonicRQSTransform,
                         shapes=[(bins,), (bins,), (bins - 1,)], **kwargs)
