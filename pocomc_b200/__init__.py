"""pocomc_b200 -- B200 (sm_100a) implementation of pocoMC's data-parallel hot path behind the
reference's own Python API (``pc.Flow``, ``pc.Sampler``, ``pc.Prior``; pocomc/__init__.py:27-33)."""
__url__ = "https://pocomc.readthedocs.io"
__license__ = "GPL-3.0"
__description__ = "Preconditioned Monte Carlo: flow preconditioner + SMC/MCMC inner loop on B200"

from ._version import version
from .flow import *          # noqa: F401,F403
from .prior import *         # noqa: F401,F403
from .sampler import *       # noqa: F401,F403
from . import config, mcmc, particles, scaler, tools, geometry  # noqa: F401

__version__ = version
