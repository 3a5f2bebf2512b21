"""TEST INFRASTRUCTURE -- minimal CPU restatement of the parts of `zuko` that pocoMC touches.

pocoMC delegates every piece of flow arithmetic to the third-party package zuko
(`zuko>=1.1.0`, reference requirements.txt:3; call sites pocomc/flow.py:55-86,97,114,131,147,162).
zuko is not vendored in /root/reference and is not installable here (no network), so this
package restates its published algorithm (MAF / NSF over a MaskedMLP hyper-network) in plain
CPU torch, importable as ``zuko`` when ``oracle/`` is on ``sys.path``.  That lets the
reference's own modules run unmodified to generate golden vectors (oracle/make_golden.py).

PARITY STATUS: *parity unpinned* with respect to real zuko -- the reference ships no numeric
golden vectors at this seam (only properties: round trip, ladj_fwd == -ladj_inv, dtype rules;
reference tests/test_flow.py:75-88,153-166).  Flow parity in this repo is defined against THIS
restatement plus those properties.  The one structural guess (placement of the residual
connection in MaskedMLP) is isolated in ``zuko.nn.MaskedMLP``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this.
"""
from . import nn, transforms, distributions, flows  # noqa: F401

__version__ = "1.1.0+oracle"
