"""CPU: the C-ABI library builds, loads and exports every symbol include/pmc_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "pmc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pocomc_b200 import _build, _lib
    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pmc_b200.h but not exported"
    # the ctypes table binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == names
    assert _lib.load().pmc_version() >= 100


def test_no_cuda_means_loud_failure():
    import torch
    import pytest
    import numpy as np
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pocomc_b200.flow import Flow
    f = Flow(4, "maf3")          # construction is host-only
    with pytest.raises(RuntimeError, match="no CUDA device"):
        f.forward(torch.zeros(3, 4))
    from pocomc_b200.scaler import Reparameterize
    s = Reparameterize(2, bounds=np.array([[0., 1.], [-np.inf, np.inf]]))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        s.fit(np.full((4, 2), 0.5))
