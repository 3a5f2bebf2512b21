"""Host-side argument checks (pocomc/input_validation.py): shape/dtype/interval -> ValueError."""
import numpy as np


def assert_array_2d(x: np.ndarray):
    if len(x.shape) != 2:
        raise ValueError(f"Input should have 2 dimensions, but got {len(x.shape)}")


def assert_array_1d(x: np.ndarray):
    if len(x.shape) != 1:
        raise ValueError(f"Input should have 1 dimension, but got {len(x.shape)}")


def assert_array_float(x: np.ndarray):
    if not np.issubdtype(x.dtype, np.floating):
        raise ValueError(f"Expected input to have dtype float, but got {x.dtype}")


def assert_array_within_interval(x: np.ndarray, left: np.ndarray, right: np.ndarray):
    """Closed-interval check; NaN bounds mean unbounded (input_validation.py:25-52)."""
    lo = np.where(np.isnan(left), -np.inf, left)
    hi = np.where(np.isnan(right), np.inf, right)
    if not np.all((lo <= x) & (x <= hi)):
        raise ValueError(f"Expected input to be within interval [{lo}, {hi}], "
                         f"but got minimum = {np.min(x)} and maximum = {np.max(x)}")
