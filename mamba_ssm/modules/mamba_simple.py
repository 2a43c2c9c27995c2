"""Re-export of :mod:`vivim_b200.mamba_simple` under the reference's module path."""
from vivim_b200.mamba_simple import Mamba  # noqa: F401
