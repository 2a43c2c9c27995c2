"""Re-export of :mod:`vivim_b200.causal_conv1d_interface` under the reference's module path."""
from vivim_b200.causal_conv1d_interface import (  # noqa: F401
    CausalConv1dFn,
    causal_conv1d_fn,
    causal_conv1d_ref,
    causal_conv1d_update,
    causal_conv1d_update_ref,
)
