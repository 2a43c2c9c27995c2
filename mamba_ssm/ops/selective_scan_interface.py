"""Re-export of :mod:`vivim_b200.selective_scan_interface` under the reference's module path."""
from vivim_b200.selective_scan_interface import (  # noqa: F401
    SelectiveScanFn,
    bimamba_inner_fn,
    bimamba_inner_ref,
    mamba_inner_fn,
    mamba_inner_fn_no_out_proj,
    mamba_inner_ref,
    selective_scan_fn,
    selective_scan_ref,
)
