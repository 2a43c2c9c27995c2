"""Drop-in ``mamba_ssm`` package backed by vivim_b200 (reference: mamba/mamba_ssm/__init__.py).
``MambaLMHeadModel`` (language-model wrapper) is out of scope and not exported."""
__version__ = "1.0.1"

from mamba_ssm.ops.selective_scan_interface import (  # noqa: F401
    bimamba_inner_fn,
    mamba_inner_fn,
    mamba_inner_fn_no_out_proj,
    selective_scan_fn,
)
from mamba_ssm.modules.mamba_simple import Mamba  # noqa: F401
