"""vivim_b200 -- B200 (sm_100a) kernels for Vivim's Temporal-Mamba hot path.

Public surface (drop-in for the reference's packages, same names and signatures):

* ``causal_conv1d``  -> :mod:`vivim_b200.causal_conv1d_interface`
* ``mamba_ssm``      -> :mod:`vivim_b200.selective_scan_interface`, :mod:`vivim_b200.mamba_simple`

The repo root also carries thin ``causal_conv1d/`` and ``mamba_ssm/`` packages that re-export these,
so ``from mamba_ssm import Mamba`` (modeling/vivim.py:19 of the reference) resolves here unchanged.
"""
__version__ = "0.1.0"
