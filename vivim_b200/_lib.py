"""ctypes binding of libvivim_b200.so (include/vivim_b200.h).  No fallback: if the library is
missing or a kernel call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libvivim_b200.so")

VV_F32, VV_F16, VV_BF16 = 0, 1, 2
VV_DIR_FWD, VV_DIR_REV, VV_DIR_FRAMES = 0, 1, 2
VV_MAX_DIRS = 4
VV_SCAN_SEGMENT = 64


class ConvArgs(Structure):
    """vv_conv1d_args"""
    _fields_ = [
        ("x", c_void_p), ("weight", c_void_p), ("bias", c_void_p), ("out", c_void_p),
        ("dout", c_void_p), ("dx", c_void_p), ("dweight", c_void_p), ("dbias", c_void_p),
        ("batch", c_int32), ("dim", c_int32), ("seqlen", c_int32), ("width", c_int32),
        ("x_bs", c_int64), ("x_ds", c_int64), ("out_bs", c_int64), ("out_ds", c_int64),
        ("dout_bs", c_int64), ("dout_ds", c_int64), ("dx_bs", c_int64), ("dx_ds", c_int64),
        ("io_dtype", c_int32), ("w_dtype", c_int32), ("silu", c_int32),
    ]


class ConvDirsArgs(Structure):
    """vv_conv1d_dirs_args"""
    _fields_ = [
        ("x", c_void_p), ("weight", c_void_p), ("bias", c_void_p), ("out", c_void_p),
        ("dout", c_void_p), ("dx", c_void_p), ("dweight", c_void_p), ("dbias", c_void_p),
        ("batch", c_int32), ("dim", c_int32), ("seqlen", c_int32), ("width", c_int32),
        ("ndirs", c_int32), ("dir_mode", c_int32 * VV_MAX_DIRS), ("nframes", c_int32),
        ("x_bs", c_int64), ("x_ds", c_int64), ("out_bs", c_int64), ("out_ds", c_int64),
        ("dout_bs", c_int64), ("dout_ds", c_int64), ("dx_bs", c_int64), ("dx_ds", c_int64),
        ("io_dtype", c_int32), ("silu", c_int32),
    ]


class ScanArgs(Structure):
    """vv_scan_args"""
    _fields_ = [
        ("u", c_void_p), ("delta", c_void_p), ("A", c_void_p), ("Bm", c_void_p), ("Cm", c_void_p),
        ("D", c_void_p), ("z", c_void_p), ("delta_bias", c_void_p),
        ("out", c_void_p), ("out_z", c_void_p), ("last_state", c_void_p),
        ("agg", c_void_p), ("chk", c_void_p), ("radj", c_void_p),
        ("dout", c_void_p), ("du", c_void_p), ("ddelta", c_void_p), ("dz", c_void_p),
        ("dA", c_void_p), ("dB", c_void_p), ("dC", c_void_p), ("dD", c_void_p),
        ("ddelta_bias", c_void_p),
        ("batch", c_int32), ("dim", c_int32), ("seqlen", c_int32), ("dstate", c_int32),
        ("ngroups", c_int32),
        ("u_bs", c_int64), ("u_ds", c_int64), ("delta_bs", c_int64), ("delta_ds", c_int64),
        ("z_bs", c_int64), ("z_ds", c_int64), ("out_bs", c_int64), ("out_ds", c_int64),
        ("outz_bs", c_int64), ("outz_ds", c_int64),
        ("A_ds", c_int64), ("A_ns", c_int64),
        ("B_bs", c_int64), ("B_gs", c_int64), ("B_ns", c_int64),
        ("C_bs", c_int64), ("C_gs", c_int64), ("C_ns", c_int64),
        ("dout_bs", c_int64), ("dout_ds", c_int64), ("du_bs", c_int64), ("du_ds", c_int64),
        ("ddelta_bs", c_int64), ("ddelta_ds", c_int64), ("dz_bs", c_int64), ("dz_ds", c_int64),
        ("io_dtype", c_int32), ("delta_softplus", c_int32),
        ("dB_io", c_void_p), ("dC_io", c_void_p), ("zero_accumulators", c_int32),
        ("ndirs", c_int32), ("dir_mode", c_int32 * VV_MAX_DIRS), ("nframes", c_int32),
        ("B_ls", c_int64), ("C_ls", c_int64),
        ("dBio_bs", c_int64), ("dBio_gs", c_int64), ("dBio_ns", c_int64), ("dBio_ls", c_int64),
        ("dCio_bs", c_int64), ("dCio_gs", c_int64), ("dCio_ns", c_int64), ("dCio_ls", c_int64),
        ("gate_rows", c_int32), ("pass_mask", c_int32),
    ]


class DwConv3dArgs(Structure):
    """vv_dwconv3d_args"""
    _fields_ = [
        ("x", c_void_p), ("weight", c_void_p), ("bias", c_void_p), ("out", c_void_p),
        ("dout", c_void_p), ("dx", c_void_p), ("dweight", c_void_p), ("dbias", c_void_p),
        ("batch", c_int32), ("frames", c_int32), ("height", c_int32), ("width", c_int32), ("channels", c_int32),
        ("io_dtype", c_int32),
    ]


class LayerNormArgs(Structure):
    """vv_layernorm_args"""
    _fields_ = [
        ("x", c_void_p), ("weight", c_void_p), ("bias", c_void_p), ("out", c_void_p),
        ("mean", c_void_p), ("rstd", c_void_p), ("dout", c_void_p), ("dx", c_void_p),
        ("dweight", c_void_p), ("dbias", c_void_p),
        ("rows", c_int64), ("channels", c_int32),
        ("x_rs", c_int64), ("out_rs", c_int64), ("dout_rs", c_int64), ("dx_rs", c_int64),
        ("io_dtype", c_int32), ("out_dtype", c_int32), ("eps", c_float),
    ]


# every symbol include/vivim_b200.h declares (checked by tests/test_cabi.py)
EXPORTS = ("vv_version", "vv_last_error", "vv_scan_num_segments", "vv_conv1d_fwd", "vv_conv1d_bwd",
           "vv_conv1d_dirs_fwd", "vv_conv1d_dirs_bwd", "vv_scan_fwd", "vv_scan_bwd", "vv_last_launch_count",
           "vv_debug_force_scalar_io", "vv_dwconv3d_fwd", "vv_dwconv3d_bwd", "vv_layernorm_fwd", "vv_layernorm_bwd")

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m vivim_b200.build` "
                "(there is no CPU or PyTorch fallback for the Vivim Mamba hot path)")
        L = ctypes.CDLL(LIB_PATH)
        L.vv_version.restype = c_int
        L.vv_last_error.restype = c_char_p
        L.vv_last_launch_count.restype = c_int
        L.vv_debug_force_scalar_io.argtypes = [c_int]
        L.vv_debug_force_scalar_io.restype = c_int
        L.vv_scan_num_segments.argtypes = [c_int]
        L.vv_scan_num_segments.restype = c_int
        for name, argt in (("vv_conv1d_fwd", ConvArgs), ("vv_conv1d_bwd", ConvArgs),
                           ("vv_conv1d_dirs_fwd", ConvDirsArgs), ("vv_conv1d_dirs_bwd", ConvDirsArgs),
                           ("vv_scan_fwd", ScanArgs), ("vv_scan_bwd", ScanArgs),
                           ("vv_dwconv3d_fwd", DwConv3dArgs), ("vv_dwconv3d_bwd", DwConv3dArgs),
                           ("vv_layernorm_fwd", LayerNormArgs), ("vv_layernorm_bwd", LayerNormArgs)):
            fn = getattr(L, name)
            fn.argtypes = [POINTER(argt), c_void_p]
            fn.restype = c_int
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vv_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
