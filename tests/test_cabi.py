"""C-ABI surface: the shared library loads without a GPU, exports every symbol that
include/vivim_b200.h declares, the ctypes structs have the C layout, and argument errors are
reported through return codes (no compute is launched here)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from vivim_b200 import _lib, build

HEADER = os.path.join(ROOT, "include", "vivim_b200.h")


@pytest.fixture(scope="module")
def L():
    build.build()
    return _lib.lib()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vv_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(L):
    names = declared_functions()
    assert names and set(names) == set(_lib.EXPORTS)
    for n in names:
        assert getattr(L, n) is not None


def test_struct_layout_matches_c(tmp_path):
    """sizeof and the offset of EVERY field of every args struct, compiled from the header with gcc."""
    structs = {"vv_scan_args": _lib.ScanArgs, "vv_conv1d_args": _lib.ConvArgs,
               "vv_conv1d_dirs_args": _lib.ConvDirsArgs, "vv_dwconv3d_args": _lib.DwConv3dArgs,
               "vv_layernorm_args": _lib.LayerNormArgs}
    lines = []
    for cname, cls in structs.items():
        lines.append(f'printf("%zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("%zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vivim_b200.h"\nint main(){' + "".join(lines)
                   + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = []
    for cls in structs.values():
        want.append(ctypes.sizeof(cls))
        want += [getattr(cls, fname).offset for fname, _ in cls._fields_]
    assert got == want


def test_version_and_units(L):
    assert L.vv_version() == 200
    assert L.vv_scan_num_segments(20480) == 320
    assert L.vv_scan_num_segments(1) == 1 and L.vv_scan_num_segments(65) == 2 and L.vv_scan_num_segments(0) == 0


def test_bad_arguments_are_reported_not_launched(L):
    assert L.vv_conv1d_fwd(None, None) == -1
    a = _lib.ConvArgs()
    assert L.vv_conv1d_fwd(ctypes.byref(a), None) == -1
    assert b"required" in L.vv_last_error()
    buf = ctypes.create_string_buffer(256)
    p = ctypes.cast(buf, ctypes.c_void_p).value
    a.x = a.weight = a.out = p
    a.batch = a.dim = 1
    a.seqlen = 8
    a.width = 5                                     # the reference rejects widths outside 2..4 too
    assert L.vv_conv1d_fwd(ctypes.byref(a), None) == -2
    assert b"width between 2 and 4" in L.vv_last_error()
    s = _lib.ScanArgs()
    assert L.vv_scan_fwd(ctypes.byref(s), None) == -1
    s.u = s.delta = s.A = s.Bm = s.Cm = s.agg = s.chk = s.out = p
    s.batch = s.dim = s.ngroups = 1
    s.seqlen = 8
    s.dstate = 300
    assert L.vv_scan_fwd(ctypes.byref(s), None) == -1 and b"<= 256" in L.vv_last_error()
    s.dstate = 64
    assert L.vv_scan_fwd(ctypes.byref(s), None) == -2  # valid for the reference, not served here
    assert L.vv_last_launch_count() == 0
    # directions: bad mode / frames that do not divide the sequence are argument errors
    s.dstate = 16
    s.ndirs, s.ngroups, s.dim = 3, 3, 3
    s.dir_mode[2] = 7
    assert L.vv_scan_fwd(ctypes.byref(s), None) == -1 and b"dir_mode" in L.vv_last_error()
    s.dir_mode[2] = _lib.VV_DIR_FRAMES
    s.nframes = 5
    assert L.vv_scan_fwd(ctypes.byref(s), None) == -1 and b"nframes" in L.vv_last_error()
    c = _lib.ConvDirsArgs()
    assert L.vv_conv1d_dirs_fwd(ctypes.byref(c), None) == -1
    c.x = c.weight = c.out = p
    c.batch = c.dim = 1
    c.seqlen, c.width, c.ndirs = 8, 4, 5
    assert L.vv_conv1d_dirs_fwd(ctypes.byref(c), None) == -1 and b"ndirs" in L.vv_last_error()
    w = _lib.DwConv3dArgs()
    assert L.vv_dwconv3d_fwd(ctypes.byref(w), None) == -1 and b"weight is required" in L.vv_last_error()
    w.weight = w.x = w.out = p
    assert L.vv_dwconv3d_fwd(ctypes.byref(w), None) == -1 and b"sizes must be positive" in L.vv_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libvivim_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()
