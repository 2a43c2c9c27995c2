"""Out-of-bounds WRITES of the kernels, caught without compute-sanitizer: every tensor the wrappers allocate while a
test body runs (outputs, gradients, the checkpoint / aggregate / carry scratch) is carved out of a larger buffer whose
8 KiB on either side hold a byte pattern; after the body (which itself checks the values against the oracle) the bands
must still hold it.  The bodies are the ragged cases of the other GPU tests: sequence lengths around the 64-position
segment, channel counts that leave warps and channel groups partly empty, frame counts on both gather routes.

`torch.empty` bodies are filled with 0xFF bytes (NaN in every floating type), so reliance on memory that happens to
be zero shows up as NaN in the body's own comparison with the oracle.

A write further than 8 KiB away is not seen; a read out of bounds is not seen either (those show up as wrong values
in the parity comparisons only if the data matters) -- this is a tripwire, not a proof."""
import contextlib
import math

import pytest
import torch

import test_conv1d_gpu
import test_dirs_gpu
import test_dwconv3d_gpu
import test_layernorm_gpu
import test_scan_gpu

pytestmark = pytest.mark.gpu

BAND = 8192
PATTERN = 0xA5


class GuardBands:
    def __init__(self):
        self.allocs = []
        self.real = {n: getattr(torch, n) for n in ("empty", "zeros", "empty_like", "zeros_like")}

    def _carve(self, shape, dtype, device, zero):
        dtype = dtype or torch.get_default_dtype()
        nbytes = math.prod(shape) * dtype.itemsize
        raw = self.real["empty"](nbytes + 2 * BAND, dtype=torch.uint8, device=device)
        raw.fill_(PATTERN)
        body = raw[BAND:BAND + nbytes].view(dtype).view(tuple(shape))
        if zero:
            body.zero_()
        else:
            raw[BAND:BAND + nbytes].fill_(0xFF)      # NaN in fp32 / bf16 / fp16: a kernel that reads (or accumulates into)
                                                     # memory nobody initialised fails the body's parity comparison
        self.allocs.append((raw, nbytes))
        return body

    @staticmethod
    def _is_cuda(device):
        return device is not None and torch.device(device).type == "cuda"

    def _plain(self, name, zero):
        def fn(*size, dtype=None, device=None, **kw):
            if not self._is_cuda(device) or kw.get("pin_memory") or kw.get("out") is not None:
                return self.real[name](*size, dtype=dtype, device=device, **kw)
            shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
            return self._carve(shape, dtype, device, zero)
        return fn

    def _like(self, name, zero):
        def fn(t, dtype=None, device=None, **kw):
            device = t.device if device is None else device
            if not self._is_cuda(device):
                return self.real[name](t, dtype=dtype, device=device, **kw)
            return self._carve(t.shape, dtype or t.dtype, device, zero)
        return fn

    def check(self):
        torch.cuda.synchronize()
        assert self.allocs, "the body allocated nothing through torch.empty / zeros: the tripwire is not armed"
        for i, (raw, nbytes) in enumerate(self.allocs):
            lo, hi = raw[:BAND], raw[BAND + nbytes:]
            assert bool((lo == PATTERN).all()), f"allocation {i} ({nbytes} B): bytes BEFORE it were written"
            assert bool((hi == PATTERN).all()), f"allocation {i} ({nbytes} B): bytes AFTER it were written"


@contextlib.contextmanager
def guarded():
    from vivim_b200 import selective_scan_cuda as ssc
    g = GuardBands()
    ssc._SCRATCH.clear()                     # the cached aggregate / carry scratch is re-allocated inside the bands
    patched = {"empty": g._plain("empty", False), "zeros": g._plain("zeros", True),
               "empty_like": g._like("empty_like", False), "zeros_like": g._like("zeros_like", True)}
    try:
        for n, f in patched.items():
            setattr(torch, n, f)
        yield g
    finally:
        for n, f in g.real.items():
            setattr(torch, n, f)
        ssc._SCRATCH.clear()


def test_tripwire_sees_a_stray_write(cuda_device):
    with guarded() as g:
        t = torch.empty((4, 6), dtype=torch.bfloat16, device="cuda")
        z = torch.zeros_like(t)
        assert t.shape == (4, 6) and t.dtype == torch.bfloat16 and t.is_contiguous() and float(z.abs().sum()) == 0.0
        g.check()
        raw, nbytes = g.allocs[0]
        raw[BAND + nbytes + 2] = 0           # one byte past the end
    with pytest.raises(AssertionError, match="AFTER"):
        g.check()


@pytest.mark.parametrize("shape", [(2, 6, 151, 8, 1), (1, 5, 7, 3, 1), (1, 4, 257, 32, 1), (1, 4, 1134, 16, 2), (3, 40, 65, 16, 1),
                                   (1, 24, 1, 16, 1), (2, 17, 63, 5, 1)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_scan_writes_stay_inside(cuda_device, shape, dtype):
    with guarded() as g:
        test_scan_gpu.test_scan_matches_oracle(cuda_device, shape, dtype)
    g.check()


@pytest.mark.parametrize("case", [test_dirs_gpu.SCAN_CASES[i] for i in (1, 2, 4, 7, 8, 9)],
                         ids=lambda c: "-".join(map(str, c[:5])) + "-" + "".join(d[0] for d in c[5]))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_scan_dirs_writes_stay_inside(cuda_device, case, dtype):
    with guarded() as g:
        test_dirs_gpu.test_scan_dirs_matches_oracle(cuda_device, case, dtype)
    g.check()


@pytest.mark.parametrize("seqlen,width", [(8, 2), (151, 4), (372, 3), (1134, 4)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_conv_writes_stay_inside(cuda_device, seqlen, width, dtype):
    with guarded() as g:
        test_conv1d_gpu.test_conv_matches_oracle(cuda_device, seqlen, width, dtype)
    g.check()


@pytest.mark.parametrize("case", [test_dirs_gpu.CONV_CASES[i] for i in (1, 2, 3, 5, 6, 9)],
                         ids=lambda c: "-".join(map(str, c[:5])) + "-" + "".join(d[0] for d in c[5]))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_conv_dirs_writes_stay_inside(cuda_device, case, dtype):
    with guarded() as g:
        test_dirs_gpu.test_conv_dirs_matches_oracle(cuda_device, case, dtype, True, True)
    g.check()


@pytest.mark.parametrize("rows,C,din,dout", [c for c in test_layernorm_gpu.CASES if c[0] < 2000],
                         ids=lambda v: str(v).replace("torch.", ""))
def test_layernorm_writes_stay_inside(cuda_device, rows, C, din, dout):
    with guarded() as g:
        test_layernorm_gpu.test_layernorm_matches_torch(cuda_device, rows, C, din, dout)
    g.check()


@pytest.mark.parametrize("shape", [(1, 5, 7, 9, 40), (2, 3, 5, 6, 12), (1, 1, 4, 4, 8), (1, 2, 3, 5, 7)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_dwconv3d_writes_stay_inside(cuda_device, shape, dtype):
    with guarded() as g:
        test_dwconv3d_gpu.test_dwconv3d_matches_conv3d(cuda_device, shape, dtype)
    g.check()


def test_fused_block_writes_stay_inside(cuda_device):
    """Mamba(v3) forward + backward on the direction-fused route, odd frame size: every intermediate of
    vivim_b200/mamba_block.py (xz, conv_out, x_dbl views, dxz, dx_dbl) is carved with bands."""
    with guarded() as g:
        test_dirs_gpu.test_fused_block_equals_three_call_route(cuda_device, "v3", 32, 5, 52, 2, True)
    g.check()
