"""LayerNorm over the channels of token tensors (csrc/layernorm.cuh, the glue of a Temporal Mamba block:
modeling/vivim.py:153-157) against torch's F.layer_norm evaluated in float64 -- forward, dx, dweight, dbias."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from gpu_util import host

pytestmark = pytest.mark.gpu

CASES = [
    # rows, channels, in dtype, out dtype
    (3 * 20480, 64, torch.float32, torch.float32),      # stage 1, batch 3: what norm1 / norm2 see
    (3 * 20480, 64, torch.float32, torch.bfloat16),     # ... handing bf16 to the GEMM that follows (autocast)
    (5120, 128, torch.float32, torch.bfloat16),
    (1280, 320, torch.float32, torch.float32),
    (320, 512, torch.bfloat16, torch.bfloat16),
    (7, 100, torch.float32, torch.float32),             # channels % 4 != 0: element-wise path
    (33, 36, torch.float16, torch.float16),
    (1, 512, torch.float32, torch.float16),
    (5, 32, torch.float32, torch.float32),              # 8 lanes per row: four rows per warp, the last trip partly empty
    (1027, 64, torch.bfloat16, torch.bfloat16),         # 16 lanes per row, odd row count
]


@pytest.mark.parametrize("rows,C,din,dout", CASES, ids=lambda v: str(v).replace("torch.", ""))
def test_layernorm_matches_torch(cuda_device, rows, C, din, dout):
    from vivim_b200.layernorm import layer_norm_tokens
    torch.manual_seed(rows + C)
    x = (torch.randn(rows, C, device="cuda") * 2 + 0.5).to(din).requires_grad_()
    w = torch.randn(C, device="cuda", requires_grad=True)
    b = torch.randn(C, device="cuda", requires_grad=True)
    g = torch.randn(rows, C, device="cuda").to(dout)
    y = layer_norm_tokens(x, w, b, 1e-5, dout)
    assert y.dtype == dout
    y.backward(g)
    xr = x.detach().double().requires_grad_()
    wr, br = w.detach().double().requires_grad_(), b.detach().double().requires_grad_()
    yr = F.layer_norm(xr, (C,), wr, br, 1e-5)
    yr.backward(g.double())
    tol = {torch.float32: 2e-5, torch.float16: 3e-3, torch.bfloat16: 2e-2}
    assert rel_err(host(y), host(yr)) < tol[dout]
    assert rel_err(host(x.grad), host(xr.grad)) < max(tol[din], tol[dout])
    # parameter gradients: fp32 sums over the rows of (rounded) dout
    assert rel_err(host(w.grad), host(wr.grad)) < 1e-4
    assert rel_err(host(b.grad), host(br.grad)) < 1e-4


def test_token_layernorm_module_is_a_drop_in(cuda_device):
    """Same parameters / state dict as nn.LayerNorm; strided (sliced) input rows; autocast semantics."""
    from vivim_b200.layernorm import TokenLayerNorm, use_token_layernorm
    torch.manual_seed(0)
    ref = torch.nn.LayerNorm(64).cuda()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.normal_()
    mine = TokenLayerNorm(64).cuda()
    mine.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, 300, 96, device="cuda")[..., 16:80]          # row stride 96, unit channel stride
    assert rel_err(host(mine(x)), host(ref(x))) < 2e-5
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y32 = mine(x)
        mine.autocast_output = True
        y16 = mine(x)
    assert y32.dtype == torch.float32 and y16.dtype == torch.bfloat16
    assert rel_err(host(y16), host(ref(x))) < 1e-2
    seq = torch.nn.Sequential(torch.nn.LayerNorm(64), torch.nn.Linear(64, 8), torch.nn.LayerNorm(2048)).cuda()
    assert use_token_layernorm(seq) == 1 and type(seq[0]) is TokenLayerNorm and type(seq[2]) is torch.nn.LayerNorm


def test_layernorm_rejects_wide_rows(cuda_device):
    from vivim_b200.layernorm import layer_norm_tokens
    x = torch.randn(4, 1024, device="cuda")
    with pytest.raises(RuntimeError, match="channels <= 512"):
        layer_norm_tokens(x, None, None)
