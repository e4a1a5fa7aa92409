"""Token GEMM `i2v_linear_fwd` (-m gpu): F.linear with the residual add in the epilogue, against fp32 math on the same
bf16 operands.  Shapes are the projections of the SD1.5 blocks plus ragged / tiny cases."""
import pytest
import torch
import torch.nn.functional as F

from i2v_adapter_unofficial_b200 import _lib, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"

# (rows, K, N, bias, residual)
CASES = [
    (4096, 320, 1536, True, False),    # packed [Wq; Wk; Wv; Wq_x] projection, augmented layout (4 x 8 x 48)
    (4096, 640, 320, False, True),     # stacked output projection [O_self | O_x] + residual (tiles of 160)
    (2048, 1288, 320, False, True),    # feed-forward output Linear with the ones column: K = 4 * 320 + 8 (K tail)
    (4096, 320, 960, True, False),     # temporal q, k, v (tiles of 192)
    (1024, 1280, 1280, True, True),    # level 2
    (2048, 640, 1920, True, False),    # temporal q, k, v at level 1
    (130, 320, 768, True, False),      # frame-0 K/V projection of a tiny clip: one ragged row pair
    (300, 72, 40, True, True),         # everything ragged: K and N below one tile
    (257, 2560, 640, False, True),     # an odd number of row blocks
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "r{}K{}N{}b{}r{}".format(*[int(v) for v in c]))
def test_linear_matches_fp32_reference(case):
    rows, K, N, with_bias, with_res = case
    g = torch.Generator().manual_seed(rows + K + N)
    x = torch.randn(rows, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(DEV, torch.bfloat16)
    b = (torch.randn(N, generator=g) * 0.5).to(DEV, torch.bfloat16) if with_bias else None
    r = torch.randn(rows, N, generator=g).to(DEV, torch.bfloat16) if with_res else None
    n0 = _lib.launch_count()
    y = ops.linear(x, w, b, r)
    assert _lib.launch_count() == n0 + 1
    ref = F.linear(x.float(), w.float(), None if b is None else b.float())
    if r is not None:
        ref = ref.to(torch.bfloat16).float() + r.float()   # the reference rounds the Linear output before the add
    err = (y.float() - ref).abs().max().item()
    assert err <= 2 ** -7 * max(1.0, ref.abs().max().item()), err   # one bf16 ulp of the output range


def test_linear_in_place_residual_strided_input_and_leading_axes():
    g = torch.Generator().manual_seed(3)
    big = torch.randn(6, 100, 648, generator=g).to(DEV, torch.bfloat16)
    x = big[..., :640]                                   # row pitch 648 != K
    w = (torch.randn(320, 640, generator=g) * 0.04).to(DEV, torch.bfloat16)
    res = torch.randn(6, 100, 320, generator=g).to(DEV, torch.bfloat16)
    want = (F.linear(x.float(), w.float()).to(torch.bfloat16).float() + res.float())
    out = ops.linear(x, w, None, res, out=res)           # accumulate into the residual stream in place
    assert out.data_ptr() == res.data_ptr() and out.shape == (6, 100, 320)
    assert (out.float() - want).abs().max().item() <= 2 ** -7 * want.abs().max().item()


def test_linear_rejects_unsupported_operands():
    x = torch.randn(64, 100, device=DEV).to(torch.bfloat16)        # K = 100 is not a multiple of 8
    w = torch.randn(64, 100, device=DEV).to(torch.bfloat16)
    with pytest.raises(ValueError):
        ops.linear(x, w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.linear(torch.randn(8, 64).to(torch.bfloat16), torch.randn(8, 64).to(torch.bfloat16))
