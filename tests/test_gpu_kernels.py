"""Parity tests proper (-m gpu): every kernel family through the C ABI against the CPU oracle on the same seeded
inputs, the golden fixtures, fp32 check mode, and size-independent properties at BASELINE.json's full sizes.

Tolerances are BASELINE.json's: max-abs <= 2e-2 on bf16 attention outputs, <= 1e-3 relative in fp32 check mode."""
import glob
import os

import numpy as np
import pytest
import torch

from i2v_adapter_unofficial_b200 import _lib, ops
from oracle.attention_oracle import attention_oracle, sdpa_oracle

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
F32_REL_TOL = 1e-3
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _rand(shape, seed, dtype=torch.bfloat16):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g).to(dtype)


def _oracle(q, k, v, kv_group=1):
    """inputs [B,S,H,d] (any dtype, CPU) -> fp32 oracle output [B,S,H,d]."""
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    if kv_group > 1:
        kf = kf.repeat_interleave(kv_group, dim=0)
        vf = vf.repeat_interleave(kv_group, dim=0)
    return sdpa_oracle(qf, kf, vf).permute(0, 2, 1, 3)


def _cuda(*ts):
    return [t.cuda() for t in ts]


# (batch, heads, Sq, Skv, d, kv_group): SURVEY.md Appendix C level shapes at reduced batch, ragged and tiny cases
DENSE_CASES = [
    (2, 8, 4096, 4096, 40, 1), (2, 8, 1024, 1024, 80, 1), (2, 8, 256, 256, 160, 1), (2, 8, 64, 64, 160, 1),
    (1, 8, 2304, 2304, 80, 1), (2, 8, 576, 576, 160, 1), (2, 8, 144, 144, 160, 1),
    (2, 8, 200, 77, 40, 1), (3, 2, 1, 1, 40, 1), (2, 4, 129, 257, 64, 1), (4, 8, 256, 256, 40, 2), (6, 2, 300, 300, 16, 3),
    (2, 2, 384, 384, 32, 1), (1, 1, 128, 128, 128, 1),
    (1, 2, 9216, 9216, 40, 1),   # configs[3]: 96x96 latent, level 0
]


@pytest.mark.parametrize("case", DENSE_CASES, ids=lambda c: "B{}H{}Sq{}Skv{}d{}g{}".format(*c))
def test_dense_sdpa_matches_oracle(case):
    B, H, Sq, Skv, d, g = case
    q, k, v = _rand((B, Sq, H, d), 1), _rand((B // g, Skv, H, d), 2), _rand((B // g, Skv, H, d), 3)
    o = ops.sdpa(*_cuda(q, k, v), g, None, ops.MODE_FAST).float().cpu()
    assert (o - _oracle(q, k, v, g)).abs().max().item() <= BF16_TOL


def test_dense_sdpa_large_magnitude_scores_online_softmax():
    # scores with a wide dynamic range and a maximum that keeps growing along the key axis exercise the lazy
    # rescale of the running maximum
    B, H, S, d = 1, 2, 1024, 40
    q = _rand((B, S, H, d), 5) * 3
    k = _rand((B, S, H, d), 6) * torch.linspace(0.2, 4.0, S).view(1, S, 1, 1).to(torch.bfloat16)
    v = _rand((B, S, H, d), 7) * 0.25  # near one-hot softmax: |o| ~ |v|; keep bf16 output quantisation below the tolerance
    o = ops.sdpa(*_cuda(q, k, v), 1, None, ops.MODE_FAST).float().cpu()
    assert (o - _oracle(q, k, v)).abs().max().item() <= BF16_TOL


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "basic_attention_*.npz"))),
                         ids=lambda p: os.path.basename(p)[16:-4])
@pytest.mark.parametrize("mode", ["bf16_fast", "f32_check"])
def test_golden_vectors_through_the_cabi(path, mode):
    """Reference-generated vectors (oracle/make_golden.py): projections on the host in fp32, attention core through
    the library, compared with the reference's output."""
    z = np.load(path)
    sd = {f"a.{k[2:]}": torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    x = torch.from_numpy(z["x"])
    ctx = torch.from_numpy(z["ctx"]) if "ctx" in z.files else x
    H = int(z["heads"])
    q = torch.nn.functional.linear(x, sd["a.to_q.weight"])
    k = torch.nn.functional.linear(ctx, sd["a.to_k.weight"])
    v = torch.nn.functional.linear(ctx, sd["a.to_v.weight"])
    B, S, C = q.shape
    d = C // H
    dt = torch.bfloat16 if mode == "bf16_fast" else torch.float32
    o = ops.sdpa(q.view(B, S, H, d).to(dt).cuda(), k.view(B, -1, H, d).to(dt).cuda(),
                 v.view(B, -1, H, d).to(dt).cuda(), 1, None,
                 ops.MODE_FAST if mode == "bf16_fast" else ops.MODE_GENERIC).float().cpu()
    y = torch.nn.functional.linear(o.reshape(B, S, C), sd["a.to_out.0.weight"], sd["a.to_out.0.bias"])
    ref = torch.from_numpy(z["y"])
    if mode == "bf16_fast":
        assert (y - ref).abs().max().item() <= BF16_TOL
    else:
        assert (y - ref).abs().max().item() <= F32_REL_TOL * ref.abs().max().item()
    assert (y - attention_oracle(sd, "a", x, None if "ctx" not in z.files else ctx, H)).abs().max().item() <= BF16_TOL


@pytest.mark.parametrize("case", [(2, 4, 8, 256, 40), (1, 16, 8, 1024, 40), (2, 3, 8, 320, 80), (2, 2, 8, 64, 160),
                                  (1, 2, 4, 100, 64)], ids=lambda c: "V{}F{}H{}S{}d{}".format(*c))
def test_fused_self_and_cross_frame_matches_oracle(case):
    V, Fr, H, S, d = case
    BF = V * Fr
    y = _rand((BF, S, 4, H, d), 11)
    kvx = _rand((V, S, 2, H, d), 12)
    yc, kc = y.cuda(), kvx.cuda()
    o = ops.fused_self_xframe(yc[:, :, 0], yc[:, :, 1], yc[:, :, 2], yc[:, :, 3], kc[:, :, 0], kc[:, :, 1], Fr, None,
                              ops.MODE_FAST).float().cpu()
    assert (o[:, :, 0] - _oracle(y[:, :, 0], y[:, :, 1], y[:, :, 2])).abs().max().item() <= BF16_TOL
    # cross-frame: every frame of video b attends to K/V row b (frame 0), reference src/modules/i2v_adapter.py:484-492
    assert (o[:, :, 1] - _oracle(y[:, :, 3], kvx[:, :, 0], kvx[:, :, 1], Fr)).abs().max().item() <= BF16_TOL


def _aug_run(q, k, v, qx, kx, vx, Fr):
    qa, ka, va = ops.augment_qkv(*_cuda(q, k, v))
    qxa, kxa, vxa = ops.augment_qkv(*_cuda(qx, kx, vx))
    return ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr).float().cpu()


@pytest.mark.parametrize("case", [(2, 4, 8, 256), (1, 16, 8, 1024), (2, 2, 8, 600), (3, 2, 2, 77), (1, 2, 8, 1537)],
                         ids=lambda c: "V{}F{}H{}S{}".format(*c))
def test_fused_augmented_layout_matches_oracle(case):
    # the augmented operand layout (scale folded into q, ones column in k and v, d = 40 padded to 48) is another
    # encoding of the same operator: same oracle, same tolerance; sizes cover ragged key tiles (S % 64 != 0) and
    # partially filled query blocks (S % 384 != 0)
    V, Fr, H, S = case
    d, BF = 40, V * Fr
    q, k, v, qx = (_rand((BF, S, H, d), 31 + i) for i in range(4))
    kx, vx = _rand((V, S, H, d), 41), _rand((V, S, H, d), 42)
    o = _aug_run(q, k, v, qx, kx, vx, Fr)
    assert (o[:, :, 0] - _oracle(q, k, v)).abs().max().item() <= BF16_TOL
    assert (o[:, :, 1] - _oracle(qx, kx, vx, Fr)).abs().max().item() <= BF16_TOL


def test_fused_augmented_layout_moving_maximum():
    # keys whose scores keep growing along the key axis move the reference maximum many times: every move rewrites the
    # max column of the query tile and sends one or two score tiles through the stale-column slow path
    V, Fr, H, S, d = 1, 2, 2, 1024, 40
    q = _rand((V * Fr, S, H, d), 51) * 2
    ramp = torch.linspace(0.2, 6.0, S).view(1, S, 1, 1)
    k = (_rand((V * Fr, S, H, d), 52) * ramp).to(torch.bfloat16)
    kx = (_rand((V, S, H, d), 53) * ramp.flip(1)).to(torch.bfloat16)
    v, vx = _rand((V * Fr, S, H, d), 54), _rand((V, S, H, d), 55)
    o = _aug_run(q, k, v, q, kx, vx, Fr)
    assert torch.isfinite(o).all()
    # scores reach +-100 octaves here, so the oracle gets the query the kernel sees (the helper rounds q * scale to
    # bf16 a second time; the product path folds the scale into the projection weights and rounds once)
    c = d ** -0.5 * 1.4426950408889634
    q_eff = (ops.augment_qkv(q, k, v)[0][..., :d].float() / c)
    assert (o[:, :, 0] - _oracle(q_eff, k, v)).abs().max().item() <= BF16_TOL
    assert (o[:, :, 1] - _oracle(q_eff, kx, vx, Fr)).abs().max().item() <= BF16_TOL


def test_fused_augmented_layout_far_below_maximum_scores_vanish():
    # one dominant key per row and all others > 126 octaves below it: the FMA-pipe exp2 must clamp, not wrap
    V, Fr, H, S, d = 1, 1, 1, 256, 40
    q = torch.zeros(1, S, H, d)
    q[..., 0] = 60.0
    k = torch.zeros(1, S, H, d)
    k[:, :, :, 0] = -60.0
    k[:, 7, :, 0] = 60.0
    v = _rand((1, S, H, d), 56)
    qb, kb = q.to(torch.bfloat16), k.to(torch.bfloat16)
    o = _aug_run(qb, kb, v, qb, kb, v, Fr)
    assert torch.isfinite(o).all()
    assert (o[:, :, 0] - v[:, 7:8].float()).abs().max().item() <= BF16_TOL


def test_kv_group_indexing_is_bit_identical_to_the_repeated_tensor():
    # reference materialises the first frame F times (einops.repeat, :485); indexing it in place must not change a bit
    V, Fr, H, S, d = 2, 4, 8, 512, 40
    q, k, v = _cuda(_rand((V * Fr, S, H, d), 21), _rand((V, S, H, d), 22), _rand((V, S, H, d), 23))
    a = ops.sdpa(q, k, v, Fr, None, ops.MODE_FAST)
    b = ops.sdpa(q, k.repeat_interleave(Fr, 0), v.repeat_interleave(Fr, 0), 1, None, ops.MODE_FAST)
    assert torch.equal(a, b)


@pytest.mark.parametrize("case", [(4, 8, 256, 40, 2, 77, 4), (2, 8, 1024, 80, 1, 77, 4), (2, 8, 64, 160, 2, 50, 14),
                                  (2, 8, 300, 40, 1, 77, 16), (2, 2, 128, 64, 1, 10, 1), (4, 8, 272, 160, 4, 77, 4),
                                  (2, 8, 1000, 40, 2, 77, 4), (2, 6, 200, 40, 1, 77, 4)],
                         ids=lambda c: "B{}H{}S{}d{}g{}n{}+{}".format(*c))
@pytest.mark.parametrize("mode", ["fast", "generic"])
def test_ip_adapter_decoupled_cross_attention_matches_oracle(case, mode):
    B, H, S, d, g, nt, ni = case
    q, k, v = _rand((B, S, H, d), 31), _rand((B // g, nt + ni, H, d), 32), _rand((B // g, nt + ni, H, d), 33)
    o = ops.ip_xattn(*_cuda(q, k, v), nt, 0.6, g, None, ops.MODE_FAST if mode == "fast" else ops.MODE_GENERIC)
    ref = _oracle(q, k[:, :nt], v[:, :nt], g) + 0.6 * _oracle(q, k[:, nt:], v[:, nt:], g)
    assert (o.float().cpu() - ref).abs().max().item() <= BF16_TOL


TEMPORAL_CASES = [(512, 16, 8, 40), (300, 16, 8, 80), (200, 16, 8, 160), (256, 8, 8, 40), (128, 32, 8, 40),
                  (64, 24, 8, 80), (77, 16, 8, 64), (50, 32, 8, 160), (33, 1, 8, 40), (40, 17, 8, 32), (9, 16, 16, 40)]


@pytest.mark.parametrize("case", TEMPORAL_CASES, ids=lambda c: "N{}F{}H{}d{}".format(*c))
def test_temporal_attention_matches_oracle(case):
    N, Fr, H, d = case
    qkv = _rand((N, Fr, 3, H, d), 41)
    c = qkv.cuda()
    o = ops.temporal_attn(c[:, :, 0], c[:, :, 1], c[:, :, 2], None, ops.MODE_FAST).float().cpu()
    assert (o - _oracle(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2])).abs().max().item() <= BF16_TOL


@pytest.mark.parametrize("case", [(2, 8, 200, 77, 40, 1), (4, 4, 65, 130, 160, 2), (3, 2, 33, 31, 16, 1),
                                  (2, 3, 50, 50, 24, 1)], ids=lambda c: "B{}H{}Sq{}Skv{}d{}g{}".format(*c))
def test_fp32_check_mode(case):
    B, H, Sq, Skv, d, g = case
    q, k, v = (_rand(s, i, torch.float32) for i, s in enumerate([(B, Sq, H, d), (B // g, Skv, H, d), (B // g, Skv, H, d)]))
    o = ops.sdpa(*_cuda(q, k, v), g, None, ops.MODE_AUTO).cpu()  # fp32 tensors select the fp32-math kernel
    ref = _oracle(q, k, v, g)
    assert (o - ref).abs().max().item() <= F32_REL_TOL * ref.abs().max().item()
    qkv = _rand((40, 16, 3, 8, 40), 9, torch.float32)
    c = qkv.cuda()
    ot = ops.temporal_attn(c[:, :, 0], c[:, :, 1], c[:, :, 2]).cpu()
    rt = _oracle(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2])
    assert (ot - rt).abs().max().item() <= F32_REL_TOL * rt.abs().max().item()


def test_auto_mode_covers_shapes_outside_the_fast_kernels():
    q, k, v = _rand((2, 50, 3, 20), 1), _rand((2, 60, 3, 20), 2), _rand((2, 60, 3, 20), 3)   # d = 20: not a multiple of 8, no tcgen05 config
    with pytest.raises(_lib.I2VLibraryError) as e:
        ops.sdpa(*_cuda(q, k, v), 1, None, ops.MODE_FAST)
    assert e.value.code == -2
    o = ops.sdpa(*_cuda(q, k, v), 1, None, ops.MODE_AUTO).float().cpu()
    assert (o - _oracle(q, k, v)).abs().max().item() <= BF16_TOL


def test_misaligned_input_is_rejected_not_silently_copied():
    base = torch.randn(2 * 64 * 2 * 40 + 8, device="cuda").to(torch.bfloat16)
    q = base[4:4 + 2 * 64 * 2 * 40].view(2, 64, 2, 40)  # 8-byte aligned only
    with pytest.raises(_lib.I2VLibraryError) as e:
        ops.sdpa(q, q, q, 1, None, ops.MODE_FAST)
    assert e.value.code == -3


# ---------------------------------------------------------------------------------------------------------------
# size-independent properties at BASELINE.json's full sizes (configs[1] level 0: BF=32, 8 heads, S=4096, d=40)
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_properties_dense():
    BF, H, S, d, Fr = 32, 8, 4096, 40, 16
    g = torch.Generator(device="cuda").manual_seed(7)
    q = torch.randn(BF, S, H, d, device="cuda", generator=g).to(torch.bfloat16)
    k = torch.randn(BF, S, H, d, device="cuda", generator=g).to(torch.bfloat16)
    v = torch.randn(BF, S, H, d, device="cuda", generator=g).to(torch.bfloat16)
    # (1) softmax rows sum to one: constant V comes back unchanged (per head-dim column)
    const = torch.linspace(-2, 2, d, device="cuda").to(torch.bfloat16).expand(BF, S, H, d).contiguous()
    o = ops.sdpa(q, k, const, 1, None, ops.MODE_FAST)
    assert (o.float() - const.float()).abs().max().item() <= 1.6e-2
    # (2) key order does not matter
    perm = torch.randperm(S, device="cuda", generator=g)
    o1 = ops.sdpa(q, k, v, 1, None, ops.MODE_FAST)
    o2 = ops.sdpa(q, k[:, perm].contiguous(), v[:, perm].contiguous(), 1, None, ops.MODE_FAST)
    assert (o1.float() - o2.float()).abs().max().item() <= BF16_TOL
    # (3) linear in V
    v2 = torch.randn(BF, S, H, d, device="cuda", generator=g).to(torch.bfloat16)
    o3 = ops.sdpa(q, k, v2, 1, None, ops.MODE_FAST)
    o4 = ops.sdpa(q, k, (v.float() + v2.float()).to(torch.bfloat16), 1, None, ops.MODE_FAST)
    assert (o4.float() - (o1.float() + o3.float())).abs().max().item() <= 3e-2
    # (4) fused launch == two separate launches, bit for bit; cross-frame half only depends on frame 0's K/V
    kx, vx = k[0::Fr].contiguous(), v[0::Fr].contiguous()
    of = ops.fused_self_xframe(q, k, v, q, kx, vx, Fr, None, ops.MODE_FAST)
    assert torch.equal(of[:, :, 0], o1)
    assert torch.equal(of[:, :, 1], ops.sdpa(q, kx, vx, Fr, None, ops.MODE_FAST))
    # (5) a checksum against torch's own SDPA on the same device (independent implementation)
    ref = torch.nn.functional.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3),
                                                           v.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    assert (o1.float() - ref.float()).abs().max().item() <= BF16_TOL


def test_full_size_properties_temporal():
    N, Fr, H, d = 8192, 16, 8, 40      # configs[1] level 0: 2 videos x 4096 positions
    g = torch.Generator(device="cuda").manual_seed(8)
    qkv = torch.randn(N, Fr, 3, H, d, device="cuda", generator=g).to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o = ops.temporal_attn(q, k, v, None, ops.MODE_FAST)
    # positions are independent: a permutation of positions permutes the output
    perm = torch.randperm(N, device="cuda", generator=g)
    op = ops.temporal_attn(q[perm].contiguous(), k[perm].contiguous(), v[perm].contiguous(), None, ops.MODE_FAST)
    assert torch.equal(op, o[perm])
    # frame (key) order does not matter
    fp = torch.randperm(Fr, device="cuda", generator=g)
    of = ops.temporal_attn(q.contiguous(), k[:, fp].contiguous(), v[:, fp].contiguous(), None, ops.MODE_FAST)
    assert (of.float() - o.float()).abs().max().item() <= BF16_TOL
    ref = torch.nn.functional.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3),
                                                           v.permute(0, 2, 1, 3)).permute(0, 2, 1, 3)
    assert (o.float() - ref.float()).abs().max().item() <= BF16_TOL


def test_reshard_kernels_roundtrip_and_match_index_definition():
    V, f, S, C, G = 2, 4, 64, 320, 4
    x = torch.randn(V, f, S, C, device="cuda").to(torch.bfloat16)
    packed = ops.reshard_pack(x, G)
    ref = torch.stack([x[:, :, r * (S // G):(r + 1) * (S // G)] for r in range(G)])
    assert torch.equal(packed, ref)
    assert torch.equal(ops.reshard_pack(packed, G, inverse=True), x)
    recv = torch.randn(G, V, f, S // G, C, device="cuda")
    y = ops.reshard_unpack(recv, G)
    assert torch.equal(y, recv.permute(1, 0, 2, 3, 4).reshape(V, G * f, S // G, C))
    assert torch.equal(ops.reshard_unpack(y, G, inverse=True), recv)


def test_launch_counter_counts_library_kernels():
    q = torch.randn(1, 128, 1, 64, device="cuda").to(torch.bfloat16)
    n0 = _lib.launch_count()
    ops.sdpa(q, q, q, 1, None, ops.MODE_FAST)
    ops.sdpa(q, q, q, 1, None, ops.MODE_GENERIC)
    assert _lib.launch_count() == n0 + 2
