"""GPU numerics of the fused Swin block tail (csrc/swin_block_tail.cu) against a plain PyTorch fp32 reference of the same
op on the same bf16-rounded operands: timm SwinTransformerV2Block's  x = x + norm2(mlp(x))  and  x = x + norm1(proj(.)).
Tolerances: the hidden activation is rounded to bf16 before fc2 (as in the un-fused path); everything after it is fp32, so
the fp32 residual stream must agree to ~1e-2 of a LayerNorm output (unit scale) and y must be exactly bf16(master)."""
import pytest
import torch
import torch.nn.functional as F

import cuda_ops as K

pytestmark = pytest.mark.gpu


def _case(M, C, HID, seed, mlp=True, K1=None):
    g = torch.Generator().manual_seed(seed)
    K1 = K1 or C
    x = (torch.randn(M, K1, generator=g)).bfloat16()
    w1 = (torch.randn(HID, K1, generator=g) / K1 ** 0.5).bfloat16() if mlp else None
    b1 = torch.randn(HID, generator=g) * 0.3 if mlp else None
    kin = HID if mlp else K1
    w2 = (torch.randn(C, kin, generator=g) / kin ** 0.5).bfloat16()
    b2 = torch.randn(C, generator=g) * 0.3
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.rand(C, generator=g) - 0.5
    master = torch.randn(M, C, generator=g) * 3
    return x, w1, b1, w2, b2, gamma, beta, master


def _ref(x, w1, b1, w2, b2, gamma, beta, master, eps=1e-5):
    h = x.float()
    if w1 is not None:
        h = F.gelu(h @ w1.float().t() + b1).bfloat16().float()       # the hidden activation is stored as bf16 (TMEM operand)
    t = h @ w2.float().t() + b2
    return master + F.layer_norm(t, (t.shape[1],), gamma, beta, eps)


def _run(case):
    x, w1, b1, w2, b2, gamma, beta, master = case
    d = lambda t: t.cuda() if t is not None else None
    md = master.cuda().clone()
    y = K.swin_block_tail(d(x), d(w2), d(b2), d(gamma), d(beta), md, d(w1), d(b1))
    torch.cuda.synchronize()
    return md.cpu(), y.cpu()


@pytest.mark.parametrize("M,C", [(256, 96), (128 * 5 + 37, 96), (4096, 96), (300, 128), (1024, 192), (777, 256), (128 * 149 + 1, 96),
                                 (128 * 300, 192)])
def test_mlp_branch_matches_torch(M, C):
    case = _case(M, C, 4 * C, seed=M + C)
    m, y = _run(case)
    ref = _ref(*case)
    err = (m - ref).abs()
    assert err.max().item() <= 3e-2 and err.mean().item() <= 3e-3, (err.max().item(), err.mean().item())
    assert torch.equal(y, m.bfloat16())


@pytest.mark.parametrize("M,C", [(256, 96), (128 * 3 + 5, 96), (2048, 128), (1000, 192), (128 * 150, 256), (128 * 300 + 64, 96)])
def test_proj_branch_matches_torch(M, C):
    case = _case(M, C, 0, seed=7 * M + C, mlp=False)
    m, y = _run(case)
    ref = _ref(*case)
    err = (m - ref).abs()
    assert err.max().item() <= 2e-3 and err.mean().item() <= 2e-4, (err.max().item(), err.mean().item())
    assert torch.equal(y, m.bfloat16())


@pytest.mark.parametrize("M,C,K1", [(1000, 384, 384), (128 * 128, 384, 384), (640, 512, 512), (128 * 3 + 9, 384, 1536),
                                    (128 * 128, 384, 1536), (128 * 37, 256, 1024), (128 * 150, 512, 2048)])
def test_wide_rows_and_long_k_stream_through_shared_memory(M, C, K1):
    """one-GEMM mode with C > 256 (N = 256 + (C - 256) columns, single TMEM buffer) and / or K1 > 512 (fc2 -> norm2 -> residual):
    activations and weights stream by K chunk."""
    case = _case(M, C, 0, seed=M + C + K1, mlp=False, K1=K1)
    m, y = _run(case)
    ref = _ref(*case)
    err = (m - ref).abs()
    assert err.max().item() <= 3e-3 and err.mean().item() <= 3e-4, (err.max().item(), err.mean().item())
    assert torch.equal(y, m.bfloat16())


def test_y_may_alias_x_and_repeated_calls_are_deterministic():
    case = _case(128 * 200, 96, 384, seed=3)
    x, w1, b1, w2, b2, gamma, beta, master = case
    d = lambda t: t.cuda()
    outs = []
    for _ in range(2):
        xd, md = d(x).clone(), d(master).clone()
        K.swin_block_tail(xd, d(w2), d(b2), d(gamma), d(beta), md, d(w1), d(b1), y=xd)      # in place: y == x
        torch.cuda.synchronize()
        outs.append((md.cpu(), xd.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    m2, y2 = _run(case)
    assert torch.equal(outs[0][0], m2) and torch.equal(outs[0][1], y2)


def test_constant_rows_have_zero_variance_and_do_not_nan():
    """x = 0 -> the branch output is the bias vector for every row; LayerNorm of it is finite and identical on all rows."""
    M, C = 640, 96
    case = list(_case(M, C, 384, seed=5))
    case[0] = torch.zeros(M, C).bfloat16()
    m, y = _run(tuple(case))
    ref = _ref(*case)
    assert torch.isfinite(m).all()
    assert (m - ref).abs().max().item() <= 3e-2
