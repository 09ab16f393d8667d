"""GPU parity of the generic tensor-core Conv1d / Linear (fac_conv_gemm_tc) against plain PyTorch fp32
(TF32 off) on the same seeded inputs, and of its operand-preparation kernels.  The split modes execute
hi*hi + lo*hi + hi*lo on tcgen05: fp16 pairs carry 22 significand bits, bf16 pairs 16."""
import pytest
import torch
import torch.nn.functional as F

from fac_via_ppg_b200 import _ext, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_weight(n_out, c_in, taps, dtype, bias=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(n_out, c_in, taps, generator=g) / (c_in * taps) ** 0.5).to(DEV)
    b = torch.randn(n_out, generator=g).to(DEV) if bias else None
    c_pad, n_pad = (c_in + 63) // 64 * 64, (n_out + 63) // 64 * 64
    wp = torch.zeros(n_pad, taps, c_pad, device=DEV)
    wp[:n_out, :, :c_in] = w.permute(0, 2, 1)
    wp = wp.reshape(n_pad, taps * c_pad)
    hi = wp.to(dtype)
    lo = (wp - hi.float()).to(dtype)
    bp = None
    if bias:
        bp = torch.zeros(n_pad, device=DEV)
        bp[:n_out] = b
    return w, b, dict(hi=hi.contiguous(), lo=lo.contiguous(), c_pad=c_pad, taps=taps, n_pad=n_pad, n_valid=n_out, bias=bp)


@pytest.fixture(autouse=True)
def no_tf32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-5), (torch.bfloat16, 2e-4)])
@pytest.mark.parametrize("B,T,c_in,n_out,taps", [(2, 200, 600, 600, 5), (3, 37, 80, 512, 5), (1, 130, 512, 80, 5),
                                                 (2, 64, 5816, 600, 1)])
def test_conv_gemm_tc_matches_torch(dtype, tol, B, T, c_in, n_out, taps):
    g = torch.Generator().manual_seed(B * 1000 + T)
    x = torch.randn(B, T, c_in, generator=g).to(DEV)
    w, b, tw = make_weight(n_out, c_in, taps, dtype, seed=T)
    mask = (torch.rand(B, T, n_out, generator=g) >= 0.5).float().mul(2.0).to(DEV)
    res = torch.randn(B, T, n_out, generator=g).to(DEV)
    ref = torch.relu(F.conv1d(x.transpose(1, 2), w, b, padding=(taps - 1) // 2)).transpose(1, 2) * mask + res
    a = ops.pad_split(x.contiguous(), tw["c_pad"], dtype=dtype)
    out = torch.empty(B, T, n_out, device=DEV)
    _, nxt = ops.conv_gemm_tc(a, tw, act=_ext.ACT_RELU, mask=mask, residual=res, out=out)
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= tol * scale
    # the operand copies written for the next layer: hi + lo reproduces the fp32 result, padding is zero
    both = nxt[0].float() + nxt[1].float()
    split_tol = 2e-6 if dtype == torch.float16 else 3e-5
    assert (both[:, :, :n_out] - out).abs().max().item() <= split_tol * scale
    assert both[:, :, n_out:].abs().max().item() == 0.0 if tw["n_pad"] > n_out else True


def test_k_chunking_reduces_the_accumulator_truncation():
    """One accumulation chain over K = 5824 truncates ~1000 times; 512-element chains meet in fp32 RN."""
    B, T, c_in, n_out = 2, 128, 5816, 600
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, T, c_in, generator=g).to(DEV)          # same-sign terms: the worst case for truncation
    w, _, tw = make_weight(n_out, c_in, 1, torch.float16, bias=False, seed=3)
    tw["hi"] = tw["hi"].abs()
    ref = F.conv1d(x.transpose(1, 2).double(), (tw["hi"].double() + tw["lo"].double())[:n_out, :c_in, None]).transpose(1, 2)
    a = ops.pad_split(x.contiguous(), tw["c_pad"])
    errs = {}
    for kc in (0, 512):
        out = torch.empty(B, T, n_out, device=DEV)
        ops.conv_gemm_tc(a, tw, out=out, want_split=False, k_chunk=kc)
        errs[kc] = ((out.double() - ref).abs().max() / ref.abs().max()).item()
    assert errs[512] <= 1e-5 and errs[512] < 0.5 * errs[0], errs


def test_transpose_split_layout():
    x = torch.randn(2, 70, 45, device=DEV)        # (B, C, T) channel-major
    hi, lo = ops.transpose_split(x.contiguous(), 128)
    both = hi.float() + lo.float()
    assert both.shape == (2, 45, 128)
    assert (both[:, :, :70] - x.transpose(1, 2)).abs().max().item() <= 1e-6
    assert both[:, :, 70:].abs().max().item() == 0.0


def test_bad_arguments_fail_loudly():
    _, _, tw = make_weight(64, 64, 1, torch.float16)
    a = ops.pad_split(torch.randn(1, 8, 64, device=DEV), 64)
    with pytest.raises(_ext.FacError):
        ops.conv_gemm_tc((a[0][:, :, :32].contiguous(), a[1][:, :, :32].contiguous()), tw, out=torch.empty(1, 8, 64, device=DEV))
    with pytest.raises(_ext.FacError):
        ops.conv_gemm_tc(a, tw, want_split=False)          # no output requested
