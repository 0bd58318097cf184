"""uvb_linear_bf16 (tcgen05 GEMM) on the GPU.  -m gpu.

The reference op is nn.Linear under bf16 autocast (model.py:119-122, :212-214), optionally followed by
nn.GELU(approximate='tanh') (:213): bf16 operands, fp32 accumulation, one rounding to bf16 (and a second one after
the activation).  The kernel is compared with a plain PyTorch fp32 evaluation of the same chain on the same bf16
operands: every element within 2 bf16 ulp (1.6e-2 relative to max(1, |ref|)) and >= 99 % of the elements bit-equal
to the rounded fp32 result (the remaining differences are 1-ulp flips from the summation order)."""
import importlib

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _ref(x, w, b, act):
    y = x.double() @ w.double().t()
    if b is not None:
        y = y + b.double()
    y = y.float().to(torch.bfloat16)
    if act:
        y = torch.nn.functional.gelu(y.float(), approximate="tanh").to(torch.bfloat16)
    return y


def _close(got, want, frac=0.99):
    got, want = got.float().cpu(), want.float().cpu()
    assert torch.isfinite(got).all()
    err = ((got - want).abs() / want.abs().clamp_min(1.0)).max().item()
    assert err <= 1.6e-2, f"max relative error {err}"
    assert (got == want).float().mean().item() >= frac


@pytest.mark.parametrize("M,N,K,bias,act", [
    (1, 1536, 1536, True, 0),          # a single row (the bias probe of the text-weighted cross-attention)
    (128, 256, 64, False, 0),          # exactly one tile, one k block
    (100, 264, 200, True, 1),          # ragged in every dimension (K not a multiple of the 64-wide k block)
    (512, 1536, 1536, True, 0),        # context projections of the 1.3B model
    (1950, 1536, 1536, True, 0),       # BASELINE configs[0] token count
    (1950, 8960, 1536, True, 1),       # ffn[0] + GELU
    (1950, 1536, 8960, True, 0),       # ffn[2]
    (700, 5120, 5120, True, 0),        # 14B width
    (300, 13824, 5120, True, 1),       # 14B ffn
])
def test_linear_matches_fp32(M, N, K, bias, act):
    from univid_b200 import _ext
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = (0.5 * torch.randn(N, generator=g)).to(torch.bfloat16).float() if bias else None
    got = _ext.linear(x.cuda(), w.cuda(), None if b is None else b.cuda(), act=act)
    assert got.shape == (M, N) and got.dtype == torch.bfloat16
    _close(got, _ref(x, w, b, act))


@pytest.mark.parametrize("ctas", ["1", "2"])
def test_both_kernel_variants_in_the_standalone_binary(ctas):
    """uvb_set_knob(UVB_KNOB_GEMM_CTAS, 1 | 2) selects one CTA per tile or CTA pairs (default): run the C battery case
    of the stand-alone binary under each setting (the test binary maps UVB_KNOBS onto uvb_set_knob; it checks
    against a naive fp32-accumulate kernel)."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "univid_b200", "csrc", "tests", "uvb_test")
    assert os.path.exists(exe), "build the test binary with `python -m univid_b200.build`"
    for case in (["gemm", "1000", "1536", "1536", "1", "0"], ["gemm", "520", "2296", "200", "0", "0"]):
        r = subprocess.run([exe] + case, env=dict(os.environ, UVB_KNOBS=f"gemm_ctas={ctas}"), capture_output=True, text=True,
                           timeout=120)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr


def test_linear_strided_rows_and_batch_shape():
    """x may be a row-strided view (a slice of a fused projection) and carry leading batch dimensions."""
    from univid_b200 import _ext
    g = torch.Generator().manual_seed(5)
    big = torch.randn(2, 70, 3 * 256, generator=g).to(torch.bfloat16).cuda()
    x = big[:, :, 256:512]                       # row stride 768, not contiguous
    w = (torch.randn(512, 256, generator=g) / 16).to(torch.bfloat16).cuda()
    got = _ext.linear(x, w)
    assert got.shape == (2, 70, 512)
    _close(got, _ref(x.cpu().reshape(-1, 256), w.cpu(), None, 0).view(2, 70, 512))


def test_linear_rejects_what_it_cannot_do():
    from univid_b200 import _ext
    x = torch.zeros(4, 64, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(64, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(NotImplementedError):
        _ext.linear(x.float(), w)
    with pytest.raises(NotImplementedError):
        _ext.linear(x[:, :60].contiguous(), w[:, :60].contiguous())
    with pytest.raises(ValueError):
        _ext.linear(x, w[:, :32].contiguous())
    with pytest.raises(RuntimeError):
        _ext.linear(x.cpu(), w)
    with pytest.raises(RuntimeError):
        with torch.enable_grad():
            _ext.linear(x.requires_grad_(), w)


def test_module_projections_use_the_gemm_and_match_cublas():
    """_lin(mod, x) == mod(x) under bf16 autocast (cuBLAS) up to summation-order flips; a wrapped projection
    (LoRA-style subclass) and a trainable one go through the module."""
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    from univid_b200 import _ext
    torch.manual_seed(0)
    lin = nn.Linear(1536, 1536).cuda()
    nn.init.normal_(lin.bias, std=0.5)
    x = torch.randn(3, 333, 1536, device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert mdl._gemm_ok(lin, x)
        n0 = _ext.launch_count
        got = mdl._lin(lin, x)
        assert _ext.launch_count == n0 + 1
        want = lin(x)
    assert got.dtype == want.dtype == torch.bfloat16
    _close(got, want, frac=0.98)

    class Wrapped(nn.Linear):
        pass
    wl = Wrapped(64, 64).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        assert not mdl._gemm_ok(wl, x[..., :64])
        n0 = _ext.launch_count
        mdl._lin(wl, x[..., :64])
        assert _ext.launch_count == n0
    with torch.autocast("cuda", dtype=torch.bfloat16):       # autograd on: the module runs, gradients flow
        assert not mdl._gemm_ok(lin, x)
        y = mdl._lin(lin, x)
        assert y.requires_grad
    # parameter update invalidates the cached bf16 copy
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        lin.weight.mul_(2.0)
        _close(mdl._lin(lin, x), lin(x), frac=0.98)


def test_ffn_fused_gelu_matches_module():
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    torch.manual_seed(1)
    ffn = nn.Sequential(nn.Linear(1536, 8960), nn.GELU(approximate="tanh"), nn.Linear(8960, 1536)).cuda()
    h = torch.randn(1, 777, 1536, device="cuda").to(torch.bfloat16)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        got = mdl._ffn_forward(ffn, h)
        want = ffn(h)
    err = (got.float() - want.float()).abs().max().item()
    assert err <= 2e-2, err
    a, b = got.double().flatten(), want.double().flatten()
    assert float(a @ b / (a.norm() * b.norm())) >= 0.99999
    exact = nn.Sequential(nn.Linear(64, 64), nn.GELU(), nn.Linear(64, 64)).cuda()     # erf GELU: not fused
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        x = torch.randn(5, 64, device="cuda")
        assert torch.equal(mdl._ffn_forward(exact, x), exact(x))


@pytest.mark.parametrize("M,N,K,groups", [(300, 1024, 256, 4), (1000, 1536, 1536, 2), (257, 5120, 512, 8), (4000, 1536, 512, 4)])
def test_linear_column_group_scatter_equals_linear_plus_head_scatter(M, N, K, groups):
    """uvb_linear_bf16_sp: the GEMM epilogue stores column group j ([M, N/groups]) through its own pointer -- on one
    GPU the 'peers' are local buffers laid out like the Ulysses exchange slots ([B=1, p, s, n, 128] with the slot of
    'rank' 1) -- bit-identical to uvb_linear_bf16 followed by the head scatter."""
    from univid_b200 import _ext
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(1, M, K, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    b = (0.5 * torch.randn(N, generator=g)).to(torch.bfloat16).float().cuda()
    n = N // groups // 128
    want = _ext.head_scatter(_ext.linear(x, w, b).view(1, M, N // 128, 128), groups)          # [groups, 1, M, n, 128]
    p, slot = 3, 1                                                                             # 3 "ranks", we are rank 1
    recv = [torch.full((1, p, M, n, 128), float("nan"), dtype=torch.bfloat16, device="cuda") for _ in range(groups)]
    ptrs = _ext.ptr_array([r.data_ptr() + slot * M * n * 128 * 2 for r in recv])
    assert _ext.linear(x, w, b, peers=(ptrs, groups, n * 128)) is None
    torch.cuda.synchronize()
    for j in range(groups):
        assert torch.equal(recv[j][0, slot], want[j, 0]), j
        assert torch.isnan(recv[j][0, 0].float()).all() and torch.isnan(recv[j][0, 2].float()).all()   # other slots untouched
