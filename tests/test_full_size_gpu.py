"""Parity at BASELINE.json's full sizes through size-independent properties.  -m gpu.

A full fp32 evaluation of 32 760 x 32 760 x 12 heads is out of reach for the CPU oracle, so the full-size
runs are checked by (i) exact re-evaluation of a sample of query rows against ALL keys in fp32 on the GPU
(plain matmul + softmax, the oracle's formula), (ii) sum-to-one (V = 1 must give exactly 1), (iii) linearity
in V, (iv) key-permutation invariance, and the prologue by (v) norm preservation of the rotation."""
import pytest
import torch

from oracle import wan_attention_oracle as orc

pytestmark = pytest.mark.gpu


def _rows_reference(q, k, v, rows, k_len=None):
    """fp32 softmax(q k^T / sqrt(d)) v for selected query rows; q/k/v [1, L, N, 128] bf16 on the GPU."""
    qf = q[0, rows].float().transpose(0, 1)                    # [N, R, D]
    kf, vf = k[0].float().transpose(0, 1), v[0].float().transpose(0, 1)
    s = torch.matmul(qf, kf.transpose(1, 2)) * 128 ** -0.5
    if k_len is not None:
        s[:, :, k_len:] = float("-inf")
    return torch.matmul(torch.softmax(s, dim=-1), vf).transpose(0, 1)      # [R, N, D]


@pytest.mark.parametrize("L,N", [(32760, 12), (75600, 5)])
def test_self_attention_full_length_sampled_rows(L, N):
    from univid_b200 import _ext
    g = torch.Generator(device="cuda").manual_seed(L)
    q, k, v = (torch.randn(1, L, N, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(3))
    out = _ext.fmha_fwd(q, k, v)
    rows = torch.tensor([0, 1, 127, 128, 255, 256, 4097, L // 2, L - 257, L - 129, L - 2, L - 1], device="cuda")
    want = _rows_reference(q, k, v, rows)
    got = out[0, rows].float()
    assert (got - want).abs().max().item() <= 2e-2
    cos = torch.nn.functional.cosine_similarity(got.flatten(), want.flatten(), dim=0).item()
    assert cos >= 0.9999, cos
    # masked tail (k_lens < L): same rows against the truncated key set
    k_len = L - 1000
    out2 = _ext.fmha_fwd(q, k, v, k_lens=torch.tensor([k_len], dtype=torch.int32, device="cuda"))
    want2 = _rows_reference(q, k, v, rows, k_len)
    assert (out2[0, rows].float() - want2).abs().max().item() <= 2e-2


def test_full_length_sum_to_one_linearity_and_permutation():
    from univid_b200 import _ext
    L, N = 32760, 4
    g = torch.Generator(device="cuda").manual_seed(1)
    q, k, v1, v2 = (torch.randn(1, L, N, 128, device="cuda", generator=g).to(torch.bfloat16) for _ in range(4))
    ones = torch.ones_like(v1)
    o = _ext.fmha_fwd(q, k, ones)
    assert (o.float() - 1).abs().max().item() <= 2 ** -7          # one bf16 ulp of 1.0
    o1, o2 = _ext.fmha_fwd(q, k, v1).float(), _ext.fmha_fwd(q, k, v2).float()
    o12 = _ext.fmha_fwd(q, k, (v1.float() + v2.float()).to(torch.bfloat16)).float()
    assert (o12 - (o1 + o2)).abs().max().item() <= 2e-2
    perm = torch.randperm(L, device="cuda", generator=g)
    op = _ext.fmha_fwd(q, k[:, perm].contiguous(), v1[:, perm].contiguous()).float()
    assert (op - o1).abs().max().item() <= 2e-2


@pytest.mark.parametrize("L,N,grid", [(32760, 12, (21, 30, 52)), (75600, 40, (21, 45, 80))])
def test_prologue_full_size_norm_preservation(L, N, grid):
    """With unit weights every output row has RMS 1 over the model width (rotations preserve norms), and the
    first rows agree with the oracle."""
    from univid_b200 import _ext
    dim = N * 128
    g = torch.Generator(device="cuda").manual_seed(2)
    q = torch.randn(1, L, dim, device="cuda", generator=g).to(torch.bfloat16)
    k = (3 * torch.randn(1, L, dim, device="cuda", generator=g)).to(torch.bfloat16)
    w = torch.ones(dim, device="cuda")
    f = orc.make_freqs(128)
    cs = torch.stack([f.real, f.imag], dim=-1).float().contiguous().cuda()
    qo, ko = _ext.qk_norm_rope(q, k, w, w, 1e-6, N, cos_sin=cs, grid_sizes=[grid])
    for o in (qo, ko):
        rms = o.float().flatten(2).square().mean(-1).sqrt()
        assert (rms - 1).abs().max().item() <= 4e-3
    rows = slice(L - 64, L)
    tail_q = q[:, rows].cpu()
    want = orc.rms_norm(tail_q, torch.ones(dim), 1e-6).view(1, 64, N, 128)
    full = orc._rotate(want[0], orc._token_rotations(*grid, f)[L - 64:L]).float().to(torch.bfloat16)
    got = qo[0, rows].cpu()
    assert ((got.float() - full.float()).abs() <= 0.0079 * full.float().abs().clamp(min=0.2) + 1e-3).all()
