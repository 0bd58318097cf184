"""Parity of the CUDA kernels (through the C ABI) with the CPU oracle on seeded inputs.  -m gpu.

Tolerance (BASELINE.json north_star): per-element max-abs <= 2e-2 and cosine >= 0.9999 against an
fp32 evaluation; the prologue is compared at bf16 resolution against the oracle's restatement of
WanRMSNorm + rope_apply (bit-exact except for isolated 1-ulp flips at rounding boundaries)."""
import pytest
import torch

from oracle import wan_attention_oracle as orc

pytestmark = pytest.mark.gpu

MAX_ABS, MIN_COS = 2e-2, 0.9999


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _check_attn(got, want, max_abs=MAX_ABS):
    got, want = got.float().cpu(), want.float().cpu()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    assert err <= max_abs, f"max-abs {err}"
    assert _cos(got, want) >= MIN_COS, f"cos {_cos(got, want)}"


def _bf16_close(got, want, frac_exact=0.98):
    """bf16 tensors equal up to 2 ulp of the rotated pair magnitude, and exactly equal almost everywhere."""
    got, want = got.float().cpu(), want.float().cpu()
    pair = want.unflatten(-1, (-1, 2)).abs().amax(-1, keepdim=True).expand(*want.shape[:-1], want.shape[-1] // 2, 2).flatten(-2)
    tol = 0.0079 * pair + 1e-3
    assert ((got - want).abs() <= tol).all(), f"max err {(got - want).abs().max()}"
    assert (got == want).float().mean().item() >= frac_exact


@pytest.fixture(scope="module")
def ext():
    from univid_b200 import _ext
    assert _ext.lib().uvb_version() == _ext.ABI_VERSION
    return _ext


def _cos_sin(dev):
    f = orc.make_freqs(128)
    return f, torch.stack([f.real, f.imag], dim=-1).float().contiguous().to(dev)


@pytest.mark.parametrize("heads,dtype", [(12, torch.bfloat16), (2, torch.float32), (40, torch.bfloat16),
                                         (24, torch.bfloat16), (3, torch.bfloat16), (16, torch.bfloat16)])
def test_qk_norm_rope_matches_oracle(ext, heads, dtype):
    dev = "cuda"
    g = torch.Generator().manual_seed(heads)
    dim = heads * 128
    b, l = 2, 130
    grid = torch.tensor([[2, 5, 13], [3, 6, 7]])       # 130 and 126 real tokens (4 pad tokens in sample 1)
    q = torch.randn(b, l, dim, generator=g).to(dtype)
    k = (0.5 * torch.randn(b, l, dim, generator=g)).to(dtype)
    wq, wk = 1 + 0.1 * torch.randn(dim, generator=g), 1 + 0.1 * torch.randn(dim, generator=g)
    freqs, cs = _cos_sin(dev)
    q_out, k_out = ext.qk_norm_rope(q.to(dev), k.to(dev), wq.to(dev), wk.to(dev), 1e-6, heads, cos_sin=cs,
                                    grid_sizes=grid)
    torch.cuda.synchronize()
    for got, x, w in ((q_out, q, wq), (k_out, k, wk)):
        want = orc.rope_apply(orc.rms_norm(x, w, 1e-6).view(b, l, heads, 128), grid, freqs).to(torch.bfloat16)
        _bf16_close(got, want)


@pytest.mark.parametrize("heads,b,l,groups", [(12, 1, 1000, 1), (12, 2, 131, 1), (24, 1, 517, 1), (40, 1, 333, 1),
                                              (12, 1, 259, 4), (40, 2, 77, 8), (12, 1, 3, 1)])
def test_streaming_prologue_is_bit_identical_to_the_row_kernels(ext, heads, b, l, groups):
    """uvb_set_knob(UVB_KNOB_PROLOGUE_PAIR, 2): the persistent bulk-copy-ring kernel.  Same arithmetic, same rounding
    points: bit-identical to the one-row-per-warp-group kernel (knob 0), which the test above pins to the oracle.
    Shapes: every product width (1536 / 3072 / 5120), ragged last stage, padding tokens, more and fewer rows than one
    stage per SM, and the Ulysses send layout (groups > 1) with a rank offset."""
    dev = "cuda"
    g = torch.Generator().manual_seed(heads * 100 + l)
    dim = heads * 128
    q = torch.randn(b, l, dim, generator=g).to(torch.bfloat16).to(dev)
    k = (0.5 * torch.randn(b, l, dim, generator=g)).to(torch.bfloat16).to(dev)
    wq, wk = (1 + 0.1 * torch.randn(dim, generator=g)).to(dev), (1 + 0.1 * torch.randn(dim, generator=g)).to(dev)
    _, cs = _cos_sin(dev)
    grid = [(max(1, (l - 2) // 12), 3, 4)] * b            # a few unrotated padding tokens at the end
    kw = dict(cos_sin=cs, grid_sizes=grid, groups=groups, tok_offset=5 if groups > 1 else 0)
    outs = {}
    for mode in (0, 1, 2):
        old = ext.set_knob("prologue_pair", mode)
        try:
            outs[mode] = ext.qk_norm_rope(q, k, wq, wk, 1e-6, heads, **kw)
        finally:
            ext.set_knob("prologue_pair", old)
    torch.cuda.synchronize()
    for mode in (1, 2):
        assert torch.equal(outs[mode][0], outs[0][0]) and torch.equal(outs[mode][1], outs[0][1]), mode
    # q alone (the query prologue of cross-attention; >= 4096 rows take the streaming kernel): norm only and norm + rope
    ql = torch.randn(1, 4500, dim, generator=g).to(torch.bfloat16).to(dev)
    gl = [(15, 10, 30)]
    for kwq in (dict(), dict(cos_sin=cs, grid_sizes=gl)):
        got = {}
        for mode in (0, 2):
            old = ext.set_knob("prologue_pair", mode)
            try:
                got[mode] = ext.qk_norm_rope(ql, None, wq, None, 1e-6, heads, **kwq)[0]
            finally:
                ext.set_knob("prologue_pair", old)
        assert torch.equal(got[0], got[2])
    # rotation only (qk_norm=False)
    old = ext.set_knob("prologue_pair", 2)
    try:
        a = ext.qk_norm_rope(q, k, None, None, 1e-6, heads, **kw)
    finally:
        ext.set_knob("prologue_pair", old)
    old = ext.set_knob("prologue_pair", 0)
    try:
        r = ext.qk_norm_rope(q, k, None, None, 1e-6, heads, **kw)
    finally:
        ext.set_knob("prologue_pair", old)
    assert torch.equal(a[0], r[0]) and torch.equal(a[1], r[1])


def test_qk_norm_without_rope_and_without_norm(ext):
    dev = "cuda"
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 77, 1536, generator=g).to(torch.bfloat16)
    w = 1 + 0.1 * torch.randn(1536, generator=g)
    _, k_out = ext.qk_norm_rope(None, x.to(dev), None, w.to(dev), 1e-6, 12)
    _bf16_close(k_out, orc.rms_norm(x, w, 1e-6).view(1, 77, 12, 128).to(torch.bfloat16))
    # qk_norm=False: rotation only
    freqs, cs = _cos_sin(dev)
    grid = torch.tensor([[1, 7, 11]])
    q_out, _ = ext.qk_norm_rope(x.to(dev), None, None, None, 1e-6, 12, cos_sin=cs, grid_sizes=grid)
    _bf16_close(q_out, orc.rope_apply(x.view(1, 77, 12, 128), grid, freqs).to(torch.bfloat16))


def test_qk_norm_rope_sequence_parallel_offset_and_send_layout(ext):
    """tok_offset reproduces the rank slice of the SP rope (sequence_parallel.py:46-55) and groups=p writes
    the Ulysses send layout [p, B, s, N/p, 128] (util.py:27 chunk on the head dimension)."""
    dev = "cuda"
    g = torch.Generator().manual_seed(9)
    heads, dim, world, s = 4, 512, 2, 32
    grid = torch.tensor([[3, 4, 5]])                    # 60 real tokens, padded to 64
    x = torch.randn(1, world * s, dim, generator=g).to(torch.bfloat16)
    w = 1 + 0.1 * torch.randn(dim, generator=g)
    freqs, cs = _cos_sin(dev)
    for r in range(world):
        xr = x[:, r * s:(r + 1) * s].contiguous()
        q_send, _ = ext.qk_norm_rope(xr.to(dev), None, w.to(dev), None, 1e-6, heads, cos_sin=cs, grid_sizes=grid,
                                     tok_offset=r * s, groups=world)
        assert q_send.shape == (world, 1, s, heads // world, 128)
        want = orc.sp_rope_apply(orc.rms_norm(xr, w, 1e-6).view(1, s, heads, 128), grid, freqs, r, world)
        want = torch.stack(want.to(torch.bfloat16).chunk(world, dim=2))
        _bf16_close(q_send, want)


def test_head_scatter(ext):
    v = torch.randn(2, 50, 6, 128, dtype=torch.bfloat16, device="cuda")
    out = ext.head_scatter(v, 3)
    assert torch.equal(out, torch.stack(v.chunk(3, dim=2)))


@pytest.mark.parametrize("b,lq,lk,n", [(1, 128, 128, 1), (1, 256, 384, 2), (2, 1950, 1950, 3), (1, 300, 77, 2),
                                       (1, 1, 1, 1), (3, 129, 513, 1), (1, 1950, 512, 12),
                                       # SURVEY 8f rank 4 shapes: Wan-Animate image branch (257 keys), ti2v-5B (24 heads)
                                       (1, 700, 257, 24), (2, 520, 520, 24)])
def test_fmha_matches_oracle(ext, b, lq, lk, n):
    g = torch.Generator().manual_seed(lq * 7 + lk)
    q, k, v = (torch.randn(b, l, n, 128, generator=g).to(torch.bfloat16) for l in (lq, lk, lk))
    got = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda())
    _check_attn(got, orc.attention_varlen(q, k, v, compute_dtype=torch.float32))


@pytest.mark.parametrize("b,lq,lk,n,lens", [(1, 2100, 2100, 1, None), (1, 4000, 2500, 2, None), (2, 513, 2200, 3, [2200, 100]),
                                             (1, 512, 2304, 1, None), (1, 130, 3000, 2, None), (1, 1025, 2049, 2, [2049]),
                                             (2, 700, 4500, 2, [0, 4500])])
def test_fmha_cta_pair_kernel(ext, b, lq, lk, n, lens):
    """Lk > 2048 runs the CTA-pair kernel (cta_group::2: 512-row units over two CTAs, K / V halves per CTA, remote
    p_full arrives): against the oracle, and against the single-CTA kernel selected through uvb_set_knob.  Shapes
    cover a ragged last unit (rows of only the leader CTA / only its first tile valid), ragged key tiles whose
    second 64-key half is entirely out of bounds, k_lens (incl. 0) and more units than CTA pairs."""
    g = torch.Generator().manual_seed(lq * 3 + lk)
    q, k, v = (torch.randn(b, l, n, 128, generator=g).to(torch.bfloat16) for l in (lq, lk, lk))
    kl = None if lens is None else torch.tensor(lens, dtype=torch.int32)
    klc = None if kl is None else kl.cuda()
    assert ext.lib().uvb_get_knob(ext.KNOBS["fmha_pair"]) == 1
    got = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=klc)
    got_nosplit = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=klc, split_units=False)
    old = ext.set_knob("fmha_pair", 0)
    try:
        single = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=klc)
    finally:
        ext.set_knob("fmha_pair", old)
    want = orc.attention_varlen(q, k, v, k_lens=kl, compute_dtype=torch.float32)
    _check_attn(got, want)
    _check_attn(got_nosplit, want)
    assert (got.float() - single.float()).abs().max().item() <= 4e-3


@pytest.mark.parametrize("b,lq,lk,n,lens", [(1, 1950, 512, 12, None), (2, 700, 257, 3, [257, 40]), (1, 513, 2048, 2, None),
                                             (1, 100, 77, 1, None), (3, 129, 513, 1, [513, 0, 1])])
def test_short_key_kernel_single_cta_and_cta_pair_variants(ext, b, lq, lk, n, lens):
    """Lk <= 2048 (cross-attention) runs the query-block-pipelined kernel; by default as CTA pairs (two Q/O buffers,
    half-size K/V stages, shared-memory P panel), with uvb_set_knob(xattn_pair, 0) as single CTAs.  Both against the
    oracle, with and without the per-key modifiers of the fused text weighting."""
    g = torch.Generator().manual_seed(lq + 31 * lk)
    q, k, v = (torch.randn(b, l, n, 128, generator=g).to(torch.bfloat16) for l in (lq, lk, lk))
    kl = None if lens is None else torch.tensor(lens, dtype=torch.int32)
    klc = None if kl is None else kl.cuda()
    w = torch.ones(lk)
    w[:lk // 4] = 1.25
    bias = 0.1 * torch.randn(n * 128, generator=g)
    want = orc.attention_varlen(q, k, v, k_lens=kl, compute_dtype=torch.float32)
    want_mod = orc.attention_varlen(q, k, v, k_lens=kl, compute_dtype=torch.float32, key_pv_weight=w, out_bias=bias)
    outs = {}
    for mode in (1, 0):
        old = ext.set_knob("xattn_pair", mode)
        try:
            outs[mode] = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=klc)
            got_mod = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=klc, key_pv_weight=w.cuda(), out_bias=bias.cuda())
        finally:
            ext.set_knob("xattn_pair", old)
        _check_attn(outs[mode], want)
        _check_attn(got_mod, want_mod)
    assert (outs[0].float() - outs[1].float()).abs().max().item() <= 4e-3


@pytest.mark.parametrize("b,lq,lk,n,lens", [(4, 6000, 512, 6, [512, 100, 0, 300]), (3, 9000, 257, 5, [257, 257, 129]),
                                             (2, 20000, 1000, 3, [128, 1000])])
def test_short_key_kernel_many_units_per_cta_pair(ext, b, lq, lk, n, lens):
    """More 512-row units than CTA pairs, so every pair walks several units back to back and the first score tiles of
    a unit are issued inside the last step of its predecessor (fmha_fwd_sm100.cuh, MMA warp).  The key lengths make
    consecutive units of one pair differ in their number of key tiles (4 -> 1 -> 0 -> 3, ...): primed and un-primed
    unit starts, units without keys, one-step units (never primed from) all occur.  Against the oracle and against
    the single-CTA variant, which starts every unit with its own prologue."""
    g = torch.Generator().manual_seed(lq + 31 * lk)
    q, k, v = (torch.randn(b, l, n, 128, generator=g).to(torch.bfloat16) for l in (lq, lk, lk))
    kl = torch.tensor(lens, dtype=torch.int32)
    bias = 0.1 * torch.randn(n * 128, generator=g)
    w = torch.ones(lk)
    w[:lk // 4] = 1.25
    got = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=kl.cuda())
    got_mod = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=kl.cuda(), key_pv_weight=w.cuda(), out_bias=bias.cuda())
    old = ext.set_knob("xattn_pair", 0)
    try:
        single = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=kl.cuda())
    finally:
        ext.set_knob("xattn_pair", old)
    _check_attn(got, orc.attention_varlen(q, k, v, k_lens=kl, compute_dtype=torch.float32))
    _check_attn(got_mod, orc.attention_varlen(q, k, v, k_lens=kl, compute_dtype=torch.float32, key_pv_weight=w, out_bias=bias))
    assert (got.float() - single.float()).abs().max().item() <= 4e-3
    again = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=kl.cuda())
    assert torch.equal(got, again)                                   # no dependence on timing


def test_fmha_randomized_shapes_all_variants_agree(ext):
    """24 seeded random problems (B 1-3, heads 1-5, 1 <= Lq <= 1500, 1 <= Lk <= 5000, random k_lens incl. 0 and Lk):
    the default kernels (CTA pairs for both key-length regimes) against the single-CTA variants and, on sampled query
    rows, against an fp32 evaluation; split and unsplit schedules."""
    import random
    rng = random.Random(1234)
    g = torch.Generator(device="cuda").manual_seed(99)
    for case in range(24):
        b, n = rng.randint(1, 3), rng.randint(1, 5)
        lq = rng.choice([1, 127, 128, 129, 255, 256, 257, 511, 512, 513, rng.randint(1, 1500)])
        lk = rng.choice([1, 64, 127, 128, 129, 2047, 2048, 2049, 2176, rng.randint(1, 5000)])
        q, k, v = (torch.randn(b, l, n, 128, device="cuda", generator=g).to(torch.bfloat16) for l in (lq, lk, lk))
        lens = [rng.choice([0, 1, lk, rng.randint(0, lk)]) for _ in range(b)] if rng.random() < 0.6 else None
        kl = None if lens is None else torch.tensor(lens, dtype=torch.int32, device="cuda")
        got = ext.fmha_fwd(q, k, v, k_lens=kl)
        got_ns = ext.fmha_fwd(q, k, v, k_lens=kl, split_units=False)
        old = (ext.set_knob("fmha_pair", 0), ext.set_knob("xattn_pair", 0))
        try:
            single = ext.fmha_fwd(q, k, v, k_lens=kl)
        finally:
            ext.set_knob("fmha_pair", old[0])
            ext.set_knob("xattn_pair", old[1])
        tag = (case, b, n, lq, lk, lens)
        assert torch.isfinite(got.float()).all(), tag
        assert (got.float() - single.float()).abs().max().item() <= 4e-3, tag
        assert (got.float() - got_ns.float()).abs().max().item() <= 4e-3, tag
        # fp32 evaluation of up to 48 sampled rows per sample
        rows = torch.tensor(sorted(set([0, lq - 1] + [rng.randrange(lq) for _ in range(46)])), device="cuda")
        for bi in range(b):
            k_len = lk if lens is None else lens[bi]
            qf = q[bi, rows].float().transpose(0, 1)                         # [N, R, D]
            sc = torch.matmul(qf, k[bi].float().permute(1, 2, 0)) * 128 ** -0.5
            sc[:, :, k_len:] = float("-inf")
            pr = torch.nan_to_num(torch.softmax(sc, dim=-1), nan=0.0)
            want = torch.matmul(pr, v[bi].float().transpose(0, 1)).transpose(0, 1)     # [R, N, D]
            err = (got[bi, rows].float() - want).abs().max().item()
            assert err <= 2e-2, (tag, bi, err)


def test_fmha_peaked_logits_exercise_the_lazy_rescale(ext):
    """Row maxima that grow by far more than 2^8 from one key tile to the next force the in-place
    rescale of the TMEM accumulator."""
    g = torch.Generator().manual_seed(1)
    lq, lk = 256, 1024
    _peaked(ext, g, lq, lk)
    _peaked(ext, g, 600, 2560)          # CTA-pair kernel


def _peaked(ext, g, lq, lk):
    q = torch.randn(1, lq, 2, 128, generator=g)
    k = torch.randn(1, lk, 2, 128, generator=g)
    k = k * torch.linspace(0.2, 6.0, lk).view(1, lk, 1, 1)      # logits grow along the key axis
    v = torch.randn(1, lk, 2, 128, generator=g)
    q, k, v = (u.to(torch.bfloat16) for u in (q, k, v))
    got = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda())
    _check_attn(got, orc.attention_varlen(q, k, v, compute_dtype=torch.float32), max_abs=4e-2)


@pytest.mark.parametrize("lens", [[1000, 1950], [1, 128], [129, 0], [1950, 1950]])
def test_fmha_key_lengths(ext, lens):
    g = torch.Generator().manual_seed(sum(lens))
    q, k, v = (torch.randn(2, 1950 if i else 300, 2, 128, generator=g).to(torch.bfloat16) for i in range(3))
    kl = torch.tensor(lens, dtype=torch.int32)
    got = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), k_lens=kl.cuda())
    _check_attn(got, orc.attention_varlen(q, k, v, k_lens=kl, compute_dtype=torch.float32))


def test_fmha_strided_views_and_scale(ext):
    """q/k/v as slices of one fused [B, L, 3, N, 128] projection and of a wider head tensor; custom scale."""
    g = torch.Generator().manual_seed(2)
    qkv = torch.randn(1, 200, 3, 4, 128, generator=g).to(torch.bfloat16).cuda()
    q, k, v = qkv[:, :, 0, 1:3], qkv[:, :, 1, 1:3], qkv[:, :, 2, 1:3]
    out = torch.zeros(1, 200, 4, 128, dtype=torch.bfloat16, device="cuda")
    ext.fmha_fwd(q, k, v, softmax_scale=0.05, out=out[:, :, 2:4])
    want = orc.attention_varlen(q.cpu(), k.cpu(), v.cpu(), softmax_scale=0.05, compute_dtype=torch.float32)
    _check_attn(out[:, :, 2:4], want)
    assert torch.count_nonzero(out[:, :, 0:2]) == 0


def test_xattn_key_modifiers(ext):
    g = torch.Generator().manual_seed(3)
    lq, lk, n = 500, 512, 3
    q, k, v = (torch.randn(1, l, n, 128, generator=g).to(torch.bfloat16) for l in (lq, lk, lk))
    temp = torch.ones(lk)
    temp[:128] = 1.3
    w = torch.ones(lk)
    w[:128] = 1.25
    bias = 0.1 * torch.randn(n * 128, generator=g)
    got = ext.fmha_fwd(q.cuda(), k.cuda(), v.cuda(), key_logit_scale=temp.cuda(), key_pv_weight=w.cuda(),
                       out_bias=bias.cuda())
    want = orc.attention_varlen(q, k, v, compute_dtype=torch.float32, key_logit_scale=temp, key_pv_weight=w,
                                out_bias=bias)
    _check_attn(got, want)
    # each modifier alone; Lk not a multiple of the key tile
    k2, v2 = k[:, :300], v[:, :300]
    got = ext.fmha_fwd(q.cuda(), k2.cuda(), v2.cuda(), key_pv_weight=w[:300].cuda())
    _check_attn(got, orc.attention_varlen(q, k2, v2, compute_dtype=torch.float32, key_pv_weight=w[:300]))
    got = ext.fmha_fwd(q.cuda(), k2.cuda(), v2.cuda(), key_logit_scale=temp[:300].cuda())
    _check_attn(got, orc.attention_varlen(q, k2, v2, compute_dtype=torch.float32, key_logit_scale=temp[:300]))


def test_unsupported_shapes_raise(ext):
    q = torch.zeros(1, 8, 2, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(NotImplementedError):
        ext.fmha_fwd(q, q, q)
    q = torch.zeros(1, 8, 2, 128, dtype=torch.float16, device="cuda")
    with pytest.raises(NotImplementedError):
        ext.fmha_fwd(q, q, q)


@pytest.mark.parametrize("p,s,heads,batch", [(2, 120, 4, 1), (4, 60, 4, 1), (2, 256, 2, 2), (8, 45, 8, 1), (2, 130, 6, 1),
                                             # > 2048 keys: the CTA-pair attention kernel does the peer stores
                                             (2, 1100, 2, 1), (4, 640, 4, 1), (8, 300, 8, 2)])
def test_sp_kernels_with_local_peers(ext, p, s, heads, batch):
    """The fused Ulysses exchange kernels (uvb_*_sp) on ONE GPU: the p 'peer' buffers are p local buffers, and the
    p ranks run one after the other.  Checks the peer-store addressing of the prologue / head scatter, the
    per-rank clipped TMA stores of the attention epilogue (tiles straddling token chunks), and the flag kernels,
    against the plain kernels on the unsharded problem."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(p * 1000 + s)
    n, L, dim = heads // p, p * s, heads * 128
    bf = torch.bfloat16
    q_lin = torch.randn(batch, L, dim, generator=g).to(bf).to(dev)
    k_lin = torch.randn(batch, L, dim, generator=g).to(bf).to(dev)
    v = torch.randn(batch, L, heads, 128, generator=g).to(bf).to(dev)
    wq = (1 + 0.1 * torch.randn(dim, generator=g)).to(dev)
    wk = (1 + 0.1 * torch.randn(dim, generator=g)).to(dev)
    _, cs = _cos_sin(dev)
    f = 2
    grid = [(f, 5, (L - 7) // (f * 5))] * batch                     # a few trailing padding tokens
    k_lens = torch.tensor([L - 3] * batch, dtype=torch.int32, device=dev)
    # unsharded reference through the plain kernels
    q_ref, k_ref = ext.qk_norm_rope(q_lin, k_lin, wq, wk, 1e-6, heads, cos_sin=cs, grid_sizes=grid)
    o_ref = ext.fmha_fwd(q_ref, k_ref, v, k_lens=k_lens)
    # "rank" j's buffers: q/k/v_recv [B, p, s, n, 128] == [B, L, n, 128]; o_recv [B, s, N, 128]
    recv = [{t: torch.zeros(batch, L, n, 128, dtype=bf, device=dev) for t in "qkv"} for _ in range(p)]
    o_recv = [torch.full((batch, s, heads, 128), float("nan"), dtype=bf, device=dev) for _ in range(p)]
    flags = torch.zeros(p, 32, dtype=torch.int32, device=dev)
    send_sb, send_sl = p * s * n * 128, n * 128
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(p):                                               # producers of rank i
        sl = slice(i * s, (i + 1) * s)
        off = i * s * n * 128 * 2
        qp = ext.ptr_array([recv[j]["q"].data_ptr() + off for j in range(p)])
        kp = ext.ptr_array([recv[j]["k"].data_ptr() + off for j in range(p)])
        vp = ext.ptr_array([recv[j]["v"].data_ptr() + off for j in range(p)])
        ext.qk_norm_rope(q_lin[:, sl].contiguous(), k_lin[:, sl].contiguous(), wq, wk, 1e-6, heads, cos_sin=cs,
                         grid_sizes=grid, tok_offset=i * s, groups=p, peers=(qp, kp, send_sb, send_sl))
        ext.head_scatter(v[:, sl].contiguous(), p, peers=(vp, send_sb, send_sl))
        ext.sp_signal(ext.ptr_array([flags[j].data_ptr() + 4 * i for j in range(p)]), p, 7, stream)
    for j in range(p):                                               # consumer of rank j
        ext.sp_wait(flags[j].data_ptr(), p, 7, stream)
        hs = slice(j * n, (j + 1) * n)
        assert torch.equal(recv[j]["q"], q_ref[:, :, hs]), "q exchange"
        assert torch.equal(recv[j]["k"], k_ref[:, :, hs]), "k exchange"
        assert torch.equal(recv[j]["v"], v[:, :, hs]), "v exchange"
        ext.fmha_fwd_sp(recv[j]["q"], recv[j]["k"], recv[j]["v"], ext.ptr_array([o.data_ptr() for o in o_recv]), p,
                        j * n, heads, k_lens=k_lens)
    torch.cuda.synchronize()
    got = torch.cat(o_recv, dim=1)                                   # [B, L, N, 128]
    assert torch.isfinite(got.float()).all(), "an output row was not written"
    if L <= 2048:
        assert torch.equal(got, o_ref)
    else:
        # long key sequences split the remainder units over the key axis (stream-K); the split points depend on the
        # number of heads in the launch, so the per-rank launches merge their partials in a different order
        assert (got.float() - o_ref.float()).abs().max().item() <= 4e-3


@pytest.mark.parametrize("b,lq,lk,n,lens", [(1, 1950, 1950, 12, None), (2, 700, 2300, 3, [2300, 130]), (1, 40000, 512, 5, None),
                                             (1, 9000, 4000, 20, None), (3, 300, 300, 2, [0, 300, 17])])
def test_fmha_split_and_unsplit_schedules_agree(ext, b, lq, lk, n, lens):
    """The persistent kernel with the stream-K split of remainder query blocks (workspace) against the schedule
    that never splits (workspace = NULL), and both against the oracle on sampled rows.  Repeated launches reuse
    the workspace flags (they must come back zeroed)."""
    g = torch.Generator().manual_seed(lq + lk)
    q, k, v = (torch.randn(b, l, n, 128, generator=g).to(torch.bfloat16).cuda() for l in (lq, lk, lk))
    kl = None if lens is None else torch.tensor(lens, dtype=torch.int32, device="cuda")
    ref = ext.fmha_fwd(q, k, v, k_lens=kl, split_units=False)
    for _ in range(3):
        got = ext.fmha_fwd(q, k, v, k_lens=kl)
        assert torch.isfinite(got.float()).all()
        assert (got.float() - ref.float()).abs().max().item() <= 4e-3      # same math, different merge order
    rows = torch.randperm(lq, generator=g)[:64]
    want = orc.attention_varlen(q[:, rows].float().cpu(), k.float().cpu(), v.float().cpu(),
                                k_lens=None if lens is None else torch.tensor(lens), compute_dtype=torch.float32)
    _check_attn(got[:, rows], want)
