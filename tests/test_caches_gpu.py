"""The host-side caches around the kernels never serve stale data (cached bf16 GEMM operands; round 2: the
prompt-side caches of WanModel / WanCrossAttention).  -m gpu."""
import importlib

import pytest
import torch

pytestmark = pytest.mark.gpu

mdl = importlib.import_module("univid_b200.wan.modules.model")


def _launches():
    from univid_b200 import _ext
    return _ext.launch_count


def _sa(dim=256, heads=2):
    torch.manual_seed(0)
    sa = mdl.WanSelfAttention(dim, heads).cuda().eval()
    for lin in (sa.q, sa.k, sa.v, sa.o):
        torch.nn.init.normal_(lin.bias, std=0.02)
    return sa


def _run(sa, x):
    from oracle import wan_attention_oracle as orc
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return sa(x, torch.tensor([x.size(1)]), torch.tensor([[2, 4, 5]]), orc.make_freqs(128).cuda()).float()


def test_gemm_operand_cache_follows_parameter_updates_and_is_dropped_by_apply():
    sa = _sa()
    x = torch.randn(1, 40, 256, device="cuda")
    y0 = _run(sa, x)
    assert all(mdl._LINEAR_ATTR in lin.__dict__ for lin in (sa.q, sa.k, sa.v, sa.o))       # fp32 params: bf16 copies cached
    assert torch.equal(_run(sa, x), y0)
    with torch.no_grad():
        sa.v.weight.mul_(2.0)                       # in-place update through the parameter: _version bumps
    y1 = _run(sa, x)
    assert (y1 - y0).abs().max() > 1e-3
    fresh = _sa()
    with torch.no_grad():
        fresh.v.weight.mul_(2.0)
    assert torch.equal(_run(fresh, x), y1)          # the updated weights were used, not the cached copy
    sa.load_state_dict(_sa().state_dict())          # copy_ into the parameters: back to the original
    assert torch.equal(_run(sa, x), y0)
    # .data edits bypass every version counter: documented, and clear_linear_cache() is the remedy
    sa.v.weight.data.mul_(2.0)
    stale = _run(sa, x)
    assert torch.equal(stale, y0)
    assert mdl.clear_linear_cache(sa) == 4
    assert torch.equal(_run(sa, x), y1)
    # moving the module drops the copies at once (offload_model=True must free the GPU memory)
    _run(sa, x)
    sa.cpu()
    assert not any(mdl._LINEAR_ATTR in lin.__dict__ for lin in (sa.q, sa.k, sa.v, sa.o))
    sa.cuda()
    assert torch.equal(_run(sa, x), y1)


def test_bf16_parameters_are_used_in_place():
    sa = _sa().to(torch.bfloat16)
    x = torch.randn(1, 40, 256, device="cuda", dtype=torch.bfloat16)
    _run(sa, x)
    w, _ = mdl._linear_operands(sa.q)
    assert w.data_ptr() == sa.q.weight.data_ptr()   # no copy of a bf16 weight


def test_context_side_cache_hits_on_the_same_tensor_and_misses_on_any_change():
    torch.manual_seed(1)
    ca = mdl.WanCrossAttention(256, 2).cuda().eval()
    for lin in (ca.q, ca.k, ca.v, ca.o):
        torch.nn.init.normal_(lin.bias, std=0.02)
    x = torch.randn(1, 300, 256, device="cuda")
    ctx = torch.randn(1, 64, 256, device="cuda")

    def run(c, **kw):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            n0 = _launches()
            out = ca(x, c, None, **kw).float()
        return out, _launches() - n0

    y0, cold = run(ctx)
    y1, warm = run(ctx)
    assert torch.equal(y0, y1) and warm < cold       # k / v projections and the k prologue were reused
    f0, cold_f = run(ctx, text_weight=1.3, text_len=16)
    f1, warm_f = run(ctx, text_weight=1.3, text_len=16)
    f2, _ = run(ctx, text_weight=1.1, text_len=16)   # another weight: same cached projections, different result
    assert torch.equal(f0, f1) and warm_f < cold_f and (f2 - f0).abs().max() > 1e-4
    # a different tensor with the same values misses (identity key) but gives the same result
    y2, n2 = run(ctx.clone())
    assert torch.equal(y2, y0) and n2 == cold
    # in-place edit of the SAME tensor bumps its version: miss, new result
    ctx.mul_(1.5)
    y3, n3 = run(ctx)
    assert n3 == cold and (y3 - y0).abs().max() > 1e-3
    fresh, _ = run(ctx.clone())
    assert torch.equal(fresh, y3)
    # weights changed: miss
    with torch.no_grad():
        ca.k.weight.mul_(0.5)
    y4, n4 = run(ctx)
    assert n4 == cold and (y4 - y3).abs().max() > 1e-4
    # disabled cache
    ca.context_cache_size = 0
    _, n5 = run(ctx)
    _, n6 = run(ctx)
    assert n5 == n6 == cold


def test_wanmodel_reuses_the_embedded_prompt_across_calls():
    torch.manual_seed(2)
    m = mdl.WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64, freq_dim=32, in_dim=4, out_dim=4)
    torch.nn.init.normal_(m.head.head.weight, std=0.02)
    m = m.cuda().eval()
    lat = [torch.randn(4, 3, 8, 12, device="cuda")]
    ctx, ctx_null = [torch.randn(20, 64, device="cuda")], [torch.zeros(20, 64, device="cuda")]
    t = torch.tensor([500.0], device="cuda")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a0 = m(lat, t, ctx, 80)[0]
        n0 = _launches()
        a1 = m(lat, t, ctx, 80)[0]
        warm = _launches() - n0
        b0 = m(lat, t, ctx_null, 80)[0]                # CFG: the second prompt has its own entry
        n0 = _launches()
        a2 = m(lat, t, ctx, 80)[0]
        b1 = m(lat, t, ctx_null, 80)[0]
        both = _launches() - n0
    assert torch.equal(a0, a1) and torch.equal(a0, a2) and torch.equal(b0, b1)
    assert (a0 - b0).abs().max() > 1e-4
    assert both == 2 * warm                            # alternating prompts keep hitting
    ctx[0].add_(1.0)                                   # edited prompt tensor: recomputed
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        c0 = m(lat, t, ctx, 80)[0]
    assert (c0 - a0).abs().max() > 1e-4
