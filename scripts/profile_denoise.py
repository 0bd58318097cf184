"""Where the denoise step's GPU time goes, by category, with CUDA events around every call of the hot-path host
functions (linears / FFN, prologue, attention, glue).  Events serialise nothing on the GPU but add small gaps, so the
sum is an upper bound of the step; use it for the SHARES.  Run under gpurun:
    python scripts/profile_denoise.py [1.3B|14B] > gpurun_out/denoise_profile.log      (UVB_LINEAR=0 for cuBLAS)"""
import collections
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from univid_b200 import _ext  # noqa: E402

mdl = importlib.import_module("univid_b200.wan.modules.model")
EVENTS = collections.defaultdict(list)


def wrap(owner, name, label=None):
    fn = getattr(owner, name)

    def timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        key = label(*a, **k) if callable(label) else (label or name)
        EVENTS[key].append((e0, e1))
        return out
    setattr(owner, name, timed)


def lin_label(mod, x, act=0):
    return f"linear {tuple(x.shape[-2:])}x{mod.out_features}" + ("+gelu" if act else "")


def main():
    cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "1.3B"]
    dev = torch.device("cuda")
    f, h, w = cfg["grid"]
    L = f * h * w
    wrap(mdl, "_lin", lin_label)
    wrap(_ext, "qk_norm_rope", "prologue (qk norm + rope)")
    wrap(_ext, "fmha_fwd", lambda q, k, v, **kw: f"attention Lk={k.shape[1]}")
    wrap(_ext, "block_glue", "block glue")
    orig_ffn = mdl._ffn_forward

    def ffn(ffn_mod, hh):
        if mdl._gemm_ok(ffn_mod[0], hh):
            return orig_ffn(ffn_mod, hh)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        y = ffn_mod[0](hh)
        e[1].record()
        y = ffn_mod[1](y)
        e[2].record()
        y = ffn_mod[2](y)
        e[3].record()
        EVENTS["cuBLAS ffn[0]"].append((e[0], e[1]))
        EVENTS["eager GELU"].append((e[1], e[2]))
        EVENTS["cuBLAS ffn[2]"].append((e[2], e[3]))
        return y
    mdl._ffn_forward = ffn
    torch.manual_seed(0)
    with torch.device(dev):
        zc = cfg.get("in_dim", 16)
        model = mdl.WanModel(model_type="t2v", dim=cfg["dim"], ffn_dim=cfg["ffn"], num_heads=cfg["heads"],
                             num_layers=cfg["layers"], text_len=cfg["text_len"], in_dim=zc, out_dim=zc).eval()
    gen = torch.Generator(device=dev).manual_seed(7)
    lat = [torch.randn(zc, f, h * 2, w * 2, device=dev, generator=gen)]
    ctx = [torch.randn(cfg["text_len"], 4096, device=dev, generator=gen)]
    tt = torch.full((1, L), 500.0, device=dev)           # per-token timesteps, like the reference loop
    if cfg.get("ti2v"):
        tt[0, :h * w] = 0.0
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        model(lat, tt, ctx, seq_len=L)
        torch.cuda.synchronize()
        EVENTS.clear()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model(lat, tt, ctx, seq_len=L)
        e1.record()
        torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    print(f"UVB_LINEAR={os.environ.get('UVB_LINEAR', '1')}  step {total:.2f} ms")
    acc = 0.0
    for key, evs in sorted(EVENTS.items(), key=lambda kv: -sum(a.elapsed_time(b) for a, b in kv[1])):
        ms = sum(a.elapsed_time(b) for a, b in evs)
        acc += ms
        print(f"  {key:44s} n={len(evs):4d}  total {ms:8.2f} ms  avg {ms / len(evs) * 1e3:8.1f} us  share {ms / total:6.1%}")
    print(f"  {'(not covered: embeddings, head, casts, gaps)':44s}        total {total - acc:8.2f} ms")


if __name__ == "__main__":
    main()
