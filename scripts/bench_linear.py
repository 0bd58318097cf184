"""uvb_linear_bf16 (tcgen05 GEMM) against cuBLAS (torch F.linear, bf16) on the projection / FFN shapes of the
BASELINE configs, one B200.  CUDA events, L2 flushed between iterations (a 256 MB memset), best and mean of N;
plus a back-to-back run of 40 launches per shape (power-capped, like inside a denoise step).
Run under gpurun:  python scripts/bench_linear.py > gpurun_out/linear.log"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from univid_b200 import _ext  # noqa: E402

SHAPES = [  # (name, M, N, K, act)
    ("1.3B q/k/v/o", 32760, 1536, 1536, 0),
    ("1.3B ffn[0]+GELU", 32760, 8960, 1536, 1),
    ("1.3B ffn[2]", 32760, 1536, 8960, 0),
    ("1.3B ctx k/v", 512, 1536, 1536, 0),
    ("14B q/k/v/o", 75600, 5120, 5120, 0),
    ("14B ffn[0]+GELU", 75600, 13824, 5120, 1),
    ("14B ffn[2]", 75600, 5120, 13824, 0),
    ("14B/8 q/k/v/o (Ulysses shard)", 9450, 5120, 5120, 0),
]


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(iters):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def sustained(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    print(f"{'shape':34s} {'uvb best':>9s} {'mean':>7s} {'TF/s':>7s} | {'cuBLAS best':>11s} {'mean':>7s} {'TF/s':>7s} | "
          f"{'uvb b2b':>8s} {'cuBLAS b2b':>10s}   (ms; cuBLAS column includes the separate GELU kernel where fused here)")
    for name, M, N, K, act in SHAPES:
        x = torch.randn(M, K, device=dev, generator=g).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).to(torch.bfloat16)
        b = torch.randn(N, device=dev, generator=g).to(torch.bfloat16)
        bf = b.float()
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        ours = lambda: _ext.linear(x, w, bf, act=act, out=out)
        if act:
            ref = lambda: F.gelu(F.linear(x, w, b), approximate="tanh")
        else:
            ref = lambda: F.linear(x, w, b)
        iters = 10 if M * N * K < 3e12 else 5
        ob, om = timed(ours, iters, flush)
        rb, rm = timed(ref, iters, flush)
        so, sr = sustained(ours, 40), sustained(ref, 40)
        flop = 2.0 * M * N * K
        print(f"{name:34s} {ob:9.3f} {om:7.3f} {flop / ob * 1e-9:7.0f} | {rb:11.3f} {rm:7.3f} {flop / rb * 1e-9:7.0f} | "
              f"{so:8.3f} {sr:10.3f}")
        err = (ours().float() - ref().float()).abs().max().item()
        print(f"{'':34s} max |uvb - cuBLAS| = {err:.4f}")
        del x, w, out


if __name__ == "__main__":
    main()
