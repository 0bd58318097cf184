#!/usr/bin/env python
"""Stand-alone timing of uvb_block_glue at the BASELINE geometries (HBM-bound): GB/s of algorithmic bytes
(read x 4 + y 2, write x 4 + h 2 per element for the full form).  Prints one line per (config, form)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from univid_b200 import _ext  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, L, dim in (("1.3B", 32760, 1536), ("14B", 75600, 5120)):
    x = torch.randn(1, L, dim, device=dev)
    y = torch.randn(1, L, dim, device=dev).to(torch.bfloat16)
    mod = 0.1 * torch.randn(1, 1, 6, dim, device=dev)
    w, b = torch.ones(dim, device=dev), torch.zeros(dim, device=dev)
    forms = {
        "ln+modulate            (r4 w2)": (6, lambda: _ext.block_glue(x, scale=mod[:, :, 1], shift=mod[:, :, 0])),
        "residual+ln+modulate   (r6 w6)": (12, lambda: _ext.block_glue(x, y=y, gate=mod[:, :, 2], scale=mod[:, :, 4], shift=mod[:, :, 3], inplace=True)),
        "residual+affine ln     (r6 w6)": (12, lambda: _ext.block_glue(x, y=y, gate=mod[:, :, 2], ln=(w, b), inplace=True)),
        "residual only          (r6 w4)": (10, lambda: _ext.block_glue(x, y=y, gate=mod[:, :, 5], want_h=False, inplace=True)),
    }
    for form, (bpe, fn) in forms.items():
        for _ in range(3):
            fn()
        best = 1e9
        for i in range(10):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"GLUE {name} {form}: {best * 1e3:8.1f} us  {bpe * L * dim / best * 1e-6:7.0f} GB/s")
