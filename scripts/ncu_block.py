"""One WanAttentionBlock forward of the 1.3B config at the full 32 760 tokens, twice (warm-up + the pass ncu should
look at): run under `ncu --metrics gpu__time_duration.sum` to get the launch list of a whole block, library kernels
included (scripts/gpu_round.sh).  The second pass starts after the marker kernel `fill_(12345)`."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

mdl = importlib.import_module("univid_b200.wan.modules.model")
cfg = bench.CONFIGS["1.3B"]
dev = torch.device("cuda")
f, h, w = cfg["grid"]
L, dim = f * h * w, cfg["dim"]
torch.manual_seed(0)
with torch.device(dev):
    blk = mdl.WanAttentionBlock(dim, cfg["ffn"], cfg["heads"], cross_attn_norm=True).eval()
d = dim // cfg["heads"]
freqs = torch.cat([mdl.rope_params(1024, d - 4 * (d // 6)), mdl.rope_params(1024, 2 * (d // 6)),
                   mdl.rope_params(1024, 2 * (d // 6))], dim=1).to(dev)
x = torch.randn(1, L, dim, device=dev)
e = 0.1 * torch.randn(1, 1, 6, dim, device=dev)
ctx = torch.randn(1, cfg["text_len"], dim, device=dev).to(torch.bfloat16)
marker = torch.empty(1, device=dev)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for i in range(2):
        marker.fill_(12345.0)
        y = blk(x, e, torch.tensor([L]), torch.tensor([[f, h, w]]), freqs, ctx, None)
torch.cuda.synchronize()
print("ok", tuple(y.shape), y.dtype)
