import torch, sys
sys.path.insert(0, '.')
from univid_b200 import _ext as ext
dev = torch.device("cuda"); bf = torch.bfloat16
p, s, heads, batch = 2, int(sys.argv[1]), 4, 1
n, L, dim = heads // p, p * s, heads * 128
q_lin = torch.randn(batch, L, dim).to(bf).to(dev); k_lin = torch.randn(batch, L, dim).to(bf).to(dev)
v = torch.randn(batch, L, heads, 128).to(bf).to(dev)
w = torch.ones(dim, device=dev)
recv = [{t: torch.zeros(batch, L, n, 128, dtype=bf, device=dev) for t in "qkv"} for _ in range(p)]
flags = torch.zeros(p, 32, dtype=torch.int32, device=dev)
send_sb, send_sl = p * s * n * 128, n * 128
stream = torch.cuda.current_stream().cuda_stream
def step(name, fn):
    try:
        fn(); torch.cuda.synchronize(); print("OK  ", name, flush=True)
    except Exception as e:
        print("FAIL", name, str(e)[:100], flush=True); sys.exit(1)
i = 0; off = 0
qp = ext.ptr_array([recv[j]["q"].data_ptr() + off for j in range(p)])
kp = ext.ptr_array([recv[j]["k"].data_ptr() + off for j in range(p)])
vp = ext.ptr_array([recv[j]["v"].data_ptr() + off for j in range(p)])
step("plain prologue", lambda: ext.qk_norm_rope(q_lin[:, :s].contiguous(), k_lin[:, :s].contiguous(), w, w, 1e-6, heads, groups=p))
step("head_scatter peers", lambda: ext.head_scatter(v[:, :s].contiguous(), p, peers=(vp, send_sb, send_sl)))
step("signal", lambda: ext.sp_signal(ext.ptr_array([flags[j].data_ptr() for j in range(p)]), p, 7, stream))
print(flags[:, :4].tolist())
step("wait", lambda: ext.sp_wait(flags[0].data_ptr(), 1, 7, stream))
step("prologue peers", lambda: ext.qk_norm_rope(q_lin[:, :s].contiguous(), k_lin[:, :s].contiguous(), w, w, 1e-6, heads, groups=p, peers=(qp, kp, send_sb, send_sl)))
from oracle import wan_attention_oracle as orc
f = orc.make_freqs(128); cs = torch.stack([f.real, f.imag], dim=-1).float().contiguous().to(dev)
grid = [(2, 5, (L - 7) // 10)]
step("prologue peers rope", lambda: ext.qk_norm_rope(q_lin[:, :s].contiguous(), k_lin[:, :s].contiguous(), w, w, 1e-6, heads, cos_sin=cs, grid_sizes=grid, tok_offset=0, groups=p, peers=(qp, kp, send_sb, send_sl)))
step("signal2", lambda: ext.sp_signal(ext.ptr_array([flags[j].data_ptr() + 4 for j in range(p)]), p, 7, stream))
step("wait2", lambda: ext.sp_wait(flags[0].data_ptr(), 2, 7, stream))
o_recv = [torch.full((batch, s, heads, 128), float("nan"), dtype=bf, device=dev) for _ in range(p)]
op = ext.ptr_array([o.data_ptr() for o in o_recv])
step("fmha plain", lambda: ext.fmha_fwd(recv[0]["q"], recv[0]["k"], recv[0]["v"]))
step("fmha sp", lambda: ext.fmha_fwd_sp(recv[0]["q"], recv[0]["k"], recv[0]["v"], op, p, 0, heads))
print(torch.isfinite(o_recv[0].float()).float().mean().item(), torch.isfinite(o_recv[1].float()).float().mean().item())
