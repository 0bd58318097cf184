#!/usr/bin/env python
"""Benchmark of the Wan DiT attention hot path (BASELINE.json metric) -- prints ONE JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 14B|1.3B|5B]
    torchrun ... bench.py --gpus N ...            (one rank per GPU, NCCL; Ulysses head sharding)

Headline workload at EVERY N: the north-star target config, Wan2.1-T2V-14B at 75 600 video tokens (40 heads shard
over 1/2/4/8 GPUs), so the per-N values form a same-config strong-scaling curve.  The 1.3B / 32 760-token config
(BASELINE configs[1..2]; N in {1, 2, 4}) and the text-weight sweep (configs[4]) ride along as sub-records of the same
line (`configs`).

A *step* is the attention stack of one DiT forward (one denoise step) of the named config: for each of the model's
layers, q/k WanRMSNorm + 3-D RoPE, self-attention over all video tokens, and the 512-key cross-attention.  `value` is
attention TFLOP/s with the layer inputs (q/k/v projections) resident in HBM, measured over the kernels of this repo
only; `e2e` is the same FLOPs divided by the time of the public module API (WanSelfAttention / WanCrossAttention
.forward, q/k/v/o linears included) fed from pinned HOST memory with the result read back to the host every step;
`denoise_step` times the full WanModel.forward harness.  At N > 1 a full-size parity check of the sequence-parallel
path (against the unsharded kernels and an fp32 re-evaluation of sampled rows) runs BEFORE the timed region and the
process exits non-zero if it fails.

--impl reference times the reference's own CPU implementation of the path: the unmodified reference modules staged
under oracle/_ref by oracle/make_ref.py (kind "reference"; the oracle port if the staging is absent, kind "port") on
all host cores, on a bounded sample of the workload.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[1..2]: Wan2.1-T2V-1.3B, 81 frames 480x832 -> token grid (21, 30, 52)
    "1.3B": dict(name="Wan2.1-T2V-1.3B denoise step, 81f 480x832", dim=1536, heads=12, layers=30, ffn=8960,
                 grid=(21, 30, 52), text_len=512),
    # BASELINE.json configs[3] = the north-star target: Wan2.1-T2V-14B, 81 frames 720x1280 -> token grid (21, 45, 80)
    "14B": dict(name="Wan2.1-T2V-14B denoise step, 81f 720x1280", dim=5120, heads=40, layers=40, ffn=13824,
                grid=(21, 45, 80), text_len=512),
    # the reference's own default model (not a BASELINE config; SURVEY.md sec. 8 geometry table, "ref native"):
    # Wan2.2 ti2v-5B (models/wan/configs/wan_ti2v_5B.py), 121 frames 704x1280 -> token grid (31, 22, 40), 48 latent
    # channels, first latent frame given (its tokens carry timestep 0, textimage2video.py:372-377)
    "5B": dict(name="Wan2.2-TI2V-5B denoise step, 121f 704x1280", dim=3072, heads=24, layers=30, ffn=14336,
               grid=(31, 22, 40), text_len=512, in_dim=48, ti2v=True),
}
HEADLINE = "14B"


def flops_per_layer(L, heads, text_len):
    f_self = 4.0 * L * L * heads * 128
    f_cross = 4.0 * L * text_len * heads * 128
    return f_self, f_cross


def linear_flops_per_layer(L, dim, ffn, text_len):
    """q/k/v/o of self-attention, q/o of cross-attention over L rows, k/v over the text rows, the two FFN GEMMs."""
    return 2.0 * dim * dim * (6 * L + 2 * text_len) + 4.0 * L * dim * ffn


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(source="measured", tflops_burst=p["bf16_tflops"],
                    tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), hbm_gbs=p["hbm_gbs"])
    return dict(source="fallback", tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], None, [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5),
                              ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []      # upper half = samples under load
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------------
def _host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core it can (rank 0 is the only rank working)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def cpu_reference_rate(cfg, seconds_budget, steps, warmup):
    """One layer of WanSelfAttention + WanCrossAttention (q/k/v/o linears included) under bf16 autocast on the CPU --
    the reference's torch-SDPA route -- on a bounded token sample (whole latent frames) sized so that steps + warmup
    runs fit the time budget.  Runs the UNMODIFIED reference modules when oracle/_ref (or /root/reference) is there
    (`kind` "reference"), else the oracle port.  Returns (tflops, ms_per_step, sample, threads, kind)."""
    from oracle import ref_loader
    from oracle import wan_attention_oracle as orc
    threads = _host_threads()
    dim, heads, text_len = cfg["dim"], cfg["heads"], cfg["text_len"]
    f, h, w = cfg["grid"]
    g = torch.Generator().manual_seed(0)
    prm_s = orc.init_attention_params(dim, g)
    prm_c = orc.init_attention_params(dim, g)
    freqs = orc.make_freqs(128)
    kind = "port"
    sa = ca = None
    if ref_loader.available():
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _, ref_model = ref_loader.load_modules()
        sa, ca = ref_model.WanSelfAttention(dim, heads, eps=1e-6), ref_model.WanCrossAttention(dim, heads, eps=1e-6)
        sa.load_state_dict(prm_s)
        ca.load_state_dict(prm_c)
        sa, ca = sa.eval(), ca.eval()
        kind = "reference"

    def run(frames):
        L = frames * h * w
        x = torch.randn(1, L, dim, generator=g).to(torch.bfloat16)
        ctx = torch.randn(1, text_len, dim, generator=g).to(torch.bfloat16)
        gs, sl = torch.tensor([[frames, h, w]]), torch.tensor([L])
        t0 = time.perf_counter()
        with torch.no_grad():
            if sa is not None:
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    y = sa(x, sl, gs, freqs)
                    ca(y, ctx, None)
            else:
                y = orc.self_attention(x, prm_s, sl, gs, freqs, heads, 1e-6, bf16=True)
                orc.cross_attention(y, ctx, prm_c, heads, None, 1e-6, bf16=True)
        return time.perf_counter() - t0, L

    t_probe, l_probe = run(1)
    t_probe2, _ = run(1)
    t_probe = min(t_probe, t_probe2)
    per_step_budget = seconds_budget / max(steps + warmup, 1)

    def pick(t_ref, l_ref):
        for cand in range(f, 0, -1):   # self-attention time grows ~quadratically with the token count
            if t_ref * (cand * h * w / l_ref) ** 2 <= per_step_budget:
                return cand
        return 1

    # one frame is a poor predictor (small problems run the host cores inefficiently): re-estimate from the size just
    # chosen until the estimate stops growing (at most two extra untimed runs)
    frames = pick(t_probe, l_probe)
    for _ in range(2):
        t_big, l_big = run(frames)
        better = pick(t_big, l_big)
        if better <= frames:
            break
        frames = better
    for _ in range(warmup):
        run(frames)
    times = [run(frames)[0] for _ in range(steps)]
    L = frames * h * w
    fs, fc = flops_per_layer(L, heads, text_len)
    sec = sum(times) / len(times)
    sample = (f"1 of {cfg['layers']} layers (WanSelfAttention + WanCrossAttention incl. q/k/v/o linears), "
              f"{frames} of {f} latent frames = {L} of {f * h * w} video tokens, CPU bf16 torch-SDPA route, "
              f"{'unmodified reference modules (oracle/_ref)' if kind == 'reference' else 'oracle port'}")
    return (fs + fc) / sec * 1e-12, sec * 1e3, sample, threads, kind


def workload_config(cfg, world, fused_exchange=True):
    """The `config` object of the JSON line -- ONE definition for both arms, so the driver's same-config check compares
    like with like: the reference arm names the workload it samples, the bounded sample itself is `cpu_baseline.sample`."""
    f, h, w = cfg["grid"]
    L = f * h * w
    sharding = "none" if world == 1 else (
        f"Ulysses heads/{world}, exchange fused into the kernels over NVLink peer memory"
        if fused_exchange else f"Ulysses heads/{world} over NCCL all-to-all")
    return {
        "workload": cfg["name"] + " -- attention stack (q/k RMSNorm + 3-D RoPE, self-attention, 512-key "
                    "cross-attention) of all layers",
        "sharding": sharding, "video_tokens": L, "text_tokens": cfg["text_len"], "dim": cfg["dim"],
        "heads": cfg["heads"], "layers": cfg["layers"],
        "l2_policy": f"inputs larger than L2: 4 rotating layer-input sets of {3 * (L // world) * cfg['dim'] * 2 / 1e6:.0f} "
                     "MB each (GPU arm; the CPU reference arm runs the bounded sample named in cpu_baseline.sample)"}


def run_reference_arm(args, cfg_key):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[cfg_key]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    tflops, ms, sample, threads, kind = cpu_reference_rate(cfg, 150.0, steps, warmup)
    line = {
        "impl": "reference", "metric": "dit_attention_tflops", "value": tflops, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(cfg, max(1, args.gpus)), "sample": sample,
        "cpu_baseline": {"value": tflops, "unit": "TFLOP/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": tflops, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = load_peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = torch.tensor([float(v)], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_max_over_ranks(self, vals):
        """vals: list of floats -> (min list, max list) over ranks."""
        if self.world == 1:
            return list(vals), list(vals)
        lo = torch.tensor(vals, device=self.dev, dtype=torch.float64)
        hi = lo.clone()
        self.dist.all_reduce(lo, op=self.dist.ReduceOp.MIN)
        self.dist.all_reduce(hi, op=self.dist.ReduceOp.MAX)
        return lo.tolist(), hi.tolist()


def _rows_reference(q, k, v, rows, k_len=None):
    """fp32 softmax(q k^T / sqrt(d)) v for selected query rows; q/k/v [1, L, N, 128] bf16 on the GPU -> [R, N, D]."""
    qf = q[0, rows].float().transpose(0, 1)
    out = torch.empty(qf.shape, dtype=torch.float32, device=q.device)
    for n in range(q.size(2)):                              # per head: bounded memory at 75 600 keys
        s = torch.matmul(qf[n], k[0, :, n].float().t()) * 128 ** -0.5
        if k_len is not None:
            s[:, k_len:] = float("-inf")
        out[n] = torch.matmul(torch.softmax(s, dim=-1), v[0, :, n].float())
    return out.transpose(0, 1)


def sp_parity_check(env, cfg, mdl, sp, _ext):
    """Full-size parity of the sequence-parallel path, before anything is timed: one layer of
    sp_attn_forward on this rank's token shard of a seeded x (identical on every rank) against
      (i)  the same rows of the unsharded WanSelfAttention.forward on this GPU (same kernels, no exchange), and
      (ii) an fp32 re-evaluation of sampled query rows against ALL keys (plain matmul + softmax, the oracle's
           formula; tests/test_full_size_gpu.py) followed by the o projection in fp32 -- rows at both ends of the
           rank's chunk, i.e. inside the 128-row output tiles that straddle two ranks' chunks.
    Returns a dict; `ok` is the AND over all ranks."""
    dim, heads = cfg["dim"], cfg["heads"]
    f, h, w = cfg["grid"]
    L = f * h * w
    s = L // env.world
    r = env.rank
    bf = torch.bfloat16
    torch.manual_seed(4321)
    sa = mdl.WanSelfAttention(dim, heads).to(env.dev).eval()
    for lin in (sa.q, sa.k, sa.v, sa.o):
        torch.nn.init.xavier_uniform_(lin.weight)
        torch.nn.init.normal_(lin.bias, std=0.02)
    # logits ~ N(0, 2.5^2): a peaked softmax, so the outputs are not just the mean of v (which would make any
    # absolute tolerance vacuous at 75 600 keys)
    with torch.no_grad():
        sa.norm_q.weight.fill_(2.5)
    gen = torch.Generator(device=env.dev).manual_seed(99)          # identical on every rank
    x = torch.randn(1, L, dim, device=env.dev, generator=gen).to(bf)
    d = 128
    table = torch.cat([mdl.rope_params(1024, d - 4 * (d // 6)), mdl.rope_params(1024, 2 * (d // 6)),
                       mdl.rope_params(1024, 2 * (d // 6))], dim=1).to(env.dev)
    grid_t, seq_lens = torch.tensor([[f, h, w]]), torch.tensor([L])
    with torch.no_grad(), torch.autocast("cuda", dtype=bf):
        mine = sp.sp_attn_forward(sa, x[:, r * s:(r + 1) * s].contiguous(), seq_lens, grid_t, table).float().clone()
        full = sa(x, seq_lens, grid_t, table)[:, r * s:(r + 1) * s].float()
        # (ii) fp32 rows: q/k after the fused prologue (bf16, what the attention kernel reads), v from the projection
        q, k = sa._prologue(mdl._lin(sa.q, x), mdl._lin(sa.k, x), mdl._cos_sin_table(table, env.dev), grid_t)
        v = mdl._lin(sa.v, x).view(1, L, heads, d)
    local = sorted(set([0, 1, 63, 127, 128, s // 2, s - 129, s - 128, s - 65, s - 2, s - 1]))
    local = [i for i in local if 0 <= i < s]
    rows = torch.tensor([r * s + i for i in local], device=env.dev)
    att = _rows_reference(q, k, v, rows)                                        # [R, N, D] fp32
    want = att.flatten(1).to(bf).float() @ sa.o.weight.float().to(bf).float().t() + sa.o.bias.float().to(bf).float()
    got = mine[0, local]
    err_ref = (got - want).abs().max().item()
    scale_ref = want.abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got.flatten().double(), want.flatten().double(), dim=0).item()
    err_unsharded = (mine - full).abs().max().item()
    # a 128-row tile straddles two ranks' chunks whenever s is not a multiple of 128
    ok_local = err_ref <= 2e-2 and cos >= 0.9999 and err_unsharded <= 8e-3 and bool(torch.isfinite(mine).all())
    flag = torch.tensor([1.0 if ok_local else 0.0, -err_ref, -err_unsharded, cos, -err_ref / max(scale_ref, 1e-30)],
                        device=env.dev, dtype=torch.float64)
    if env.world > 1:
        env.dist.all_reduce(flag, op=env.dist.ReduceOp.MIN)
    del sa, x, q, k, v, full, mine
    torch.cuda.empty_cache()
    return {"ok": bool(flag[0].item() == 1.0), "max_abs_vs_fp32_rows": -flag[1].item(), "cos_vs_fp32_rows": flag[3].item(),
            "max_abs_vs_unsharded_kernel": -flag[2].item(), "max_abs_over_output_max": -flag[4].item(),
            "rows_per_rank": len(local),
            "tokens_per_rank": s, "tile_straddles_ranks": s % 128 != 0, "tolerance": "max_abs <= 2e-2, cos >= 0.9999",
            "what": "one layer of sp_attn_forward at the full benchmark size, every rank, worst over ranks"}


def measure(env, args, cfg_key, steps, warmup, *, denoise=True, gemm_roofline=True, e2e=True):
    """All measurements of one config on the current world.  Returns a dict (identical on every rank except for
    rank-local diagnostics)."""
    from univid_b200 import _ext
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    sp = importlib.import_module("univid_b200.wan.distributed.sequence_parallel")
    uly = importlib.import_module("univid_b200.wan.distributed.ulysses")
    dist, world, rank, dev, peaks = env.dist, env.world, env.rank, env.dev, env.peaks

    cfg = CONFIGS[cfg_key]
    dim, heads, layers, text_len = cfg["dim"], cfg["heads"], cfg["layers"], cfg["text_len"]
    f, h, w = cfg["grid"]
    L = f * h * w
    if heads % world != 0 or L % world != 0:
        raise SystemExit(f"{heads} heads / {L} tokens cannot be sharded over {world} ranks")
    s = L // world
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    bf = torch.bfloat16
    res = {"config_key": cfg_key}

    # ---------------- parity first (N > 1): nothing is timed on a path that is not proven on this box
    ctx = None
    if world > 1:
        p2p = importlib.import_module("univid_b200.wan.distributed.p2p")
        res["parity_check"] = sp_parity_check(env, cfg, mdl, sp, _ext)
        if not res["parity_check"]["ok"]:
            if rank == 0:
                print(json.dumps({"error": "sequence-parallel parity check failed", "config": cfg_key,
                                  "parity_check": res["parity_check"]}))
            env.barrier()
            sys.exit(3)
        ctx = p2p.context(1, s, heads, dev)

    # ---------------- device-resident layer inputs (rotating sets, each far larger than the 126 MB L2)
    n_sets = 4
    sets = []
    for _ in range(n_sets):
        sets.append(dict(
            q=torch.randn(1, s, dim, device=dev, generator=g).to(bf),
            k=torch.randn(1, s, dim, device=dev, generator=g).to(bf),
            v=torch.randn(1, s, heads, 128, device=dev, generator=g).to(bf),
            qc=torch.randn(1, s, dim, device=dev, generator=g).to(bf),
            kc=torch.randn(1, text_len, dim, device=dev, generator=g).to(bf),
            vc=torch.randn(1, text_len, heads, 128, device=dev, generator=g).to(bf)))
    wn = torch.ones(dim, device=dev)
    d = 128
    table = torch.cat([mdl.rope_params(1024, d - 4 * (d // 6)), mdl.rope_params(1024, 2 * (d // 6)),
                       mdl.rope_params(1024, 2 * (d // 6))], dim=1)
    cs = mdl._cos_sin_table(table, dev)
    grid = [(f, h, w)]
    seq_lens = torch.tensor([L])
    marks = []              # [(name, event)] of the recorded steps, in stream order

    def kernel_step(record):
        """P + S + X of every layer through the C ABI; with world > 1 the Ulysses exchange around S."""
        def mark(name):
            if record:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))
        for layer in range(layers):
            t = sets[layer % n_sets]
            mark("begin")
            if world == 1:
                q, k = _ext.qk_norm_rope(t["q"], t["k"], wn, wn, 1e-6, heads, cos_sin=cs, grid_sizes=grid)
                mark("prologue")
                _ext.fmha_fwd(q, k, t["v"])
                mark("self_attention")
            elif ctx is not None:
                # fused exchange: producers store into the peers' buffers, attention stores o into the owners'
                ctx.next_epoch()
                _ext.qk_norm_rope(t["q"], t["k"], wn, wn, 1e-6, heads, cos_sin=cs, grid_sizes=grid,
                                  tok_offset=rank * s, groups=world,
                                  peers=(ctx.q_peers, ctx.k_peers, ctx.send_sb, ctx.send_sl))
                mark("prologue")
                _ext.head_scatter(t["v"], world, peers=(ctx.v_peers, ctx.send_sb, ctx.send_sl))
                mark("head_scatter")
                ctx.attend(None, mark=mark)          # marks: qkv_signal_wait, self_attention, o_signal_wait
            else:
                q_send, k_send = _ext.qk_norm_rope(t["q"], t["k"], wn, wn, 1e-6, heads, cos_sin=cs,
                                                   grid_sizes=grid, tok_offset=rank * s, groups=world)
                mark("prologue")
                v_send = _ext.head_scatter(t["v"], world)
                mark("head_scatter")
                uly.attend_exchanged(q_send, k_send, v_send, seq_lens)
                mark("nccl_exchange_and_attention")
            qc, _ = _ext.qk_norm_rope(t["qc"], None, wn, None, 1e-6, heads)
            _, kc = _ext.qk_norm_rope(None, t["kc"], None, wn, 1e-6, heads)
            mark("cross_prologue")
            _ext.fmha_fwd(qc, kc, t["vc"])
            mark("cross_attention")

    def timed(fn, steps, warmup, sample_clocks=False, record_last=False):
        for _ in range(warmup):
            fn(False)
        env.barrier()
        sampler = ClockSampler(env.local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _ext.launch_count
        e0.record()
        for i in range(steps):
            fn(record_last and i == steps - 1)
        e1.record()
        env.barrier()
        clocks = sampler.stop() if sampler else None
        ms = env.max_over_ranks(e0.elapsed_time(e1) / steps)
        return ms, (_ext.launch_count - n0), clocks

    fs, fc = flops_per_layer(L, heads, text_len)
    flop_step = (fs + fc) * layers
    with torch.no_grad():
        ms_kernel, launches, clocks = timed(kernel_step, steps, warmup, sample_clocks=True, record_last=True)
    res.update(value=flop_step / (ms_kernel * 1e-3) * 1e-12, ms_per_step=ms_kernel, gpu_launches=launches, clocks=clocks,
               attention_flop_per_step=flop_step)

    # ---------------- per-kernel split of the last timed step (CUDA events on the launching stream) ----------
    seg = {}
    for (n0_, e0_), (n1_, e1_) in zip(marks[:-1], marks[1:]):
        if n1_ == "begin":
            continue
        seg.setdefault(n1_, []).append(e0_.elapsed_time(e1_))
    names = list(seg.keys())
    avg = [sum(seg[n]) / len(seg[n]) for n in names]
    lo, hi = env.min_max_over_ranks(avg)
    split = {n: {"avg_ms_per_layer": a, "min_over_ranks": l_, "max_over_ranks": h_}
             for n, a, l_, h_ in zip(names, avg, lo, hi)}
    res["kernel_split"] = {"of": "last timed step, per layer, CUDA events", "segments": split,
                           "sum_ms_per_step_this_rank": sum(avg) * layers}

    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    traffic = json.load(open(tpath)).get(cfg_key, {}) if os.path.exists(tpath) else {}
    heads_local = heads // world
    if "self_attention" in seg:
        a = sum(seg["self_attention"]) / len(seg["self_attention"])
        fs_local = fs / world                                  # this rank's head shard
        achieved = fs_local / (a * 1e-3) * 1e-12
        res["roofline"] = {
            "kernel": "fmha_fwd_kernel (self-attention" + (", CTA pairs" if _ext.lib().uvb_get_knob(0) == 1 else "") + ")",
            "bound": "tensor", "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
            "frac": achieved / peaks["tflops_sustained"], "frac_of_burst_peak": achieved / peaks["tflops_burst"],
            "peak_source": peaks["source"] + " sustained bf16 (kernel timed inside a long step)",
            "traffic": traffic.get("fmha_dram_bytes_per_launch") if world == 1 else None,
            "avg_launch_ms": a, "flop_per_launch": fs_local, "heads_per_launch": heads_local,
            "kernel_share_of_step": a * layers / ms_kernel}
    if "prologue" in seg and world == 1:
        pavg = sum(seg["prologue"]) / len(seg["prologue"])
        pbytes = 8.0 * L * dim
        res["roofline_prologue"] = {
            "kernel": "qk_norm_rope_kernel (q/k RMSNorm + 3-D RoPE)", "bound": "hbm",
            "achieved": pbytes / (pavg * 1e-3) * 1e-9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": pbytes / (pavg * 1e-3) * 1e-9 / peaks["hbm_gbs"], "peak_source": peaks["source"] + " copy bandwidth",
            "avg_launch_ms": pavg, "bytes_per_launch": pbytes, "traffic": traffic.get("prologue_dram_bytes_per_launch")}

    # ---------------- the prologue kernel back to back (rotating input sets >> L2): its own sustained rate, without
    # the event gaps of a 0.1-0.7 ms kernel between two attention launches and at the clock a memory-bound kernel gets
    if "roofline_prologue" in res:
        n_p = 24
        for i in range(3):
            _ext.qk_norm_rope(sets[i % n_sets]["q"], sets[i % n_sets]["k"], wn, wn, 1e-6, heads, cos_sin=cs, grid_sizes=grid)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for i in range(n_p):
            t = sets[i % n_sets]
            _ext.qk_norm_rope(t["q"], t["k"], wn, wn, 1e-6, heads, cos_sin=cs, grid_sizes=grid)
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1) / n_p
        pb = res["roofline_prologue"]["bytes_per_launch"]
        res["roofline_prologue"]["in_step"] = {k: res["roofline_prologue"][k] for k in ("achieved", "frac", "avg_launch_ms")}
        res["roofline_prologue"]["back_to_back"] = {
            "achieved": pb / (pms * 1e-3) * 1e-9, "frac": pb / (pms * 1e-3) * 1e-9 / peaks["hbm_gbs"], "avg_launch_ms": pms,
            "launches": n_p, "what": "same kernel, same shapes, 4 rotating input sets (each >> L2), no other kernel in between"}

    # ---------------- the block's largest GEMM (ffn[0] + tanh-GELU), timed live -----
    if gemm_roofline and world == 1:
        ffn = cfg["ffn"]
        xg = sets[0]["q"].view(s, dim)
        wg = (torch.randn(ffn, dim, device=dev, generator=g) / dim ** 0.5).to(bf)
        bg = torch.zeros(ffn, device=dev)
        og = torch.empty(s, ffn, dtype=bf, device=dev)          # x + out >> L2: every launch streams from HBM
        for _ in range(3):
            _ext.linear(xg, wg, bg, act=_ext.ACT_GELU_TANH, out=og)
        torch.cuda.synchronize()
        n_g = 20 if cfg_key != "14B" else 8
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(n_g):
            _ext.linear(xg, wg, bg, act=_ext.ACT_GELU_TANH, out=og)
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / n_g
        gflop = 2.0 * s * ffn * dim
        res["roofline_gemm"] = {
            "kernel": "gemm_bf16_kernel (ffn[0] + bias + tanh-GELU)", "bound": "tensor",
            "achieved": gflop / (gms * 1e-3) * 1e-12, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
            "frac": gflop / (gms * 1e-3) * 1e-12 / peaks["tflops_sustained"],
            "frac_of_burst_peak": gflop / (gms * 1e-3) * 1e-12 / peaks["tflops_burst"],
            "peak_source": peaks["source"] + f" sustained bf16 ({n_g} launches back to back)",
            "avg_launch_ms": gms, "flop_per_launch": gflop, "shape_mnk": [s, ffn, dim],
            "traffic": traffic.get("gemm_ffn0_dram_bytes_per_launch")}
        del wg, og

    # ---------------- e2e: public module API from pinned host memory ---------------------------------
    if e2e:
        torch.manual_seed(0)
        sa = mdl.WanSelfAttention(dim, heads).to(dev).eval()
        ca = mdl.WanCrossAttention(dim, heads).to(dev).eval()
        for m in (sa, ca):
            for lin in (m.q, m.k, m.v, m.o):
                torch.nn.init.xavier_uniform_(lin.weight)
                torch.nn.init.zeros_(lin.bias)
        x_host = torch.randn(1, s, dim).to(bf).pin_memory()
        ctx_host = torch.randn(1, text_len, dim).to(bf).pin_memory()
        out_host = torch.empty(1, s, dim, dtype=bf).pin_memory()
        grid_t = torch.tensor([[f, h, w]])
        table_dev = table.to(dev)

        def e2e_step(_record):
            x = x_host.to(dev, non_blocking=True)
            cx = ctx_host.to(dev, non_blocking=True)
            with torch.autocast("cuda", dtype=bf):
                for _ in range(layers):
                    if world == 1:
                        y = sa(x, seq_lens, grid_t, table_dev)
                    else:
                        y = sp.sp_attn_forward(sa, x, seq_lens, grid_t, table_dev)
                    x = y + ca(y, cx, None)
            out_host.copy_(x, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        with torch.no_grad():
            ms_e2e, _, _ = timed(e2e_step, steps, warmup)
        res["e2e"] = {"value": flop_step / (ms_e2e * 1e-3) * 1e-12, "unit": "TFLOP/s", "ms_per_step": ms_e2e,
                      "steps": steps, "warmup": warmup,
                      "h2d_bytes_per_step": (x_host.numel() + ctx_host.numel()) * 2 * world,
                      "d2h_bytes_per_step": out_host.numel() * 2 * world,
                      "api": "WanSelfAttention.forward + WanCrossAttention.forward per layer (q/k/v/o linears included), "
                             "x/context from pinned host memory, result copied back"}
        del sa, ca

    # ---------------- full denoise step through the WanModel harness ------------------------------------
    # N > 1: sp_attn_forward / sp_dit_forward are bound onto the instance exactly like the reference does
    # with use_sp=True (textimage2video.py:143-147): tokens sharded over the ranks, Ulysses around attention.
    if denoise:
        import types
        del sets
        torch.cuda.empty_cache()
        torch.manual_seed(0)
        with torch.device(dev):
            zc = cfg.get("in_dim", 16)
            model = mdl.WanModel(model_type="ti2v" if cfg.get("ti2v") else "t2v", dim=dim, ffn_dim=cfg["ffn"],
                                 num_heads=heads, num_layers=layers, text_len=text_len, in_dim=zc, out_dim=zc)
        model = model.eval()
        if world > 1:
            for block in model.blocks:
                block.self_attn.forward = types.MethodType(sp.sp_attn_forward, block.self_attn)
            model.forward = types.MethodType(sp.sp_dit_forward, model)
        gen = torch.Generator(device=dev).manual_seed(7)      # identical inputs on every rank
        lat = [torch.randn(zc, f, h * 2, w * 2, device=dev, generator=gen)]
        ctx_in = [torch.randn(text_len, 4096, device=dev, generator=gen)]
        # per-token timesteps [1, seq_len], the way the reference sampling loop calls the DiT (textimage2video.py:372-377)
        tt = torch.full((1, L), 500.0, device=dev)
        if cfg.get("ti2v"):
            tt[0, :h * w] = 0.0
        n_den = 2
        with torch.no_grad(), torch.autocast("cuda", dtype=bf):
            model(lat, tt, ctx_in, seq_len=L)
            env.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = _ext.launch_count
            e0.record()
            for _ in range(n_den):
                model(lat, tt, ctx_in, seq_len=L)
            e1.record()
            env.barrier()
        den_ms = env.max_over_ranks(e0.elapsed_time(e1) / n_den)
        den_flop = flop_step + linear_flops_per_layer(L, dim, cfg["ffn"], text_len) * layers
        den_tf = den_flop / (den_ms * 1e-3) * 1e-12
        res["denoise_step"] = {
            "ms": den_ms, "steps": n_den, "warmup": 1, "flop": den_flop, "tflops": den_tf,
            "frac_of_sustained_peak": den_tf / (peaks["tflops_sustained"] * world),
            "frac_of_burst_peak": den_tf / (peaks["tflops_burst"] * world),
            "launches_per_step": (_ext.launch_count - n0) // n_den,
            "what": "WanModel.forward, all blocks (attention + q/k/v/o + FFN GEMMs + glue) with per-token timesteps; "
                    "flop = attention + block linears"}
        del model
        torch.cuda.empty_cache()
    res["config"] = workload_config(cfg, world, fused_exchange=ctx is not None)
    res["sharding"] = res["config"]["sharding"]
    res["geometry"] = {k: res["config"][k] for k in ("video_tokens", "text_tokens", "dim", "heads", "layers", "l2_policy")}
    return res


def run_native_arm(args, cfg_key):
    env = Env(args)
    cfg = CONFIGS[cfg_key]
    if args.knob:
        from univid_b200 import _ext
        for kv in args.knob:
            name, value = kv.split("=")
            _ext.set_knob(name, int(value))
    main = measure(env, args, cfg_key, args.steps, args.warmup, denoise=not args.skip_denoise)
    subs = {}
    if not args.no_sub_records and cfg_key == HEADLINE:
        # BASELINE configs[1..2]: the 1.3B model at 32 760 tokens (12 heads: 1, 2 or 4 GPUs)
        if CONFIGS["1.3B"]["heads"] % env.world == 0:
            sub = measure(env, args, "1.3B", min(args.steps, 10), 3, denoise=not args.skip_denoise, gemm_roofline=False)
            subs["1.3B"] = {k: sub.get(k) for k in ("value", "ms_per_step", "e2e", "denoise_step", "roofline",
                                                     "roofline_prologue", "kernel_split", "parity_check", "sharding",
                                                     "geometry", "gpu_launches")}
            subs["1.3B"]["unit"] = "TFLOP/s"
            subs["1.3B"]["workload"] = CONFIGS["1.3B"]["name"] + " -- attention stack"
        # BASELINE configs[4]: the text-weight sweep (query rows sharded, context replicated, no communication)
        subs["tma_sweep"] = tma_sweep(env)
    if env.rank == 0:
        cpu = None
        if env.world == 1 and not args.skip_cpu:
            tf, ms, sample, threads, kind = cpu_reference_rate(cfg, 20.0, 2, 1)
            cpu = {"value": tf, "unit": "TFLOP/s", "cores": threads, "kind": kind, "sample": sample, "ms_per_sample": ms}
        line = {
            "metric": "dit_attention_tflops", "value": main["value"], "unit": "TFLOP/s", "n_gpus": env.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": main["config"],
            "attention_flop_per_step": main["attention_flop_per_step"],
            "denoise_step_ms": (main.get("denoise_step") or {}).get("ms"),
            "denoise_step": main.get("denoise_step"),
            "e2e": main.get("e2e"), "gpu_launches": main["gpu_launches"], "clocks": main["clocks"],
            "parity_check": main.get("parity_check"),
            "roofline": main.get("roofline"), "roofline_prologue": main.get("roofline_prologue"),
            "roofline_gemm": main.get("roofline_gemm"), "kernel_split": main.get("kernel_split"),
            "configs": subs, "cpu_baseline": cpu,
        }
        if args.knob:
            line["knobs"] = args.knob
        print(json.dumps(line))
    if env.world > 1:
        env.dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: Temperature Modality Alignment cross-attention sweep
# ------------------------------------------------------------------------------------------------------
def tma_sweep(env):
    """512 text tokens x {32 760, 75 600} video tokens across the 50-step flow schedule (2 DiT calls per step
    with classifier-free guidance -> call index c = 0..99, weight w(c) from univid_b200.tma).  Per call: the
    text-weighted k-norm prologue on the 512 context rows + the fused cross-attention kernel (per-key
    post-softmax weight, value bias) -- `kernel` -- and the public WanCrossAttention.forward(text_weight=,
    text_len=) with its q/k/v/o linears -- `module`.  With several ranks every rank takes L/p query rows (no
    communication: the context is replicated)."""
    from univid_b200 import _ext, tma
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    dist, world, dev = env.dist, env.world, env.dev
    bf = torch.bfloat16
    tcfg = tma.TextWeightConfig()
    calls = 2 * tcfg.total_sampling_steps
    weights = [tma.calculate_text_weight(c, tcfg) for c in range(calls)]
    res = {}
    for key in ("1.3B", "14B"):
        cfg = CONFIGS[key]
        dim, heads, text_len = cfg["dim"], cfg["heads"], cfg["text_len"]
        f, h, w_ = cfg["grid"]
        L = f * h * w_
        s = L // world
        torch.manual_seed(0)
        ca = mdl.WanCrossAttention(dim, heads).to(dev).eval()
        for lin in (ca.q, ca.k, ca.v, ca.o):
            torch.nn.init.xavier_uniform_(lin.weight)
            torch.nn.init.normal_(lin.bias, std=0.02)
        x = torch.randn(1, s, dim, device=dev).to(bf)
        ctx = torch.randn(1, text_len, dim, device=dev).to(bf)
        tl = tma.text_len_for(ctx, tcfg)
        with torch.no_grad(), torch.autocast("cuda", dtype=bf):
            q, _ = ca._prologue(ca.q(x), None, None, None)
            zero = ctx.new_zeros(1, 1, dim)
            b_k, b_v = ca.k(zero).flatten().float(), ca.v(zero).flatten().float()
            k_lin = (ca.k(ctx).float() - b_k).to(bf).contiguous()
            v_lin = (ca.v(ctx).float() - b_v).to(bf).view(1, text_len, heads, 128)
            wk = ca.norm_k.weight.float()
            w_of = {}
            for wt in sorted(set(weights)):
                v_ = torch.ones(text_len, dtype=torch.float32, device=dev)
                v_[:tl] = wt
                w_of[wt] = v_
            out = torch.empty(1, s, heads, 128, dtype=bf, device=dev)
            v_f32 = v_lin.float()

            def kernel_call(c):
                wv = w_of[weights[c]]
                _, k = _ext.qk_norm_rope(None, k_lin, None, wk, 1e-6, heads, row_scale=wv, pre_bias=b_k)
                v_w = (v_f32 * wv.view(1, text_len, 1, 1)).to(bf)       # 512 rows: the weight rides on V
                _ext.fmha_fwd(q, k, v_w, out=out, out_bias=b_v)

            def module_call(c):
                ca(x, ctx, None, text_weight=weights[c], text_len=tl)

            timing = {}
            for name, fn in (("kernel", kernel_call), ("module", module_call)):
                for c in range(3):
                    fn(c)
                env.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for c in range(calls):
                    fn(c)
                e1.record()
                torch.cuda.synchronize()
                timing[name] = env.max_over_ranks(e0.elapsed_time(e1) / calls)
        flop = 4.0 * L * text_len * heads * 128
        bytes_x = 4.0 * L * dim + 4.0 * text_len * dim
        res[key] = {"video_tokens": L, "heads": heads, "text_len_weighted": tl, "calls": calls,
                    "kernel_ms_per_call": timing["kernel"], "module_ms_per_call": timing["module"],
                    "kernel_tflops": flop / (timing["kernel"] * 1e-3) * 1e-12,
                    "kernel_qo_stream_gbs": bytes_x / (timing["kernel"] * 1e-3) * 1e-9,
                    "module_tflops_attn_only": flop / (timing["module"] * 1e-3) * 1e-12}
        del ca, x, q, out
        torch.cuda.empty_cache()
    res["workload"] = ("Temperature Modality Alignment cross-attention sweep: 512 text tokens x {32760, 75600} video "
                       "tokens, 50 flow steps x 2 CFG calls, cosine 1.3 -> 1.0")
    res["weights_first_last"] = [weights[0], weights[-1]]
    return res


def run_tma_sweep(args):
    env = Env(args)
    res = tma_sweep(env)
    if env.rank == 0:
        peaks = env.peaks
        line = {"metric": "tma_cross_attention_sweep_tflops", "value": res["1.3B"]["kernel_tflops"], "unit": "TFLOP/s",
                "n_gpus": env.world, "steps": res["1.3B"]["calls"], "warmup": 3,
                "ms_per_step": res["1.3B"]["kernel_ms_per_call"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": res["workload"], "weights_first_last": res["weights_first_last"],
                           "transition_calls": int(50 * 0.4)},
                "sweep": {k: v for k, v in res.items() if k in ("1.3B", "14B")},
                "peak_tflops_burst": peaks["tflops_burst"], "hbm_gbs": peaks["hbm_gbs"],
                "gpu_launches": 2 * res["1.3B"]["calls"] * 2}
        print(json.dumps(line))
    if env.world > 1:
        env.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default=None, choices=list(CONFIGS),
                    help="headline config (default: 14B at every N -- the north-star target, same config at 1/2/4/8 GPUs)")
    ap.add_argument("--workload", default="attention", choices=["attention", "tma-sweep"],
                    help="attention = the denoise-step attention stack (default); tma-sweep = the text-weighted "
                         "cross-attention sweep (BASELINE configs[4]) alone")
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg")
    ap.add_argument("--skip-denoise", action="store_true", help="omit the full WanModel denoise-step timing")
    ap.add_argument("--no-sub-records", action="store_true", help="omit the 1.3B and text-weight-sweep sub-records")
    ap.add_argument("--knob", action="append", default=[], metavar="NAME=VALUE",
                    help="A/B runs only: uvb_set_knob before measuring (e.g. fmha_pair=0, prologue_pair=1); recorded in the line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    cfg_key = args.config or HEADLINE
    if args.impl == "reference":
        run_reference_arm(args, cfg_key)
    elif args.workload == "tma-sweep":
        run_tma_sweep(args)
    else:
        run_native_arm(args, cfg_key)


if __name__ == "__main__":
    main()
