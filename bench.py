#!/usr/bin/env python
"""Benchmark of the Wan DiT attention hot path (BASELINE.json metric) -- prints ONE JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1.3B|14B]
    torchrun ... bench.py --gpus N ...            (one rank per GPU, NCCL; Ulysses head sharding)

A *step* is the attention stack of one DiT forward (one denoise step) of the named config: for each of
the model's layers, q/k WanRMSNorm + 3-D RoPE, self-attention over all video tokens, and the 512-key
cross-attention.  `value` is attention TFLOP/s with the layer inputs (q/k/v projections) resident in
HBM, measured over the kernels of this repo only; `e2e` is the same FLOPs divided by the time of the
public module API (WanSelfAttention / WanCrossAttention .forward, q/k/v/o linears included) fed from
pinned HOST memory with the result read back to the host every step.  `denoise_step_ms` additionally
times the full WanModel.forward harness (all blocks with PyTorch linears / FFN) once.

--impl reference times the reference's CPU path for the same metric: the oracle port of the reference
modules (oracle/wan_attention_oracle.py; the reference itself is Python and cannot travel to the GPU
box) on the host cores, on a bounded sample of the workload.
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
    # BASELINE.json configs[1..3]: Wan2.1-T2V-1.3B, 81 frames 480x832 -> token grid (21, 30, 52)
    "1.3B": dict(name="Wan2.1-T2V-1.3B denoise step, 81f 480x832", dim=1536, heads=12, layers=30, ffn=8960,
                 grid=(21, 30, 52), text_len=512),
    # BASELINE.json configs[3]: Wan2.1-T2V-14B, 81 frames 720x1280 -> token grid (21, 45, 80)
    "14B": dict(name="Wan2.1-T2V-14B denoise step, 81f 720x1280", dim=5120, heads=40, layers=40, ffn=13824,
                grid=(21, 45, 80), text_len=512),
    # the reference's own default model (not a BASELINE config; SURVEY.md sec. 8 geometry table, "ref native"):
    # Wan2.2 ti2v-5B (models/wan/configs/wan_ti2v_5B.py), 121 frames 704x1280 -> token grid (31, 22, 40), 48 latent
    # channels, first latent frame given (its tokens carry timestep 0, textimage2video.py:372-377)
    "5B": dict(name="Wan2.2-TI2V-5B denoise step, 121f 704x1280", dim=3072, heads=24, layers=30, ffn=14336,
               grid=(31, 22, 40), text_len=512, in_dim=48, ti2v=True),
}


def flops_per_layer(L, heads, text_len):
    f_self = 4.0 * L * L * heads * 128
    f_cross = 4.0 * L * text_len * heads * 128
    return f_self, f_cross


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
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5),
                              ("sw_power_cap", 6)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []      # upper half = samples under load
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, seconds_budget, steps, warmup):
    """Times oracle.self_attention + oracle.cross_attention (bf16 autocast-equivalent, torch-SDPA route:
    what the reference runs on CPU) for ONE layer on a bounded token sample.  Returns (tflops, ms_per_step,
    sample description, threads)."""
    from oracle import wan_attention_oracle as orc
    dim, heads, text_len = cfg["dim"], cfg["heads"], cfg["text_len"]
    f, h, w = cfg["grid"]
    threads = torch.get_num_threads()
    # probe at a small size to choose the largest frame count that fits the time budget
    g = torch.Generator().manual_seed(0)
    prm_s = orc.init_attention_params(dim, g)
    prm_c = orc.init_attention_params(dim, g)
    freqs = orc.make_freqs(128)

    def run(frames):
        L = frames * h * w
        x = torch.randn(1, L, dim, generator=g).to(torch.bfloat16)
        ctx = torch.randn(1, text_len, dim, generator=g).to(torch.bfloat16)
        gs, sl = torch.tensor([[frames, h, w]]), torch.tensor([L])
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.self_attention(x, prm_s, sl, gs, freqs, heads, 1e-6, bf16=True)
            orc.cross_attention(x, ctx, prm_c, heads, None, 1e-6, bf16=True)
        return time.perf_counter() - t0, L

    t_probe, l_probe = run(1)
    per_step_budget = seconds_budget / max(steps + warmup, 1)
    frames = 1
    for cand in range(f, 0, -1):       # self-attention time grows ~quadratically with the token count
        est = t_probe * (cand * h * w / l_probe) ** 2
        if est <= per_step_budget:
            frames = cand
            break
    for _ in range(warmup):
        run(frames)
    times = []
    for _ in range(steps):
        t, L = run(frames)
        times.append(t)
    L = frames * h * w
    fs, fc = flops_per_layer(L, heads, text_len)
    sec = sum(times) / len(times)
    sample = (f"1 of {cfg['layers']} layers (WanSelfAttention + WanCrossAttention incl. q/k/v/o linears), "
              f"{frames} of {f} latent frames = {L} of {f * h * w} video tokens, CPU bf16 torch-SDPA route")
    return (fs + fc) / sec * 1e-12, sec * 1e3, sample, threads


def run_reference_arm(args, cfg_key):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[cfg_key]
    steps, warmup = max(1, min(args.steps, 5)), max(0, min(args.warmup, 1))
    tflops, ms, sample, threads = cpu_reference_rate(cfg, 150.0, steps, warmup)
    line = {
        "impl": "reference", "metric": "dit_attention_tflops", "value": tflops, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": cfg["name"] + " -- attention stack", "sample": sample},
        "cpu_baseline": {"value": tflops, "unit": "TFLOP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": tflops, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
def run_native_arm(args, cfg_key):
    import torch.distributed as dist
    from univid_b200 import _ext
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    sp = importlib.import_module("univid_b200.wan.distributed.sequence_parallel")
    uly = importlib.import_module("univid_b200.wan.distributed.ulysses")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cfg = CONFIGS[cfg_key]
    dim, heads, layers, text_len = cfg["dim"], cfg["heads"], cfg["layers"], cfg["text_len"]
    f, h, w = cfg["grid"]
    L = f * h * w
    if heads % world != 0 or L % world != 0:
        raise SystemExit(f"{heads} heads / {L} tokens cannot be sharded over {world} ranks")
    s = L // world
    peaks = load_peaks()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    bf = torch.bfloat16

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    fmha_events, prol_events = [], []
    ctx = None
    if world > 1:
        p2p = importlib.import_module("univid_b200.wan.distributed.p2p")
        ctx = p2p.context(1, s, heads, dev)

    def kernel_step(record):
        """P + S + X of every layer through the C ABI; with world > 1 the Ulysses exchange around S."""
        for layer in range(layers):
            t = sets[layer % n_sets]
            if world == 1:
                if record:
                    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    ev[0].record()
                q, k = _ext.qk_norm_rope(t["q"], t["k"], wn, wn, 1e-6, heads, cos_sin=cs, grid_sizes=grid)
                if record:
                    ev[1].record()
                _ext.fmha_fwd(q, k, t["v"])
                if record:
                    ev[2].record()
                    prol_events.append((ev[0], ev[1]))
                    fmha_events.append((ev[1], ev[2]))
            elif ctx is not None:
                # fused exchange: producers store into the peers' buffers, attention stores o into the owners'
                ctx.next_epoch()
                _ext.qk_norm_rope(t["q"], t["k"], wn, wn, 1e-6, heads, cos_sin=cs, grid_sizes=grid,
                                  tok_offset=rank * s, groups=world,
                                  peers=(ctx.q_peers, ctx.k_peers, ctx.send_sb, ctx.send_sl))
                _ext.head_scatter(t["v"], world, peers=(ctx.v_peers, ctx.send_sb, ctx.send_sl))
                ctx.attend(None)
            else:
                q_send, k_send = _ext.qk_norm_rope(t["q"], t["k"], wn, wn, 1e-6, heads, cos_sin=cs,
                                                   grid_sizes=grid, tok_offset=rank * s, groups=world)
                v_send = _ext.head_scatter(t["v"], world)
                uly.attend_exchanged(q_send, k_send, v_send, seq_lens)
            qc, _ = _ext.qk_norm_rope(t["qc"], None, wn, None, 1e-6, heads)
            _, kc = _ext.qk_norm_rope(None, t["kc"], None, wn, 1e-6, heads)
            _ext.fmha_fwd(qc, kc, t["vc"])

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn(False)
        barrier()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _ext.launch_count
        e0.record()
        for _ in range(steps):
            fn(True)
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms, (_ext.launch_count - n0), clocks

    fs, fc = flops_per_layer(L, heads, text_len)
    flop_step = (fs + fc) * layers
    with torch.no_grad():
        ms_kernel, launches, clocks = timed(kernel_step, args.steps, args.warmup, sample_clocks=True)
    value = flop_step / (ms_kernel * 1e-3) * 1e-12

    roofline, roofline_prologue = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if world == 1 and fmha_events:
        durs = [a.elapsed_time(b) for a, b in fmha_events]
        avg = sum(durs) / len(durs)
        achieved = fs / (avg * 1e-3) * 1e-12
        traffic = None
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(cfg_key, {}).get("fmha_dram_bytes_per_launch")
        roofline = {"kernel": "fmha_fwd_kernel (self-attention)", "bound": "tensor", "achieved": achieved,
                    "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_sustained"],
                    "frac_of_burst_peak": achieved / peaks["tflops_burst"], "peak_source": peaks["source"] +
                    " sustained bf16 (kernel timed inside a long step)", "traffic": traffic,
                    "avg_launch_ms": avg, "flop_per_launch": fs,
                    "kernel_share_of_step": avg * layers / ms_kernel}
        pavg = sum(a.elapsed_time(b) for a, b in prol_events) / len(prol_events)
        pbytes = 8.0 * L * dim
        roofline_prologue = {"kernel": "qk_norm_rope_kernel (q/k RMSNorm + 3-D RoPE)", "bound": "hbm",
                             "achieved": pbytes / (pavg * 1e-3) * 1e-9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": pbytes / (pavg * 1e-3) * 1e-9 / peaks["hbm_gbs"],
                             "peak_source": peaks["source"] + " copy bandwidth", "avg_launch_ms": pavg,
                             "bytes_per_launch": pbytes,
                             "traffic": (json.load(open(tpath)).get(cfg_key, {}).get("prologue_dram_bytes_per_launch")
                                         if os.path.exists(tpath) else None)}

    # ---------------- the block's largest GEMM (ffn[0] + tanh-GELU), timed live -----
    roofline_gemm = None
    if world == 1:
        ffn = cfg["ffn"]
        xg = sets[0]["q"].view(s, dim)
        wg = (torch.randn(ffn, dim, device=dev, generator=g) / dim ** 0.5).to(bf)
        bg = torch.zeros(ffn, device=dev)
        og = torch.empty(s, ffn, dtype=bf, device=dev)          # x + out >> L2: every launch streams from HBM
        for _ in range(3):
            _ext.linear(xg, wg, bg, act=_ext.ACT_GELU_TANH, out=og)
        torch.cuda.synchronize()
        n_g = 20
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(n_g):
            _ext.linear(xg, wg, bg, act=_ext.ACT_GELU_TANH, out=og)
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / n_g
        gflop = 2.0 * s * ffn * dim
        roofline_gemm = {"kernel": "gemm_bf16_kernel (ffn[0] + bias + tanh-GELU)", "bound": "tensor",
                         "achieved": gflop / (gms * 1e-3) * 1e-12, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                         "frac": gflop / (gms * 1e-3) * 1e-12 / peaks["tflops_sustained"],
                         "frac_of_burst_peak": gflop / (gms * 1e-3) * 1e-12 / peaks["tflops_burst"],
                         "peak_source": peaks["source"] + " sustained bf16 (20 launches back to back)",
                         "avg_launch_ms": gms, "flop_per_launch": gflop, "shape_mnk": [s, ffn, dim],
                         "traffic": (json.load(open(tpath)).get(cfg_key, {}).get("gemm_ffn0_dram_bytes_per_launch")
                                     if os.path.exists(tpath) else None)}
        del wg, og

    # ---------------- e2e: public module API from pinned host memory ---------------------------------
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
        ctx = ctx_host.to(dev, non_blocking=True)
        with torch.autocast("cuda", dtype=bf):
            for _ in range(layers):
                if world == 1:
                    y = sa(x, seq_lens, grid_t, table_dev)
                else:
                    y = sp.sp_attn_forward(sa, x, seq_lens, grid_t, table_dev)
                x = y + ca(y, ctx, None)
        out_host.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_steps = max(1, min(args.steps, 5))
    with torch.no_grad():
        ms_e2e, _, _ = timed(e2e_step, e2e_steps, 1)
    e2e = {"value": flop_step / (ms_e2e * 1e-3) * 1e-12, "unit": "TFLOP/s", "ms_per_step": ms_e2e,
           "steps": e2e_steps,
           "h2d_bytes_per_step": (x_host.numel() + ctx_host.numel()) * 2 * world,
           "d2h_bytes_per_step": out_host.numel() * 2 * world,
           "api": "WanSelfAttention.forward + WanCrossAttention.forward per layer (q/k/v/o linears included), "
                  "x/context from pinned host memory, result copied back"}

    # ---------------- full denoise step through the WanModel harness ------------------------------------
    # N > 1: sp_attn_forward / sp_dit_forward are bound onto the instance exactly like the reference does
    # with use_sp=True (textimage2video.py:143-147): tokens sharded over the ranks, Ulysses around attention.
    denoise_ms = None
    if not args.skip_denoise:
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
        with torch.no_grad(), torch.autocast("cuda", dtype=bf):
            model(lat, tt, ctx_in, seq_len=L)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(2):
                model(lat, tt, ctx_in, seq_len=L)
            e1.record()
            barrier()
        denoise_ms = e0.elapsed_time(e1) / 2
        if world > 1:
            tms = torch.tensor([denoise_ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            denoise_ms = float(tms.item())
        del model

    if rank == 0:
        cpu = None
        if world == 1 and not args.skip_cpu:
            tf, ms, sample, threads = cpu_reference_rate(cfg, 20.0, 2, 1)
            cpu = {"value": tf, "unit": "TFLOP/s", "cores": threads, "kind": "port", "sample": sample,
                   "ms_per_sample": ms}
        line = {
            "metric": "dit_attention_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_kernel, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": cfg["name"] + " -- attention stack (q/k RMSNorm + 3-D RoPE, self-attention, "
                       "512-key cross-attention) of all layers", "video_tokens": L, "text_tokens": text_len,
                       "dim": dim, "heads": heads, "layers": layers,
                       "sharding": "none" if world == 1 else (
                           f"Ulysses heads/{world}, exchange fused into the kernels over NVLink peer memory"
                           if ctx is not None else f"Ulysses heads/{world} over NCCL all-to-all"),
                       "l2_policy": "inputs larger than L2: 4 rotating layer-input sets of "
                                    f"{3 * s * dim * 2 / 1e6:.0f} MB each"},
            "attention_flop_per_step": flop_step,
            "denoise_step_ms": denoise_ms,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_prologue": roofline_prologue,
            "roofline_gemm": roofline_gemm,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: Temperature Modality Alignment cross-attention sweep
# ------------------------------------------------------------------------------------------------------
def run_tma_sweep(args):
    """512 text tokens x {32 760, 75 600} video tokens across the 50-step flow schedule (2 DiT calls per step
    with classifier-free guidance -> call index c = 0..99, weight w(c) from univid_b200.tma).  Per call: the
    text-weighted k-norm prologue on the 512 context rows + the fused cross-attention kernel (per-key
    post-softmax weight, value bias) -- `kernel` -- and the public WanCrossAttention.forward(text_weight=,
    text_len=) with its q/k/v/o linears -- `module`.  Under torchrun every rank takes L/p query rows (no
    communication: the context is replicated)."""
    import torch.distributed as dist
    from univid_b200 import _ext, tma
    mdl = importlib.import_module("univid_b200.wan.modules.model")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
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
            w_vecs = []
            for wt in sorted(set(weights)):
                v_ = torch.ones(text_len, dtype=torch.float32, device=dev)
                v_[:tl] = wt
                w_vecs.append((wt, v_))
            w_of = dict(w_vecs)
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
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for c in range(calls):
                    fn(c)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / calls
                if world > 1:
                    t = torch.tensor([ms], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                timing[name] = ms
        flop = 4.0 * L * text_len * heads * 128
        bytes_x = 4.0 * L * dim + 4.0 * text_len * dim
        res[key] = {"video_tokens": L, "heads": heads, "text_len_weighted": tl, "calls": calls,
                    "kernel_ms_per_call": timing["kernel"], "module_ms_per_call": timing["module"],
                    "kernel_tflops": flop / (timing["kernel"] * 1e-3) * 1e-12,
                    "kernel_qo_stream_gbs": bytes_x / (timing["kernel"] * 1e-3) * 1e-9,
                    "module_tflops_attn_only": flop / (timing["module"] * 1e-3) * 1e-12}
        del ca, x, q, out
        torch.cuda.empty_cache()
    if rank == 0:
        peaks = load_peaks()
        line = {"metric": "tma_cross_attention_sweep_tflops", "value": res["1.3B"]["kernel_tflops"], "unit": "TFLOP/s",
                "n_gpus": world, "steps": calls, "warmup": 3, "ms_per_step": res["1.3B"]["kernel_ms_per_call"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "Temperature Modality Alignment cross-attention sweep: 512 text tokens x "
                                       "{32760, 75600} video tokens, 50 flow steps x 2 CFG calls, cosine 1.3 -> 1.0",
                           "weights_first_last": [weights[0], weights[-1]], "transition_calls": int(50 * 0.4)},
                "sweep": res, "peak_tflops_burst": peaks["tflops_burst"], "hbm_gbs": peaks["hbm_gbs"],
                "gpu_launches": 2 * calls * 2}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default=None, choices=list(CONFIGS))
    ap.add_argument("--workload", default="attention", choices=["attention", "tma-sweep"],
                    help="attention = the denoise-step attention stack (default, BASELINE configs[1..3]); "
                         "tma-sweep = the text-weighted cross-attention sweep (configs[4])")
    ap.add_argument("--skip-cpu", action="store_true", help="omit the cpu_baseline leg")
    ap.add_argument("--skip-denoise", action="store_true", help="omit the full WanModel denoise-step timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    # 12 heads shard over 2 / 4 GPUs (BASELINE configs[2]); 8 GPUs need the 40-head 14B model (configs[3])
    cfg_key = args.config or ("14B" if args.gpus == 8 else "1.3B")
    if args.impl == "reference":
        run_reference_arm(args, cfg_key)
    elif args.workload == "tma-sweep":
        run_tma_sweep(args)
    else:
        run_native_arm(args, cfg_key)


if __name__ == "__main__":
    main()
