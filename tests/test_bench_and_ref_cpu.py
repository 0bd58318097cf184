"""CPU checks of the measurement plumbing: bench.py's flop accounting and reference arm (on a toy geometry that runs in
seconds), and the reference staging recipe oracle/make_ref.py."""
import hashlib
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import make_ref, ref_loader  # noqa: E402


def test_flop_accounting_matches_the_definitions():
    L, heads, text = 32760, 12, 512
    fs, fc = bench.flops_per_layer(L, heads, text)
    assert fs == 4.0 * L * L * heads * 128 and fc == 4.0 * L * text * heads * 128
    # judge's round-1 arithmetic: 2.009e14 attention flop per 1.3B step, 8.19e13 linear flop
    assert abs((fs + fc) * 30 / 2.009e14 - 1) < 2e-3
    assert abs(bench.linear_flops_per_layer(L, 1536, 8960, text) * 30 / 8.19e13 - 1) < 5e-3
    cfg = bench.CONFIGS[bench.HEADLINE]
    assert cfg["heads"] % 8 == 0 and (cfg["grid"][0] * cfg["grid"][1] * cfg["grid"][2]) % 8 == 0      # shards over 1/2/4/8 GPUs
    assert cfg["grid"][0] * cfg["grid"][1] * cfg["grid"][2] == 75600


def test_reference_arm_runs_the_reference_modules_on_all_threads():
    toy = dict(name="toy", dim=256, heads=2, layers=3, ffn=512, grid=(4, 6, 8), text_len=64)
    tf, ms, sample, threads, kind = bench.cpu_reference_rate(toy, 4.0, 2, 1)
    assert tf > 0 and ms > 0 and threads >= 1
    assert kind == ("reference" if ref_loader.available() else "port")
    assert "of 4 latent frames" in sample and ("unmodified reference modules" in sample) == (kind == "reference")


def test_reference_arm_cli_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")          # what torchrun exports: the arm must undo it
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1.3B", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "TFLOP/s" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["warmup"] == 0                                  # honours --steps / --warmup
    assert line["e2e"] == {"value": line["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["kind"] in ("reference", "port") and cb["cores"] == len(os.sched_getaffinity(0))
    # the arm reports on the GPU arm's config object (one definition for both); the bounded sample is named beside it
    assert line["config"] == bench.workload_config(bench.CONFIGS["1.3B"], 1) and line["sample"] == cb["sample"]


def test_both_arms_share_one_config_object():
    """workload_config is what both arms print: same keys as the recorded GPU lines, per-rank working set in l2_policy."""
    rec = json.loads([l for l in open(os.path.join(ROOT, "profiles/r02i/bench_8gpu_r02i.json")).read().strip().splitlines()
                      if l.startswith("{")][-1])["config"]
    now = bench.workload_config(bench.CONFIGS[bench.HEADLINE], 8)
    assert set(now) == set(rec)
    assert {k: v for k, v in now.items() if k != "l2_policy"} == {k: v for k, v in rec.items() if k != "l2_policy"}
    assert now["l2_policy"].startswith(rec["l2_policy"])
    assert bench.workload_config(bench.CONFIGS[bench.HEADLINE], 1)["sharding"] == "none"


@pytest.mark.skipif(not os.path.isfile("/root/reference/models/wan/utils/modules/model.py"), reason="/root/reference not mounted")
def test_make_ref_stages_the_reference_byte_for_byte(tmp_path):
    man = make_ref.stage(src="/root/reference", dest=str(tmp_path / "stage"), verbose=False)
    out = make_ref.unpack(str(tmp_path / "unpacked"), str(tmp_path / "stage"))
    for rel in make_ref.FILES:
        a = open(os.path.join("/root/reference", rel), "rb").read()
        b = open(os.path.join(out, rel), "rb").read()
        assert a == b and man["files"][rel] == hashlib.sha256(a).hexdigest()
    cut = open(os.path.join(out, make_ref.PIPELINE)).read()
    src = open(os.path.join("/root/reference", make_ref.PIPELINE)).read().splitlines(keepends=True)
    lo, hi = man["files"][make_ref.PIPELINE]["lines"]
    assert "".join(src[lo - 1:hi]) in cut and cut.count("\nclass Wan22ContextWrapper") == 1
    # a tampered archive is refused
    import json as _json
    mpath = tmp_path / "stage" / "MANIFEST.json"
    m = _json.loads(mpath.read_text())
    m["files"][make_ref.FILES[0]] = "0" * 64
    mpath.write_text(_json.dumps(m))
    with pytest.raises(RuntimeError, match="checksum"):
        make_ref.unpack(str(tmp_path / "again"), str(tmp_path / "stage"))


def test_staged_reference_is_what_the_loader_uses_when_the_tree_is_absent(monkeypatch):
    """On the GPU box there is no /root/reference: ref_loader unpacks oracle/_ref/reference_hotpath.tar.gz."""
    monkeypatch.delenv("UNIVID_REFERENCE", raising=False)
    picked = ref_loader._pick_root()
    if os.path.isdir("/root/reference"):
        assert picked == "/root/reference"
    staged = ref_loader._staged_root()
    if os.path.isfile(os.path.join(ref_loader.STAGED_DIR, "MANIFEST.json")):
        assert os.path.isfile(os.path.join(staged, "models/wan/utils/modules/model.py"))
        if not os.path.isdir("/root/reference"):
            assert picked == staged


@pytest.mark.parametrize("path", ["profiles/r02i/bench_r02i.json", "profiles/r02f/bench_8gpu_r02f.json"])
def test_recorded_bench_lines_satisfy_the_contract(path):
    """The committed bench lines (one GPU with the driver's --steps 20 --warmup 5, and 8 GPUs) carry every key the
    bench contract names, with consistent arithmetic."""
    line = json.loads([l for l in open(os.path.join(ROOT, path)).read().strip().splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in line, k
    assert line["unit"] == "TFLOP/s" and line["dtype"] == "bf16" and line["data"] == "synthetic"
    assert line["scaling"] == "strong" and line["vs_baseline"] is None and line["higher_is_better"] is True
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line["config"]["video_tokens"] == 75600 and line["config"]["heads"] == 40       # same config at every N
    # value = attention flop of the step / time
    assert abs(line["attention_flop_per_step"] / (line["ms_per_step"] * 1e-3) * 1e-12 / line["value"] - 1) < 1e-6
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] < line["value"]
    rf = line["roofline"]
    assert rf["bound"] == "tensor" and abs(rf["achieved"] / rf["peak"] - rf["frac"]) < 1e-9 and rf["frac"] < 1.0
    assert line["gpu_launches"] > 0 and set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if line["n_gpus"] == 1:
        cb = line["cpu_baseline"]
        assert cb["kind"] == "reference" and cb["cores"] >= 1 and "sample" in cb
        assert rf["traffic"] and 0.9 < rf["traffic"] / (4 * 75600 * 5120 * 2) < 1.2     # DRAM bytes vs algorithmic bytes
        assert line["roofline_prologue"]["bound"] == "hbm" and "back_to_back" in line["roofline_prologue"]
    else:
        pc = line["parity_check"]
        assert pc["ok"] is True and pc["max_abs_vs_fp32_rows"] <= 2e-2 and pc["cos_vs_fp32_rows"] >= 0.9999
        assert "self_attention" in line["kernel_split"]["segments"]
    den = line["denoise_step"]
    assert den["ms"] > 0 and den["flop"] > line["attention_flop_per_step"] and 0 < den["frac_of_sustained_peak"] < 1.05
