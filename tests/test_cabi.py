"""The C-ABI boundary on a machine without a GPU: the library builds, loads, and exports every symbol
include/univid_b200.h declares.  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "univid_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uvb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    from univid_b200 import build
    return build.build(with_tests=False)


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for name in ("uvb_qk_norm_rope", "uvb_fmha_fwd_bf16", "uvb_xattn_fwd_bf16", "uvb_head_scatter_bf16",
                 "uvb_last_error", "uvb_version", "uvb_block_glue", "uvb_linear_bf16", "uvb_unipc_step"):
        assert name in syms


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_python_binding_lists_the_same_symbols(libpath):
    from univid_b200 import _ext
    assert sorted(_ext.EXPORTS) == _declared_symbols()
    assert _ext.lib().uvb_version() == _ext.ABI_VERSION
    assert _ext.lib().uvb_last_error() == b""


def test_header_cites_reference_call_sites():
    src = open(HEADER).read()
    for cite in ("attention.py:96,113,175", "model.py:77-85", "model.py:38-66", "util.py:27",
                 "model_pipeline.py:1756-1803", "model.py:119-122", "model.py:212-214", "fm_solvers_unipc.py:657-741",
                 "textimage2video.py:385-386"):
        assert cite in src


def test_library_contains_blackwell_instructions(libpath):
    """SASS evidence that the attention kernel is tcgen05/TMEM/TMA code (B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", libpath], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass and "HGMMA" not in sass
    # the GEMM's CTA-pair variant: cta_group::2 MMA, pair TMA loads, multicast commit
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in sass, mnemonic


def test_missing_library_fails_loudly(monkeypatch):
    from univid_b200 import _ext
    monkeypatch.setattr(_ext, "_lib", None)
    monkeypatch.setattr(_ext, "LIB_PATH", "/nonexistent/libunivid_b200.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        _ext.lib()
