"""Load the UNMODIFIED reference hot-path sources -- TEST INFRASTRUCTURE.

From /root/reference when it is mounted (the build container: tests/golden/make_*.py, tests/test_reference_compat.py
pin the oracle with it), otherwise from the archive oracle/_ref/reference_hotpath.tar.gz that oracle/make_ref.py
writes from the same files (git-ignored; it travels to the GPU box like a built .so, is unpacked into a temporary
directory per process and verified against its sha256 manifest; bench.py's reference arm and
tests/test_reference_binding_gpu.py use it).  The files are imported as they are (recipe: SURVEY.md sec. 8c).

  * a 4-symbol stub stands in for `diffusers` (only ConfigMixin / register_to_config / ModelMixin are
    used, model.py:6-7);
  * models/wan/utils/modules/{attention,model}.py are loaded into a synthetic package so the
    relative import `from .attention import flash_attention` resolves;
  * flash_attention is routed to attention()'s torch-SDPA branch (attention.py:164-179) -- "the
    reference's torch SDPA path" of BASELINE.json -- because the flash branch asserts CUDA;
  * models/wan/distributed/{util,ulysses,sequence_parallel}.py load into a synthetic `wan` package;
  * class Wan22ContextWrapper is cut out of models/model_pipeline.py with `ast` (the module itself
    pip-installs and writes files at import time, model_pipeline.py:42-132).
"""
import ast
import importlib.util
import logging
import os
import sys
import types

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED_DIR = os.path.join(_HERE, "_ref")           # written by oracle/make_ref.py (git-ignored, travels to the GPU box)
_UNPACKED = None


def _staged_root():
    """The staged archive unpacked into a per-process temporary directory (removed at exit), or a path that does not
    exist when nothing was staged."""
    global _UNPACKED
    if _UNPACKED is None:
        if not os.path.isfile(os.path.join(STAGED_DIR, "MANIFEST.json")):
            return os.path.join(STAGED_DIR, "missing")
        import atexit
        import shutil
        import tempfile
        from oracle import make_ref
        _UNPACKED = make_ref.unpack(tempfile.mkdtemp(prefix="uvb_ref_"), STAGED_DIR)
        atexit.register(shutil.rmtree, _UNPACKED, True)
    return _UNPACKED


def _pick_root():
    env = os.environ.get("UNIVID_REFERENCE")
    if env:
        return env
    if os.path.isfile("/root/reference/models/wan/utils/modules/model.py"):
        return "/root/reference"
    return _staged_root()


REF_ROOT = _pick_root()
MODULES_DIR = os.path.join(REF_ROOT, "models/wan/utils/modules")
DIST_DIR = os.path.join(REF_ROOT, "models/wan/distributed")
PIPELINE = os.path.join(REF_ROOT, "models/model_pipeline.py")
ANIMATE = os.path.join(MODULES_DIR, "animate/model_animate.py")
UNIPC = os.path.join(REF_ROOT, "models/wan/utils/fm_solvers_unipc.py")


def available():
    return os.path.isfile(os.path.join(MODULES_DIR, "model.py"))


def _install_diffusers_stub():
    if "diffusers" in sys.modules:
        return
    d = types.ModuleType("diffusers")
    cu = types.ModuleType("diffusers.configuration_utils")
    m = types.ModuleType("diffusers.models")
    mu = types.ModuleType("diffusers.models.modeling_utils")

    class ConfigMixin:
        pass

    def register_to_config(fn):
        return fn

    class ModelMixin(nn.Module):
        pass

    cu.ConfigMixin, cu.register_to_config, mu.ModelMixin = ConfigMixin, register_to_config, ModelMixin
    d.configuration_utils, d.models, m.modeling_utils = cu, m, mu
    sys.modules.update({"diffusers": d, "diffusers.configuration_utils": cu,
                        "diffusers.models": m, "diffusers.models.modeling_utils": mu})


def _load(pkg_name, mod_name, path):
    full = f"{pkg_name}.{mod_name}"
    spec = importlib.util.spec_from_file_location(full, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def load_modules():
    """Returns (attention_module, model_module) of the reference, SDPA-routed."""
    if "modules" in _CACHE:
        return _CACHE["modules"]
    _install_diffusers_stub()
    for name, path in (("uvref_wan", None), ("uvref_wan.modules", MODULES_DIR)):
        pkg = types.ModuleType(name)
        pkg.__path__ = [path] if path else []
        sys.modules[name] = pkg
    att = _load("uvref_wan.modules", "attention", os.path.join(MODULES_DIR, "attention.py"))
    att.FLASH_ATTN_2_AVAILABLE = False
    att.FLASH_ATTN_3_AVAILABLE = False
    model = _load("uvref_wan.modules", "model", os.path.join(MODULES_DIR, "model.py"))

    def flash_attention_via_sdpa(*args, version=None, **kwargs):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return att.attention(*args, **kwargs)

    model.flash_attention = flash_attention_via_sdpa
    _CACHE["modules"] = (att, model)
    return att, model


def load_distributed():
    """Returns (util, ulysses, sequence_parallel) of the reference loaded under `uvref_wan`."""
    if "dist" in _CACHE:
        return _CACHE["dist"]
    load_modules()
    pkg = types.ModuleType("uvref_wan.distributed")
    pkg.__path__ = [DIST_DIR]
    sys.modules["uvref_wan.distributed"] = pkg
    util = _load("uvref_wan.distributed", "util", os.path.join(DIST_DIR, "util.py"))
    uly = _load("uvref_wan.distributed", "ulysses", os.path.join(DIST_DIR, "ulysses.py"))
    sp = _load("uvref_wan.distributed", "sequence_parallel", os.path.join(DIST_DIR, "sequence_parallel.py"))
    _CACHE["dist"] = (util, uly, sp)
    return util, uly, sp


def load_context_wrapper():
    """The reference's Wan22ContextWrapper class, cut out of model_pipeline.py by AST."""
    if "wrapper" in _CACHE:
        return _CACHE["wrapper"]
    src = open(PIPELINE).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Wan22ContextWrapper")
    code = compile(ast.Module(body=[node], type_ignores=[]), PIPELINE, "exec")
    ns = {"torch": torch, "logging": logging, "ContextProjector": object, "CrossAttentionConfig": object}
    exec(code, ns)
    _CACHE["wrapper"] = ns["Wan22ContextWrapper"]
    return _CACHE["wrapper"]


def load_animate_attention():
    """(WanAnimateSelfAttention, WanAnimateCrossAttention, namespace) of the reference, cut out of
    models/wan/utils/modules/animate/model_animate.py by AST (the module itself imports face/motion encoders,
    diffusers.loaders and a relative path that does not exist in the tree) and bound to the reference's own
    WanSelfAttention / WanRMSNorm / rope_apply and the SDPA-routed flash_attention of load_modules()."""
    if "animate" in _CACHE:
        return _CACHE["animate"]
    _, model = load_modules()
    tree = ast.parse(open(ANIMATE).read())
    want = ("WanAnimateSelfAttention", "WanAnimateCrossAttention")
    nodes = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in want]
    code = compile(ast.Module(body=nodes, type_ignores=[]), ANIMATE, "exec")
    ns = {"torch": torch, "nn": nn, "WanSelfAttention": model.WanSelfAttention, "WanRMSNorm": model.WanRMSNorm,
          "rope_apply": model.rope_apply, "flash_attention": model.flash_attention}
    exec(code, ns)
    # the namespace is returned too: rebinding ns["flash_attention"] switches the attention route of both classes
    _CACHE["animate"] = (ns["WanAnimateSelfAttention"], ns["WanAnimateCrossAttention"], ns)
    return _CACHE["animate"]


def load_unipc_scheduler():
    """The reference's FlowUniPCMultistepScheduler (models/wan/utils/fm_solvers_unipc.py), loaded unmodified.  The
    `diffusers` names it imports are stood in for by the minimum that gives them their documented behaviour:
    `register_to_config` records the constructor arguments (defaults included) in `self.config`,
    `ConfigMixin.register_to_config(**kw)` updates it, SchedulerOutput carries `prev_sample`."""
    if "unipc" in _CACHE:
        return _CACHE["unipc"]
    import dataclasses
    import enum
    import functools
    import inspect
    _install_diffusers_stub()

    class ConfigMixin:
        def register_to_config(self, **kw):
            if not hasattr(self, "config"):
                self.config = types.SimpleNamespace()
            for k, v in kw.items():
                setattr(self.config, k, v)

    def register_to_config(init):
        sig = inspect.signature(init)

        @functools.wraps(init)
        def wrapped(self, *args, **kwargs):
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            ConfigMixin.register_to_config(self, **cfg)
            init(self, *args, **kwargs)
        return wrapped

    class SchedulerMixin:
        pass

    @dataclasses.dataclass
    class SchedulerOutput:
        prev_sample: torch.Tensor

    class KarrasDiffusionSchedulers(enum.Enum):
        UniPCMultistepScheduler = 1

    cu = types.ModuleType("diffusers.configuration_utils")
    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
    sch = types.ModuleType("diffusers.schedulers")
    su = types.ModuleType("diffusers.schedulers.scheduling_utils")
    su.KarrasDiffusionSchedulers, su.SchedulerMixin, su.SchedulerOutput = KarrasDiffusionSchedulers, SchedulerMixin, SchedulerOutput
    ut = types.ModuleType("diffusers.utils")
    ut.deprecate = lambda *a, **k: None
    ut.is_scipy_available = lambda: False
    saved = {k: sys.modules.get(k) for k in ("diffusers.configuration_utils", "diffusers.schedulers",
                                             "diffusers.schedulers.scheduling_utils", "diffusers.utils")}
    sys.modules.update({"diffusers.configuration_utils": cu, "diffusers.schedulers": sch,
                        "diffusers.schedulers.scheduling_utils": su, "diffusers.utils": ut})
    try:
        spec = importlib.util.spec_from_file_location("uvref_fm_solvers_unipc", UNIPC)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():       # the model loader's identity-decorator stub stays in place for model.py
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE["unipc"] = mod.FlowUniPCMultistepScheduler
    return _CACHE["unipc"]
