"""Stage the UNMODIFIED reference hot-path sources under oracle/_ref/ -- TEST INFRASTRUCTURE.

    python oracle/make_ref.py            (needs /root/reference; __graft_entry__.build() runs it when present)

`/root/reference` does not exist on the GPU box, and the reference is not an installable package (no setup.py /
pyproject; importing the tree pip-installs and writes files).  This recipe copies, byte for
byte, the few Python files the hot path consists of into the git-ignored (but not gpurun-ignored) directory
oracle/_ref/, mirroring their relative paths, so that they travel to the GPU box the way a built .so does:

    models/wan/utils/modules/{model,attention}.py          WanSelfAttention / WanCrossAttention / WanModel, attention()
    models/wan/utils/modules/animate/model_animate.py      Wan-Animate attention classes (cut out by AST at load time)
    models/wan/distributed/{util,ulysses,sequence_parallel}.py
    models/wan/utils/fm_solvers_unipc.py                   FlowUniPCMultistepScheduler
    models/model_pipeline.py                               ONLY the source text of class Wan22ContextWrapper (the module
                                                           itself pip-installs at import time, model_pipeline.py:42-132)

oracle/ref_loader.py loads from /root/reference when it is mounted and from oracle/_ref/ otherwise; bench.py's
`--impl reference` arm and `cpu_baseline` leg then time the reference's own code (`kind: "reference"`), and the
`-m gpu` binding test (tests/test_reference_binding_gpu.py) runs the reference WanAttentionBlock / WanModel with this
repo's attention modules swapped in.  Nothing under oracle/_ref/ is ever committed or imported by univid_b200/.
A MANIFEST.json with the sha256 of every staged file is written next to them.
"""
import ast
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("UNIVID_REFERENCE", "/root/reference")

FILES = (
    "models/wan/utils/modules/model.py",
    "models/wan/utils/modules/attention.py",
    "models/wan/utils/modules/animate/model_animate.py",
    "models/wan/distributed/util.py",
    "models/wan/distributed/ulysses.py",
    "models/wan/distributed/sequence_parallel.py",
    "models/wan/utils/fm_solvers_unipc.py",
)
PIPELINE = "models/model_pipeline.py"
PIPELINE_CLASS = "Wan22ContextWrapper"


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def stage(src=SRC, dest=DEST, verbose=True):
    if not os.path.isfile(os.path.join(src, FILES[0])):
        raise FileNotFoundError(f"reference tree not found at {src}")
    manifest = {"source": src, "files": {}}
    for rel in FILES:
        out = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
        manifest["files"][rel] = _sha(out)
    # the one class of model_pipeline.py the path needs, as the exact source lines of the reference
    text = open(os.path.join(src, PIPELINE)).read()
    node = next(n for n in ast.parse(text).body if isinstance(n, ast.ClassDef) and n.name == PIPELINE_CLASS)
    lines = text.splitlines(keepends=True)[node.lineno - 1:node.end_lineno]
    out = os.path.join(dest, PIPELINE)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        f.write(f"# class {PIPELINE_CLASS}: lines {node.lineno}-{node.end_lineno} of the reference's {PIPELINE}, verbatim\n")
        f.writelines(lines)
    manifest["files"][PIPELINE] = {"class": PIPELINE_CLASS, "lines": [node.lineno, node.end_lineno], "sha256": _sha(out)}
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print(f"staged {len(manifest['files'])} reference files under {dest}")
    return manifest


if __name__ == "__main__":
    stage()
    sys.exit(0)
