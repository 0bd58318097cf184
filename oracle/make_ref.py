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

The staging is ONE archive, oracle/_ref/reference_hotpath.tar.gz (+ MANIFEST.json with the sha256 of every member): a
build artefact like a compiled oracle/_ref/*.so would be, not a source tree inside the repo.  oracle/ref_loader.py
loads from /root/reference when it is mounted and otherwise unpacks the archive into a per-process temporary directory; bench.py's
`--impl reference` arm and `cpu_baseline` leg then time the reference's own code (`kind: "reference"`), and the
`-m gpu` binding test (tests/test_reference_binding_gpu.py) runs the reference WanAttentionBlock / WanModel with this
repo's attention modules swapped in.  Nothing under oracle/_ref/ is ever committed or imported by univid_b200/.
A MANIFEST.json with the sha256 of every staged file is written next to them.
"""
import ast
import hashlib
import io
import json
import os
import sys
import tarfile

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


ARCHIVE = "reference_hotpath.tar.gz"


def _sha(data):
    return hashlib.sha256(data).hexdigest()


def stage(src=SRC, dest=DEST, verbose=True):
    if not os.path.isfile(os.path.join(src, FILES[0])):
        raise FileNotFoundError(f"reference tree not found at {src}")
    os.makedirs(dest, exist_ok=True)
    manifest = {"source": src, "archive": ARCHIVE, "files": {}}
    members = []
    for rel in FILES:
        data = open(os.path.join(src, rel), "rb").read()
        members.append((rel, data))
        manifest["files"][rel] = _sha(data)
    # the one class of model_pipeline.py the path needs, as the exact source lines of the reference
    text = open(os.path.join(src, PIPELINE)).read()
    node = next(n for n in ast.parse(text).body if isinstance(n, ast.ClassDef) and n.name == PIPELINE_CLASS)
    lines = text.splitlines(keepends=True)[node.lineno - 1:node.end_lineno]
    cut = (f"# class {PIPELINE_CLASS}: lines {node.lineno}-{node.end_lineno} of the reference's {PIPELINE}, verbatim\n"
           + "".join(lines)).encode()
    members.append((PIPELINE, cut))
    manifest["files"][PIPELINE] = {"class": PIPELINE_CLASS, "lines": [node.lineno, node.end_lineno], "sha256": _sha(cut)}
    with tarfile.open(os.path.join(dest, ARCHIVE), "w:gz") as tar:
        for rel, data in members:
            info = tarfile.TarInfo(rel)
            info.size, info.mtime, info.mode = len(data), 0, 0o644
            tar.addfile(info, io.BytesIO(data))
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print(f"staged {len(members)} reference files in {os.path.join(dest, ARCHIVE)}")
    return manifest


def unpack(dest, archive_dir=DEST):
    """Extract the staged archive into `dest` (verifying the manifest); returns dest.  Used by oracle/ref_loader.py."""
    man = json.load(open(os.path.join(archive_dir, "MANIFEST.json")))
    with tarfile.open(os.path.join(archive_dir, man["archive"]), "r:gz") as tar:
        for m in tar.getmembers():
            if m.name.startswith("/") or ".." in m.name.split("/"):
                raise RuntimeError(f"unsafe member {m.name}")
            data = tar.extractfile(m).read()
            want = man["files"][m.name]
            want = want["sha256"] if isinstance(want, dict) else want
            if _sha(data) != want:
                raise RuntimeError(f"{m.name}: checksum mismatch with MANIFEST.json")
            out = os.path.join(dest, m.name)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            with open(out, "wb") as f:
                f.write(data)
    return dest


if __name__ == "__main__":
    stage()
    sys.exit(0)
