"""Recipe: mirror the UNMODIFIED reference files of the hot path into oracle/_ref/ (test infrastructure).

    python -m oracle.make_ref            # needs /root/reference (the build container); a no-op elsewhere

The reference is pure Python; it is never copied into the git history (oracle/_ref/ is git-ignored) but, like the built
libkmunet.so, the mirror travels to the GPU box with the gpurun snapshot, so that

  * `bench.py --impl reference` times the reference's own model on the box's host cores (and on the B200, eager fp16
    autocast, for the `gpu_eager_reference` field),
  * `tests/test_gpu_reference_dropin.py` runs the reference's KM_UNetV3_SH.py on top of the drop-in operators.

Only the files the path imports are mirrored: the two model files, convKAN/, vim_block_init/, WPL/,
DySample_md.py, DAGEM_md.py, metrics.py.  Files are copied byte for byte; a MANIFEST with their sha256 is written next to
them so a test can tell a stale or edited mirror from the real thing.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("KMU_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("KM_UNetV3_SH.py", "KM_UNetV3_LAPS.py", "DySample_md.py", "DAGEM_md.py", "metrics.py")
DIRS = ("convKAN", "vim_block_init", "WPL")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make(verbose=False):
    """Returns the mirror path, or None when the reference tree is not present (GPU box: the prebuilt mirror is used)."""
    if not os.path.isfile(os.path.join(SRC, "convKAN", "KANlayers.py")):
        return DST if os.path.isfile(os.path.join(DST, "MANIFEST.json")) else None
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    rels = list(FILES)
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(SRC, d))):
            if name.endswith(".py"):
                rels.append(os.path.join(d, name))
    for rel in rels:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"mirrored {len(rels)} reference files into {DST}")
    return DST


def verify():
    """True when every mirrored file still has the sha256 recorded at mirror time."""
    path = os.path.join(DST, "MANIFEST.json")
    if not os.path.isfile(path):
        return False
    with open(path) as f:
        files = json.load(f)["files"]
    return all(os.path.isfile(os.path.join(DST, rel)) and _sha(os.path.join(DST, rel)) == h for rel, h in files.items())


if __name__ == "__main__":
    out = make(verbose=True)
    print(out if out else "reference tree not present and no mirror", file=sys.stderr if out is None else sys.stdout)
