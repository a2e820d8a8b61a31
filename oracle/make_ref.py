"""TEST INFRASTRUCTURE -- recipe that materialises the UNMODIFIED reference under oracle/_ref/ (git-ignored).

    python oracle/make_ref.py            # needs /root/reference (build container only)

Copies /root/reference/Testing (test.py, dataloader.py, model/, data/vid1/*.png) verbatim to oracle/_ref/Testing and
writes oracle/_ref/MANIFEST.json with the sha256 of every file, so that the GPU box -- where /root/reference does not
exist -- can run the reference's own script and model package next to the drop-in (tests/test_reference_script_gpu.py)
and time the reference itself as the CPU baseline (bench.py, cpu_baseline.kind = "reference").  Nothing is edited:
the copy is byte-identical (the manifest is checked by tests/test_oracle.py when both trees are present).
oracle/_ref/ is listed in .gitignore (never part of the history) but not in .gpurunignore (it travels with the
snapshot).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs execute it.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/Testing"
DST = os.path.join(ROOT, "oracle", "_ref", "Testing")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def manifest(tree):
    out = {}
    for base, _, files in os.walk(tree):
        for name in sorted(files):
            if name.endswith(".pyc"):
                continue
            p = os.path.join(base, name)
            out[os.path.relpath(p, tree)] = sha256(p)
    return out


def make(force=False):
    """Returns True when oracle/_ref is present and matches the source tree (or the source tree is absent)."""
    man_path = os.path.join(os.path.dirname(DST), "MANIFEST.json")
    if not os.path.isdir(SRC):
        return os.path.isfile(man_path)
    want = manifest(SRC)
    if not force and os.path.isfile(man_path) and json.load(open(man_path)).get("files") == want:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    assert manifest(DST) == want
    with open(man_path, "w") as f:
        json.dump({"source": SRC, "files": want}, f, indent=1, sort_keys=True)
    print(f"[make_ref] copied {len(want)} files to {os.path.relpath(DST, ROOT)}", file=sys.stderr)
    return True


if __name__ == "__main__":
    make(force="--force" in sys.argv)
