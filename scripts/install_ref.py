#!/usr/bin/env python
"""Put the UNMODIFIED reference next to the repo so it travels to the GPU box.

EleMisi/VAEL has no setup.py / pyproject.toml, so `pip install --target baseline/_ref` has nothing to
install; the reference is a directory of scripts run from its root (`python run_VAEL.py`).  This recipe
copies that directory verbatim (Python sources, the four Mario ProbLog tables, the two classifier
checkpoints; no .git, no README/licence/data icons) from /root/reference (read-only, present in the
build container only) to baseline/_ref/ (git-ignored, NOT gpurun-ignored), and writes a manifest with the
sha256 of every file so a test can show nothing was edited.  Run by __graft_entry__.build() whenever
/root/reference exists; on the GPU box the prebuilt copy is used as it arrived.

Users: tests/test_gpu_reference_dropin.py (the reference's own classes driving the CUDA path through
vael_b200.compat) and `bench.py --impl reference` (the reference's CPU path, timed).
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("VAEL_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP_EXT = (".py", ".pt")
TOP = ("models", "utils", "config.py", "const_define.py", "run_VAEL.py")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for blk in iter(lambda: fh.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def install(src=SRC, dst=DST, quiet=False):
    """-> manifest dict, or None when the reference is not available (GPU box)."""
    if not os.path.isdir(src):
        return None
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    manifest = {}
    for top in TOP:
        p = os.path.join(src, top)
        files = [p] if os.path.isfile(p) else [os.path.join(d, f) for d, _, fs in os.walk(p) for f in fs]
        for f in sorted(files):
            if not f.endswith(KEEP_EXT):
                continue
            rel = os.path.relpath(f, src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(f, out)
            os.chmod(out, 0o644)
            manifest[rel] = sha256(out)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump({"source": "EleMisi/VAEL (copied verbatim from " + src + ")", "files": manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print(f"installed {len(manifest)} reference files under {dst}")
    return manifest


if __name__ == "__main__":
    if install() is None:
        print(f"{SRC} not present: nothing installed", file=sys.stderr)
        sys.exit(1)
