"""Recipe that vendors the REAL reference next to the oracle  --  TEST / BENCH INFRASTRUCTURE.

    python oracle/build_ref.py            (also run by __graft_entry__.build())

The reference (Harshs27/uGLAD) is pure Python: there is nothing to compile and it cannot be
pip-installed in this image (build backend `hatchling` absent, `requires-python >= 3.13`,
DESIGN.md).  This recipe therefore copies the package tree `/root/reference/uglad` UNMODIFIED
into `oracle/_ref/uglad` -- a build output: `oracle/_ref/` is git-ignored (no reference source
ever enters the history) but not gpurun-ignored, so it travels to the GPU box exactly like the
built `libuglad_b200.so` does.  `oracle/ref_loader.py` imports it (with import-only stand-ins
for the plotting packages the image lacks).

Consumers: `bench.py --impl reference` / `cpu_baseline` (kind "reference") and
`tests/test_reference_vendored.py` (oracle port == vendored reference).  Nothing under
`uglad_b200/` may import it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/uglad"
DST = os.path.join(HERE, "_ref")


def build(verbose: bool = True) -> bool:
    """Copy the reference package into oracle/_ref.  Returns True when oracle/_ref is usable
    afterwards (freshly copied, or left over from an earlier run when /root/reference is absent,
    as on the GPU box)."""
    if not os.path.isdir(SRC):
        ok = os.path.isfile(os.path.join(DST, "uglad", "main.py"))
        if verbose:
            print(f"build_ref: {SRC} not present; existing oracle/_ref {'kept' if ok else 'absent'}")
        return ok
    pkg = os.path.join(DST, "uglad")
    if os.path.isdir(pkg):
        shutil.rmtree(pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(SRC, pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for root, _, files in os.walk(pkg):
        for f in sorted(files):
            p = os.path.join(root, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"build_ref: copied {len(manifest)} files from {SRC} to {pkg}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
