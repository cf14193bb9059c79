#!/usr/bin/env python
"""Test infrastructure -- recipe for `oracle/_ref/`: the UNMODIFIED reference, made importable where /root/reference is absent.

The reference is plain Python with no packaging (no setup.py / pyproject), so "building" it is copying the two packages the
sweep path lives in -- `tensor/` (node, network, bregman, layers, utils, module, ...) and `models/` (tensor_train.py, tnml.py)
-- byte for byte from where they lie under /root/reference into the git-ignored `oracle/_ref/`, which travels to the GPU box
with the snapshot like the built `.so`.  Nothing is edited; `MANIFEST.json` records the sha256 of every file so a test can
prove it.  What the authors' environment has and this image lacks (opt_einsum, matplotlib) comes from `oracle/standins/`
(our own files, committed).

    python oracle/make_ref.py            # (re)creates oracle/_ref/ ; a no-op message when /root/reference is absent

Users: `bench.py --impl reference` and the `cpu_baseline` leg (kind "reference"), `tests/` (checker only).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TN_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PACKAGES = ("tensor", "models")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    if not os.path.isdir(os.path.join(REF, "tensor")):
        print(f"make_ref: {REF} is not mounted here; keeping whatever oracle/_ref/ holds", file=sys.stderr)
        return 0
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    manifest = {}
    for pkg in PACKAGES:
        for name in sorted(os.listdir(os.path.join(REF, pkg))):
            if not name.endswith(".py"):
                continue
            src = os.path.join(REF, pkg, name)
            os.makedirs(os.path.join(OUT, pkg), exist_ok=True)
            dst = os.path.join(OUT, pkg, name)
            shutil.copyfile(src, dst)
            manifest[f"{pkg}/{name}"] = sha(dst)
    json.dump({"source": REF, "files": manifest}, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    print(f"make_ref: {len(manifest)} files -> {OUT}", file=sys.stderr)
    return 0


def ref_root():
    """Directory to put on sys.path to import the reference: oracle/_ref when built, else /root/reference, else None."""
    if os.path.isfile(os.path.join(OUT, "tensor", "network.py")):
        return OUT
    if os.path.isfile(os.path.join(REF, "tensor", "network.py")):
        return REF
    return None


def activate():
    """Put the stand-ins and the reference on sys.path.  Must run BEFORE `import torch` for the opt_einsum stand-in to be seen.
    Returns the reference root or None."""
    root = ref_root()
    if root is None:
        return None
    sd = os.path.join(HERE, "standins")
    try:
        import matplotlib  # noqa: F401
        have_mpl = True
    except Exception:
        have_mpl = False
    try:
        import opt_einsum  # noqa: F401
        have_oe = True
    except Exception:
        have_oe = False
    if not (have_mpl and have_oe):
        # one directory holds both stand-ins; a real package, when installed, was imported above and wins through sys.modules
        sys.path.insert(0, sd)
    sys.path.insert(0, root)
    return root


if __name__ == "__main__":
    sys.exit(main())
