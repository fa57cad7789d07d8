"""Put the UNMODIFIED reference under baseline/_ref/ so it travels to the GPU box (TEST / BASELINE INFRASTRUCTURE ONLY).

    python oracle/install_ref.py            # build container only: /root/reference does not exist on the GPU box

The reference is a plain Python tree without setup.py / pyproject.toml, so `pip install --target baseline/_ref
/root/reference` has nothing to build; the "install" is a byte-for-byte copy of the importable packages the hot path
needs (models/, scheduler/, utils/, config/ and the reference's own smoke test).  baseline/_ref/ is git-ignored
(never part of the history, never part of the product) but NOT gpurun-ignored, so `bench.py --impl reference`, the
`torch_cuda_reference` leg of `bench.py` and the `-m gpu` tests can run the reference's own modules on the box.
Nothing under controlnet-pytorch_b200/ imports it.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CNB_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PARTS = ("models", "scheduler", "utils", "config", "test_distribution_matching.py", "LICENSE")


def tree_digest(root):
    h = hashlib.sha256()
    for part in PARTS:
        p = os.path.join(root, part)
        files = [p] if os.path.isfile(p) else sorted(
            os.path.join(d, f) for d, _, fs in os.walk(p) for f in fs if not f.endswith(".pyc"))
        for f in files:
            h.update(os.path.relpath(f, root).encode())
            with open(f, "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def install(force=False):
    """Copy SRC -> DST when SRC exists; returns DST if a usable copy is present afterwards, else None."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(os.path.join(DST, "models")) else None
    stamp = os.path.join(DST, ".digest")
    want = tree_digest(SRC)
    if not force and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == want:
                return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            shutil.copy2(s, d)
    with open(stamp, "w") as f:
        f.write(want)
    return DST


def ref_path():
    """Where the reference can be imported from on THIS box: the live checkout, else the installed copy, else None."""
    if os.path.isdir(os.path.join(SRC, "models")):
        return SRC
    if os.path.isdir(os.path.join(DST, "models")):
        return DST
    return None


if __name__ == "__main__":
    p = install(force="--force" in sys.argv)
    print(p if p else "reference not available: nothing installed")
