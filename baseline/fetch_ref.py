"""Materialise the UNMODIFIED reference Python under ``baseline/_ref/`` (git-ignored; it travels to
the GPU box with the gpurun snapshot, where ``/root/reference`` does not exist).

    python baseline/fetch_ref.py            # no-op when /root/reference is absent

Only the files that the model-level parity test, ``bench.py``'s full-model leg and the reference
arm import are taken (Python sources of ``models/``, ``pointnet2/``, ``data/`` and the two
scripts); nothing is edited and nothing here enters the git history. The reference's CUDA sources
are NOT copied -- they are compiled where they lie by ``oracle/Makefile``.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MOCOPCI_REFERENCE", "/root/reference")
TAKE_DIRS = ("models", "pointnet2", "data")
TAKE_FILES = ("train.py", "test.py", "LICENSE", "README.md")


def fetch(src=SRC, dest=DEST, verbose=False):
    if not os.path.isdir(os.path.join(src, "models")):
        return None
    n = 0
    for d in TAKE_DIRS:
        for root, _dirs, files in os.walk(os.path.join(src, d)):
            for f in files:
                if not f.endswith((".py", ".txt")):
                    continue
                s = os.path.join(root, f)
                t = os.path.join(dest, os.path.relpath(s, src))
                os.makedirs(os.path.dirname(t), exist_ok=True)
                if not os.path.exists(t) or os.path.getmtime(t) < os.path.getmtime(s):
                    shutil.copy2(s, t)
                n += 1
    for f in TAKE_FILES:
        s = os.path.join(src, f)
        if os.path.exists(s):
            os.makedirs(dest, exist_ok=True)
            shutil.copy2(s, os.path.join(dest, f))
            n += 1
    if verbose:
        print(f"baseline/_ref: {n} reference files from {src}")
    return dest


def root():
    """Path of a reference checkout usable for imports: the live one if present, else the copy."""
    if os.path.isdir(os.path.join(SRC, "models")):
        return SRC
    if os.path.isdir(os.path.join(DEST, "models")):
        return DEST
    return None


if __name__ == "__main__":
    sys.exit(0 if fetch(verbose=True) or True else 1)
