"""Recipe: make the UNMODIFIED Python reference importable on the GPU box (test / bench infrastructure).

`pip install --no-index --target ... /root/reference` cannot work offline: the reference's build backend is
hatchling + hatch-vcs (pyproject.toml:1-3), neither of which is in the image or the wheelhouse.  The package is pure
Python, so an install IS a copy of `src/pygenray/`: this script copies that directory, byte for byte, to
`oracle/_ref/pygenray/` (git-ignored -- never part of the history -- but shipped to the GPU box with the working
tree) and records the sha256 of every file.  It is used by bench.py's `python_reference` baseline only, together
with the xarray / matplotlib stand-ins of oracle/stubs/ (those two packages are not in the image).

    python oracle/install_ref.py        # no-op when /root/reference is absent (the GPU box) or already copied
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/pygenray"
DST = os.path.join(HERE, "_ref", "pygenray")


def install(force=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    if os.path.isdir(DST) and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    sums = {}
    for name in sorted(os.listdir(DST)):
        with open(os.path.join(DST, name), "rb") as f:
            sums[name] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": sums}, f, indent=1)
    return True


if __name__ == "__main__":
    print("reference available:", install())
