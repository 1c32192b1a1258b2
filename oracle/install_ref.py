"""Recipe: make the UNMODIFIED Python reference importable on the GPU box (test / bench infrastructure).

`pip install --no-index --target ... /root/reference` cannot work offline: the reference's build backend is
hatchling + hatch-vcs (pyproject.toml:1-3), neither of which is in the image or the wheelhouse.  The package is pure
Python, so an install IS a copy of `src/pygenray/`: this script copies that directory, byte for byte, to
`oracle/_ref/pygenray/` (git-ignored -- never part of the history -- but shipped to the GPU box with the working
tree) and records the sha256 of every file.  It is used by bench.py's `python_reference` baseline, together
with the xarray / matplotlib stand-ins of oracle/stubs/ (those two packages are not in the image).  The reference's
own test files and fixture (`tests/`) are copied the same way to `oracle/_ref/suite/tests/`, so that
tests/test_reference_suite.py can run them UNCHANGED against the product (also on the GPU box).

    python oracle/install_ref.py        # no-op when /root/reference is absent (the GPU box) or already copied
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/pygenray"
DST = os.path.join(HERE, "_ref", "pygenray")
SRC_TESTS = "/root/reference/tests"
# one directory down: the reference's tests are a package (tests/__init__.py), so pytest puts the PARENT of tests/ on
# sys.path -- which must not be oracle/_ref/ itself, where the reference package would shadow the module under test
DST_TESTS = os.path.join(HERE, "_ref", "suite", "tests")


def install(force=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    if os.path.isdir(DST) and os.path.isdir(DST_TESTS) and not force:
        return True
    sums = {}
    stale = os.path.join(HERE, "_ref", "tests")  # location used by an earlier revision of this script
    if os.path.isdir(stale):
        shutil.rmtree(stale)
    for src, dst in ((SRC, DST), (SRC_TESTS, DST_TESTS)):
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".pytest_cache"))
        for root, _, files in sorted(os.walk(dst)):
            for name in sorted(files):
                path = os.path.join(root, name)
                with open(path, "rb") as f:
                    sums[os.path.relpath(path, os.path.join(HERE, "_ref"))] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as f:
        json.dump({"source": [SRC, SRC_TESTS], "sha256": sums}, f, indent=1)
    return True


if __name__ == "__main__":
    print("reference available:", install())
