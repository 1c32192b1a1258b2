"""The reference's OWN test files, run UNCHANGED against pygenray_b200 (drop-in check at the Python boundary).

`oracle/install_ref.py` copies /root/reference/tests byte for byte to the git-ignored `oracle/_ref/suite/tests/` (shipped to
the GPU box with the working tree, never part of the history).  `tests/dropin_shims/` aliases `pygenray` -> the
product and `xarray` -> the product's minimal DataArray; `matplotlib` is the import stand-in of `oracle/stubs/`
(plotting is out of scope: the `*plot*` tests are deselected, nothing else).

* CPU (`-m "not gpu"`): tests/test_environment.py + tests/test_ray_objects.py of the reference -- containers,
  environment validation ladder, eflat round trips, save_mat.
* GPU (`-m gpu`): tests/test_physics.py of the reference -- the analytic tests A-D (Snell invariant, straight-line
  travel time and depth, turning depth, Hamiltonian), the regression fixture E with the reference's own tolerances,
  the near-vertical no-crash tests F and the mirrored / backwards shots G, all through `shoot_ray` / `shoot_rays` of
  the product, i.e. through libprt.so on the device.
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "suite", "tests")
SHIMS = os.path.join(ROOT, "tests", "dropin_shims")
STUBS = os.path.join(ROOT, "oracle", "stubs")  # matplotlib stand-in (its xarray is shadowed by the shim: SHIMS comes first)

needs_ref_tests = pytest.mark.skipif(not os.path.isdir(REF_TESTS),
                                     reason="oracle/_ref/suite/tests absent (made by oracle/install_ref.py where /root/reference exists)")


def _run(files, timeout):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([SHIMS, STUBS, ROOT])
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "-p", "dropin_report", "-s", "-k", "not plot", "--rootdir", REF_TESTS,
           "-c", os.devnull] + [os.path.join(REF_TESTS, f) for f in files]
    out = subprocess.run(cmd, cwd=REF_TESTS, env=env, capture_output=True, text=True, timeout=timeout)
    tail = (out.stdout + out.stderr)[-4000:]
    m = re.search(r"(\d+) passed", out.stdout)
    # the module under test was the product (tests/dropin_shims/pygenray -> pygenray_b200), not the reference copy
    who = re.search(r"DROPIN pygenray=(\S+) launch_rays=(\S+)", out.stdout)
    assert who and who.group(1).startswith(SHIMS) and os.sep + "pygenray_b200" + os.sep in who.group(2), tail
    return out.returncode, int(m.group(1)) if m else 0, tail


@needs_ref_tests
def test_reference_container_and_environment_tests_pass_on_the_product():
    rc, passed, tail = _run(["test_environment.py", "test_ray_objects.py"], 600)
    assert rc == 0, tail
    assert passed >= 50, tail  # 56 tests, 6 of them plotting


@pytest.mark.gpu
@needs_ref_tests
def test_reference_physics_tests_pass_on_the_product():
    rc, passed, tail = _run(["test_physics.py"], 900)
    print(tail[-600:])
    assert rc == 0, tail
    assert passed >= 22, tail  # all of test_physics.py (22 tests incl. parametrised cases)
