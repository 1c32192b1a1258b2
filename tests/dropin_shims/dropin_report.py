"""pytest plugin (-p dropin_report): prints which module the reference's tests imported as `pygenray`, so that
tests/test_reference_suite.py can assert it was the alias of the product and not a copy of the reference."""
import sys


def pytest_sessionfinish(session, exitstatus):
    mod = sys.modules.get("pygenray")
    lr = sys.modules.get("pygenray.launch_rays")
    print("\nDROPIN pygenray=%s launch_rays=%s" % (getattr(mod, "__file__", None), getattr(lr, "__file__", None)))
