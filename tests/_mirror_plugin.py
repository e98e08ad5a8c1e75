"""pytest plugin (test infrastructure): installs the pymbar_b200 mirror over pymbar.mbar_solvers with the
oracle-backed DeviceProblem stand-in, so the reference's OWN test files exercise the mirror's driver layer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle", "ref_shim"), "/root/reference"):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ["PYMBAR_B200_CACHE"] = "0"


def pytest_configure(config):
    import pymbar_b200
    from pymbar_b200 import mbar_solvers as ours
    from tests.test_driver_logic_cpu import OracleProblem

    ours.DeviceProblem = OracleProblem
    pymbar_b200._lib.load = lambda: None
    pymbar_b200.install()
