"""pytest plugin (test infrastructure): routes the reference's `pymbar.mbar_solvers` through the REAL
pymbar_b200 backend (libmbar_b200.so on cuda:0) before the reference's own test files are collected.
Counterpart of tests/_mirror_plugin.py, which does the same with the CPU stand-in."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CALLS = {"solve": 0, "logW": 0}


def pytest_configure(config):
    import pymbar.mbar_solvers as ref_ms

    import pymbar_b200
    from pymbar_b200 import mbar_solvers as ours

    pymbar_b200._lib.load()                      # fails loudly without the compiled library
    pymbar_b200.install()
    assert ref_ms.solve_mbar_for_all_states is ours.solve_mbar_for_all_states
    assert ref_ms.mbar_log_W_nk is ours.mbar_log_W_nk
    # count the calls MBAR makes into the backend so the log proves the kernels served them
    solve, logw = ref_ms.solve_mbar_for_all_states, ref_ms.mbar_log_W_nk

    def counted_solve(*a, **k):
        CALLS["solve"] += 1
        return solve(*a, **k)

    def counted_logw(*a, **k):
        CALLS["logW"] += 1
        return logw(*a, **k)

    ref_ms.solve_mbar_for_all_states = counted_solve
    ref_ms.mbar_log_W_nk = counted_logw


def pytest_terminal_summary(terminalreporter):
    import pymbar_b200

    loaded = [ln.split()[-1] for ln in open("/proc/self/maps") if "libmbar_b200.so" in ln]
    from pymbar_b200 import facade

    terminalreporter.write_line(f"pymbar_b200 facade: {facade.STATS} (Log_W_nk tickets issued inside MBAR.__init__ / "
                                "downloaded on first read; estimators and expectations served from device moments)")
    terminalreporter.write_line(
        f"pymbar_b200 backend: solve_mbar_for_all_states calls={CALLS['solve']} mbar_log_W_nk calls={CALLS['logW']} "
        f"native library mapped={sorted(set(loaded))}")
