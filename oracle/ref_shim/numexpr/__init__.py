"""Stand-in for the (absent) numexpr package — TEST INFRASTRUCTURE ONLY.

pymbar/utils.py:29 imports numexpr at module top; it is only used inside
pymbar.utils.logsumexp (utils.py:324-330), which is not on the solver path.
This stub lets `import pymbar` succeed in the build container so that
oracle/make_golden.py can run the unmodified reference from /root/reference.
It never travels into the product path.
"""
import sys

import numpy as np

__version__ = "2.10.2"   # (pandas probes the version of an installed numexpr; any valid PEP 440 string)


def evaluate(expr, local_dict=None, global_dict=None, **_kw):
    frame = sys._getframe(1)
    ns = dict(frame.f_globals if global_dict is None else global_dict)
    ns.update(frame.f_locals if local_dict is None else local_dict)
    ns.update(exp=np.exp, log=np.log)
    return eval(expr, {"__builtins__": {}}, ns)
