"""Pack the UNMODIFIED reference package into oracle/_ref/pymbar_ref.zip — TEST INFRASTRUCTURE ONLY.

Why: the acceptance set for this path is the reference's own test files
(pymbar/tests/test_mbar_solvers.py:25-41, pymbar/tests/test_mbar.py).
/root/reference does not exist on the GPU box, so the only way to run real `pymbar.MBAR` against the
real kernels is to let the package travel with the gpurun snapshot.  oracle/_ref/ is git-ignored (the
history stays free of reference sources) but not gpurun-ignored.  Nothing under pymbar_b200/ ever
reads this archive; tests/test_gpu_reference_suite.py unpacks it into a temporary directory.

    python oracle/vendor_reference.py            # no-op (exit 0) when /root/reference is absent
"""
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/pymbar"
DST = os.path.join(HERE, "_ref", "pymbar_ref.zip")


def vendor(force=False):
    if not os.path.isdir(SRC):
        return None
    newest = 0.0
    files = []
    for root, _dirs, names in os.walk(SRC):
        if "__pycache__" in root:
            continue
        for n in names:
            if n.endswith((".pyc", ".pyo")):
                continue
            p = os.path.join(root, n)
            files.append(p)
            newest = max(newest, os.path.getmtime(p))
    if not force and os.path.exists(DST) and os.path.getmtime(DST) >= newest:
        return DST
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    with zipfile.ZipFile(DST, "w", zipfile.ZIP_DEFLATED) as z:
        for p in sorted(files):
            z.write(p, os.path.join("pymbar", os.path.relpath(p, SRC)))
    return DST


if __name__ == "__main__":
    out = vendor(force="--force" in sys.argv)
    print(out if out else "reference checkout not present: nothing vendored")
