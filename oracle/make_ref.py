#!/usr/bin/env python
"""Recipe for ``oracle/_ref``: a verbatim, git-ignored copy of the reference's Python package.

TEST INFRASTRUCTURE.  The reference (bcwarner/physicl) is pure Python: there is nothing to compile,
so "building" it means placing ``/root/reference/physicl/*.py`` where ``bench.py --impl reference``
and ``cpu_baseline`` can import it on the GPU box (``/root/reference`` does not exist there).
``oracle/_ref/`` is listed in ``.gitignore`` (never committed) but not in ``.gpurunignore`` (it travels
with the snapshot, like the built ``.so`` files).  Nothing in ``physicl_b200`` imports it.

    python oracle/make_ref.py            # run in the build container; no-op where /root/reference is absent
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/physicl"
DST = os.path.join(HERE, "_ref", "physicl")


def make(verbose=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    digest = hashlib.sha256()
    for name in sorted(os.listdir(SRC)):
        if name.endswith(".py"):
            shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
            with open(os.path.join(SRC, name), "rb") as f:
                digest.update(f.read())
    with open(os.path.join(HERE, "_ref", "SOURCE"), "w") as f:
        f.write("verbatim copy of %s/*.py (sha256 of the concatenation: %s)\n" % (SRC, digest.hexdigest()))
    if verbose:
        print("oracle/_ref/physicl <- %s (%s)" % (SRC, digest.hexdigest()[:16]))
    return True


if __name__ == "__main__":
    sys.exit(0 if make(verbose=True) else 1)
