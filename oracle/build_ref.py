"""Build the reference's one native module into oracle/_ref/ (build container only).

The reference's hot path is Python + one Cython file
(/root/reference/decision_tree/np_stream_median.pyx).  Its own setup.py does
not work with g++ (``-stdlib=libc++``, ref: setup.py:12), so we run cython and
g++ directly on the file where it lies.  Outputs go ONLY to oracle/_ref/
(git-ignored).  Nothing here is copied into the repository history, and
nothing in the product, the GPU tests, smoke() or bench.py depends on it:
it exists so tests/golden/make_golden.py can run the unmodified reference.
"""
import os
import subprocess
import sys
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("B200DT_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build():
    pyx = os.path.join(REF, "decision_tree", "np_stream_median.pyx")
    if not os.path.exists(pyx):
        raise FileNotFoundError(pyx)
    os.makedirs(OUT, exist_ok=True)
    cpp = os.path.join(OUT, "np_stream_median.cpp")
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(OUT, "np_stream_median" + ext)
    if os.path.exists(so):
        return so
    subprocess.check_call(
        [sys.executable, "-m", "cython", "--cplus", "-3", "-o", cpp, pyx])
    subprocess.check_call(
        ["g++", "-O2", "-std=c++11", "-shared", "-fPIC", "-w",
         "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
         cpp, "-o", so])
    return so


if __name__ == "__main__":
    print(build())
