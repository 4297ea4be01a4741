"""In-tree build of libb200tree.so (hand-written CUDA for sm_100a, no fallback arch).

`python -m np_decision_tree_b200._build` or `__graft_entry__.build()`.  nvcc cross-compiles
without a GPU; the .so is git-ignored but travels to the GPU box with the source snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200tree.so")
SOURCES = ["sort.cu", "l2.cu", "weighted.cu", "partition.cu", "relabel.cu", "children.cu", "predict.cu", "l1.cu", "classes.cu", "rows.cu", "fit.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function,-Wno-maybe-uninitialized", "--expt-relaxed-constexpr",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200tree.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "b200dt.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=(), checked=False):
    """checked=True builds libb200tree_checked.so with -DB200DT_BOUNDS_CHECK: device-side index assertions in the
    hot kernels (load it with B200DT_CHECKED=1)."""
    lib = LIB.replace("libb200tree.so", "libb200tree_checked.so") if checked else LIB
    if checked:
        extra = tuple(extra) + ("-DB200DT_BOUNDS_CHECK",)
    elif not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    obj_dir = os.path.join(HERE, "build_checked" if checked else "build")
    os.makedirs(obj_dir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src} ---\n{out}\n")
        elif out.strip() and verbose:
            print(out)
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-lcudart"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True,
                extra=("-Xptxas", "-v") if "--ptxas" in sys.argv else (), checked="--checked" in sys.argv))
