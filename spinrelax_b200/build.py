"""Build libspinrelax_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m spinrelax_b200.build [--force] [--verbose]

The .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libspinrelax_b200.so")
STAMP = LIB + ".stamp"
UFUNC_EXT = os.path.join(HERE, "_npufunc_ext.so")       # CPython extension: the real numpy.ufunc `npufunc.Jomega`

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--fmad=true",
]
# per-file extras: the histogram fast path is float32 candidate arithmetic whose results are always verified
# with a margin, so flushing float32 denormals there only removes the MUFU range fix-ups
PER_FILE_FLAGS = {"hist.cu": ["-ftz=true"]}


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))
    ) + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".inc", ".c"))) + \
        [os.path.join(HERE, "..", "include", "spinrelax_b200.h")]
    for p in files:
        with open(p, "rb") as fp:
            h.update(os.path.basename(p).encode())      # not the absolute path: the snapshot on the GPU box lives elsewhere
            h.update(fp.read())
    h.update((" ".join(NVCC_FLAGS) + repr(sorted(PER_FILE_FLAGS.items()))).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    return "nvcc"


def _compile_one(args):
    src, obj, verbose = args
    cmd = [nvcc_path()] + NVCC_FLAGS + PER_FILE_FLAGS.get(os.path.basename(src), []) + \
        (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build_tuning(out, sources=("ct.cu", "api.cu"), verbose=False):
    """Tuning build (-DSR_TUNING): the experimental K1 configurations + sr_ct_lag_sums_variant, linked into `out`
    (a path outside the package; tools/tune_ct.py).  Never loaded by the product."""
    objs = []
    for name in sources:
        obj = out + "." + name[:-3] + ".o"
        cmd = [nvcc_path()] + NVCC_FLAGS + ["-DSR_TUNING"] + (["-Xptxas", "-v"] if verbose else []) + \
            ["-c", "-o", obj, os.path.join(CSRC, name)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s (tuning build)" % name)
        objs.append(obj)
    res = subprocess.run([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking the tuning library")
    return out


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ (in parallel) and link them into one shared library. Returns its path."""
    from concurrent.futures import ThreadPoolExecutor
    import fcntl
    dig = _digest()

    def fresh():
        if os.path.exists(LIB) and os.path.exists(UFUNC_EXT) and os.path.exists(STAMP):
            with open(STAMP) as fp:
                return fp.read().strip() == dig
        return False

    if not force and fresh():
        return LIB
    if not os.path.exists(nvcc_path()) and os.path.sep in nvcc_path():
        raise RuntimeError("libspinrelax_b200.so is missing or stale and nvcc is not available to rebuild it")
    # one builder at a time (torchrun starts every rank at once): the others wait here and find a fresh library
    lock = open(LIB + ".lock", "w")
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and fresh():
            return LIB
        return _build_locked(dig, verbose, ThreadPoolExecutor)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(dig, verbose, ThreadPoolExecutor):
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = [(s, os.path.join(objdir, os.path.basename(s)[:-3] + ".o"), verbose) for s in _sources()]
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
        results = list(pool.map(_compile_one, jobs))
    for src, rc, out in results:
        if verbose or rc != 0:
            sys.stderr.write("== %s\n%s" % (os.path.basename(src), out))
        if rc != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    link = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [j[1] for j in jobs]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libspinrelax_b200.so")
    _build_ufunc_extension()
    with open(STAMP, "w") as fp:
        fp.write(dig)
    return LIB


def _build_ufunc_extension():
    """gcc -shared csrc/npufunc_module.c against CPython + NumPy headers, linked to libspinrelax_b200.so ($ORIGIN rpath)."""
    import sysconfig
    import numpy
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-o", UFUNC_EXT, os.path.join(CSRC, "npufunc_module.c"),
           "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), "-L" + HERE, "-l:libspinrelax_b200.so",
           "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed on npufunc_module.c")


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(LIB)
