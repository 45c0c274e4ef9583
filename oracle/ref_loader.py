"""TEST INFRASTRUCTURE ONLY -- loads the real SpinRelax sources from /root/reference (build container only).

Used by tests/golden/make_golden.py to generate the committed golden vectors and by the `not gpu`
tests that re-validate the oracle against the reference when the tree is present.  Nothing here is
imported by the product package; /root/reference does not exist on the GPU box.

Recipe: put stand-ins for the absent third-party modules `transforms3d` and
`mdtraj` on sys.path (oracle/_stubs), add the gcc-built `npufunc` from oracle/_ref, and exec the
pre-`__main__` part of the hyphenated, self-exiting stage scripts into module objects.
"""
import importlib
import os
import sys
import types

REFERENCE = os.environ.get("SPINRELAX_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = os.path.join(HERE, "_stubs")
REF_BUILD = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REFERENCE, "calculate-Ct-from-traj.py"))


def _paths():
    for p in (REFERENCE, REF_BUILD, STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)


def module(name):
    """Import a plain reference module (transforms3d_supplement, general_maths, fitting_Ct_functions, ...)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE)
    _paths()
    return importlib.import_module(name)


def script(filename):
    """Load the function definitions of a hyphenated stage script (everything before `if __name__ ==`)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE)
    _paths()
    path = os.path.join(REFERENCE, filename)
    with open(path) as fp:
        src = fp.read()
    cut = src.find("if __name__ == '__main__':")
    if cut < 0:
        cut = src.find('if __name__ == "__main__":')
    if cut >= 0:
        src = src[:cut]
    mod = types.ModuleType("ref_" + filename.replace("-", "_").replace(".py", ""))
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod
