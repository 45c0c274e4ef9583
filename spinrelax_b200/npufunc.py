"""Drop-in for the reference's compiled NumPy ufunc module `npufunc` (Jomega/Jomega.c).

    npufunc.Jomega(x, y)        -> x / (x*x + y*y): a real numpy.ufunc (broadcasting, out=, where=, .outer, .types)
    npufunc.Jomega.outer(a, b)  -> the (len(a), len(b)) table used by spectral_densities._do_Jsum (:1971)
    npufunc.Jomega.nin == 2, .nout == 1, .types == ['ff->f', 'dd->d']

`Jomega` is the ufunc object registered by the extension module csrc/npufunc_module.c (built next to
libspinrelax_b200.so by spinrelax_b200.build); its inner loops run the CUDA kernels sr_jomega_f32 / sr_jomega_f64
through sr_jomega_host_f32 / _f64.  The reference's 'ee->e' loop is broken upstream (writes a float into a half slot,
Jomega.c:100) and its 'gg->g' long-double loop has no GPU counterpart; both raise TypeError here, as for any ufunc
without a matching loop.  There is no CPU fallback: without a CUDA device a call raises RuntimeError.
"""
from . import _lib

_lib.load()                                  # builds the library and the extension if needed
from ._npufunc_ext import Jomega  # noqa: E402,F401
