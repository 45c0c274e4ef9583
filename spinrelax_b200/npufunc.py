"""Drop-in for the reference's compiled NumPy ufunc module `npufunc` (Jomega/Jomega.c).

    npufunc.Jomega(x, y)        -> x / (x*x + y*y), NumPy broadcasting, dtype preserved (f -> f, d -> d)
    npufunc.Jomega.outer(a, b)  -> the (len(a), len(b)) table used by spectral_densities._do_Jsum (:1971)
    npufunc.Jomega.nin == 2, .nout == 1, .types == ['ff->f', 'dd->d']

The arithmetic runs in the CUDA kernels sr_jomega_f32 / sr_jomega_f64; broadcasting is done on the host.
The reference's 'ee->e' loop is broken upstream (writes a float into a half slot, Jomega.c:100) and its
'gg->g' long-double loop has no GPU counterpart; both raise TypeError here.
"""
import numpy as np

from . import _lib


class _JomegaUfunc:
    nin, nout, nargs = 2, 1, 3
    types = ['ff->f', 'dd->d']
    __name__ = "Jomega"

    def _run(self, x, y):
        x = np.asarray(x)
        y = np.asarray(y)
        dt = np.result_type(x, y)
        if dt == np.float32:
            fn = "sr_jomega_f32"
        elif dt == np.float64 or dt.kind in "iub":
            dt, fn = np.dtype(np.float64), "sr_jomega_f64"
        else:
            raise TypeError("ufunc 'Jomega' not supported for the input types %s, %s" % (x.dtype, y.dtype))
        xb, yb = np.broadcast_arrays(x.astype(dt, copy=False), y.astype(dt, copy=False))
        shape = xb.shape
        torch = _lib.require_cuda()
        lib = _lib.load()
        xd = torch.from_numpy(np.ascontiguousarray(xb).reshape(-1)).cuda()
        yd = torch.from_numpy(np.ascontiguousarray(yb).reshape(-1)).cuda()
        od = torch.empty_like(xd)
        _lib.check(getattr(lib, fn)(xd.data_ptr(), yd.data_ptr(), od.data_ptr(), xd.numel(), _lib.current_stream_ptr()), fn)
        out = od.cpu().numpy().reshape(shape)
        return out[()] if shape == () else out

    def __call__(self, x, y):
        return self._run(x, y)

    def outer(self, a, b):
        a = np.asarray(a)
        b = np.asarray(b)
        return self._run(a.reshape(a.shape + (1,) * b.ndim), b)


Jomega = _JomegaUfunc()
