/* `npufunc`: a real numpy.ufunc object `Jomega(x, y) = x / (x*x + y*y)` whose inner loops run on the GPU.
 *
 * Replaces the reference's compiled module Jomega/Jomega.c (loops :30-104, registration :135-156): same module and
 * ufunc name, same `ff->f` and `dd->d` loops, so `npufunc.Jomega(a, b)`, `.outer`, `out=`, `where=`, broadcasting and
 * `.types` behave as NumPy defines them for any ufunc.  The reference's `ee->e` loop writes a float into a half slot
 * (:100) and its `gg->g` long-double loop has no GPU counterpart; neither is registered.
 * The loops hand NumPy's strided buffers to sr_jomega_host_f64 / _f32 (include/spinrelax_b200.h).  A failing launch (no
 * CUDA device: there is no CPU fallback) fills the output with NaN and sets a Python RuntimeError.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>

#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/ndarraytypes.h>
#include <numpy/ufuncobject.h>

#include "../../include/spinrelax_b200.h"

static void fail(char* out, npy_intp so, npy_intp n, int is_double) {
  npy_intp i;
  PyGILState_STATE st;
  for (i = 0; i < n; ++i) {
    if (is_double) *(double*)(out + i * so) = NAN;
    else *(float*)(out + i * so) = NAN;
  }
  st = PyGILState_Ensure();
  if (!PyErr_Occurred()) PyErr_Format(PyExc_RuntimeError, "npufunc.Jomega: %s", sr_last_error());
  PyGILState_Release(st);
}

static void double_Jomega(char** args, const npy_intp* dimensions, const npy_intp* steps, void* data) {
  (void)data;
  if (sr_jomega_host_f64(args[0], (long long)steps[0], args[1], (long long)steps[1], args[2], (long long)steps[2],
                         (long long)dimensions[0]) != SR_OK)
    fail(args[2], steps[2], dimensions[0], 1);
}

static void float_Jomega(char** args, const npy_intp* dimensions, const npy_intp* steps, void* data) {
  (void)data;
  if (sr_jomega_host_f32(args[0], (long long)steps[0], args[1], (long long)steps[1], args[2], (long long)steps[2],
                         (long long)dimensions[0]) != SR_OK)
    fail(args[2], steps[2], dimensions[0], 0);
}

static PyUFuncGenericFunction funcs[2] = {&float_Jomega, &double_Jomega};
static const char types[6] = {NPY_FLOAT, NPY_FLOAT, NPY_FLOAT, NPY_DOUBLE, NPY_DOUBLE, NPY_DOUBLE};
static void* const loop_data[2] = {NULL, NULL};

static PyMethodDef methods[] = {{NULL, NULL, 0, NULL}};
static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_npufunc_ext", NULL, -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__npufunc_ext(void) {
  PyObject *m, *jomega, *d;
  import_array();
  import_umath();
  m = PyModule_Create(&moduledef);
  if (!m) return NULL;
  jomega = PyUFunc_FromFuncAndData(funcs, loop_data, types, 2, 2, 1, PyUFunc_None, "Jomega",
                                   "Jomega(x, y) = x / (x*x + y*y), evaluated on the GPU", 0);
  d = PyModule_GetDict(m);
  PyDict_SetItemString(d, "Jomega", jomega);
  Py_DECREF(jomega);
  return m;
}
