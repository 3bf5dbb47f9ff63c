/* Host glue for ParsePredictor.parse_batch: nested-tuple parse trees from the CKY kernel's backpointer table.
 *
 * The reference returns trees as nested tuples of word positions (cliora/analysis/cky.py:101-109,
 * follow_backpointers); callers str() them.  Building 256 trees of 59 nodes in Python costs more than the
 * CKY kernel and the device->host copy together, so this does it with the CPython C API, straight from the int32
 * buffer (no .tolist()).  Pure host code: no CUDA here.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <stdlib.h>

static PyObject* subtree(const int32_t* row, const int64_t* off, int level, int pos, PyObject** leaves) {
  if (level == 0) {
    Py_INCREF(leaves[pos]);
    return leaves[pos];
  }
  const int k = row[off[level] + pos];
  if (k < 0 || k >= level) {
    PyErr_Format(PyExc_ValueError, "backpointer %d out of range at level %d, position %d", k, level, pos);
    return NULL;
  }
  PyObject* l = subtree(row, off, k, pos, leaves);
  if (l == NULL) return NULL;
  PyObject* r = subtree(row, off, level - 1 - k, pos + k + 1, leaves);
  if (r == NULL) {
    Py_DECREF(l);
    return NULL;
  }
  PyObject* t = PyTuple_New(2);
  if (t == NULL) {
    Py_DECREF(l);
    Py_DECREF(r);
    return NULL;
  }
  PyTuple_SET_ITEM(t, 0, l);
  PyTuple_SET_ITEM(t, 1, r);
  return t;
}

/* build(backpointers: buffer of int32 [B, n(n+1)/2], B: int, n: int) -> list of B nested tuples */
static PyObject* build(PyObject* self, PyObject* args) {
  (void)self;
  Py_buffer buf;
  int B, n;
  if (!PyArg_ParseTuple(args, "y*ii", &buf, &B, &n)) return NULL;
  PyObject* result = NULL;
  PyObject** leaves = NULL;
  int64_t* off = NULL;
  const int64_t C = (int64_t)n * (n + 1) / 2;
  if (B < 0 || n < 1 || (int64_t)buf.len < (int64_t)B * C * 4) {
    PyErr_SetString(PyExc_ValueError, "backpointer buffer smaller than B * n(n+1)/2 int32");
    goto done;
  }
  leaves = (PyObject**)calloc((size_t)n, sizeof(PyObject*));
  off = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  if (leaves == NULL || off == NULL) {
    PyErr_NoMemory();
    goto done;
  }
  for (int l = 0; l < n; ++l) off[l] = (int64_t)l * n - (int64_t)l * (l - 1) / 2;
  for (int i = 0; i < n; ++i) {
    leaves[i] = PyLong_FromLong(i);
    if (leaves[i] == NULL) goto done;
  }
  result = PyList_New(B);
  if (result == NULL) goto done;
  for (int b = 0; b < B; ++b) {
    PyObject* t = subtree((const int32_t*)buf.buf + (int64_t)b * C, off, n - 1, 0, leaves);
    if (t == NULL) {
      Py_CLEAR(result);
      goto done;
    }
    PyList_SET_ITEM(result, b, t);
  }
done:
  if (leaves != NULL) {
    for (int i = 0; i < n; ++i) Py_XDECREF(leaves[i]);
    free(leaves);
  }
  free(off);
  PyBuffer_Release(&buf);
  return result;
}

static PyMethodDef methods[] = {{"build", build, METH_VARARGS, "nested-tuple trees from a backpointer table"},
                                {NULL, NULL, 0, NULL}};
static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_pytrees", NULL, -1, methods, NULL, NULL, NULL, NULL};
PyMODINIT_FUNC PyInit__pytrees(void) { return PyModule_Create(&module); }
