/* _hostpack -- CPython helper for the HOST side of the drop-in boundary (no device code, no arithmetic on group
 * elements): turns the Python lists the reference's API hands over (witness x, coefficients of L, exponents --
 * pivot.py:139-145, compressed_pivot.py:89-145: ints and field-element objects) into the packed 32-byte little-endian
 * residues libvmsm.so takes, in one C loop instead of three Python calls per element (65 535 elements x 3 vectors per
 * N = 2^16 proof: ~30 ms of a 78 ms prove call).  Optional: verifiable_mpc_b200/hostpack.py falls back to the same
 * conversion in pure Python when this module is not built. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <string.h>

static PyObject *str_value;

static int lt_le32(const unsigned char *a, const unsigned char *b) { /* a < b, 32-byte little-endian */
    for (int i = 31; i >= 0; i--) {
        if (a[i] != b[i]) return a[i] < b[i];
    }
    return 0;
}

/* pack_residues(seq, cls, order, allow_int=True) -> bytes | None
 * Every element must be an exact int or an instance of exactly `cls` (a prime-field element class whose modulus is
 * `order`, carrying its residue in `.value`); anything else -> None (the caller takes its generic path); with
 * allow_int false, ints are "anything else" too (the transcript text of a plain int differs from a field element's).
 * cls = True: the class is taken from the first non-int element (hostpack.field_class_for's rule, without a separate
 * Python pass over the list).
 * Values outside
 * [0, order) -- negative ints, unreduced products, compressed_pivot.py:66,134 -- are reduced with Python's %. */
static PyObject *pack_residues(PyObject *self, PyObject *args) {
    PyObject *seq, *cls, *order;
    int allow_int = 1;
    if (!PyArg_ParseTuple(args, "OOO|p", &seq, &cls, &order, &allow_int)) return NULL;
    if (!PyLong_Check(order)) {
        PyErr_SetString(PyExc_TypeError, "order must be an int");
        return NULL;
    }
    unsigned char ob[32];
    if (_PyLong_Sign(order) <= 0 || _PyLong_NumBits(order) > 256 ||
        _PyLong_AsByteArray((PyLongObject *)order, ob, 32, 1, 0) < 0) {
        PyErr_Clear();
        PyErr_SetString(PyExc_ValueError, "order must be in (0, 2^256)");
        return NULL;
    }
    PyObject *fast = PySequence_Fast(seq, "expected a sequence");
    if (!fast) return NULL;
    Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject *out = PyBytes_FromStringAndSize(NULL, 32 * n);
    if (!out) {
        Py_DECREF(fast);
        return NULL;
    }
    unsigned char *buf = (unsigned char *)PyBytes_AS_STRING(out);
    for (Py_ssize_t i = 0; i < n; i++, buf += 32) {
        PyObject *item = PySequence_Fast_GET_ITEM(fast, i);
        PyObject *v = NULL;
        int owned = 0;
        if (allow_int && PyLong_CheckExact(item)) {
            v = item;
        } else if (cls == Py_True) {
            /* auto: the first non-int element fixes the class -- a prime-field class with this modulus and a .value */
            PyObject *t = (PyObject *)Py_TYPE(item);
            PyObject *m = PyObject_GetAttrString(t, "modulus");
            int same = m ? PyObject_RichCompareBool(m, order, Py_EQ) : 0;
            Py_XDECREF(m);
            if (same <= 0 || !PyObject_HasAttr(item, str_value)) {
                PyErr_Clear();
                goto unsupported;
            }
            cls = t;
            i--, buf -= 32; /* again, now as an instance of cls */
            continue;
        } else if (cls != Py_None && (PyObject *)Py_TYPE(item) == cls) {
            v = PyObject_GetAttr(item, str_value);
            owned = 1;
            if (!v) goto fail;
            if (!PyLong_Check(v)) {
                Py_DECREF(v);
                goto unsupported;
            }
        } else {
            goto unsupported;
        }
        int ok = 0;
        if (_PyLong_Sign(v) >= 0 && _PyLong_NumBits(v) <= 256) {
            if (_PyLong_AsByteArray((PyLongObject *)v, buf, 32, 1, 0) < 0) {
                if (owned) Py_DECREF(v);
                goto fail;
            }
            ok = lt_le32(buf, ob);
        }
        if (!ok) {
            PyObject *r = PyNumber_Remainder(v, order);
            if (!r || _PyLong_AsByteArray((PyLongObject *)r, buf, 32, 1, 0) < 0) {
                Py_XDECREF(r);
                if (owned) Py_DECREF(v);
                goto fail;
            }
            Py_DECREF(r);
        }
        if (owned) Py_DECREF(v);
    }
    Py_DECREF(fast);
    return out;
unsupported:
    Py_DECREF(fast);
    Py_DECREF(out);
    Py_RETURN_NONE;
fail:
    Py_DECREF(fast);
    Py_DECREF(out);
    return NULL;
}

static PyMethodDef methods[] = {
    {"pack_residues", pack_residues, METH_VARARGS,
     "pack_residues(seq, cls, order) -> n*32 bytes little-endian residues, or None for unsupported element types"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_hostpack", "host-side packing helper", -1, methods};

PyMODINIT_FUNC PyInit__hostpack(void) {
    str_value = PyUnicode_InternFromString("value");
    if (!str_value) return NULL;
    return PyModule_Create(&moddef);
}
