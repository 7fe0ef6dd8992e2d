/* _hostpack -- CPython helper for the HOST side of the drop-in boundary (no device code, no arithmetic on group
 * elements): turns the Python lists the reference's API hands over (witness x, coefficients of L, exponents --
 * pivot.py:139-145, compressed_pivot.py:89-145: ints and field-element objects) into the packed 32-byte little-endian
 * residues libvmsm.so takes, in one C loop instead of three Python calls per element (65 535 elements x 3 vectors per
 * N = 2^16 proof: ~30 ms of a 78 ms prove call).  Optional: verifiable_mpc_b200/hostpack.py falls back to the same
 * conversion in pure Python when this module is not built. */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
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

/* ---- the reference's announcement draws, r = [prng.randrange(order) for _ in range(n)] (compressed_pivot.py:103-105,
 * pivot.py:160), for a seeded random.Random: the same Mersenne-twister stream consumed in the same order -- randrange is
 * getrandbits(bits) repeated until the value is below `order`; getrandbits(bits) takes ceil(bits / 32) outputs, least
 * significant word first, the last one shifted down to the remaining bits -- written straight into packed residues.
 * mt_randbelow_packed(key: 2496 bytes = 624 state words, pos, order, n) -> (n * 32 bytes, new key bytes, new pos).
 * The loop runs without the GIL.  Marshalling of the host-side API, like pack_residues: no group arithmetic. */
#define MT_N 624
#define MT_M 397
static void mt_refill(uint32_t *mt) {
    static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
    int kk;
    uint32_t y;
    for (kk = 0; kk < MT_N - MT_M; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
    }
    for (; kk < MT_N - 1; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
}
static inline uint32_t mt_next(uint32_t *mt, int *pos) {
    if (*pos >= MT_N) {
        mt_refill(mt);
        *pos = 0;
    }
    uint32_t y = mt[(*pos)++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

static PyObject *mt_randbelow_packed(PyObject *self, PyObject *args) {
    Py_buffer key;
    int pos;
    PyObject *order;
    Py_ssize_t n;
    if (!PyArg_ParseTuple(args, "y*iOn", &key, &pos, &order, &n)) return NULL;
    unsigned char ob[32];
    if (key.len != MT_N * 4 || pos < 0 || pos > MT_N || n < 0 || !PyLong_Check(order) || _PyLong_Sign(order) <= 0 ||
        _PyLong_NumBits(order) > 256 || _PyLong_AsByteArray((PyLongObject *)order, ob, 32, 1, 0) < 0) {
        PyErr_Clear();
        PyBuffer_Release(&key);
        PyErr_SetString(PyExc_ValueError, "mt_randbelow_packed(key[2496], pos, order < 2^256, n)");
        return NULL;
    }
    const int bits = (int)_PyLong_NumBits(order);
    const int words = (bits - 1) / 32 + 1, top_bits = bits - 32 * (words - 1);
    uint32_t mt[MT_N];
    memcpy(mt, key.buf, sizeof mt);
    PyBuffer_Release(&key);
    PyObject *out = PyBytes_FromStringAndSize(NULL, 32 * n);
    if (!out) return NULL;
    unsigned char *buf = (unsigned char *)PyBytes_AS_STRING(out);
    Py_BEGIN_ALLOW_THREADS
    uint32_t ow[8];
    for (int w = 0; w < 8; w++)
        ow[w] = (uint32_t)ob[4 * w] | (uint32_t)ob[4 * w + 1] << 8 | (uint32_t)ob[4 * w + 2] << 16 | (uint32_t)ob[4 * w + 3] << 24;
    for (Py_ssize_t i = 0; i < n; i++, buf += 32) {
        uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (;;) {
            for (int w = 0; w < words; w++) v[w] = mt_next(mt, &pos);
            if (top_bits < 32) v[words - 1] >>= (32 - top_bits);
            int below = 0;
            for (int w = 7; w >= 0; w--)
                if (v[w] != ow[w]) {
                    below = v[w] < ow[w];
                    break;
                }
            if (below) break;
        }
        for (int w = 0; w < 8; w++) {
            buf[4 * w] = (unsigned char)v[w], buf[4 * w + 1] = (unsigned char)(v[w] >> 8);
            buf[4 * w + 2] = (unsigned char)(v[w] >> 16), buf[4 * w + 3] = (unsigned char)(v[w] >> 24);
        }
    }
    Py_END_ALLOW_THREADS
    PyObject *nk = PyBytes_FromStringAndSize((const char *)mt, sizeof mt);
    if (!nk) {
        Py_DECREF(out);
        return NULL;
    }
    return Py_BuildValue("(NNi)", out, nk, pos);
}

static PyMethodDef methods[] = {
    {"mt_randbelow_packed", mt_randbelow_packed, METH_VARARGS,
     "mt_randbelow_packed(key, pos, order, n) -> (n*32 bytes of randrange(order) draws, new key, new pos)"},
    {"pack_residues", pack_residues, METH_VARARGS,
     "pack_residues(seq, cls, order) -> n*32 bytes little-endian residues, or None for unsupported element types"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_hostpack", "host-side packing helper", -1, methods};

PyMODINIT_FUNC PyInit__hostpack(void) {
    str_value = PyUnicode_InternFromString("value");
    if (!str_value) return NULL;
    return PyModule_Create(&moddef);
}
