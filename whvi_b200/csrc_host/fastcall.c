/* _fastcall: the per-call hot path of whvi_b200.fwht_ without ctypes.
 *
 * At the sizes of the reference's published benchmark (batch 512, D = 2^6..2^11, benchmarks/walsh_plot.py:43-54) one FWHT
 * call is bound by host-side launch latency, and ctypes' argument marshalling was 3.5 us of its 9.5 us (the reference's
 * extension goes through pybind11, src/fwht/cuda/fwht_cuda.cpp:16-18).  This CPython extension calls the SAME C-ABI entry
 * points of libwhvi_b200.so (addresses taken from the ctypes handle by whvi_b200/_lib.py) through a METH_FASTCALL
 * function: plain integers in, status code out.  No torch headers, no CUDA headers, no computation here.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

/* calli(address, a0, a1, ...) -> status: a C-ABI entry point whose parameters are all pointers / integers (device pointers,
 * sizes, flags, the stream: every launch call of include/whvi_b200.h except the few with float parameters), each argument a
 * Python int or None (= NULL).  On x86-64 SysV and AArch64 every such parameter travels in a 64-bit register or stack slot, so
 * one prototype per argument count serves all of them. */
typedef int64_t I;
#define A(i) a[i]
static PyObject* fc_calli(PyObject* self, PyObject* const* args, Py_ssize_t nargs)
{
    I a[28];
    if (nargs < 1 || nargs > 29) {
        PyErr_SetString(PyExc_TypeError, "calli(address, up to 28 integer arguments)");
        return NULL;
    }
    const unsigned long long addr = PyLong_AsUnsignedLongLong(args[0]);
    const int n = (int)nargs - 1;
    for (int i = 0; i < n; ++i) {
        PyObject* o = args[i + 1];
        if (o == Py_None) {
            a[i] = 0;
        } else {
            a[i] = (I)PyLong_AsLongLong(o);
            if (a[i] == -1 && PyErr_Occurred()) {   /* pointers above 2^63 do not occur; sizes fit */
                PyErr_Clear();
                a[i] = (I)PyLong_AsUnsignedLongLong(o);
            }
        }
    }
    if (PyErr_Occurred() || !addr) {
        if (!PyErr_Occurred()) PyErr_SetString(PyExc_RuntimeError, "whvi_b200._fastcall: null entry point");
        return NULL;
    }
    void* f = (void*)(uintptr_t)addr;
    int rc;
    switch (n) {
    case 4: rc = ((int (*)(I, I, I, I))f)(A(0), A(1), A(2), A(3)); break;
    case 5: rc = ((int (*)(I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4)); break;
    case 6: rc = ((int (*)(I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5)); break;
    case 7: rc = ((int (*)(I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6)); break;
    case 8: rc = ((int (*)(I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7)); break;
    case 9: rc = ((int (*)(I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8)); break;
    case 10: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9)); break;
    case 11: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10)); break;
    case 12: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11)); break;
    case 13: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12)); break;
    case 14: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13)); break;
    case 15: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14)); break;
    case 16: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15)); break;
    case 17: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16)); break;
    case 18: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17)); break;
    case 19: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18)); break;
    case 20: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19)); break;
    case 21: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19), A(20)); break;
    case 22: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19), A(20), A(21)); break;
    case 23: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19), A(20), A(21), A(22)); break;
    case 24: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19), A(20), A(21), A(22), A(23)); break;
    case 25: rc = ((int (*)(I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I, I))f)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19), A(20), A(21), A(22), A(23), A(24)); break;
    default:
        PyErr_SetString(PyExc_TypeError, "calli: unsupported argument count");
        return NULL;
    }
    return PyLong_FromLong(rc);
}

static PyMethodDef methods[] = {
    {"calli", (PyCFunction)(void (*)(void))fc_calli, METH_FASTCALL, "calli(address, int...) -> status of an all-integer C-ABI entry point"},
    {NULL, NULL, 0, NULL}};
static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_fastcall", "fast host path of the FWHT call", -1, methods};
PyMODINIT_FUNC PyInit__fastcall(void) { return PyModule_Create(&module); }
