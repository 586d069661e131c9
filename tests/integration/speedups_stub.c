/*
 * speedups_stub.c -- the binding INTEGRATION.md section 1 describes, compiled and
 * run by the tests: a CPython extension in the shape of the reference's
 * safelife/speedups_src/module.c whose advance_board / alive_counts call
 * libsafelife_b200.so through its C ABI (include/safelife_b200.h) instead of
 * running the CPU loops.  It stands for the edit a maintainer of the reference
 * makes inside module.c:20-49 (advance_board_py) and :99-132 (alive_counts_py);
 * it uses the buffer protocol where module.c uses the numpy C API, which is
 * irrelevant to the boundary.
 *
 *   advance_board(board: uint16 buffer [H, W], spawn_prob=0.3, n_step=1, rng=None) -> bytes
 *       rng: writable buffer of 4 uint64 (PCG64 state_hi, state_lo, inc_hi, inc_lo),
 *       advanced in place by the number of doubles the CPU loop would have drawn.
 *   alive_counts(board, goals) -> bytes (int64[8][9])
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include "safelife_b200.h"

static PyObject *raise_status(int rc) {
    PyErr_SetString(rc == SL_ERR_ARG ? PyExc_ValueError : PyExc_RuntimeError, sl_last_error());
    return NULL;
}

static int board_shape(Py_buffer *view, int *H, int *W) {
    if (view->ndim != 2 || view->itemsize != 2 || !PyBuffer_IsContiguous(view, 'C')) {
        PyErr_SetString(PyExc_ValueError, "board must be a C-contiguous 2-D uint16 array");
        return -1;
    }
    *H = (int)view->shape[0];
    *W = (int)view->shape[1];
    return 0;
}

static PyObject *advance_board_py(PyObject *self, PyObject *args, PyObject *kw) {
    static char *names[] = {"board", "spawn_prob", "n_step", "rng", NULL};
    Py_buffer board, rng = {0};
    PyObject *board_obj, *rng_obj = Py_None;
    float spawn_prob = 0.3f;
    int n_step = 1, H, W, rc;
    if (!PyArg_ParseTupleAndKeywords(args, kw, "O|fiO", names, &board_obj, &spawn_prob, &n_step, &rng_obj))
        return NULL;
    if (PyObject_GetBuffer(board_obj, &board, PyBUF_ND | PyBUF_C_CONTIGUOUS) < 0) return NULL;
    if (board_shape(&board, &H, &W) < 0) {
        PyBuffer_Release(&board);
        return NULL;
    }
    if (rng_obj != Py_None &&
        (PyObject_GetBuffer(rng_obj, &rng, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0 || rng.len != 32)) {
        if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "rng must hold 4 uint64 words");
        if (rng.obj) PyBuffer_Release(&rng);
        PyBuffer_Release(&board);
        return NULL;
    }
    PyObject *out = PyBytes_FromStringAndSize(NULL, board.len);
    if (out != NULL) {
        Py_BEGIN_ALLOW_THREADS                       /* as module.c:38 */
        rc = sl_advance_board_host((const uint16_t *)board.buf, (uint16_t *)PyBytes_AS_STRING(out),
                                   1, H, W, spawn_prob, n_step, rng.obj ? (uint64_t *)rng.buf : NULL);
        Py_END_ALLOW_THREADS
        if (rc != SL_OK) {
            Py_DECREF(out);
            out = raise_status(rc);
        }
    }
    if (rng.obj) PyBuffer_Release(&rng);
    PyBuffer_Release(&board);
    return out;
}

static PyObject *alive_counts_py(PyObject *self, PyObject *args) {
    Py_buffer board, goals;
    PyObject *board_obj, *goals_obj;
    int H, W, rc;
    if (!PyArg_ParseTuple(args, "OO", &board_obj, &goals_obj)) return NULL;
    if (PyObject_GetBuffer(board_obj, &board, PyBUF_ND | PyBUF_C_CONTIGUOUS) < 0) return NULL;
    if (PyObject_GetBuffer(goals_obj, &goals, PyBUF_ND | PyBUF_C_CONTIGUOUS) < 0) {
        PyBuffer_Release(&board);
        return NULL;
    }
    PyObject *out = NULL;
    if (board_shape(&board, &H, &W) == 0) {
        if (goals.len != board.len) {
            PyErr_SetString(PyExc_ValueError, "Board and goals must have same size.");   /* module.c:112-114 */
        } else if ((out = PyBytes_FromStringAndSize(NULL, 72 * sizeof(int64_t))) != NULL) {
            rc = sl_alive_counts_host((const uint16_t *)board.buf, (const uint16_t *)goals.buf,
                                      (int64_t *)PyBytes_AS_STRING(out), 1, H * W);
            if (rc != SL_OK) {
                Py_DECREF(out);
                out = raise_status(rc);
            }
        }
    }
    PyBuffer_Release(&board);
    PyBuffer_Release(&goals);
    return out;
}

static PyMethodDef methods[] = {
    {"advance_board", (PyCFunction)advance_board_py, METH_VARARGS | METH_KEYWORDS, "advance_board via libsafelife_b200"},
    {"alive_counts", alive_counts_py, METH_VARARGS, "alive_counts via libsafelife_b200"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "speedups_stub", NULL, -1, methods};

PyMODINIT_FUNC PyInit_speedups_stub(void) {
    PyObject *m = PyModule_Create(&moduledef);
    if (m) PyModule_AddIntConstant(m, "abi_version", sl_abi_version());
    return m;
}
