/*
 * Host-side record assembly for the drop-in API, as a CPython extension (eagle_b200._assemble).
 *
 * The kernels leave per-frame arrays (keypoints, inlier masks, projected foot points, boundaries); the
 * reference returns one Python dict per frame (eagle/models/coordinate_model.py:359-362, 369-392, 405-415;
 * docs/data.md:20-42).  Building ~150 Python objects per frame from interpreted code costs 50-150 us per
 * frame, which caps the public API at a few thousand frames/s whatever the GPUs do; the same construction
 * through the C API costs a few microseconds.  The statements mirrored here are exactly those of
 * eagle_b200/coordinate_model.py::assemble_frames_py (kept as the readable statement and as the checker in
 * the tests): same keys, same insertion order, same Python value types.
 *
 * Arrays arrive through the buffer protocol (C-contiguous numpy arrays); no numpy headers are needed.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject *s_bbox, *s_conf, *s_tc, *s_ibc, *s_bc, *s_coords, *s_time, *s_kps, *s_bounds;

static int get_buf(PyObject* o, Py_buffer* b, Py_ssize_t itemsize, const char* name) {
    if (PyObject_GetBuffer(o, b, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) return -1;
    if (b->itemsize != itemsize) {
        PyErr_Format(PyExc_TypeError, "%s: expected items of %zd bytes, got %zd", name, itemsize, b->itemsize);
        PyBuffer_Release(b);
        return -1;
    }
    return 0;
}

/* The records are plain acyclic data (dicts / lists / tuples of numbers and strings) created here and handed over
 * whole, millions of containers for a full match.  Left tracked, the first allocation after the assembly sends the cyclic
 * collector over every one of them (seconds for 135 k frames, measured); CPython itself untracks dicts and tuples of
 * atoms for this reason, and the same is done here for every container this file creates, once it is complete.  A dict
 * re-tracks itself when a collectable value is stored into it later, so user code that extends the records is unaffected. */
#define UNTRACK(o) PyObject_GC_UnTrack((PyObject*)(o))

/* np.array(bbox, dtype=np.uint16).tolist() for the common case of a list of four in-range Python ints;
 * anything else goes through the Python fallback (which makes the numpy round trip). */
static PyObject* bbox_list(PyObject* bbox, PyObject* fallback) {
    if (PyList_CheckExact(bbox) && PyList_GET_SIZE(bbox) == 4) {
        int ok = 1;
        for (int k = 0; k < 4 && ok; ++k) {
            PyObject* v = PyList_GET_ITEM(bbox, k);
            if (!PyLong_CheckExact(v)) { ok = 0; break; }
            int overflow = 0;
            long x = PyLong_AsLongAndOverflow(v, &overflow);
            if (overflow || x < 0 || x > 65535) ok = 0;
        }
        if (ok) return PyList_GetSlice(bbox, 0, 4);
    }
    return PyObject_CallOneArg(fallback, bbox);
}

static PyObject* as_int_key(PyObject* k) {
    if (PyLong_CheckExact(k)) { Py_INCREF(k); return k; }
    return PyNumber_Long(k);
}

/* max_objects(objects_per_frame) -> largest number of detections in one frame */
static PyObject* py_max_objects(PyObject* self, PyObject* arg) {
    if (!PyList_Check(arg)) { PyErr_SetString(PyExc_TypeError, "objects_per_frame must be a list"); return NULL; }
    Py_ssize_t best = 0;
    for (Py_ssize_t i = 0; i < PyList_GET_SIZE(arg); ++i) {
        PyObject* o = PyList_GET_ITEM(arg, i);
        if (!PyDict_Check(o)) { PyErr_SetString(PyExc_TypeError, "every frame's detections must be a dict"); return NULL; }
        Py_ssize_t pos = 0, n = 0;
        PyObject *k, *v;
        while (PyDict_Next(o, &pos, &k, &v)) {
            if (!PyDict_Check(v)) { PyErr_SetString(PyExc_TypeError, "every class entry must be a dict"); return NULL; }
            n += PyDict_GET_SIZE(v);
        }
        if (n > best) best = n;
    }
    return PyLong_FromSsize_t(best);
}

/* pack_foot_points(objects_per_frame, foot (F,P,2) float32, count (F,) int32, P):
 * Bottom_center of every detection in the reference's iteration order (class dict order, then id order) */
static PyObject* py_pack_foot_points(PyObject* self, PyObject* args) {
    PyObject *objs, *foot_o, *count_o;
    Py_ssize_t P;
    if (!PyArg_ParseTuple(args, "O!OOn", &PyList_Type, &objs, &foot_o, &count_o, &P)) return NULL;
    Py_buffer foot, count;
    if (PyObject_GetBuffer(foot_o, &foot, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE) < 0) return NULL;
    if (PyObject_GetBuffer(count_o, &count, PyBUF_C_CONTIGUOUS | PyBUF_WRITABLE) < 0) { PyBuffer_Release(&foot); return NULL; }
    const Py_ssize_t F = PyList_GET_SIZE(objs);
    PyObject* ret = NULL;
    if (foot.len != F * P * 2 * 4 || count.len != F * 4) {
        PyErr_SetString(PyExc_ValueError, "pack_foot_points: buffer sizes do not match (F, P)");
        goto done;
    }
    float* fp = (float*)foot.buf;
    int32_t* cp = (int32_t*)count.buf;
    for (Py_ssize_t i = 0; i < F; ++i) {
        PyObject* o = PyList_GET_ITEM(objs, i);
        if (!PyDict_Check(o)) { PyErr_SetString(PyExc_TypeError, "every frame's detections must be a dict"); goto done; }
        Py_ssize_t pos = 0, n = 0;
        PyObject *cls, *cd;
        while (PyDict_Next(o, &pos, &cls, &cd)) {
            if (!PyDict_Check(cd)) { PyErr_SetString(PyExc_TypeError, "every class entry must be a dict"); goto done; }
            Py_ssize_t pos2 = 0;
            PyObject *id, *d;
            while (PyDict_Next(cd, &pos2, &id, &d)) {
                if (n >= P) { PyErr_Format(PyExc_ValueError, "frame %zd: more than %zd objects", i, P); goto done; }
                PyObject* bc = PyObject_GetItem(d, s_bc);
                if (!bc) goto done;
                PyObject* seq = PySequence_Fast(bc, "Bottom_center must be a sequence of two numbers");
                Py_DECREF(bc);
                if (!seq) goto done;
                if (PySequence_Fast_GET_SIZE(seq) != 2) {
                    Py_DECREF(seq);
                    PyErr_SetString(PyExc_ValueError, "Bottom_center must have two entries");
                    goto done;
                }
                const double x = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(seq, 0));
                const double y = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(seq, 1));
                Py_DECREF(seq);
                if ((x == -1.0 || y == -1.0) && PyErr_Occurred()) goto done;
                fp[(i * P + n) * 2] = (float)x;
                fp[(i * P + n) * 2 + 1] = (float)y;
                ++n;
            }
        }
        cp[i] = (int32_t)n;
    }
    ret = Py_None;
    Py_INCREF(ret);
done:
    PyBuffer_Release(&foot);
    PyBuffer_Release(&count);
    return ret;
}

/*
 * assemble(out, objects_per_frame, fps, first_index, names, kp_xy, kp_order, kp_count, inlier_mask, fitted, h_index,
 *          coords_i, in_bounds, bounds, kp_src, off_plane_mask, pitch_width, bbox_fallback, np_int64,
 *          tag_py_int, tag_float)
 *   kp_xy (F,57,2) int32; kp_order (F,S) uint8; kp_count (F,2) int32; inlier_mask (F,) int64 / None; fitted (F,) uint8 / None;
 *   h_index (F,) int32; coords_i (F,P,2) int64; in_bounds (F,P) uint8; bounds (F,4) float64; kp_src (F,S) uint8 / None.
 * Fills out[first_index + k] for every frame k.
 */
static PyObject* py_assemble(PyObject* self, PyObject* args) {
    PyObject *out, *objs, *names, *xy_o, *order_o, *count_o, *inl_o, *fit_o, *hi_o, *ci_o, *ib_o, *bd_o, *src_o, *fallback, *i64;
    Py_ssize_t fps, first;
    unsigned long long off_mask;
    long pitch_w, tag_py_int, tag_float;
    if (!PyArg_ParseTuple(args, "O!O!nnO!OOOOOOOOOOKlOOll", &PyDict_Type, &out, &PyList_Type, &objs, &fps, &first, &PyTuple_Type, &names,
                          &xy_o, &order_o, &count_o, &inl_o, &fit_o, &hi_o, &ci_o, &ib_o, &bd_o, &src_o, &off_mask, &pitch_w, &fallback,
                          &i64, &tag_py_int, &tag_float))
        return NULL;
    if (fps <= 0) { PyErr_SetString(PyExc_ValueError, "fps must be positive"); return NULL; }
    const Py_ssize_t F = PyList_GET_SIZE(objs);
    const int with_src = src_o != Py_None;
    Py_buffer xy, order, count, fit, hi, ci, ib, bd, src;  /* src doubles as the inlier-mask buffer in the dense mode */
    int nb = 0;
    PyObject* ret = NULL;
    PyObject *zero = NULL, *wobj = NULL;
    if (get_buf(xy_o, &xy, 4, "kp_xy") < 0) goto fail;
    nb = 1;
    if (get_buf(order_o, &order, 1, "kp_order") < 0) goto fail;
    nb = 2;
    if (get_buf(count_o, &count, 4, "kp_count") < 0) goto fail;
    nb = 3;
    if (get_buf(hi_o, &hi, 4, "h_index") < 0) goto fail;
    nb = 4;
    if (get_buf(ci_o, &ci, 8, "coords_i") < 0) goto fail;
    nb = 5;
    if (get_buf(ib_o, &ib, 1, "in_bounds") < 0) goto fail;
    nb = 6;
    if (get_buf(bd_o, &bd, 8, "bounds") < 0) goto fail;
    nb = 7;
    if (with_src) { if (get_buf(src_o, &src, 1, "kp_src") < 0) goto fail; }
    else { if (get_buf(inl_o, &src, 8, "inlier_mask") < 0) goto fail; }
    nb = 8;
    if (!with_src) { if (get_buf(fit_o, &fit, 1, "fitted") < 0) goto fail; nb = 9; }
    {
        const Py_ssize_t n_names = PyTuple_GET_SIZE(names);
        if (F == 0) { ret = Py_None; Py_INCREF(ret); goto fail; }
        if (xy.len % (F * 8) || order.len % F || ib.len % F) { PyErr_SetString(PyExc_ValueError, "array sizes do not match F"); goto fail; }
        const Py_ssize_t C = xy.len / (F * 8), S = order.len / F, P = ib.len / F;
        if (count.len != F * 8 || hi.len != F * 4 || ci.len != F * P * 16 || bd.len != F * 32 || C > n_names ||
            (with_src ? src.len != F * S : (src.len != F * 8 || fit.len != F))) {
            PyErr_SetString(PyExc_ValueError, "array shapes are inconsistent");
            goto fail;
        }
        const int32_t* xyp = (const int32_t*)xy.buf;
        const uint8_t* orderp = (const uint8_t*)order.buf;
        const int32_t* countp = (const int32_t*)count.buf;
        const int32_t* hip = (const int32_t*)hi.buf;
        const int64_t* cip = (const int64_t*)ci.buf;
        const uint8_t* ibp = (const uint8_t*)ib.buf;
        const double* bdp = (const double*)bd.buf;
        const uint8_t* srcp = with_src ? (const uint8_t*)src.buf : NULL;
        const int64_t* inlp = with_src ? NULL : (const int64_t*)src.buf;
        const uint8_t* fitp = with_src ? NULL : (const uint8_t*)fit.buf;
        zero = PyLong_FromLong(0);
        wobj = PyLong_FromLong(pitch_w);
        if (!zero || !wobj) goto fail;
        for (Py_ssize_t k = 0; k < F; ++k) {
            const Py_ssize_t i = first + k;
            /* ---- "Keypoints" ---- */
            PyObject* kps = PyDict_New();
            if (!kps) goto fail;
            Py_ssize_t n = countp[2 * k];
            if (n > S) n = S;
            const int fitted = with_src ? 0 : fitp[k];
            const uint64_t inl_bits = with_src ? 0 : (uint64_t)inlp[k];
            for (Py_ssize_t j = 0; j < n; ++j) {
                const int c = orderp[k * S + j];
                if (c >= C) continue;
                const long x = xyp[(k * C + c) * 2], y = xyp[(k * C + c) * 2 + 1];
                PyObject* val;
                int as_float, as_np = 0;
                if (with_src) {
                    const int t = srcp[k * S + c];
                    as_float = t == tag_float;
                    as_np = !as_float && t != tag_py_int;
                } else {
                    if (fitted && (((off_mask >> c) & 1ull) || !((inl_bits >> c) & 1ull))) continue;
                    as_float = fitted;
                }
                if (as_float) {
                    PyObject *a = PyFloat_FromDouble((double)x), *b = PyFloat_FromDouble((double)y);
                    val = (a && b) ? PyList_New(2) : NULL;
                    if (!val) { Py_XDECREF(a); Py_XDECREF(b); Py_DECREF(kps); goto fail; }
                    PyList_SET_ITEM(val, 0, a);
                    PyList_SET_ITEM(val, 1, b);
                    UNTRACK(val);
                } else {
                    PyObject *a = PyLong_FromLong(x), *b = PyLong_FromLong(y);
                    if (as_np && a && b) {  /* the reference holds numpy integers here (they reach json.dump's default=) */
                        PyObject *a2 = PyObject_CallOneArg(i64, a), *b2 = PyObject_CallOneArg(i64, b);
                        Py_DECREF(a); Py_DECREF(b);
                        a = a2; b = b2;
                    }
                    val = (a && b) ? PyTuple_New(2) : NULL;
                    if (!val) { Py_XDECREF(a); Py_XDECREF(b); Py_DECREF(kps); goto fail; }
                    PyTuple_SET_ITEM(val, 0, a);
                    PyTuple_SET_ITEM(val, 1, b);
                    UNTRACK(val);
                }
                const int rc = PyDict_SetItem(kps, PyTuple_GET_ITEM(names, c), val);
                Py_DECREF(val);
                if (rc < 0) { Py_DECREF(kps); goto fail; }
            }
            UNTRACK(kps);
            /* ---- "Coordinates" ---- */
            PyObject* objects = PyList_GET_ITEM(objs, k);
            PyObject* indiv = PyDict_New();
            if (!indiv || !PyDict_Check(objects)) {
                if (indiv) PyErr_SetString(PyExc_TypeError, "every frame's detections must be a dict");
                Py_XDECREF(indiv); Py_DECREF(kps);
                goto fail;
            }
            const int have_h = hip[k] >= 0;
            Py_ssize_t pos = 0, p = 0;
            PyObject *cls, *cd;
            int bad = 0;
            while (!bad && PyDict_Next(objects, &pos, &cls, &cd)) {
                if (!PyDict_Check(cd)) { PyErr_SetString(PyExc_TypeError, "every class entry must be a dict"); bad = 1; break; }
                if (PyDict_GET_SIZE(cd) == 0) continue;
                PyObject* d = PyDict_GetItemWithError(indiv, cls);  /* borrowed */
                if (!d) {
                    if (PyErr_Occurred()) { bad = 1; break; }
                    d = PyDict_New();
                    if (!d || PyDict_SetItem(indiv, cls, d) < 0) { Py_XDECREF(d); bad = 1; break; }
                    Py_DECREF(d);  /* indiv keeps it alive */
                }
                Py_ssize_t pos2 = 0;
                PyObject *id, *obj;
                while (PyDict_Next(cd, &pos2, &id, &obj)) {
                    if (p >= P) { PyErr_Format(PyExc_ValueError, "frame %zd: more than %zd objects", i, P); bad = 1; break; }
                    PyObject* rec = PyDict_New();
                    PyObject* key = as_int_key(id);
                    PyObject* bb_in = PyObject_GetItem(obj, s_bbox);
                    PyObject* bb = bb_in ? bbox_list(bb_in, fallback) : NULL;
                    PyObject* conf = PyObject_GetItem(obj, s_conf);
                    Py_XDECREF(bb_in);
                    if (!rec || !key || !bb || !conf) { Py_XDECREF(rec); Py_XDECREF(key); Py_XDECREF(bb); Py_XDECREF(conf); bad = 1; break; }
                    if (PyList_CheckExact(bb) && Py_REFCNT(bb) == 1) UNTRACK(bb);   /* a fresh list of four ints */
                    int rc = PyDict_SetItem(rec, s_bbox, bb) | PyDict_SetItem(rec, s_conf, conf);
                    Py_DECREF(bb); Py_DECREF(conf);
                    if (have_h && ibp[k * P + p]) {
                        PyObject *a = PyLong_FromLongLong(cip[(k * P + p) * 2]), *b = PyLong_FromLongLong(cip[(k * P + p) * 2 + 1]);
                        PyObject* tc = (a && b) ? PyList_New(2) : NULL;
                        if (!tc) { Py_XDECREF(a); Py_XDECREF(b); rc = -1; }
                        else {
                            PyList_SET_ITEM(tc, 0, a);
                            PyList_SET_ITEM(tc, 1, b);
                            UNTRACK(tc);
                            rc |= PyDict_SetItem(rec, s_tc, tc);
                            Py_DECREF(tc);
                        }
                    } else {
                        PyObject* foot = PyObject_GetItem(obj, s_bc);
                        if (!foot) rc = -1;
                        else {
                            rc |= PyDict_SetItem(rec, s_tc, Py_None) | PyDict_SetItem(rec, s_ibc, foot);
                            Py_DECREF(foot);
                        }
                    }
                    UNTRACK(rec);
                    if (rc == 0) rc = PyDict_SetItem(d, key, rec);
                    Py_DECREF(rec); Py_DECREF(key);
                    if (rc) { bad = 1; break; }
                    ++p;
                }
            }
            if (bad) { Py_DECREF(indiv); Py_DECREF(kps); goto fail; }
            {   /* the class dicts are complete now */
                Py_ssize_t pos3 = 0;
                PyObject *ck, *cv;
                while (PyDict_Next(indiv, &pos3, &ck, &cv)) UNTRACK(cv);
                UNTRACK(indiv);
            }
            /* ---- "Boundaries" ---- */
            PyObject* bl = PyList_New(4);
            if (!bl) { Py_DECREF(indiv); Py_DECREF(kps); goto fail; }
            const double* b = bdp + 4 * k;
            if (have_h && b[0] == b[0]) {
                for (int q = 0; q < 4; ++q) {
                    PyObject* t = PyTuple_New(2);
                    PyObject* v = PyFloat_FromDouble(b[q]);
                    if (!t || !v) { Py_XDECREF(t); Py_XDECREF(v); Py_DECREF(bl); Py_DECREF(indiv); Py_DECREF(kps); goto fail; }
                    PyObject* e = (q == 1 || q == 2) ? wobj : zero;
                    Py_INCREF(e);
                    PyTuple_SET_ITEM(t, 0, v);
                    PyTuple_SET_ITEM(t, 1, e);
                    UNTRACK(t);
                    PyList_SET_ITEM(bl, q, t);
                }
            } else {
                for (int q = 0; q < 4; ++q) { Py_INCREF(Py_None); PyList_SET_ITEM(bl, q, Py_None); }
            }
            UNTRACK(bl);
            /* ---- the frame's record ---- */
            const long long sec = (long long)(i / fps);
            PyObject* tm = PyUnicode_FromFormat("%02lld:%02lld", sec / 60, sec % 60);
            PyObject* rec = PyDict_New();
            PyObject* idx = PyLong_FromSsize_t(i);
            int rc = (tm && rec && idx) ? 0 : -1;
            if (rc == 0)
                rc = PyDict_SetItem(rec, s_coords, indiv) | PyDict_SetItem(rec, s_time, tm) | PyDict_SetItem(rec, s_kps, kps) |
                     PyDict_SetItem(rec, s_bounds, bl);
            if (rc == 0) UNTRACK(rec);
            if (rc == 0) rc = PyDict_SetItem(out, idx, rec);
            Py_XDECREF(tm); Py_XDECREF(rec); Py_XDECREF(idx);
            Py_DECREF(bl); Py_DECREF(indiv); Py_DECREF(kps);
            if (rc) goto fail;
        }
        ret = Py_None;
        Py_INCREF(ret);
    }
fail:
    Py_XDECREF(zero);
    Py_XDECREF(wobj);
    if (nb >= 9) PyBuffer_Release(&fit);
    if (nb >= 8) PyBuffer_Release(&src);
    if (nb >= 7) PyBuffer_Release(&bd);
    if (nb >= 6) PyBuffer_Release(&ib);
    if (nb >= 5) PyBuffer_Release(&ci);
    if (nb >= 4) PyBuffer_Release(&hi);
    if (nb >= 3) PyBuffer_Release(&count);
    if (nb >= 2) PyBuffer_Release(&order);
    if (nb >= 1) PyBuffer_Release(&xy);
    return ret;
}

static PyMethodDef methods[] = {
    {"assemble", py_assemble, METH_VARARGS, "fill out[first_index + k] with the reference-format record of every frame"},
    {"pack_foot_points", py_pack_foot_points, METH_VARARGS, "detections -> (F,P,2) float32 foot points and (F,) int32 counts"},
    {"max_objects", py_max_objects, METH_O, "largest number of detections in one frame"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_assemble", "record assembly for eagle_b200.CoordinateModel", -1, methods};

PyMODINIT_FUNC PyInit__assemble(void) {
    s_bbox = PyUnicode_InternFromString("BBox");
    s_conf = PyUnicode_InternFromString("Confidence");
    s_tc = PyUnicode_InternFromString("Transformed_Coordinates");
    s_ibc = PyUnicode_InternFromString("Image_Bottom_center");
    s_bc = PyUnicode_InternFromString("Bottom_center");
    s_coords = PyUnicode_InternFromString("Coordinates");
    s_time = PyUnicode_InternFromString("Time");
    s_kps = PyUnicode_InternFromString("Keypoints");
    s_bounds = PyUnicode_InternFromString("Boundaries");
    return PyModule_Create(&moddef);
}
