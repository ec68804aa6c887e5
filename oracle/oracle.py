"""ctypes wrapper of the CPU oracle (oracle/gs_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py — never by the product package.
PARITY UNPINNED: see the header of gs_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgs_oracle.so")

GAUSSIAN = np.dtype([("rot", "<f4", 4), ("pos", "<f4", 3), ("color", "u1", 4), ("sh", "<f4", 45), ("scale", "<f4", 3)])
PLY = np.dtype([("pos", "<f4", 3), ("normal", "<f4", 3), ("f_dc", "<f4", 3), ("f_rest", "<f4", 45),
                ("opacity", "<f4"), ("scale", "<f4", 3), ("rot", "<f4", 4)])
EDIT = np.dtype([("flag", "<u4"), ("color", "<f4", 3), ("contrast", "<f4"), ("exposure", "<f4"),
                 ("gamma", "<f4"), ("alpha", "<f4")])
SPLAT_F32 = np.dtype([("mx", "<f4"), ("my", "<f4"), ("radius", "<f4"), ("opacity", "<f4"), ("r", "<f4"), ("g", "<f4"),
                      ("b", "<f4"), ("ca", "<f4"), ("cb", "<f4"), ("cc", "<f4"), ("flags", "<u4")])
SPLAT = np.dtype([("mx", "<f4"), ("my", "<f4"), ("radius", "<u2"), ("opacity_h", "<f2"), ("r_h", "<f2"),
                  ("g_h", "<f2"), ("ca", "<f4"), ("cb", "<f4"), ("cc", "<f4"), ("b_h", "<f2"), ("flags", "<u2")])
MASK_SHAPE = np.dtype([("kind", "<u4"), ("pos", "<f4", 3), ("quat", "<f4", 4), ("scale", "<f4", 3)])
MASK_OP = np.dtype([("kind", "<u4"), ("arg", "<u4")])
assert GAUSSIAN.itemsize == 224 and PLY.itemsize == 248 and EDIT.itemsize == 32 and SPLAT.itemsize == 32
assert MASK_SHAPE.itemsize == 44 and SPLAT_F32.itemsize == 44


class EditPod(C.Structure):
    _fields_ = [("flag", C.c_uint32), ("color", C.c_float * 3), ("contrast", C.c_float), ("exposure", C.c_float),
                ("gamma", C.c_float), ("alpha", C.c_float)]


class QueryPod(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("op", C.c_uint32), ("p0", C.c_float * 2), ("p1", C.c_float * 2),
                ("radius", C.c_float), ("_pad", C.c_uint32)]


class Frame(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("size", C.c_float * 2),
                ("gaussian_size", C.c_float), ("display_mode", C.c_uint32), ("sh_deg", C.c_uint32),
                ("no_sh0", C.c_uint32), ("selection_edit", EditPod), ("highlight", C.c_float * 4),
                ("background", C.c_float * 4), ("query", QueryPod), ("query_tex", C.c_void_p),
                ("query_tex_w", C.c_uint32), ("query_tex_h", C.c_uint32)]


class Model(C.Structure):
    _fields_ = [("sh", C.c_uint32), ("cov3d", C.c_uint32), ("packed", C.c_void_p), ("n", C.c_uint64),
                ("pos", C.c_float * 3), ("quat", C.c_float * 4), ("scale", C.c_float * 3),
                ("mask", C.c_void_p), ("selection", C.c_void_p), ("edits", C.c_void_p)]


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc -O2 -ffp-contract=off -fopenmp)."""
    src = [os.path.join(_HERE, f) for f in ("gs_oracle.c", "gs_oracle.h", "Makefile")] + \
          [os.path.join(_HERE, "..", "include", "b200gs.h")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_record_bytes.restype = C.c_uint32
        L.orc_f32_to_f16.restype = C.c_uint16
        L.orc_f32_to_f16.argtypes = [C.c_float]
        L.orc_f16_to_f32.restype = C.c_float
        L.orc_f16_to_f32.argtypes = [C.c_uint16]
        L.orc_preprocess.restype = C.c_uint64
        L.orc_preprocess_f32.restype = C.c_uint64
        L.orc_composite_f2b.restype = C.c_uint64
        L.orc_composite_f2b_f32.restype = C.c_uint64
        L.orc_render_frame.restype = C.c_uint64
        L.orc_render_frame_f32.restype = C.c_uint64
        L.orc_num_threads.restype = C.c_int
        L.orc_export_edited.restype = C.c_uint64
        _lib = L
    return _lib


def _p(a):
    # data_as keeps a reference to the array, so temporaries stay alive for the call
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_edit():
    e = EditPod()
    e.flag = 0
    e.color[:] = [0.0, 1.0, 1.0]
    e.contrast, e.exposure, e.gamma, e.alpha = 0.0, 0.0, 1.0, 1.0
    return e


def edit_pod(flag=1, color=(0.0, 1.0, 1.0), contrast=0.0, exposure=0.0, gamma=1.0, alpha=1.0):
    e = EditPod()
    e.flag = flag
    e.color[:] = list(color)
    e.contrast, e.exposure, e.gamma, e.alpha = contrast, exposure, gamma, alpha
    return e


def record_bytes(sh, cov3d):
    return int(lib().orc_record_bytes(C.c_uint32(sh), C.c_uint32(cov3d)))


def synth_scene(seed, count, start=0):
    out = np.zeros(count, dtype=PLY)
    lib().orc_synth_scene(C.c_uint64(seed), C.c_uint64(start), C.c_uint64(count), _p(out))
    return out


def gaussian_from_ply(ply):
    ply = np.ascontiguousarray(ply, dtype=PLY)
    out = np.zeros(len(ply), dtype=GAUSSIAN)
    lib().orc_gaussian_from_ply(_p(ply), C.c_uint64(len(ply)), _p(out))
    return out


def pack(sh, cov3d, gaussians):
    g = np.ascontiguousarray(gaussians, dtype=GAUSSIAN)
    out = np.zeros(len(g) * record_bytes(sh, cov3d), dtype=np.uint8)
    lib().orc_pack(C.c_uint32(sh), C.c_uint32(cov3d), _p(g), C.c_uint64(len(g)), _p(out))
    return out


def look_at_rh(eye, target=(0, 0, 0), up=(0, 1, 0)):
    out = np.zeros(16, np.float32)
    lib().orc_look_at_rh(_p(np.asarray(eye, np.float32)), _p(np.asarray(target, np.float32)),
                         _p(np.asarray(up, np.float32)), _p(out))
    return out


def perspective_rh(vfov, aspect, z_near=0.1, z_far=1e4):
    out = np.zeros(16, np.float32)
    lib().orc_perspective_rh(C.c_float(vfov), C.c_float(aspect), C.c_float(z_near), C.c_float(z_far), _p(out))
    return out


def quat_from_euler_zyx_deg(rot_deg):
    out = np.zeros(4, np.float32)
    lib().orc_quat_from_euler_zyx_deg(_p(np.asarray(rot_deg, np.float32)), _p(out))
    return out


def make_frame(view, proj, width, height, gaussian_size=1.0, display_mode=0, sh_deg=3, no_sh0=0,
               selection_edit=None, highlight=(0, 0, 0, 0), background=(0, 0, 0, 0), query=None, query_texture=None):
    f = Frame()
    f.view[:] = [float(x) for x in view]
    f.proj[:] = [float(x) for x in proj]
    f.size[:] = [float(width), float(height)]
    f.gaussian_size = gaussian_size
    f.display_mode, f.sh_deg, f.no_sh0 = display_mode, sh_deg, no_sh0
    f.selection_edit = selection_edit if selection_edit is not None else default_edit()
    f.highlight[:] = [float(x) for x in highlight]
    f.background[:] = [float(x) for x in background]
    if query is not None:
        f.query = query
    if query_texture is not None:   # (h, w) uint8, non-zero = painted; kept alive on the frame object
        f._tex = np.ascontiguousarray(query_texture, dtype=np.uint8)
        f.query_tex = f._tex.ctypes.data
        f.query_tex_h, f.query_tex_w = f._tex.shape
    return f


class ModelRef:
    """Keeps the numpy arrays alive next to the ctypes struct."""

    def __init__(self, sh, cov3d, packed, n, pos=(0, 0, 0), quat=(0, 0, 0, 1), scale=(1, 1, 1), mask=None,
                 selection=None, edits=None):
        self.packed = np.ascontiguousarray(packed, dtype=np.uint8)
        self.mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
        self.selection = None if selection is None else np.ascontiguousarray(selection, dtype=np.uint32)
        self.edits = None if edits is None else np.ascontiguousarray(edits, dtype=EDIT)
        m = Model()
        m.sh, m.cov3d, m.n = sh, cov3d, n
        m.packed = self.packed.ctypes.data
        m.pos[:] = [float(x) for x in pos]
        m.quat[:] = [float(x) for x in quat]
        m.scale[:] = [float(x) for x in scale]
        m.mask = None if self.mask is None else self.mask.ctypes.data
        m.selection = None if self.selection is None else self.selection.ctypes.data
        m.edits = None if self.edits is None else self.edits.ctypes.data
        self.c = m
        self.n = n


def preprocess(frame, model):
    idx = np.zeros(model.n, np.uint32)
    keys = np.zeros(model.n, np.uint32)
    spl = np.zeros(model.n, SPLAT)
    v = int(lib().orc_preprocess(C.byref(frame), C.byref(model.c), _p(idx), _p(keys), _p(spl)))
    return idx[:v].copy(), keys[:v].copy(), spl[:v].copy()


def preprocess_f32(frame, model):
    """The oracle's own fp32 projected splats (nothing rounded to the product's f16 record)."""
    idx = np.zeros(model.n, np.uint32)
    keys = np.zeros(model.n, np.uint32)
    spl = np.zeros(model.n, SPLAT_F32)
    v = int(lib().orc_preprocess_f32(C.byref(frame), C.byref(model.c), _p(idx), _p(keys), _p(spl)))
    return idx[:v].copy(), keys[:v].copy(), spl[:v].copy()


def sort(keys, idx, splats=None):
    keys, idx = keys.copy(), idx.copy()
    splats = None if splats is None else splats.copy()
    if splats is not None and splats.dtype == SPLAT_F32:
        lib().orc_sort_f32(C.c_uint64(len(keys)), _p(keys), _p(idx), _p(splats))
    else:
        lib().orc_sort(C.c_uint64(len(keys)), _p(keys), _p(idx), _p(splats))
    return keys, idx, splats


def sort_pairs(keys, values, bits=32):
    keys, values = keys.copy(), values.copy()
    lib().orc_sort_pairs(C.c_uint64(len(keys)), _p(keys), _p(values), C.c_uint32(bits))
    return keys, values


def composite(frame, splats, front_to_back=False, want_float=False):
    w, h = int(frame.size[0]), int(frame.size[1])
    fp32 = splats.dtype == SPLAT_F32
    splats = np.ascontiguousarray(splats, dtype=SPLAT_F32 if fp32 else SPLAT)
    img = np.zeros((h, w, 4), np.uint8)
    imf = np.zeros((h, w, 4), np.float32) if want_float else None
    evals = 0
    f2b = lib().orc_composite_f2b_f32 if fp32 else lib().orc_composite_f2b
    b2f = lib().orc_composite_b2f_f32 if fp32 else lib().orc_composite_b2f
    if front_to_back:
        evals = int(f2b(C.byref(frame), _p(splats), C.c_uint64(len(splats)), _p(imf), _p(img)))
    else:
        b2f(C.byref(frame), _p(splats), C.c_uint64(len(splats)), _p(imf), _p(img))
    return (img, imf, evals) if want_float else (img, evals)


def order_models(frame, models, centers):
    n = len(models)
    arr = (Model * n)(*[m.c for m in models])
    order = np.zeros(n, np.uint32)
    lib().orc_order_models(C.byref(frame), arr, _p(np.ascontiguousarray(centers, np.float32)), C.c_uint32(n), _p(order))
    return order


def render_frame(frame, models_far_to_near, front_to_back=False, fp32=False):
    """Whole frame; fp32=True keeps the projected splats in fp32 end to end (orc_render_frame_f32)."""
    n = len(models_far_to_near)
    arr = (Model * n)(*[m.c for m in models_far_to_near])
    w, h = int(frame.size[0]), int(frame.size[1])
    img = np.zeros((h, w, 4), np.uint8)
    st = (C.c_double * 3)()
    fn = lib().orc_render_frame_f32 if fp32 else lib().orc_render_frame
    v = int(fn(C.byref(frame), arr, C.c_uint32(n), C.c_int(1 if front_to_back else 0), _p(img), st))
    return img, v, list(st)


def eval_mask(model, ops, shapes):
    ops = np.ascontiguousarray(ops, dtype=MASK_OP)
    shapes = np.ascontiguousarray(shapes, dtype=MASK_SHAPE)
    words = np.zeros((model.n + 31) // 32, np.uint32)
    lib().orc_eval_mask(C.byref(model.c), _p(ops), C.c_uint32(len(ops)), _p(shapes), C.c_uint32(len(shapes)), _p(words))
    return words


def query_pod(kind, op=0, p0=(0, 0), p1=(0, 0), radius=0.0):
    """kind: 0 none, 1 hit, 2 rect, 3 brush; op: 0 set, 1 add, 2 remove (include/b200gs.h)."""
    q = QueryPod()
    q.kind, q.op, q.radius = kind, op, radius
    q.p0[:] = [float(x) for x in p0]
    q.p1[:] = [float(x) for x in p1]
    return q


def query_selection(frame, model):
    words = np.zeros((model.n + 31) // 32, np.uint32)
    lib().orc_query_selection(C.byref(frame), C.byref(model.c), _p(words))
    return words


def query_texture_paint(tex, stroke):
    """Paints one rect / brush stroke (a QueryPod of kind 2 / 3) into the (h, w) uint8 texture, in place."""
    assert tex.dtype == np.uint8 and tex.flags.c_contiguous
    lib().orc_query_texture_paint(_p(tex), C.c_uint32(tex.shape[1]), C.c_uint32(tex.shape[0]), C.byref(stroke))
    return tex


def postprocess(selection, edits, selection_edit):
    """Commits the selection edit into the edit pods of the selected Gaussians (returns a new array)."""
    e = np.ascontiguousarray(edits, dtype=EDIT).copy()
    sel = np.ascontiguousarray(selection, dtype=np.uint32)
    lib().orc_postprocess(C.c_uint64(len(e)), _p(sel), _p(e), C.byref(selection_edit))
    return e


def apply_edit(edit, rgb, opacity):
    c = (C.c_float * 3)(*[float(x) for x in rgb])
    o = C.c_float(opacity)
    lib().orc_apply_edit(C.byref(edit), c, C.byref(o))
    return np.array(list(c), np.float32), float(o.value)


def export_edited(gaussians, edits=None, mask=None):
    """PLY vertices of the model as exported with edits and mask (orc_export_edited)."""
    g = np.ascontiguousarray(gaussians, dtype=GAUSSIAN)
    e = None if edits is None else np.ascontiguousarray(edits, dtype=EDIT)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
    out = np.zeros(len(g), dtype=PLY)
    k = int(lib().orc_export_edited(_p(g), C.c_uint64(len(g)), _p(e), _p(m), _p(out)))
    return out[:k].copy()


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n=None):
    """Use n OpenMP threads (default: every host core), whatever OMP_NUM_THREADS the launcher exported."""
    lib().orc_set_num_threads(C.c_int(int(n or os.cpu_count() or 1)))
    return num_threads()


def orbit_camera(radius=4.5, elev_deg=20.0, azim_deg=35.0, width=1920, height=1080, vfov_deg=60.0,
                 z_near=0.1, z_far=1e4):
    """Camera of SURVEY.md §8d: orbit around the origin (src/app.rs:1236-1244 conventions)."""
    el, az = np.float32(np.deg2rad(elev_deg)), np.float32(np.deg2rad(azim_deg))
    eye = np.array([radius * np.cos(el) * np.sin(az), radius * np.sin(el), radius * np.cos(el) * np.cos(az)], np.float32)
    view = look_at_rh(eye)
    proj = perspective_rh(np.float32(np.deg2rad(vfov_deg)), np.float32(width) / np.float32(height), z_near, z_far)
    return view, proj


def view_batch(width, height, n_az=32, n_el=8, radii=(3.0, 4.5, 6.0, 8.0), el_range=(-10.0, 60.0)):
    """(view, proj) of the 1024-view batch of SURVEY.md §8d in its fixed order (radius fastest, then elevation, then
    azimuth, the azimuths in bit-reversed order) — the oracle's own copy, so that the reference arm of bench.py never
    imports the product."""
    bits = max(1, (n_az - 1).bit_length())
    az = [a for a in (int(format(i, "0%db" % bits)[::-1], 2) for i in range(1 << bits)) if a < n_az]
    out = []
    for a in az:
        for e in range(n_el):
            el = el_range[0] + (el_range[1] - el_range[0]) * e / max(n_el - 1, 1)
            for r in radii:
                out.append(orbit_camera(r, el, 360.0 * a / n_az, width, height))
    return out
