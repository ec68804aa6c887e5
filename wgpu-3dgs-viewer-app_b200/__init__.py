"""b200gs — host-side harness over the C ABI of the B200-native 3DGS render core.

The product is `libb200gs.so` (CUDA kernels + C ABI, include/b200gs.h).  This module is the thin
ctypes binding used by the tests, bench.py and headless callers; it mirrors the names of the
reference's `gs::` API (crate wgpu-3dgs-viewer as used by src/tab/scene.rs): `MultiModelViewer`
(`Viewer`), per-model `GaussiansBuffer.update_range`, `preprocessor.preprocess`,
`radix_sorter.sort`, `renderer.render`, `Camera`, `Gaussians.read_ply`.  There is no CPU
fallback: compute calls raise `GsError` when the library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200gs.so")

SH_SINGLE, SH_HALF, SH_NORM8, SH_NONE = 0, 1, 2, 3
COV3D_SINGLE, COV3D_HALF = 0, 1
DISPLAY_SPLAT, DISPLAY_ELLIPSE, DISPLAY_POINT = 0, 1, 2
EDIT_ENABLED, EDIT_HIDDEN, EDIT_OVERRIDE_COLOR = 1, 2, 4
MASK_BOX, MASK_ELLIPSOID = 0, 1
MASKOP_SHAPE, MASKOP_UNION, MASKOP_INTERSECTION, MASKOP_DIFFERENCE, MASKOP_SYMDIFF, MASKOP_COMPLEMENT, MASKOP_RESET = range(7)

GAUSSIAN = np.dtype([("rot", "<f4", 4), ("pos", "<f4", 3), ("color", "u1", 4), ("sh", "<f4", 45), ("scale", "<f4", 3)])
PLY = np.dtype([("pos", "<f4", 3), ("normal", "<f4", 3), ("f_dc", "<f4", 3), ("f_rest", "<f4", 45),
                ("opacity", "<f4"), ("scale", "<f4", 3), ("rot", "<f4", 4)])
EDIT = np.dtype([("flag", "<u4"), ("color", "<f4", 3), ("contrast", "<f4"), ("exposure", "<f4"),
                 ("gamma", "<f4"), ("alpha", "<f4")])
SPLAT = np.dtype([("mx", "<f4"), ("my", "<f4"), ("radius", "<u2"), ("opacity_h", "<f2"), ("r_h", "<f2"),
                  ("g_h", "<f2"), ("ca", "<f4"), ("cb", "<f4"), ("cc", "<f4"), ("b_h", "<f2"), ("flags", "<u2")])
MASK_SHAPE = np.dtype([("kind", "<u4"), ("pos", "<f4", 3), ("quat", "<f4", 4), ("scale", "<f4", 3)])
MASK_OP = np.dtype([("kind", "<u4"), ("arg", "<u4")])
HIT = np.dtype([("model", "<u4"), ("index", "<u4"), ("alpha", "<f4"), ("depth", "<f4")])


class EditPod(C.Structure):
    """gs::GaussianEditPod (reference src/app.rs:1556-1563)."""
    _fields_ = [("flag", C.c_uint32), ("color", C.c_float * 3), ("contrast", C.c_float), ("exposure", C.c_float),
                ("gamma", C.c_float), ("alpha", C.c_float)]

    @staticmethod
    def default():
        return EditPod.new(0)

    @staticmethod
    def new(flag, color=(0.0, 1.0, 1.0), contrast=0.0, exposure=0.0, gamma=1.0, alpha=1.0):
        e = EditPod()
        e.flag = flag
        e.color[:] = list(color)
        e.contrast, e.exposure, e.gamma, e.alpha = contrast, exposure, gamma, alpha
        return e


class QueryPod(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("op", C.c_uint32), ("p0", C.c_float * 2), ("p1", C.c_float * 2),
                ("radius", C.c_float), ("_pad", C.c_uint32)]


QUERY_NONE, QUERY_HIT, QUERY_RECT, QUERY_BRUSH, QUERY_TEXTURE = 0, 1, 2, 3, 4
SELECT_SET, SELECT_ADD, SELECT_REMOVE = 0, 1, 2


def query_pod(kind=QUERY_NONE, op=SELECT_SET, p0=(0, 0), p1=(0, 0), radius=0.0):
    """gs::Query*Pod / QueryToolset in immediate mode (reference src/tab/scene.rs:758-791, 1224-1263):
    rect = [p0, p1] in viewport pixels (top-left origin), brush = segment p0->p1 with `radius`."""
    q = QueryPod()
    q.kind, q.op, q.radius = kind, op, radius
    q.p0[:] = [float(x) for x in p0]
    q.p1[:] = [float(x) for x in p1]
    return q


class Timings(C.Structure):
    _fields_ = [("preprocess_ms", C.c_float), ("sort_ms", C.c_float), ("bin_ms", C.c_float),
                ("composite_ms", C.c_float), ("total_ms", C.c_float), ("visible", C.c_uint64),
                ("tile_entries", C.c_uint64), ("evals", C.c_uint64), ("staged_entries", C.c_uint64),
                ("overflow", C.c_uint32), ("_pad", C.c_uint32)]


ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_IO, ERR_FORMAT, ERR_OVERFLOW = 1, 2, 3, 4, 5, 6


class GsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200gs error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load libb200gs.so (built in-tree by build.py).  Fails loudly if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GsError(-1, "%s not found: run `python wgpu-3dgs-viewer-app_b200/build.py` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.b200gs_last_error.restype = C.c_char_p
        L.b200gs_version.restype = C.c_char_p
        L.b200gs_record_bytes.restype = C.c_uint32
        L.b200gs_model_len.restype = C.c_uint64
        L.b200gs_stream.restype = C.c_void_p
        L.b200gs_image_device.restype = C.c_void_p
        L.b200gs_model_find.restype = C.c_void_p
        for n in ("b200gs_stream", "b200gs_image_device", "b200gs_viewer_destroy", "b200gs_sync"):
            getattr(L, n).argtypes = [C.c_void_p]
        L.b200gs_model_len.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _ck(code):
    if code != 0:
        raise GsError(code, lib().b200gs_last_error().decode("utf-8", "replace"))


def _p(a):
    # data_as keeps a reference to the array, so temporaries stay alive for the call
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f(a, n):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    assert a.size == n, "expected %d floats" % n
    return a


def version():
    return lib().b200gs_version().decode()


def device_count():
    n = C.c_int(0)
    _ck(lib().b200gs_device_count(C.byref(n)))
    return n.value


def record_bytes(sh, cov3d):
    return int(lib().b200gs_record_bytes(C.c_uint32(sh), C.c_uint32(cov3d)))


# ------------------------------------------------------------------ host side (no GPU needed)
def synth_scene(seed, count, start=0):
    """Deterministic synthetic scene of SURVEY.md §8d as Inria-format PLY vertices."""
    out = np.zeros(count, dtype=PLY)
    _ck(lib().b200gs_synth_scene(C.c_uint64(seed), C.c_uint64(start), C.c_uint64(count), _p(out)))
    return out


def gaussian_from_ply(ply):
    """gs::Gaussian::from(PlyGaussianPod) (reference src/app.rs:1066)."""
    ply = np.ascontiguousarray(ply, dtype=PLY)
    out = np.zeros(len(ply), dtype=GAUSSIAN)
    _ck(lib().b200gs_gaussian_from_ply(_p(ply), C.c_uint64(len(ply)), _p(out)))
    return out


def gaussian_to_ply(g):
    g = np.ascontiguousarray(g, dtype=GAUSSIAN)
    out = np.zeros(len(g), dtype=PLY)
    _ck(lib().b200gs_gaussian_to_ply(_p(g), C.c_uint64(len(g)), _p(out)))
    return out


def pack_gaussians(sh, cov3d, gaussians, out=None):
    """Host half of GaussiansBuffer::update_range (reference src/tab/scene.rs:2069-2085)."""
    g = np.ascontiguousarray(gaussians, dtype=GAUSSIAN)
    rb = record_bytes(sh, cov3d)
    if out is None:
        out = np.zeros(len(g) * rb, dtype=np.uint8)
    _ck(lib().b200gs_pack_gaussians(C.c_uint32(sh), C.c_uint32(cov3d), _p(g), C.c_uint64(len(g)), _p(out)))
    return out


def unpack_gaussians(sh, cov3d, packed):
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    n = packed.size // record_bytes(sh, cov3d)
    out = np.zeros(n, dtype=GAUSSIAN)
    _ck(lib().b200gs_unpack_gaussians(C.c_uint32(sh), C.c_uint32(cov3d), _p(packed), C.c_uint64(n), _p(out)))
    return out


def look_at_rh(eye, target=(0, 0, 0), up=(0, 1, 0)):
    out = np.zeros(16, np.float32)
    lib().b200gs_look_at_rh(_p(_f(eye, 3)), _p(_f(target, 3)), _p(_f(up, 3)), _p(out))
    return out


def perspective_rh(vfov, aspect, z_near=0.1, z_far=1e4):
    out = np.zeros(16, np.float32)
    lib().b200gs_perspective_rh(C.c_float(vfov), C.c_float(aspect), C.c_float(z_near), C.c_float(z_far), _p(out))
    return out


def quat_from_euler_zyx_deg(rot_deg):
    out = np.zeros(4, np.float32)
    lib().b200gs_quat_from_euler_zyx_deg(_p(_f(rot_deg, 3)), _p(out))
    return out


class OrbitCamera:
    """CameraOrbitControl + gs::CameraTrait (reference src/app.rs:1174-1244)."""

    def __init__(self, target=(0, 0, 0), pos=(0, 0, -1), z=(0.1, 1e4), vertical_fov=np.deg2rad(60.0)):
        self.target, self.pos, self.z, self.vertical_fov = target, pos, z, np.float32(vertical_fov)

    @staticmethod
    def orbit(radius=4.5, elev_deg=20.0, azim_deg=35.0, **kw):
        el, az = np.float32(np.deg2rad(elev_deg)), np.float32(np.deg2rad(azim_deg))
        pos = np.array([radius * np.cos(el) * np.sin(az), radius * np.sin(el), radius * np.cos(el) * np.cos(az)], np.float32)
        return OrbitCamera(pos=pos, **kw)

    def view(self):
        return look_at_rh(self.pos, self.target, (0, 1, 0))

    def projection(self, aspect_ratio):
        return perspective_rh(self.vertical_fov, np.float32(aspect_ratio), self.z[0], self.z[1])


def azimuth_order(n_az):
    """Order in which view_batch visits its azimuths: bit-reversed indices (0, n/2, n/4, 3n/4, ...), so that every
    contiguous run of azimuths is spread around the whole orbit."""
    bits = max(1, (n_az - 1).bit_length())
    order = [int(format(i, "0%db" % bits)[::-1], 2) for i in range(1 << bits)]
    return [a for a in order if a < n_az]


def view_batch(n_az=32, n_el=8, radii=(3.0, 4.5, 6.0, 8.0), el_range=(-10.0, 60.0)):
    """The 1024-view batch of SURVEY.md §8d: 32 azimuths x 8 elevations x 4 radii, fixed order.

    Radius varies fastest, then elevation, then azimuth, and the azimuths are visited in bit-reversed order
    (0, 180, 90, 270, 45, ... degrees): every run of 32 consecutive views holds all radii and elevations, and the
    contiguous per-rank blocks of a 2 / 4 / 8-GPU run each hold azimuths from all around the orbit.  A frame's cost
    follows the azimuth smoothly (0.81 - 0.88 ms on the bench scene, tools/view_costs.py); with the azimuths in
    natural order the slowest of eight contiguous blocks carried 3 % more work than the average one, which capped the
    8-GPU efficiency at 0.97 whatever the gather did; in this order the blocks are within 0.8 %."""
    cams = []
    for a in azimuth_order(n_az):
        for e in range(n_el):
            el = el_range[0] + (el_range[1] - el_range[0]) * e / max(n_el - 1, 1)
            for r in radii:
                cams.append(OrbitCamera.orbit(r, el, 360.0 * a / n_az))
    return cams


def partition_views(n_views, world_size, rank):
    """Contiguous block [lo, hi) of a view batch owned by `rank` (views are the
    independent units; every GPU holds a full scene replica)."""
    per, extra = divmod(n_views, world_size)
    lo = rank * per + min(rank, extra)
    return lo, lo + per + (1 if rank < extra else 0)


def read_ply(path, chunk=1 << 20):
    """Gaussians::read_ply_header + read_ply_gaussians (reference src/app.rs:1056-1070): yields
    chunks of PLY vertices as they are parsed (streaming)."""
    r, n = C.c_void_p(), C.c_uint64(0)
    _ck(lib().b200gs_ply_open(path.encode(), C.byref(r), C.byref(n)))
    try:
        left = n.value
        while left:
            buf = np.zeros(min(chunk, left), dtype=PLY)
            got = C.c_uint64(0)
            _ck(lib().b200gs_ply_read(r, _p(buf), C.c_uint64(len(buf)), C.byref(got)))
            if got.value == 0:
                break
            left -= got.value
            yield buf[:got.value]
    finally:
        lib().b200gs_ply_close(r)


def ply_count(path):
    r, n = C.c_void_p(), C.c_uint64(0)
    _ck(lib().b200gs_ply_open(path.encode(), C.byref(r), C.byref(n)))
    lib().b200gs_ply_close(r)
    return n.value


def read_ply_bytes(data):
    buf = np.frombuffer(data, dtype=np.uint8)
    r, n = C.c_void_p(), C.c_uint64(0)
    _ck(lib().b200gs_ply_open_memory(_p(buf), C.c_size_t(buf.size), C.byref(r), C.byref(n)))
    try:
        out = np.zeros(n.value, dtype=PLY)
        got = C.c_uint64(0)
        _ck(lib().b200gs_ply_read(r, _p(out), C.c_uint64(n.value), C.byref(got)))
        return out[:got.value]
    finally:
        lib().b200gs_ply_close(r)


def write_ply(path, verts):
    verts = np.ascontiguousarray(verts, dtype=PLY)
    _ck(lib().b200gs_ply_write(path.encode(), _p(verts), C.c_uint64(len(verts))))


def write_ply_edited(path, gaussians, edits=None, mask=None):
    """Gaussians::write_ply(writer, Some(&edits), Some(mask)) (reference src/app.rs:904-914): masked-out and hidden
    Gaussians are dropped, enabled edit pods are baked into colour / opacity."""
    g = np.ascontiguousarray(gaussians, dtype=GAUSSIAN)
    e = None if edits is None else np.ascontiguousarray(edits, dtype=EDIT)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint32)
    if e is not None and len(e) != len(g):
        raise ValueError("one edit pod per Gaussian")
    if m is not None and len(m) != (len(g) + 31) // 32:
        raise ValueError("mask must hold ceil(N/32) words")
    _ck(lib().b200gs_ply_write_edited(path.encode(), _p(g), C.c_uint64(len(g)), _p(e), _p(m)))


def hit_pos_by_closest(hits, view, proj, size, px, py):
    """gs::query::hit_pos_by_closest (reference src/tab/scene.rs:668-672)."""
    hits = np.ascontiguousarray(hits, dtype=HIT)
    out = np.zeros(3, np.float32)
    _ck(lib().b200gs_hit_pos_by_closest(_p(hits), C.c_uint64(len(hits)), _p(_f(view, 16)), _p(_f(proj, 16)), _p(_f(size, 2)),
                                        C.c_uint32(px), C.c_uint32(py), _p(out)))
    return out


def hit_pos_by_alpha_range(hits, alpha_threshold, view, proj, size, px, py):
    """gs::query::hit_pos_by_alpha_range(.., 0.05) (reference src/tab/scene.rs:659-667)."""
    hits = np.ascontiguousarray(hits, dtype=HIT)
    out = np.zeros(3, np.float32)
    _ck(lib().b200gs_hit_pos_by_alpha_range(_p(hits), C.c_uint64(len(hits)), C.c_float(alpha_threshold), _p(_f(view, 16)),
                                            _p(_f(proj, 16)), _p(_f(size, 2)), C.c_uint32(px), C.c_uint32(py), _p(out)))
    return out


# ------------------------------------------------------------------ pinned host memory
class PinnedBuffer:
    def __init__(self, nbytes):
        p = C.c_void_p()
        _ck(lib().b200gs_host_alloc(C.c_size_t(nbytes), C.byref(p)))
        self.ptr, self.nbytes = p.value, nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            lib().b200gs_host_free(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ------------------------------------------------------------------ viewer / model
class Model:
    """One entry of gs::MultiModelViewer::models (reference src/tab/scene.rs:2111-2139)."""

    def __init__(self, viewer, key, handle, capacity):
        self.viewer, self.key, self.h, self.capacity = viewer, key, handle, capacity

    def __len__(self):
        return int(lib().b200gs_model_len(self.h))

    def update_range(self, start, gaussians):
        """gaussians_buffer.update_range(queue, start, &[Gaussian]) — scene.rs:2076-2084."""
        g = np.ascontiguousarray(gaussians, dtype=GAUSSIAN)
        _ck(lib().b200gs_model_update_range(self.h, C.c_uint64(start), _p(g), C.c_uint64(len(g))))

    def upload_packed(self, start, packed):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        n = packed.size // self.viewer.record_bytes
        _ck(lib().b200gs_model_upload_packed(self.h, C.c_uint64(start), _p(packed), C.c_uint64(n)))

    def upload_packed_device(self, start, dev_ptr, count):
        _ck(lib().b200gs_model_upload_packed_device(self.h, C.c_uint64(start), C.c_void_p(dev_ptr), C.c_uint64(count)))

    def set_transform(self, pos=(0, 0, 0), quat=(0, 0, 0, 1), scale=(1, 1, 1)):
        """viewer.update_model_transform(queue, key, pos, quat, scale) — scene.rs:796-802."""
        _ck(lib().b200gs_model_set_transform(self.h, _p(_f(pos, 3)), _p(_f(quat, 4)), _p(_f(scale, 3))))

    def upload_mask(self, words):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        _ck(lib().b200gs_model_upload_mask(self.h, _p(w), C.c_uint64(len(w))))

    def upload_selection(self, words):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        _ck(lib().b200gs_model_upload_selection(self.h, _p(w), C.c_uint64(len(w))))

    def upload_edits(self, start, pods):
        e = np.ascontiguousarray(pods, dtype=EDIT)
        _ck(lib().b200gs_model_upload_edits(self.h, C.c_uint64(start), _p(e), C.c_uint64(len(e))))

    def eval_mask(self, ops, shapes):
        """mask_evaluator.evaluate(...) — scene.rs:2124-2131, 2201-2209 (postfix op list)."""
        ops = np.ascontiguousarray(ops, dtype=MASK_OP)
        shapes = np.ascontiguousarray(shapes, dtype=MASK_SHAPE)
        _ck(lib().b200gs_model_eval_mask(self.h, _p(ops), C.c_uint32(len(ops)), _p(shapes), C.c_uint32(len(shapes))))

    def postprocess(self):
        _ck(lib().b200gs_model_postprocess(self.h))

    def preprocess(self, use_unedited=False):
        """viewer.preprocessor.preprocess(encoder, bind_group, N) — scene.rs:856-863."""
        _ck(lib().b200gs_model_preprocess(self.h, C.c_int(1 if use_unedited else 0)))

    def sort(self):
        """viewer.radix_sorter.sort(encoder, bind_group, indirect_args) — scene.rs:865-869."""
        _ck(lib().b200gs_model_sort(self.h))

    def visible_count(self):
        n = C.c_uint64(0)
        _ck(lib().b200gs_model_visible_count(self.h, C.byref(n)))
        return n.value

    def _dl(self, fn, dtype, cap):
        out = np.zeros(max(cap, 1), dtype=dtype)
        n = C.c_uint64(0)
        _ck(fn(self.h, _p(out), C.c_uint64(cap), C.byref(n)))
        return out[:n.value]

    def depth_keys(self):
        return self._dl(lib().b200gs_model_download_depth_keys, np.uint32, self.visible_count())

    def indices(self):
        return self._dl(lib().b200gs_model_download_indices, np.uint32, self.visible_count())

    def splats(self):
        return self._dl(lib().b200gs_model_download_splats, SPLAT, self.visible_count())

    def download_mask(self):
        return self._dl(lib().b200gs_model_download_mask, np.uint32, (self.capacity + 31) // 32)

    def download_selection(self):
        return self._dl(lib().b200gs_model_download_selection, np.uint32, (self.capacity + 31) // 32)

    def download_edits(self):
        return self._dl(lib().b200gs_model_download_edits, EDIT, self.capacity)

    def download_packed(self, start=0, count=None):
        count = self.capacity - start if count is None else count
        out = np.zeros(count * self.viewer.record_bytes, dtype=np.uint8)
        _ck(lib().b200gs_model_download_packed(self.h, C.c_uint64(start), _p(out), C.c_uint64(count)))
        return out


def set_tuning(name, value):
    _ck(lib().b200gs_set_tuning(name.encode(), C.c_int64(int(value))))


class Viewer:
    """gs::MultiModelViewer::<G>::new_with(device, format, depth_stencil, size) — scene.rs:1969-1980.

    `sh`/`cov3d` choose G among the 8 GaussianPod layouts (reference src/app.rs:250-257; default
    Norm8 + Half as in src/app.rs:398-417)."""

    def __init__(self, width, height, sh=SH_NORM8, cov3d=COV3D_HALF, device=0):
        h = C.c_void_p()
        _ck(lib().b200gs_viewer_create(C.c_int(device), C.c_uint32(sh), C.c_uint32(cov3d), C.c_uint32(width),
                                       C.c_uint32(height), C.byref(h)))
        self.h, self.sh, self.cov3d, self.width, self.height = h, sh, cov3d, width, height
        self.record_bytes = record_bytes(sh, cov3d)
        self.models = {}
        self._pinned = None

    def close(self):
        if self.h:
            lib().b200gs_viewer_destroy(self.h)
            self.h = None
            self.models = {}
        if self._pinned is not None:
            self._pinned.free()
            self._pinned = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def resize(self, width, height):
        _ck(lib().b200gs_resize(self.h, C.c_uint32(width), C.c_uint32(height)))
        self.width, self.height = width, height

    def update_camera(self, camera, size=None):
        """viewer.update_camera(queue, &impl CameraTrait, size) — scene.rs:795."""
        size = (self.width, self.height) if size is None else size
        self.update_camera_matrices(camera.view(), camera.projection(np.float32(size[0]) / np.float32(size[1])), size)

    def update_camera_matrices(self, view, proj, size=None):
        size = (self.width, self.height) if size is None else size
        _ck(lib().b200gs_set_camera(self.h, _p(_f(view, 16)), _p(_f(proj, 16)), _p(_f(size, 2))))
        self.width, self.height = int(size[0]), int(size[1])

    def update_gaussian_transform(self, size=1.0, display_mode=DISPLAY_SPLAT, sh_deg=3, no_sh0=False):
        """viewer.update_gaussian_transform(queue, size, display_mode, sh_deg, no_sh0) — scene.rs:803-809."""
        _ck(lib().b200gs_set_gaussian_transform(self.h, C.c_float(size), C.c_uint32(display_mode), C.c_uint32(sh_deg),
                                                C.c_uint32(1 if no_sh0 else 0)))

    def update_selection_edit(self, pod):
        _ck(lib().b200gs_set_selection_edit(self.h, C.byref(pod)))

    def update_selection_highlight(self, rgba):
        _ck(lib().b200gs_set_selection_highlight(self.h, _p(_f(rgba, 4))))

    def update_query(self, pod):
        _ck(lib().b200gs_set_query(self.h, C.byref(pod)))

    def query_texture_clear(self):
        _ck(lib().b200gs_query_texture_clear(self.h))

    def query_texture_paint(self, stroke):
        """One rect / brush stroke (a QueryPod of kind QUERY_RECT / QUERY_BRUSH) painted into the query texture."""
        _ck(lib().b200gs_query_texture_paint(self.h, C.byref(stroke)))

    def query_texture_upload(self, texels):
        t = np.ascontiguousarray(texels, dtype=np.uint8)
        _ck(lib().b200gs_query_texture_upload(self.h, _p(t), C.c_uint32(t.shape[1]), C.c_uint32(t.shape[0])))

    def query_texture_download(self):
        out = np.zeros((self.height, self.width), np.uint8)
        _ck(lib().b200gs_query_texture_download(self.h, _p(out), C.c_size_t(out.size)))
        return out

    def set_background(self, rgba):
        _ck(lib().b200gs_set_background(self.h, _p(_f(rgba, 4))))

    def info(self, name):
        out = C.c_int64(0)
        _ck(lib().b200gs_get_info(self.h, name.encode(), C.byref(out)))
        return int(out.value)

    def set_tile_entry_capacity(self, entries):
        _ck(lib().b200gs_set_tile_entry_capacity(self.h, C.c_uint64(entries)))

    def enable_timings(self, on=True, count_evals=False):
        _ck(lib().b200gs_enable_timings(self.h, C.c_int(1 if on else 0), C.c_int(1 if count_evals else 0)))

    def sync(self):
        _ck(lib().b200gs_sync(self.h))

    def stream(self):
        return lib().b200gs_stream(self.h)

    def image_device(self):
        return lib().b200gs_image_device(self.h)

    def add_shared_model(self, key, source):
        """A model over the packed records already resident in `source` (a Model of another viewer on the same
        device): reference-counted, nothing is copied."""
        h = C.c_void_p()
        _ck(lib().b200gs_model_create_shared(self.h, key.encode(), source.h, C.byref(h)))
        m = Model(self, key, h, source.capacity)
        self.models[key] = m
        return m

    def add_model(self, key, capacity):
        h = C.c_void_p()
        _ck(lib().b200gs_model_create(self.h, key.encode(), C.c_uint64(capacity), C.byref(h)))
        m = Model(self, key, h, capacity)
        self.models[key] = m
        return m

    def remove_model(self, key):
        """viewer.remove_model(key) — scene.rs:2176."""
        m = self.models.pop(key)
        _ck(lib().b200gs_model_destroy(self.h, m.h))

    def order_models(self, models, centers):
        """Farthest-first model order of scene.rs:533-558."""
        n = len(models)
        arr = (C.c_void_p * n)(*[m.h for m in models])
        order = np.zeros(n, np.uint32)
        _ck(lib().b200gs_order_models(self.h, arr, _p(_f(centers, 3 * n)), C.c_uint32(n), _p(order)))
        return [models[i] for i in order]

    def _handles(self, models):
        return (C.c_void_p * len(models))(*[m.h for m in models])

    def render(self, models_far_to_near, out_dev, pitch=None):
        """renderer.render_with_pass over model_render_keys — scene.rs:2302-2314 (device target)."""
        pitch = self.width * 4 if pitch is None else pitch
        _ck(lib().b200gs_render(self.h, self._handles(models_far_to_near), C.c_uint32(len(models_far_to_near)),
                                C.c_void_p(out_dev), C.c_size_t(pitch)))

    def render_frame(self, models_far_to_near, out_dev=None, pitch=None):
        """preprocess + sort for each model, then render; enqueue only (device target)."""
        out_dev = self.image_device() if out_dev is None else out_dev
        pitch = self.width * 4 if pitch is None else pitch
        _ck(lib().b200gs_render_frame(self.h, self._handles(models_far_to_near), C.c_uint32(len(models_far_to_near)),
                                      C.c_void_p(out_dev), C.c_size_t(pitch)))

    def render_frame_host(self, models_far_to_near, camera=None, out=None):
        """Whole frame, host camera in, host RGBA8 image out (H, W, 4) — the end-to-end call."""
        nbytes = self.width * self.height * 4
        if out is None:
            if self._pinned is None or self._pinned.nbytes != nbytes:
                if self._pinned is not None:
                    self._pinned.free()
                self._pinned = PinnedBuffer(nbytes)
            out = self._pinned.array
        view = proj = None
        if camera is not None:
            view = _f(camera.view(), 16)
            proj = _f(camera.projection(np.float32(self.width) / np.float32(self.height)), 16)
        _ck(lib().b200gs_render_frame_host(self.h, self._handles(models_far_to_near), C.c_uint32(len(models_far_to_near)),
                                           _p(view), _p(proj), C.c_void_p(out.ctypes.data)))
        return out.reshape(self.height, self.width, 4)

    def render_frame_host_begin(self, models_far_to_near, camera, out):
        """Pipelined end-to-end frame: returns at once; `out` (pinned uint8 array) is valid after the
        matching render_frame_host_end().  At most two frames in flight."""
        view = _f(camera.view(), 16)
        proj = _f(camera.projection(np.float32(self.width) / np.float32(self.height)), 16)
        _ck(lib().b200gs_render_frame_host_begin(self.h, self._handles(models_far_to_near), C.c_uint32(len(models_far_to_near)),
                                                 _p(view), _p(proj), C.c_void_p(out.ctypes.data)))

    def render_frame_host_end(self):
        _ck(lib().b200gs_render_frame_host_end(self.h))

    def launch_count(self):
        n = C.c_uint64(0)
        _ck(lib().b200gs_launch_count(self.h, C.byref(n)))
        return n.value

    def query_hits(self, models_far_to_near, px, py, cap=4096):
        """gs::QueryHitPod + query::download (reference src/tab/scene.rs:617-657): ordered hit list of a pixel
        of the last rendered frame."""
        out = np.zeros(cap, dtype=HIT)
        n = C.c_uint64(0)
        _ck(lib().b200gs_query_hits(self.h, self._handles(models_far_to_near), C.c_uint32(len(models_far_to_near)),
                                    C.c_uint32(px), C.c_uint32(py), _p(out), C.c_uint64(cap), C.byref(n)))
        return out[:min(n.value, cap)]

    def last_timings(self):
        t = Timings()
        _ck(lib().b200gs_last_timings(self.h, C.byref(t)))
        return t

    def sort_pairs(self, keys, values, bits=32, wide=False):
        """Raw sort: the 8-bit-digit depth sort (bits 16 / 32), or with wide=True the 11-bit cluster sort (bits 1..32)."""
        keys = np.ascontiguousarray(keys, dtype=np.uint32).copy()
        values = np.ascontiguousarray(values, dtype=np.uint32).copy()
        fn = lib().b200gs_sort_pairs_wide_host if wide else lib().b200gs_sort_pairs_host
        _ck(fn(self.h, _p(keys), _p(values), C.c_uint64(len(keys)), C.c_uint32(bits)))
        return keys, values

    def sort_pairs_device(self, keys_ptr, values_ptr, n, bits=32):
        _ck(lib().b200gs_sort_pairs_device(self.h, C.c_void_p(keys_ptr), C.c_void_p(values_ptr), C.c_uint64(n),
                                           C.c_uint32(bits)))


MultiModelViewer = Viewer
