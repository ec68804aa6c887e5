"""CPU-side tests of the product library: host code (packing, PLY, camera, synthetic scene)
against the oracle, the C ABI surface, and error behaviour without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from util import pack_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(G):
    """include/b200gs.h is the contract: every B200GS_API declaration must be exported."""
    hdr = open(os.path.join(ROOT, "include", "b200gs.h")).read()
    names = sorted(set(re.findall(r"B200GS_API[^;(]*?\b(b200gs_\w+)\s*\(", hdr)))
    assert len(names) >= 55, names
    L = G.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert G.version().startswith("b200gs")


def test_host_matches_oracle_bytes(G, O):
    """Scene generator, Gaussian::from(PlyGaussianPod) and the 8 packed layouts are byte-identical
    between the product's host code and the oracle's independent restatement."""
    for seed, n in ((1, 257), (0xB2000001, 20000)):
        a, b = G.synth_scene(seed, n), O.synth_scene(seed, n)
        assert a.tobytes() == b.tobytes()
        ga, gb = G.gaussian_from_ply(a), O.gaussian_from_ply(b)
        assert ga.tobytes() == gb.tobytes()
        for sh in range(4):
            for cov in range(2):
                assert G.record_bytes(sh, cov) == O.record_bytes(sh, cov)
                assert G.pack_gaussians(sh, cov, ga).tobytes() == O.pack(sh, cov, gb).tobytes(), (sh, cov)
    assert G.synth_scene(5, 100, start=40).tobytes() == G.synth_scene(5, 140)[40:].tobytes()


def test_camera_and_euler_match_oracle(G, O):
    rng = np.random.default_rng(3)
    for _ in range(50):
        eye = rng.uniform(-8, 8, 3).astype(np.float32)
        tgt = rng.uniform(-1, 1, 3).astype(np.float32)
        assert np.array_equal(G.look_at_rh(eye, tgt), O.look_at_rh(eye, tgt))
        fov, asp = np.float32(rng.uniform(0.3, 2.0)), np.float32(rng.uniform(0.5, 2.5))
        assert np.array_equal(G.perspective_rh(fov, asp, 0.1, 1e4), O.perspective_rh(fov, asp, 0.1, 1e4))
        rot = rng.uniform(-180, 180, 3).astype(np.float32)
        assert np.array_equal(G.quat_from_euler_zyx_deg(rot), O.quat_from_euler_zyx_deg(rot))
    cam = G.OrbitCamera.orbit()
    v, p = O.orbit_camera()
    assert np.array_equal(cam.view(), v) and np.array_equal(cam.projection(np.float32(1920) / np.float32(1080)), p)
    # the 1024-view batch: same cameras in the same order from the product and from the oracle's own copy (the reference
    # arm of bench.py uses the latter), azimuths in bit-reversed order so that contiguous per-rank blocks are balanced
    cams, ocams = G.view_batch(), O.view_batch(1920, 1080)
    assert len(cams) == 1024 == len(ocams)
    asp = np.float32(1920) / np.float32(1080)
    for i in range(0, 1024, 37):
        assert np.array_equal(cams[i].view(), ocams[i][0]) and np.array_equal(cams[i].projection(asp), ocams[i][1])
    assert G.azimuth_order(32)[:4] == [0, 16, 8, 24] and sorted(G.azimuth_order(32)) == list(range(32))
    assert G.azimuth_order(3) == [0, 2, 1] and G.azimuth_order(1) == [0]
    for n_az, kw in ((3, dict(n_el=2, radii=(4.5,))), (5, dict(n_el=1, radii=(3.0, 6.0)))):
        a, b = G.view_batch(n_az=n_az, **kw), O.view_batch(64, 64, n_az=n_az, **kw)
        assert len(a) == len(b) and all(np.array_equal(x.view(), y[0]) for x, y in zip(a, b))


def test_unpack_recovers_what_the_kernels_decode(G):
    ply = G.synth_scene(9, 500)
    g = G.gaussian_from_ply(ply)
    for sh in range(4):
        for cov in range(2):
            u = G.unpack_gaussians(sh, cov, G.pack_gaussians(sh, cov, g))
            assert np.array_equal(u["pos"], g["pos"]) and np.array_equal(u["color"], g["color"])
            tol = {0: 0.0, 1: 1e-3, 2: 1.0 / 255 + 1e-6}.get(sh)
            if sh == 3:
                assert not u["sh"].any()
            else:
                assert np.max(np.abs(u["sh"] - np.clip(g["sh"], -1, 1) if sh == 2 else u["sh"] - g["sh"])) <= tol


def test_ply_roundtrip_binary_ascii_and_reordered(G, tmp_path):
    ply = G.synth_scene(11, 1234)
    path = str(tmp_path / "scene.ply")
    G.write_ply(path, ply)
    assert G.ply_count(path) == 1234
    chunks = list(G.read_ply(path, chunk=500))                   # streaming reader
    assert [len(c) for c in chunks] == [500, 500, 234]
    assert np.concatenate(chunks).tobytes() == ply.tobytes()
    assert G.read_ply_bytes(open(path, "rb").read()).tobytes() == ply.tobytes()

    # ascii with a different property order, an extra property and no f_rest / normals
    sub = ply[:7]
    lines = ["ply", "format ascii 1.0", "comment test", "element vertex 7"]
    props = ["opacity", "x", "z", "y", "extra", "rot_0", "rot_1", "rot_2", "rot_3", "scale_0", "scale_1", "scale_2",
             "f_dc_0", "f_dc_1", "f_dc_2"]
    lines += ["property float %s" % p for p in props] + ["end_header"]
    for v in sub:
        vals = [v["opacity"], v["pos"][0], v["pos"][2], v["pos"][1], 42.0, *v["rot"], *v["scale"], *v["f_dc"]]
        lines.append(" ".join(repr(float(x)) for x in vals))
    got = G.read_ply_bytes(("\n".join(lines) + "\n").encode())
    assert np.array_equal(got["pos"], sub["pos"]) and np.array_equal(got["rot"], sub["rot"])
    assert np.array_equal(got["opacity"], sub["opacity"]) and np.array_equal(got["f_dc"], sub["f_dc"])
    assert not got["f_rest"].any() and not got["normal"].any()

    # binary with a double and a uchar column mixed in
    hdr = ("ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty double x\nproperty float y\n"
           "property float z\nproperty uchar tag\nproperty float opacity\nend_header\n").encode()
    rec = np.dtype([("x", "<f8"), ("y", "<f4"), ("z", "<f4"), ("tag", "u1"), ("opacity", "<f4")])
    body = np.zeros(3, rec)
    body["x"], body["y"], body["z"], body["opacity"] = [1.5, 2.5, 3.5], [4, 5, 6], [7, 8, 9], [0.1, 0.2, 0.3]
    got = G.read_ply_bytes(hdr + body.tobytes())
    assert np.array_equal(got["pos"], np.array([[1.5, 4, 7], [2.5, 5, 8], [3.5, 6, 9]], np.float32))
    assert np.allclose(got["opacity"], [0.1, 0.2, 0.3])


def test_ply_errors(G, tmp_path):
    with pytest.raises(G.GsError) as e:
        list(G.read_ply(str(tmp_path / "missing.ply")))
    assert e.value.code == 4                                     # B200GS_ERR_IO (gs::Error::Io)
    with pytest.raises(G.GsError) as e:
        G.read_ply_bytes(b"not a ply file\n")
    assert e.value.code == 5
    with pytest.raises(G.GsError) as e:                          # truncated body
        G.read_ply_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\n"
                         b"property float y\nproperty float z\nend_header\n" + b"\0" * 12)
    assert e.value.code == 4
    with pytest.raises(G.GsError):
        G.read_ply_bytes(b"ply\nformat binary_big_endian 1.0\nelement vertex 0\nend_header\n")


def test_gaussian_ply_roundtrip(G):
    ply = G.synth_scene(13, 300)
    g = G.gaussian_from_ply(ply)
    back = G.gaussian_from_ply(G.gaussian_to_ply(g))
    assert np.array_equal(back["pos"], g["pos"]) and np.array_equal(back["sh"], g["sh"])
    assert np.max(np.abs(back["color"].astype(int) - g["color"].astype(int))) <= 1
    assert np.allclose(back["scale"], g["scale"], rtol=1e-5) and np.allclose(back["rot"], g["rot"], atol=1e-6)


def test_export_with_edits_and_mask_matches_oracle(G, O, tmp_path):
    """Gaussians::write_ply(writer, Some(&edits), Some(mask)) — reference src/app.rs:904-914, 935-943: masked-out and
    hidden Gaussians are not written, enabled edit pods are baked into colour / opacity.  The file written by the
    product equals the oracle's export byte for byte, for every (edits, mask) combination the app can pass."""
    n = 5003
    g = G.gaussian_from_ply(G.synth_scene(0xB2000099, n))
    rng = np.random.default_rng(5)
    edits = np.zeros(n, dtype=G.EDIT)
    edits["color"] = (0.0, 1.0, 1.0)
    edits["gamma"] = 1.0
    edits["alpha"] = 1.0
    sel = rng.random(n) < 0.4
    edits["flag"][sel] = G.EDIT_ENABLED
    edits["color"][sel] = np.c_[rng.random(sel.sum()), rng.random(sel.sum()) * 2, rng.random(sel.sum()) * 2]
    edits["contrast"][sel] = rng.uniform(-1, 1, sel.sum())
    edits["exposure"][sel] = rng.uniform(-2, 2, sel.sum())
    edits["gamma"][sel] = rng.uniform(0.2, 3, sel.sum())
    edits["alpha"][sel] = rng.uniform(0, 2, sel.sum())
    hid = rng.random(n) < 0.1
    edits["flag"][hid] = G.EDIT_ENABLED | G.EDIT_HIDDEN
    ovr = (rng.random(n) < 0.1) & ~hid
    edits["flag"][ovr] = G.EDIT_ENABLED | G.EDIT_OVERRIDE_COLOR
    edits["color"][ovr] = rng.random((ovr.sum(), 3))
    shown = rng.random(n) < 0.7
    mask = pack_bits(shown)
    for e, m in ((edits, mask), (edits, None), (None, mask), (None, None)):
        path = str(tmp_path / "edited.ply")
        G.write_ply_edited(path, g, e, m)
        got = np.concatenate(list(G.read_ply(path))) if G.ply_count(path) else np.zeros(0, G.PLY)
        want = O.export_edited(g, e, m)
        keep = np.ones(n, bool) if m is None else shown.copy()
        if e is not None:
            keep &= ~hid
        assert len(got) == len(want) == int(keep.sum())
        assert got.tobytes() == want.tobytes()
        assert np.array_equal(got["pos"], g["pos"][keep])              # the survivors, in model order
    # a no-op export is the plain write_ply of the model
    plain = str(tmp_path / "plain.ply")
    G.write_ply(plain, G.gaussian_to_ply(g))
    assert open(plain, "rb").read() == open(str(tmp_path / "edited.ply"), "rb").read()
    with pytest.raises(ValueError):
        G.write_ply_edited(plain, g, edits[:10], None)


def test_invalid_arguments_fail_with_status(G):
    L = G.lib()
    assert G.record_bytes(4, 0) == 0 and G.record_bytes(2, 2) == 0
    with pytest.raises(G.GsError) as e:
        G.pack_gaussians(7, 0, np.zeros(1, G.GAUSSIAN))
    assert e.value.code == 1
    h = C.c_void_p()
    assert L.b200gs_viewer_create(0, 9, 0, 64, 64, C.byref(h)) == 1            # invalid layout
    assert L.b200gs_viewer_create(0, 2, 1, 0, 64, C.byref(h)) == 1             # invalid size
    assert L.b200gs_viewer_create(0, 2, 1, 64, 64, None) == 1
    assert b"layout" in L.b200gs_last_error() or b"null" in L.b200gs_last_error()
    assert L.b200gs_set_camera(None, None, None, None) == 1
    assert L.b200gs_model_preprocess(None, 0) == 1


def test_no_cpu_fallback_without_gpu(G):
    """The product path must fail loudly, not fall back, when no CUDA device is usable."""
    if G.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(G.GsError) as e:
        G.Viewer(64, 64)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "wgpu-3dgs-viewer-app_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "gs_oracle" not in src and "from oracle" not in src and "import oracle" not in src, fn


def _build_gs_hpp_check(tmp_path):
    import subprocess
    exe = str(tmp_path / "gs_hpp_check")
    pkg = os.path.join(ROOT, "wgpu-3dgs-viewer-app_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", os.path.join(ROOT, "tests", "gs_hpp_check.cpp"), "-o", exe,
                    "-L" + pkg, "-lb200gs", "-Wl,-rpath," + pkg], check=True)
    return exe


def test_cpp_gs_mirror_host_side(G, tmp_path):
    """host/gs.hpp — the C++ mirror of the reference's gs:: names — compiles against the C ABI and its
    host half (PLY streaming, Gaussian::from, layout constants, error type) works without a GPU."""
    import subprocess
    exe = _build_gs_hpp_check(tmp_path)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host-only ok" in r.stdout


def test_rust_sys_binding_is_current_and_complete(G):
    """integration/rust/b200gs-sys/src/lib.rs is generated from include/b200gs.h (tools/gen_rust_sys.py): it must be up to
    date and declare every function the header declares (there is no Rust toolchain here to compile it)."""
    import importlib.util, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(root, "tools", "gen_rust_sys.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    text, names = mod.generate()
    assert open(mod.OUT).read() == text, "stale binding: run python tools/gen_rust_sys.py"
    declared = set(re.findall(r"B200GS_API[^;(]*?\b(b200gs_\w+)\s*\(", open(mod.HDR).read()))
    assert declared and declared == set(names)
    assert len(re.findall(r"pub fn b200gs_", text)) == len(declared)
