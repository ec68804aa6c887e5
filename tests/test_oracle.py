"""Pins the CPU oracle with analytic known-answer cases: the reference tree
holds no tests or golden vectors for this path (parity unpinned), so the closed forms below and
the committed fixtures in tests/golden/ are what the oracle is held to."""
import math
import os

import numpy as np
import pytest

from util import bits_set, make_gaussians, pack_bits, psnr

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ------------------------------------------------------------------ number formats
def test_f16_conversion_matches_ieee(O):
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.standard_normal(20000).astype(np.float32),
        (rng.standard_normal(5000) * 1e-6).astype(np.float32),
        (rng.standard_normal(5000) * 7e4).astype(np.float32),
        np.array([0.0, -0.0, 65504.0, 65520.0, 1e-8, 5.96e-8, 6.1e-5, np.inf, -np.inf, 0.333251953125], np.float32)])
    L = O.lib()
    for v in vals:
        h = L.orc_f32_to_f16(float(v))
        assert h == int(np.float32(v).astype(np.float16).view(np.uint16)), v
    # every half value converts back exactly
    allh = np.arange(0, 65536, 7, dtype=np.uint16)
    for h in allh:
        f = L.orc_f16_to_f32(int(h))
        ref = np.uint16(h).view(np.float16).astype(np.float32)
        assert (np.isnan(ref) and math.isnan(f)) or np.float32(f) == ref


def test_norm8_roundtrip_table(O):
    g = make_gaussians(O.GAUSSIAN, [[0, 0, 0]] * 256)
    xs = np.linspace(-1.2, 1.2, 256).astype(np.float32)
    g["sh"][:, 0] = xs
    packed = O.pack(2, 1, g).reshape(256, 76)
    q = packed[:, 16]
    expect = np.floor(np.clip((xs.astype(np.float32) + np.float32(1)) * np.float32(0.5), 0, 1) * np.float32(255) + np.float32(0.5))
    assert np.array_equal(q, expect.astype(np.uint8))
    dec = q.astype(np.float32) * np.float32(2.0 / 255.0) - np.float32(1)
    assert np.max(np.abs(dec - np.clip(xs, -1, 1))) <= 1.0 / 255.0 + 1e-6


def test_record_sizes(O):
    # 16 + SH{180,92,48,0} + Cov{24,12}  (reference src/tab/scene.rs:907-978)
    sizes = {(0, 0): 220, (0, 1): 208, (1, 0): 132, (1, 1): 120, (2, 0): 88, (2, 1): 76, (3, 0): 40, (3, 1): 28}
    for (sh, cov), s in sizes.items():
        assert O.record_bytes(sh, cov) == s
    assert O.record_bytes(4, 0) == 0 and O.record_bytes(0, 2) == 0


# ------------------------------------------------------------------ synthetic scene
def _mix64(z):
    M = (1 << 64) - 1
    z = (z + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)


def _u01(seed, i, k):
    M = (1 << 64) - 1
    h = _mix64((_mix64(seed ^ ((i * 0xD1342543DE82EF95) & M)) + k) & M)
    return ((h >> 11) + 0.5) / 9007199254740992.0


def test_synth_scene_pure_python_crosscheck(O):
    """the generator must be reproducible outside C — restated with Python ints."""
    seed, n = 0xB2000001, 64
    ply = O.synth_scene(seed, n)
    cl = []
    for c in range(64):
        s2 = seed ^ 0xC1A57E2500000000
        cl.append(((2 * _u01(s2, c, 0) - 1) * 4, (2 * _u01(s2, c, 1) - 1) * 1.5, (2 * _u01(s2, c, 2) - 1) * 4,
                   0.05 + 0.35 * _u01(s2, c, 3)))
    for i in range(n):
        nrm = []
        for k in range(32):
            u1, u2 = _u01(seed, i, 2 * k), _u01(seed, i, 2 * k + 1)
            r = math.sqrt(-2.0 * math.log(u1))
            t = 6.283185307179586476925286766559 * u2
            nrm += [r * math.cos(t), r * math.sin(t)]
        if _u01(seed, i, 100) < 0.1:
            pos = [(2 * _u01(seed, i, 102) - 1) * 4, (2 * _u01(seed, i, 103) - 1) * 1.5, (2 * _u01(seed, i, 104) - 1) * 4]
        else:
            c = min(int(_u01(seed, i, 101) * 64), 63)
            pos = [cl[c][a] + cl[c][3] * nrm[a] for a in range(3)]
        assert np.array_equal(ply["pos"][i], np.array(pos, np.float64).astype(np.float32))
        assert np.array_equal(ply["scale"][i], (math.log(0.006) + 0.6 * np.array(nrm[4:7])).astype(np.float32))
        assert np.array_equal(ply["rot"][i], np.array(nrm[8:12]).astype(np.float32))
        assert ply["opacity"][i] == np.float32(0.5 + 2.0 * nrm[12])
        assert ply["f_rest"][i][16] == np.float32((0.15 / 1.0) * nrm[18 + 16])   # channel G, k = 1, band 1
        assert ply["f_rest"][i][44] == np.float32((0.15 / 3.0) * nrm[18 + 44])   # channel B, k = 14, band 3


def test_synth_scene_is_counter_based(O):
    whole = O.synth_scene(0xB2000002, 5000)
    part = O.synth_scene(0xB2000002, 1000, start=3000)
    assert whole[3000:4000].tobytes() == part.tobytes()
    assert O.synth_scene(0xB2000003, 10).tobytes() != whole[:10].tobytes()


# ------------------------------------------------------------------ camera
def test_camera_matches_glam_conventions(O):
    # look_at_rh: camera at +z looking at the origin sees -z forward, x right, y up
    v = O.look_at_rh([0, 0, 5]).reshape(4, 4).T        # row-major
    assert np.allclose(v @ np.array([0, 0, 0, 1]), [0, 0, -5, 1])
    assert np.allclose(v @ np.array([1, 2, 0, 1]), [1, 2, -5, 1])
    p = O.perspective_rh(np.float32(np.deg2rad(60)), np.float32(16 / 9), 0.1, 1e4).reshape(4, 4).T
    near = p @ np.array([0, 0, -0.1, 1.0])
    far = p @ np.array([0, 0, -1e4, 1.0])
    assert abs(near[2] / near[3]) < 1e-6 and abs(far[2] / far[3] - 1) < 1e-4   # depth 0..1
    assert np.isclose(p[1, 1], 1 / math.tan(math.radians(30)), rtol=1e-6)
    assert np.isclose(p[0, 0], p[1, 1] / (16 / 9), rtol=1e-6)
    assert p[3, 2] == -1


def test_quat_from_euler_zyx(O):
    q = O.quat_from_euler_zyx_deg([0, 0, 90])          # 90 degrees about z
    assert np.allclose(q, [0, 0, math.sin(math.pi / 4), math.cos(math.pi / 4)], atol=1e-6)
    q = O.quat_from_euler_zyx_deg([90, 0, 0])
    assert np.allclose(q, [math.sin(math.pi / 4), 0, 0, math.cos(math.pi / 4)], atol=1e-6)
    # composition order: q = qz * qy * qx
    def qmul(a, b):
        ax, ay, az, aw = a
        bx, by, bz, bw = b
        return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                         aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])
    qx, qy, qz = O.quat_from_euler_zyx_deg([25, 0, 0]), O.quat_from_euler_zyx_deg([0, -40, 0]), O.quat_from_euler_zyx_deg([0, 0, 70])
    assert np.allclose(O.quat_from_euler_zyx_deg([25, -40, 70]), qmul(qmul(qz, qy), qx), atol=1e-6)


# ------------------------------------------------------------------ known-answer renders
def _frame(O, W, H, eye=(0, 0, 5), **kw):
    view = O.look_at_rh(eye)
    proj = O.perspective_rh(np.float32(np.deg2rad(60)), np.float32(W) / np.float32(H), 0.1, 1e4)
    return O.make_frame(view, proj, W, H, **kw)


def _render(O, f, g, sh=0, cov=0, front_to_back=False, **mk):
    m = O.ModelRef(sh, cov, O.pack(sh, cov, g), len(g), **mk)
    idx, keys, spl = O.preprocess(f, m)
    keys, idx, spl = O.sort(keys, idx, spl)
    img, _ = O.composite(f, spl, front_to_back)
    return img, idx, keys, spl


def test_single_isotropic_splat_closed_form(O):
    """One opaque isotropic Gaussian on the optical axis: the image has a closed form."""
    W = H = 65
    s, D, o8 = 0.2, 5.0, 230
    g = make_gaussians(O.GAUSSIAN, [[0, 0, 0]], scale=s, color=(255, 128, 0, o8))
    f = _frame(O, W, H)
    img, idx, keys, spl = _render(O, f, g)
    assert len(idx) == 1
    fy = (1 / math.tan(math.radians(30))) * H / 2
    var = (fy * s / D) ** 2 + 0.3
    assert np.isclose(spl["mx"][0], (W - 1) / 2, atol=1e-4) and np.isclose(spl["my"][0], (H - 1) / 2, atol=1e-4)
    assert np.isclose(spl["ca"][0], 1 / var, rtol=1e-4) and abs(spl["cb"][0]) < 1e-6
    assert spl["radius"][0] == math.ceil(3 * math.sqrt(var + math.sqrt(0.1)))
    # depth key = bits(ndc.z) for perspective_rh depth 0..1
    r = 1e4 / (0.1 - 1e4)
    ndcz = (r * (-D) + r * 0.1) / D
    assert np.isclose(keys.view(np.float32)[0], ndcz, rtol=1e-6)
    yy, xx = np.mgrid[0:H, 0:W]
    r2 = (xx - (W - 1) / 2) ** 2 + (yy - (H - 1) / 2) ** 2
    alpha = np.minimum(0.99, (o8 / 255) * np.exp(-0.5 * r2 / var))
    rad = int(spl["radius"][0])
    alpha[(np.abs(xx - (W - 1) / 2) > rad) | (np.abs(yy - (H - 1) / 2) > rad) | (alpha < 1 / 255)] = 0
    col = np.array([1.0, 128 / 255, 0.0])
    ref = np.concatenate([alpha[..., None] * col, alpha[..., None]], -1)
    assert np.max(np.abs(img.astype(np.float64) / 255 - ref)) <= 1.01 / 255
    assert img[0, 0].tolist() == [0, 0, 0, 0]          # transparent black clear colour


def test_two_splat_overlap_both_depth_orders(O):
    W = H = 33
    f = _frame(O, W, H)
    for near_first in (True, False):
        z_red, z_green = (1.0, -1.0) if near_first else (-1.0, 1.0)   # camera at z = +5: larger z is nearer
        g = make_gaussians(O.GAUSSIAN, [[0, 0, z_red], [0, 0, z_green]], scale=0.5, color=(255, 0, 0, 200))
        g["color"][1] = (0, 255, 0, 200)
        img, idx, keys, spl = _render(O, f, g)
        assert list(idx) == ([0, 1] if near_first else [1, 0])         # ascending key = near -> far
        c = img[H // 2, W // 2].astype(float) / 255
        a = min(0.99, 200 / 255)
        front, back = a, a * (1 - a)
        exp = [front, back] if near_first else [back, front]
        assert abs(c[0] - exp[0]) < 2 / 255 and abs(c[1] - exp[1]) < 2 / 255
        assert abs(c[3] - (1 - (1 - a) ** 2)) < 2 / 255
        img2, *_ = _render(O, f, g, front_to_back=True)
        assert np.max(np.abs(img.astype(int) - img2.astype(int))) <= 1


def test_frustum_edge_straddlers(O):
    """Cull rule (§8c.5): visible iff 0 < ndc.z < 1 and |ndc.x|,|ndc.y| <= 1.3."""
    W, H = 64, 48
    f = _frame(O, W, H)
    D = 5.0
    tx = math.tan(math.radians(30)) * (W / H)    # |x| / depth at ndc.x = 1
    pts, expect = [], []
    for k, e in [(1.29, True), (1.31, False), (-1.29, True), (-1.31, False), (0.999, True), (1.0, True)]:
        pts.append([k * tx * D, 0, 0]); expect.append(e)
    ty = math.tan(math.radians(30))
    for k, e in [(1.29, True), (1.31, False), (-1.31, False)]:
        pts.append([0, k * ty * D, 0]); expect.append(e)
    for z, e in [(5.0 - 0.0999, False), (5.0 - 0.1001, True), (5.0 + 1.0, False), (5.0 - 9000.0, True), (5.0 - 10001.0, False)]:
        pts.append([0, 0, z]); expect.append(e)
    g = make_gaussians(O.GAUSSIAN, pts, scale=0.01)
    m = O.ModelRef(0, 0, O.pack(0, 0, g), len(g))
    idx, keys, spl = O.preprocess(f, m)
    got = np.zeros(len(pts), bool)
    got[idx] = True
    assert got.tolist() == expect
    assert np.all(np.diff(idx.astype(np.int64)) > 0)      # compaction keeps ascending index order


def test_mask_hidden_and_selection(O):
    W = H = 32
    f = _frame(O, W, H)
    g = make_gaussians(O.GAUSSIAN, [[-0.5, 0, 0], [0, 0, 0], [0.5, 0, 0], [0, 0.5, 0]], scale=0.1)
    packed = O.pack(0, 0, g)
    mask = pack_bits([True, False, True, True])
    edits = np.zeros(4, O.EDIT)
    edits["color"] = (0, 1, 1); edits["gamma"] = 1; edits["alpha"] = 1
    edits["flag"][2] = 1 | 2                               # ENABLED | HIDDEN
    edits["flag"][3] = 2                                   # HIDDEN without ENABLED: ignored
    idx, _, _ = O.preprocess(f, O.ModelRef(0, 0, packed, 4, mask=mask, edits=edits))
    assert idx.tolist() == [0, 3]
    # selection + a hiding selection edit removes the selected Gaussian; a highlight recolours it
    sel = pack_bits([True, False, False, False])
    f2 = _frame(O, W, H, selection_edit=O.edit_pod(flag=1 | 2))
    idx, _, _ = O.preprocess(f2, O.ModelRef(0, 0, packed, 4, selection=sel))
    assert idx.tolist() == [1, 2, 3]
    f3 = _frame(O, W, H, highlight=(1.0, 0.0, 1.0, 0.5))
    idx, _, spl = O.preprocess(f3, O.ModelRef(0, 0, packed, 4, selection=sel))
    assert spl["flags"].tolist() == [1, 0, 0, 0]
    assert np.allclose([spl["r_h"][0], spl["g_h"][0], spl["b_h"][0]], [1.0, 0.5, 1.0], atol=2e-3)
    assert np.allclose([spl["r_h"][1], spl["g_h"][1], spl["b_h"][1]], [1.0, 1.0, 1.0], atol=2e-3)


def _sh_basis_f64(d):
    x, y, z = d
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    C1 = 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]
    return [-C1 * y, C1 * z, -C1 * x,
            C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz, C2[4] * (xx - yy),
            C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
            C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
            C3[6] * x * (xx - 3 * yy)]


def test_sh_probe_one_coefficient_at_a_time(O):
    """Inria computeColorFromSH sign/constant pattern: one coefficient x 6 axis view directions
    (+ one oblique), all SH degrees."""
    W = H = 16
    eyes = [(3, 0, 0), (-3, 0, 0), (0, 3, 0.001), (0, -3, 0.001), (0, 0, 3), (0, 0, -3), (1.5, 2.0, -2.5)]
    for eye in eyes:
        view = O.look_at_rh(eye)
        proj = O.perspective_rh(np.float32(1.0), np.float32(1.0), 0.1, 100.0)
        d = -np.array(eye, np.float64) / np.linalg.norm(eye)        # camera -> Gaussian at the origin
        basis = _sh_basis_f64(d)
        for deg in (0, 1, 2, 3):
            f = O.make_frame(view, proj, W, H, sh_deg=deg)
            g = make_gaussians(O.GAUSSIAN, [[0, 0, 0]] * 15, scale=0.05, color=(102, 102, 102, 255))
            for k in range(15):
                g["sh"][k, 3 * k + 1] = 0.35                        # green channel only
            _, _, spl = O.preprocess(f, O.ModelRef(0, 0, O.pack(0, 0, g), 15))
            ncoef = [0, 3, 8, 15][deg]
            for k in range(15):
                exp_g = 0.4 + (0.35 * basis[k] if k < ncoef else 0.0)
                assert abs(float(spl["g_h"][k]) - min(max(exp_g, 0), 1)) < 1.5e-3, (eye, deg, k)
                assert abs(float(spl["r_h"][k]) - 0.4) < 1e-3
    # no_sh0 drops the baked SH0 colour (reference src/tab/transform.rs:142-145)
    f = O.make_frame(O.look_at_rh((0, 0, 3)), O.perspective_rh(np.float32(1.0), np.float32(1.0), 0.1, 100.0), W, H, no_sh0=1)
    g = make_gaussians(O.GAUSSIAN, [[0, 0, 0]], color=(200, 200, 200, 255))
    _, _, spl = O.preprocess(f, O.ModelRef(0, 0, O.pack(0, 0, g), 1))
    assert float(spl["r_h"][0]) == 0.0


def test_model_transform_and_cov_projection(O):
    """world = q*(s⊙p)+t (reference src/app.rs:1044-1046) and Σ' = (R S)Σ(R S)^T·size²."""
    W = H = 64
    f = _frame(O, W, H, gaussian_size=1.5)
    g = make_gaussians(O.GAUSSIAN, [[1.0, 0.0, 0.0]], scale=(0.3, 0.1, 0.1))
    q = O.quat_from_euler_zyx_deg([0, 0, 90])                # x axis -> y axis
    m = O.ModelRef(0, 0, O.pack(0, 0, g), 1, pos=(0.0, 0.5, 0.0), quat=q, scale=(2.0, 1.0, 1.0))
    idx, keys, spl = O.preprocess(f, m)
    fy = (1 / math.tan(math.radians(30))) * H / 2
    # world position = Rz90 * (2,0,0) + (0,0.5,0) = (0, 2.5, 0); depth 5
    assert np.isclose(spl["mx"][0], (W - 1) / 2, atol=1e-3)
    assert np.isclose(spl["my"][0], (H - 1) / 2 - fy * 2.5 / 5, atol=1e-3)
    # long axis (0.3 * model scale 2 * size 1.5) now points along screen y (ignoring the tiny
    # off-axis Jacobian term in x); short axis 0.1 * 1.5 along screen x
    var_x = (fy * 0.1 * 1.5 / 5) ** 2 + 0.3
    assert np.isclose(1 / spl["ca"][0], var_x, rtol=2e-3)
    assert 1 / spl["cc"][0] > (fy * 0.3 * 2 * 1.5 / 5) ** 2      # plus perspective stretch


def test_model_rank_layering(O):
    """Models are drawn whole, farthest CENTRE first (reference src/tab/scene.rs:533-558): a
    splat of the 'far' model that is actually nearer than a splat of the 'near' model still ends up
    underneath it."""
    W = H = 33
    f = _frame(O, W, H)
    red = make_gaussians(O.GAUSSIAN, [[0, 0, 2.0]], scale=0.5, color=(255, 0, 0, 220))     # nearer splat
    green = make_gaussians(O.GAUSSIAN, [[0, 0, 0.0]], scale=0.5, color=(0, 255, 0, 220))
    m_red = O.ModelRef(0, 0, O.pack(0, 0, red), 1)
    m_green = O.ModelRef(0, 0, O.pack(0, 0, green), 1)
    # centres: red model's centre far away, green model's centre near
    order = O.order_models(f, [m_red, m_green], [[0, 0, -50.0], [0, 0, 0.0]])
    assert order.tolist() == [0, 1]                        # red model is "farther" -> drawn first
    img, v, _ = O.render_frame(f, [m_red, m_green], front_to_back=False)
    a = min(0.99, 220 / 255)
    c = img[H // 2, W // 2].astype(float) / 255
    assert abs(c[1] - a) < 2 / 255 and abs(c[0] - a * (1 - a)) < 2 / 255    # green on top of red
    img2, _, _ = O.render_frame(f, [m_red, m_green], front_to_back=True)
    assert np.max(np.abs(img.astype(int) - img2.astype(int))) <= 1


def test_edit_maths(O):
    # default pod is a no-op (reference src/tab/scene.rs:821, 848)
    rgb, op = O.apply_edit(O.default_edit(), (0.2, 0.4, 0.6), 0.5)
    assert np.allclose(rgb, (0.2, 0.4, 0.6)) and op == 0.5
    # enabled identity edit (hsv (0,1,1), contrast 0, exposure 0, gamma 1, alpha 1) is a no-op too
    rgb, op = O.apply_edit(O.edit_pod(flag=1), (0.2, 0.4, 0.6), 0.5)
    assert np.allclose(rgb, (0.2, 0.4, 0.6), atol=1e-6) and op == 0.5
    rgb, _ = O.apply_edit(O.edit_pod(flag=1 | 4, color=(0.9, 0.1, 0.3)), (0.2, 0.4, 0.6), 0.5)
    assert np.allclose(rgb, (0.9, 0.1, 0.3), atol=1e-6)
    rgb, _ = O.apply_edit(O.edit_pod(flag=1, color=(0.5, 1.0, 1.0)), (1.0, 0.0, 0.0), 1.0)   # hue +180°: red -> cyan
    assert np.allclose(rgb, (0.0, 1.0, 1.0), atol=1e-5)
    rgb, _ = O.apply_edit(O.edit_pod(flag=1, color=(0.0, 0.0, 1.0)), (0.8, 0.2, 0.4), 1.0)   # saturation 0: grey = V
    assert np.allclose(rgb, (0.8, 0.8, 0.8), atol=1e-6)
    rgb, op = O.apply_edit(O.edit_pod(flag=1, exposure=1.0, alpha=0.5), (0.1, 0.2, 0.3), 0.8)
    assert np.allclose(rgb, (0.2, 0.4, 0.6), atol=1e-6) and abs(op - 0.4) < 1e-7
    rgb, _ = O.apply_edit(O.edit_pod(flag=1, contrast=0.5), (0.5, 0.7, 0.1), 1.0)
    assert np.allclose(rgb, (0.5, 0.8, -0.1 if False else 0.0), atol=1e-6) or np.allclose(rgb, (0.5, 0.8, 0.0), atol=1e-6)
    rgb, _ = O.apply_edit(O.edit_pod(flag=1, gamma=2.0), (0.5, 0.25, 1.0), 1.0)
    assert np.allclose(rgb, (0.25, 0.0625, 1.0), atol=1e-6)


def test_mask_box_union_ellipsoid_minus_box_on_lattice(O):
    """`0 | 1 - 2` with the app's precedence (- binds tighter than |, src/app.rs:1660-1783):
    box0 ∪ (ellipsoid1 − box2), evaluated on a lattice against a float64 restatement."""
    ax = np.linspace(-2, 2, 21, dtype=np.float32)
    pts = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3) + np.float32(0.013)
    g = make_gaussians(O.GAUSSIAN, pts, scale=0.01)
    q = O.quat_from_euler_zyx_deg([0, 0, 30])
    shapes = np.zeros(3, O.MASK_SHAPE)
    shapes["kind"] = [0, 1, 0]
    shapes["pos"] = [(-1, 0, 0), (0.8, 0.2, 0), (0.8, 0.2, 0)]
    shapes["quat"] = [q, (0, 0, 0, 1), (0, 0, 0, 1)]
    shapes["scale"] = [(1.5, 1.0, 2.0), (2.0, 3.0, 1.6), (0.8, 0.8, 4.0)]
    ops = np.array([(0, 0), (0, 1), (0, 2), (3, 0), (1, 0)], O.MASK_OP)      # 0 1 2 - |
    m = O.ModelRef(0, 0, O.pack(0, 0, g), len(g))
    words = O.eval_mask(m, ops, shapes)
    got = bits_set(words, len(g))
    p = pts.astype(np.float64)
    c, s = math.cos(math.radians(30)), math.sin(math.radians(30))
    Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    l0 = ((p - shapes["pos"][0]) @ Rz) / shapes["scale"][0]
    in0 = np.all(np.abs(l0) <= 0.5, -1)
    l1 = (p - shapes["pos"][1]) / shapes["scale"][1]
    in1 = (l1 ** 2).sum(-1) <= 0.25
    l2 = (p - shapes["pos"][2]) / shapes["scale"][2]
    in2 = np.all(np.abs(l2) <= 0.5, -1)
    ref = in0 | (in1 & ~in2)
    assert 0 < ref.sum() < len(ref)
    assert (got != ref).sum() == 0
    # Reset = everything shown (reference src/tab/scene.rs:2124-2131)
    allw = O.eval_mask(m, np.array([(6, 0)], O.MASK_OP), shapes[:0])
    assert bits_set(allw, len(g)).all()
    # complement / intersection / symmetric difference
    w = O.eval_mask(m, np.array([(0, 0), (5, 0)], O.MASK_OP), shapes)
    assert np.array_equal(bits_set(w, len(g)), ~in0)
    w = O.eval_mask(m, np.array([(0, 1), (0, 2), (2, 0)], O.MASK_OP), shapes)
    assert np.array_equal(bits_set(w, len(g)), in1 & in2)
    w = O.eval_mask(m, np.array([(0, 1), (0, 2), (4, 0)], O.MASK_OP), shapes)
    assert np.array_equal(bits_set(w, len(g)), in1 ^ in2)


def test_sort_is_stable_and_matches_numpy(O):
    rng = np.random.default_rng(7)
    for n in (0, 1, 17, 4096, 100003):
        keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
        keys[: n // 3] &= 0xFF            # many ties
        vals = np.arange(n, dtype=np.uint32)
        k2, v2 = O.sort_pairs(keys, vals)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k2, keys[order]) and np.array_equal(v2, vals[order])
    keys = rng.integers(0, 1 << 32, 5000, dtype=np.uint64).astype(np.uint32)
    k2, v2 = O.sort_pairs(keys, np.arange(5000, dtype=np.uint32), bits=16)
    order = np.argsort(keys & 0xFFFF, kind="stable")
    assert np.array_equal(v2, order.astype(np.uint32))


def test_display_modes(O):
    W = H = 41
    g = make_gaussians(O.GAUSSIAN, [[0, 0, 0]], scale=(0.4, 0.15, 0.15), color=(255, 255, 255, 180))
    img_s, *_ = _render(O, _frame(O, W, H, display_mode=0), g)
    img_e, *_ = _render(O, _frame(O, W, H, display_mode=1), g)
    img_p, *_ = _render(O, _frame(O, W, H, display_mode=2), g)
    a = round(min(0.99, 180 / 255) * 255)
    assert img_e[H // 2, W // 2, 3] == a and set(np.unique(img_e[..., 3])) == {0, a}     # flat ellipse
    assert (img_e[..., 3] > 0).sum() < (img_s[..., 3] > 0).sum()
    assert (img_p[..., 3] > 0).sum() == 9                                                # 3x3 point


# ------------------------------------------------------------------ committed fixtures
def test_golden_fixture(O):
    """tests/golden/oracle_small.npz was written by tests/golden/make_golden.py from this oracle;
    it pins the oracle's outputs against accidental drift (the analytic tests above pin meaning)."""
    path = os.path.join(GOLDEN, "oracle_small.npz")
    z = np.load(path)
    n, W, H = int(z["n"]), int(z["W"]), int(z["H"])
    ply = O.synth_scene(int(z["seed"]), n)
    g = O.gaussian_from_ply(ply)
    for sh, cov in ((2, 1), (1, 0), (0, 0), (3, 1)):
        packed = O.pack(sh, cov, g)
        view, proj = O.orbit_camera(width=W, height=H)
        f = O.make_frame(view, proj, W, H)
        idx, keys, spl = O.preprocess(f, O.ModelRef(sh, cov, packed, n))
        keys, idx, spl = O.sort(keys, idx, spl)
        img, _ = O.composite(f, spl, False)
        tag = "%d%d" % (sh, cov)
        assert np.array_equal(idx, z["idx_" + tag])
        assert np.array_equal(keys, z["keys_" + tag])
        assert np.array_equal(img, z["img_" + tag])
        assert psnr(img, O.composite(f, spl, True)[0]) > 50


def test_selection_query_rect_and_brush(O):
    """Immediate-mode selection queries (reference src/tab/scene.rs:758-791, 1224-1263; ops src/app.rs:1453):
    a Gaussian is hit when its projected centre lies in the shape; Set / Add / Remove on the bitset."""
    import math
    W = H = 64
    D = 5.0
    fy = (1 / math.tan(math.radians(30))) * H / 2
    # a 5x5 lattice whose centres project to pixel centres x = 32 + 8*i (i = -2..2), same for y
    step = 8.0 * D / fy
    pts = [[i * step, -j * step, 0.0] for j in range(-2, 3) for i in range(-2, 3)]
    g = make_gaussians(O.GAUSSIAN, pts, scale=0.01)
    packed = O.pack(0, 0, g)
    view = O.look_at_rh((0, 0, D))
    proj = O.perspective_rh(np.float32(np.deg2rad(60)), np.float32(1.0), 0.1, 1e4)
    sx = np.array([32 + 8 * i for j in range(-2, 3) for i in range(-2, 3)], float)
    sy = np.array([32 + 8 * j for j in range(-2, 3) for i in range(-2, 3)], float)

    f = O.make_frame(view, proj, W, H, query=O.query_pod(2, 0, (20, 20), (44, 36)))          # rect, Set
    sel = bits_set(O.query_selection(f, O.ModelRef(0, 0, packed, 25)), 25)
    assert np.array_equal(sel, (sx >= 20) & (sx <= 44) & (sy >= 20) & (sy <= 36)) and sel.sum() == 6
    old = pack_bits(np.arange(25) % 2 == 0)
    f = O.make_frame(view, proj, W, H, query=O.query_pod(2, 1, (20, 20), (44, 36)))          # Add
    got = bits_set(O.query_selection(f, O.ModelRef(0, 0, packed, 25, selection=old)), 25)
    assert np.array_equal(got, sel | (np.arange(25) % 2 == 0))
    f = O.make_frame(view, proj, W, H, query=O.query_pod(2, 2, (20, 20), (44, 36)))          # Remove
    got = bits_set(O.query_selection(f, O.ModelRef(0, 0, packed, 25, selection=old)), 25)
    assert np.array_equal(got, ~sel & (np.arange(25) % 2 == 0))
    # brush: capsule of radius 9 around the diagonal segment (16,16)-(48,48)
    f = O.make_frame(view, proj, W, H, query=O.query_pod(3, 0, (16, 16), (48, 48), 9.0))
    got = bits_set(O.query_selection(f, O.ModelRef(0, 0, packed, 25)), 25)
    t = np.clip(((sx - 16) * 32 + (sy - 16) * 32) / (2 * 32 * 32), 0, 1)
    d2 = (sx - 16 - 32 * t) ** 2 + (sy - 16 - 32 * t) ** 2
    assert np.array_equal(got, d2 <= 81) and 5 <= got.sum() <= 15
    # the frame shows the NEW selection: highlighted splats carry the flag
    f = O.make_frame(view, proj, W, H, query=O.query_pod(2, 0, (20, 20), (44, 36)), highlight=(1, 0, 1, 0.5))
    idx, _, spl = O.preprocess(f, O.ModelRef(0, 0, packed, 25))
    assert np.array_equal(spl["flags"].astype(bool), sel[idx])


# ---------------------------------------------------------------------------------------------- cross-checks
def test_fp32_splat_path_agrees_with_the_f16_record_path(O):
    """The oracle's own fp32 splats (orc_preprocess_f32: nothing rounded to the product's 32-byte record) against
    its f16-record path on BASELINE config 1: same visible set / order, images within 1/255 — i.e. the record's
    f16 colour and opacity cost at most one 8-bit step."""
    n, W, H = 100_000, 1280, 720
    packed = O.pack(2, 1, O.gaussian_from_ply(O.synth_scene(0xB2000001, n)))
    view, proj = O.orbit_camera(width=W, height=H)
    f = O.make_frame(view, proj, W, H)
    m = O.ModelRef(2, 1, packed, n)
    i1, k1, s1 = O.preprocess(f, m)
    i2, k2, s2 = O.preprocess_f32(f, m)
    assert np.array_equal(i1, i2) and np.array_equal(k1, k2)
    for a, b in (("mx", "mx"), ("my", "my"), ("ca", "ca"), ("cb", "cb"), ("cc", "cc")):
        assert np.array_equal(s1[a], s2[b])
    assert np.array_equal(s1["radius"].astype(np.float32), s2["radius"])
    assert np.max(np.abs(s1["r_h"].astype(np.float32) - s2["r"])) <= 2.0 ** -11
    k1s, i1s, s1s = O.sort(k1, i1, s1)
    k2s, i2s, s2s = O.sort(k2, i2, s2)
    assert np.array_equal(i1s, i2s)
    a, _ = O.composite(f, s1s)
    b, _ = O.composite(f, s2s)
    assert np.abs(a.astype(np.int32) - b.astype(np.int32)).max() <= 1
    img, v, _ = O.render_frame(f, [m], fp32=True)
    assert v == len(i2) and np.array_equal(img, b)


def test_numpy_float64_restatement_agrees_on_config1(O):
    """SURVEY.md §7 step 2's "slower NumPy cross-check": tests/numpy_ref.py restates §8(c) in float64 NumPy without
    looking at the C code.  On BASELINE config 1 (100k @ 1280x720): the visible set is identical; the oracle's f32
    depths are the float64 depths to within a few f32 ulps and its sorted order is a sort of the float64 depths up
    to that rounding; blended in that order, the float64 image is within 1/255 of the oracle's."""
    import numpy_ref as NR
    n, W, H = 100_000, 1280, 720
    packed = O.pack(2, 1, O.gaussian_from_ply(O.synth_scene(0xB2000001, n)))
    view, proj = O.orbit_camera(width=W, height=H)
    f = O.make_frame(view, proj, W, H)
    m = O.ModelRef(2, 1, packed, n)
    oi, ok, osp = O.preprocess_f32(f, m)
    ok2, oi2, osp2 = O.sort(ok, oi, osp)
    ref, ref_f, _ = O.composite(f, osp2, want_float=True)
    idx, z, order, img = NR.render(packed, n, view, proj, W, H, near_to_far=oi2)
    assert np.array_equal(idx, oi)                                         # visible set (and compaction order)
    ulp = float(np.spacing(np.float32(0.9)))
    assert np.abs(ok.view(np.float32).astype(np.float64) - z).max() <= 8 * ulp      # depth keys
    zs = z[np.searchsorted(idx, oi2)]
    assert np.maximum(0.0, zs[:-1] - zs[1:]).max() <= 8 * ulp              # the oracle's order sorts the f64 depths
    own = NR.render(packed, n, view, proj, W, H)[2]
    assert np.mean(own == oi2) > 0.8                                       # (free-running f64 order: near-ties swap)
    q = np.floor(np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.int32)
    assert np.abs(q - ref.astype(np.int32)).max() <= 1
    assert np.abs(img - ref_f).max() <= 1.5 / 255.0   # (1/255 + an alpha-cut flip between f32 and f64)


def test_numpy_float64_restatement_with_model_transform(O):
    """Same cross-check under a TRS model transform (BASELINE config 4's second transform) on a 20k scene."""
    import numpy_ref as NR
    n, W, H = 20_000, 640, 360
    packed = O.pack(2, 1, O.gaussian_from_ply(O.synth_scene(0xB2000042, n)))
    view, proj = O.orbit_camera(width=W, height=H)
    f = O.make_frame(view, proj, W, H, gaussian_size=1.3, sh_deg=2)
    quat = O.quat_from_euler_zyx_deg((10.0, 0.0, 45.0))
    m = O.ModelRef(2, 1, packed, n, pos=(0.3, -0.2, 0.1), quat=quat, scale=(1.2, 0.8, 1.0))
    oi, ok, osp = O.preprocess_f32(f, m)
    ok2, oi2, osp2 = O.sort(ok, oi, osp)
    ref, ref_f, _ = O.composite(f, osp2, want_float=True)
    idx, z, order, img = NR.render(packed, n, view, proj, W, H, sh_deg=2, size=1.3, model_pos=(0.3, -0.2, 0.1),
                                   model_quat=quat, model_scale=(1.2, 0.8, 1.0), near_to_far=oi2)
    assert np.array_equal(idx, oi)
    # (a fragment sitting exactly on the alpha >= 1/255 cut can flip in or out between f32 and f64: one cut step)
    assert np.abs(img - ref_f).max() <= 1.5 / 255.0
    q = np.floor(np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.int32)
    assert np.abs(q - ref.astype(np.int32)).max() <= 1
