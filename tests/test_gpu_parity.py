"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against
the CPU oracle on the same seeded inputs.  Bars (BASELINE.json north_star):
  * visible count, compaction order, depth keys, sorted index order: BIT-EXACT
  * RGBA: max |Δ| <= 2/255 per channel and PSNR >= 50 dB (tests/util.py)."""
import os

import numpy as np
import pytest

from util import (SEED_100K, SEED_1M, SEED_6M, SEEDS_CFG4, assert_image_close, bits_set, make_gaussians, pack_bits, psnr)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SPLAT_FIELDS = ("mx", "my", "radius", "ca", "cb", "cc", "opacity_h", "r_h", "g_h", "b_h", "flags")


def _splats_equal(a, b):
    """EXACT class (pixel centre, extent, conic: same float ops in the same order on both sides) must
    be bit-identical; TOLERANCE class (SH colour and opacity, stored as f16; the kernel uses FMA and
    rsqrt there) within 2e-3, far inside the 2/255 image budget."""
    assert len(a) == len(b)
    for f in ("mx", "my", "radius", "ca", "cb", "cc", "flags"):
        x, y = np.ascontiguousarray(a[f]), np.ascontiguousarray(b[f])
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), "splat field %s differs" % f
    for f in ("opacity_h", "r_h", "g_h", "b_h"):
        assert np.max(np.abs(a[f].astype(np.float32) - b[f].astype(np.float32)), initial=0) <= 2e-3, f


def _splats_close(a, b, bounds_exact=True):
    """colour / conic / opacity are floating point: tolerance; bounds inputs stay exact"""
    assert len(a) == len(b)
    if bounds_exact:
        for f in ("mx", "my", "radius"):
            assert np.array_equal(np.ascontiguousarray(a[f]).view(np.uint8), np.ascontiguousarray(b[f]).view(np.uint8)), f
    for f in ("ca", "cb", "cc"):
        assert np.allclose(a[f], b[f], rtol=1e-5, atol=1e-7), f
    for f in ("opacity_h", "r_h", "g_h", "b_h"):
        assert np.max(np.abs(a[f].astype(np.float32) - b[f].astype(np.float32)), initial=0) <= 2e-3, f


def _full_parity(G, O, n, seed, W, H, sh=2, cov=1, cam=None, check_image=True, **frame_kw):
    ply = G.synth_scene(seed, n)
    g = G.gaussian_from_ply(ply)
    packed = G.pack_gaussians(sh, cov, g)
    cam = cam or G.OrbitCamera.orbit()
    view, proj = cam.view(), cam.projection(np.float32(W) / np.float32(H))
    f = O.make_frame(view, proj, W, H, **frame_kw)
    om = O.ModelRef(sh, cov, packed, n)
    oi, ok, osp = O.preprocess(f, om)
    with G.Viewer(W, H, sh, cov) as v:
        m = v.add_model("scene", n)
        m.upload_packed(0, packed)
        v.update_camera(cam)
        v.update_gaussian_transform(frame_kw.get("gaussian_size", 1.0), frame_kw.get("display_mode", 0),
                                    frame_kw.get("sh_deg", 3), bool(frame_kw.get("no_sh0", 0)))
        if "background" in frame_kw:
            v.set_background(frame_kw["background"])
        m.preprocess()
        assert m.visible_count() == len(oi)                       # visible-splat count: exact
        assert np.array_equal(m.indices(), oi)                    # order-preserving compaction
        assert np.array_equal(m.depth_keys(), ok)                 # depth keys: bit-exact
        _splats_equal(m.splats(), osp)
        m.sort()
        ok2, oi2, osp2 = O.sort(ok, oi, osp)
        assert np.array_equal(m.depth_keys(), ok2)
        assert np.array_equal(m.indices(), oi2)                   # sorted index order: bit-exact
        _splats_equal(m.splats(), osp2)
        img = v.render_frame_host([m]).copy()
        if check_image:
            ref, _ = O.composite(f, osp2, front_to_back=False)    # reference-style back-to-front
            assert_image_close(img, ref)
            # ... and against the oracle's OWN fp32 splats (colour / opacity never rounded to the product's f16
            # record), so that the 2/255 / 50 dB bar includes the product's quantisation
            fi, fk, fsp = O.preprocess_f32(f, om)
            fk2, fi2, fsp2 = O.sort(fk, fi, fsp)
            assert np.array_equal(fi2, oi2)
            ref32, _ = O.composite(f, fsp2, front_to_back=False)
            assert_image_close(img, ref32)
        t = v.last_timings()
        assert t.overflow == 0
    return img


def test_config1_100k_720p_all_stages(G, O):
    """BASELINE.json configs[0]: 100k-Gaussian SH3 scene at 1280x720 (Norm8 SH + Half Cov3d)."""
    _full_parity(G, O, 100_000, SEED_100K, 1280, 720)


def test_config2_1m_1080p_all_stages(G, O):
    """BASELINE.json configs[1]: 1M-Gaussian SH3 scene, single view 1920x1080."""
    _full_parity(G, O, 1_000_000, SEED_1M, 1920, 1080)


def test_config3_6m_1080p_all_stages(G, O):
    """BASELINE.json configs[2] at full size: 6M Gaussians at 1920x1080 — every stage still compared
    with the oracle (it finishes in seconds on the host cores)."""
    _full_parity(G, O, 6_000_000, SEED_6M, 1920, 1080)


def test_config3_6m_4k_full_size(G, O):
    """BASELINE.json configs[2], 3840x2160 leg at FULL size: 6M Gaussians, every stage and the image."""
    _full_parity(G, O, 6_000_000, SEED_6M, 3840, 2160)


def test_config3_4k_image(G, O):
    """BASELINE.json configs[2], 3840x2160 leg, on a 1M subset so that the oracle image is quick."""
    _full_parity(G, O, 1_000_000, SEED_6M, 3840, 2160)


@pytest.mark.parametrize("sh,cov", [(s, c) for s in range(4) for c in range(2)])
def test_all_eight_layouts(G, O, sh, cov):
    """The 8 GaussianPod layouts of reference src/app.rs:250-257."""
    _full_parity(G, O, 20_011, 0xB2000077, 640, 360, sh=sh, cov=cov)


@pytest.mark.parametrize("kw", [dict(sh_deg=0), dict(sh_deg=1), dict(sh_deg=2), dict(no_sh0=1), dict(gaussian_size=0.5),
                                dict(gaussian_size=2.0), dict(display_mode=1), dict(display_mode=2),
                                dict(background=(0.2, 0.4, 0.6, 1.0))])
def test_gaussian_transform_settings(G, O, kw):
    """update_gaussian_transform(size, display_mode, sh_deg, no_sh0) — reference scene.rs:803-809;
    UI ranges src/tab/transform.rs:110-145."""
    _full_parity(G, O, 30_000, 0xB2000078, 800, 450, **kw)


def test_other_cameras_and_odd_viewport(G, O):
    for cam, (W, H) in [(G.OrbitCamera.orbit(3.0, -10.0, 200.0), (333, 217)), (G.OrbitCamera.orbit(8.0, 60.0, 90.0), (1000, 16)),
                        (G.OrbitCamera.orbit(0.7, 5.0, 10.0), (640, 480))]:
        _full_parity(G, O, 50_000, SEED_100K, W, H, cam=cam)


def test_golden_fixture_matches_gpu(G):
    """Committed oracle bytes (tests/golden/oracle_small.npz) vs the CUDA path — no oracle run."""
    z = np.load(os.path.join(GOLDEN, "oracle_small.npz"))
    n, W, H = int(z["n"]), int(z["W"]), int(z["H"])
    g = G.gaussian_from_ply(G.synth_scene(int(z["seed"]), n))
    for sh, cov in ((2, 1), (1, 0), (0, 0), (3, 1)):
        with G.Viewer(W, H, sh, cov) as v:
            m = v.add_model("g", n)
            m.update_range(0, g)                                   # host packs, then H2D
            v.update_camera(G.OrbitCamera.orbit())
            img = v.render_frame_host([m]).copy()
            tag = "%d%d" % (sh, cov)
            assert np.array_equal(m.indices(), z["idx_" + tag])
            assert np.array_equal(m.depth_keys(), z["keys_" + tag])
            assert_image_close(img, z["img_" + tag])


def test_single_splat_closed_form_on_gpu(G):
    """The analytic known-answer of tests/test_oracle.py, rendered by the CUDA path."""
    import math
    W = H = 65
    s, D, o8 = 0.2, 5.0, 230
    g = make_gaussians(G.GAUSSIAN, [[0, 0, 0]], scale=s, color=(255, 128, 0, o8))
    with G.Viewer(W, H, G.SH_SINGLE, G.COV3D_SINGLE) as v:
        m = v.add_model("one", 1)
        m.update_range(0, g)
        v.update_camera(G.OrbitCamera(pos=(0, 0, D)))
        img = v.render_frame_host([m]).copy()
        assert m.visible_count() == 1
    fy = (1 / math.tan(math.radians(30))) * H / 2
    var = (fy * s / D) ** 2 + 0.3
    yy, xx = np.mgrid[0:H, 0:W]
    r2 = (xx - (W - 1) / 2) ** 2 + (yy - (H - 1) / 2) ** 2
    alpha = np.minimum(0.99, (o8 / 255) * np.exp(-0.5 * r2 / var))
    rad = math.ceil(3 * math.sqrt(var + math.sqrt(0.1)))
    alpha[(np.abs(xx - (W - 1) / 2) > rad) | (np.abs(yy - (H - 1) / 2) > rad) | (alpha < 1 / 255)] = 0
    ref = np.concatenate([alpha[..., None] * np.array([1.0, 128 / 255, 0.0]), alpha[..., None]], -1)
    assert np.max(np.abs(img.astype(np.float64) / 255 - ref)) <= 1.01 / 255


# ------------------------------------------------------------------ config 4: multi-model
def _config4(G, O, n_per_model, W, H):
    """BASELINE.json configs[3] / SURVEY.md §8d: three models with transforms, a colour edit on a
    selected half of model 2, and the composite mask `0 | 1 - 2` on model 1."""
    xf = [((-2, 0, 0), (0, 30, 0), (1, 1, 1)), ((0, 0, 0), (10, 0, 45), (1.2, 0.8, 1)), ((2, 0.5, 0), (0, -60, 0), (0.7, 0.7, 0.7))]
    cam = G.OrbitCamera.orbit()
    view, proj = cam.view(), cam.projection(np.float32(W) / np.float32(H))
    f = O.make_frame(view, proj, W, H)
    shapes = np.zeros(3, G.MASK_SHAPE)
    shapes["kind"] = [G.MASK_BOX, G.MASK_ELLIPSOID, G.MASK_BOX]
    shapes["pos"] = [(-1, 0, 0), (1, 0, 0.5), (1, 0, 0)]
    shapes["quat"] = [G.quat_from_euler_zyx_deg([0, 20, 0]), (0, 0, 0, 1), (0, 0, 0, 1)]
    shapes["scale"] = [(3, 3, 6), (4, 3, 5), (1.5, 4, 1.5)]
    ops = np.array([(0, 0), (0, 1), (0, 2), (3, 0), (1, 0)], G.MASK_OP)         # 0 1 2 - |
    edit = dict(flag=1, color=(0.5, 1.2, 0.9), contrast=0.2, exposure=0.5, gamma=1.2, alpha=0.8)
    v = G.Viewer(W, H)
    gm, om, centers = [], [], []
    for k in range(3):
        g = G.gaussian_from_ply(G.synth_scene(SEEDS_CFG4[k], n_per_model))
        packed = G.pack_gaussians(2, 1, g)
        q = G.quat_from_euler_zyx_deg(xf[k][1])
        m = v.add_model("model%d" % k, n_per_model)
        m.upload_packed(0, packed)
        m.set_transform(xf[k][0], q, xf[k][2])
        kw = {}
        if k == 1:
            m.eval_mask(ops, shapes)
            mask_gpu = m.download_mask()
            ref_model = O.ModelRef(2, 1, packed, n_per_model, pos=xf[k][0], quat=q, scale=xf[k][2])
            mask_ref = O.eval_mask(ref_model, ops.astype(O.MASK_OP), shapes.astype(O.MASK_SHAPE))
            assert np.array_equal(mask_gpu, mask_ref)             # mask bitset: bit-exact
            assert 0 < bits_set(mask_ref, n_per_model).sum() < n_per_model
            kw["mask"] = mask_ref
        if k == 2:
            sel = pack_bits(g["pos"][:, 0] > 0)                    # "rect-selected half"
            edits = np.zeros(n_per_model, G.EDIT)
            edits["color"], edits["gamma"], edits["alpha"] = (0, 1, 1), 1, 1
            chosen = bits_set(sel, n_per_model)
            for key, val in edit.items():
                edits[key][chosen] = val
            m.upload_edits(0, edits)
            kw["edits"] = edits.astype(O.EDIT)
        gm.append(m)
        om.append(O.ModelRef(2, 1, packed, n_per_model, pos=xf[k][0], quat=q, scale=xf[k][2], **kw))
        centers.append(g["pos"].mean(0))
    v.update_camera(cam)
    order_g = v.order_models(gm, np.array(centers, np.float32))
    order_o = O.order_models(f, om, np.array(centers, np.float32))
    assert [gm.index(x) for x in order_g] == order_o.tolist()     # farthest centre first
    far_to_near_o = [om[i] for i in order_o]
    return v, f, order_g, far_to_near_o


def test_config4_three_models_edits_mask(G, O):
    v, f, gms, oms = _config4(G, O, 150_000, 1920, 1080)
    try:
        img = v.render_frame_host(gms).copy()
        for gmod, omod in zip(gms, oms):
            oi, ok, osp = O.preprocess(f, omod)
            ok, oi, osp = O.sort(ok, oi, osp)
            assert np.array_equal(gmod.indices(), oi) and np.array_equal(gmod.depth_keys(), ok)
            _splats_close(gmod.splats(), osp)                      # edited colours: powf/exp2f tolerance
        ref, total, _ = O.render_frame(f, oms, front_to_back=False)
        assert v.last_timings().visible == total
        assert_image_close(img, ref)
        # layering is per model, not a global depth sort: reversing the order changes the image
        img_rev = v.render_frame_host(gms[::-1]).copy()
        assert psnr(img_rev, ref) < 45
    finally:
        v.close()


def test_config4_full_size_properties(G, O):
    """3 x 2M at 1080p: visible counts and sorted order per model exact; image within tolerance."""
    v, f, gms, oms = _config4(G, O, 2_000_000, 1920, 1080)
    try:
        img = v.render_frame_host(gms).copy()
        for gmod, omod in zip(gms, oms):
            oi, ok, _ = O.preprocess(f, omod)
            ok, oi, _ = O.sort(ok, oi)
            assert np.array_equal(gmod.indices(), oi) and np.array_equal(gmod.depth_keys(), ok)
        ref, _, _ = O.render_frame(f, oms, front_to_back=False)
        assert_image_close(img, ref)
        assert np.array_equal(img, v.render_frame_host(gms))       # idempotent: same bytes again
    finally:
        v.close()


def test_selection_highlight_hidden_and_unedited(G, O):
    W, H, n = 640, 360, 40_000
    g = G.gaussian_from_ply(G.synth_scene(0xB2000079, n))
    packed = G.pack_gaussians(2, 1, g)
    sel = pack_bits(g["pos"][:, 1] > 0)
    mask = pack_bits(np.arange(n) % 3 != 0)
    cam = G.OrbitCamera.orbit()
    view, proj = cam.view(), cam.projection(np.float32(W) / np.float32(H))
    with G.Viewer(W, H) as v:
        m = v.add_model("m", n)
        m.upload_packed(0, packed)
        m.upload_selection(sel)
        m.upload_mask(mask)
        v.update_camera(cam)
        assert np.array_equal(m.download_mask(), mask) and np.array_equal(m.download_selection(), sel)
        # highlight only (reference scene.rs:822-829)
        v.update_selection_highlight((1.0, 0.0, 1.0, 0.5))
        f = O.make_frame(view, proj, W, H, highlight=(1.0, 0.0, 1.0, 0.5))
        om = O.ModelRef(2, 1, packed, n, mask=mask, selection=sel)
        img = v.render_frame_host([m]).copy()
        oi, ok, osp = O.preprocess(f, om)
        ok, oi, osp = O.sort(ok, oi, osp)
        assert np.array_equal(m.indices(), oi)
        _splats_close(m.splats(), osp)
        assert_image_close(img, O.composite(f, osp)[0])
        # a live selection edit that hides the selection (GaussianEditFlag::HIDDEN, app.rs:1548-1551)
        pod = G.EditPod.new(G.EDIT_ENABLED | G.EDIT_HIDDEN)
        v.update_selection_edit(pod)
        v.update_selection_highlight((0, 0, 0, 0))
        f2 = O.make_frame(view, proj, W, H, selection_edit=O.edit_pod(flag=3))
        m.preprocess()
        oi2, _, _ = O.preprocess(f2, om)
        assert np.array_equal(m.indices(), oi2) and len(oi2) < len(oi)
        # show_unedited (reference scene.rs:843-849, 858-861): blank edit buffer + default pod
        m.preprocess(use_unedited=True)
        f3 = O.make_frame(view, proj, W, H)
        oi3, _, _ = O.preprocess(f3, om)
        assert np.array_equal(m.indices(), oi3)
        # postprocess commits the selection edit into the per-Gaussian edit buffer (scene.rs:604-610)
        pod2 = G.EditPod.new(G.EDIT_ENABLED | G.EDIT_OVERRIDE_COLOR, color=(0.1, 0.9, 0.2), alpha=0.7)
        v.update_selection_edit(pod2)
        m.postprocess()
        ed = m.download_edits()
        chosen = bits_set(sel, n)
        assert np.all(ed["flag"][chosen] == 5) and np.all(ed["flag"][~chosen] == 0)
        assert np.allclose(ed["color"][chosen], (0.1, 0.9, 0.2)) and np.allclose(ed["alpha"][chosen], 0.7)


# ------------------------------------------------------------------ sort
# (6143 / 6144 / 6145 / 12289: one key short of, exactly, and one past one and two of the sort's 6144-key tiles)
@pytest.mark.parametrize("n", [1, 2, 255, 4095, 4096, 4097, 6143, 6144, 6145, 12289, 100_003, 1 << 20, 5_000_000])
def test_sort_pairs_matches_stable_sort(G, O, n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    vals = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    with G.Viewer(16, 16) as v:
        k, val = v.sort_pairs(keys, vals)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(val, vals[order])
    if n <= 200_000:
        ko, vo = O.sort_pairs(keys, vals)
        assert np.array_equal(k, ko) and np.array_equal(val, vo)


def test_sort_ties_collisions_and_16bit(G):
    rng = np.random.default_rng(5)
    n = 300_000
    with G.Viewer(16, 16) as v:
        for keys in (np.zeros(n, np.uint32), np.full(n, 0xFFFFFFFF, np.uint32),
                     rng.integers(0, 4, n).astype(np.uint32) << 24, rng.integers(0, 7, n).astype(np.uint32),
                     np.arange(n, dtype=np.uint32)[::-1].copy(),
                     np.float32(rng.uniform(0.9, 1.0, n)).view(np.uint32)):      # depth-like keys: shared top bytes
            vals = np.arange(n, dtype=np.uint32)
            k, val = v.sort_pairs(keys, vals)
            order = np.argsort(keys, kind="stable")
            assert np.array_equal(k, keys[order]) and np.array_equal(val, order.astype(np.uint32))   # ties keep input order
        keys = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
        k, val = v.sort_pairs(keys, np.arange(n, dtype=np.uint32), bits=16)
        order = np.argsort(keys & 0xFFFF, kind="stable")
        assert np.array_equal(val, order.astype(np.uint32))


# ------------------------------------------------------------------ edge cases & ABI behaviour
def test_edge_cases(G, O):
    W, H = 320, 200
    cam = G.OrbitCamera.orbit()
    with G.Viewer(W, H) as v:
        v.update_camera(cam)
        # nothing uploaded yet: capacity is rendered as zero records (streaming-upload contract,
        # reference scene.rs:341-380): zero opacity, so a transparent image
        m = v.add_model("empty", 1000)
        img = v.render_frame_host([m]).copy()
        assert not img.any()
        # a model entirely behind the camera: V = 0
        g = make_gaussians(G.GAUSSIAN, np.tile(np.array(cam.pos, np.float32) * 2.0, (513, 1)), scale=0.01)
        m2 = v.add_model("behind", 513)
        m2.update_range(0, g)
        img = v.render_frame_host([m2]).copy()
        assert m2.visible_count() == 0 and not img.any()
        # no models at all: background only
        v.set_background((1.0, 0.5, 0.25, 1.0))
        img = v.render_frame_host([]).copy()
        assert np.all(img == np.array([255, 128, 64, 255], np.uint8))
        v.set_background((0, 0, 0, 0))
        # ragged sizes around the 256-Gaussian chunk
        for n in (1, 255, 256, 257, 1023):
            ply = G.synth_scene(77, n)
            gg = G.gaussian_from_ply(ply)
            packed = G.pack_gaussians(2, 1, gg)
            mm = v.add_model("r%d" % n, n)
            mm.upload_packed(0, packed)
            mm.preprocess()
            f = O.make_frame(cam.view(), cam.projection(np.float32(W) / np.float32(H)), W, H)
            oi, ok, _ = O.preprocess(f, O.ModelRef(2, 1, packed, n))
            assert np.array_equal(mm.indices(), oi) and np.array_equal(mm.depth_keys(), ok)
            v.remove_model("r%d" % n)
        # partial upload: update_range into the middle of the buffer
        n = 5000
        gg = G.gaussian_from_ply(G.synth_scene(78, n))
        mm = v.add_model("partial", n)
        mm.update_range(1000, gg[1000:3000])
        back = mm.download_packed(1000, 2000)
        assert back.tobytes() == G.pack_gaussians(2, 1, gg[1000:3000]).tobytes()


def test_huge_splat_and_tile_capacity_overflow(G, O):
    """A splat covering the whole screen touches every tile; a tiny entry capacity must flag
    overflow instead of corrupting memory."""
    W, H = 640, 360
    g = make_gaussians(G.GAUSSIAN, [[0, 0, 0], [0.1, 0, 0.5]], scale=3.0, color=(200, 100, 50, 128))
    cam = G.OrbitCamera(pos=(0, 0, 2.0))
    with G.Viewer(W, H, G.SH_NONE, G.COV3D_SINGLE) as v:
        m = v.add_model("big", 2)
        m.update_range(0, g)
        v.update_camera(cam)
        img = v.render_frame_host([m]).copy()
        t = v.last_timings()
        assert t.tile_entries == 2 * ((W + 31) // 32) * ((H + 31) // 32) and t.overflow == 0   # one entry per 32-pixel bin
        f = O.make_frame(cam.view(), cam.projection(np.float32(W) / np.float32(H)), W, H)
        ref, _, _ = O.render_frame(f, [O.ModelRef(3, 0, G.pack_gaussians(3, 0, g), 2)])
        assert_image_close(img, ref)
        v.set_tile_entry_capacity(100)
        with pytest.raises(G.GsError) as err:          # the (truncated) frame is delivered, the status says so
            v.render_frame_host([m])
        assert err.value.code == G.ERR_OVERFLOW
        assert v.last_timings().overflow == 1


def test_viewport_with_more_than_65536_bins(G, O):
    """8208 x 8208 pixels = 257 x 257 = 66049 bins of 32 pixels: bin ids need 17 bits, so the bin sort takes a third
    onesweep pass; a screen-filling splat exercises the huge class (a side of more than 32 tiles), the scene the
    regular one."""
    W = H = 8208
    n = 20_000
    g = G.gaussian_from_ply(G.synth_scene(SEED_100K, n))
    big = make_gaussians(G.GAUSSIAN, [[0, 0, 0.2]], scale=2.0, color=(40, 160, 220, 90))
    allg = np.concatenate([g, big])
    packed = G.pack_gaussians(2, 1, allg)
    cam = G.OrbitCamera.orbit()
    with G.Viewer(W, H) as v:
        m = v.add_model("m", len(allg))
        m.upload_packed(0, packed)
        v.update_camera(cam)
        img = v.render_frame_host([m]).copy()
        t = v.last_timings()
        assert t.overflow == 0 and t.tile_entries >= 257 * 257
        f = O.make_frame(cam.view(), cam.projection(np.float32(W) / np.float32(H)), W, H)
        ref, _, _ = O.render_frame(f, [O.ModelRef(2, 1, packed, len(allg))])
        assert_image_close(img, ref)


def test_call_order_errors(G):
    with G.Viewer(64, 64) as v:
        m = v.add_model("m", 10)
        with pytest.raises(G.GsError):
            m.sort()                                               # sort before preprocess
        m.preprocess()
        with pytest.raises(G.GsError):
            v.render([m], v.image_device())                        # render before sort
        with pytest.raises(G.GsError):
            v.add_model("m", 10)                                   # duplicate key
        with pytest.raises(G.GsError):
            m.upload_packed(5, np.zeros(10 * v.record_bytes, np.uint8))   # range exceeds capacity
        with pytest.raises(G.GsError):
            m.upload_mask(np.zeros(5, np.uint32))                  # wrong bitset length
        with pytest.raises(G.GsError):
            v.update_gaussian_transform(1.0, 0, 4, False)          # GaussianShDegree::new(4) is None
        with pytest.raises(G.GsError):
            m.eval_mask(np.array([(1, 0)], G.MASK_OP), np.zeros(0, G.MASK_SHAPE))   # malformed tree


def test_resize_and_determinism(G, O):
    n = 60_000
    packed = G.pack_gaussians(2, 1, G.gaussian_from_ply(G.synth_scene(SEED_100K, n)))
    cam = G.OrbitCamera.orbit()
    with G.Viewer(320, 180) as v:
        m = v.add_model("m", n)
        m.upload_packed(0, packed)
        imgs = {}
        for (W, H) in [(320, 180), (1280, 720), (320, 180)]:
            v.resize(W, H)
            v.update_camera(cam)
            img = v.render_frame_host([m]).copy()
            f = O.make_frame(cam.view(), cam.projection(np.float32(W) / np.float32(H)), W, H)
            ref, _, _ = O.render_frame(f, [O.ModelRef(2, 1, packed, n)])
            assert_image_close(img, ref)
            if (W, H) in imgs:
                assert np.array_equal(img, imgs[(W, H)])           # same view -> same bytes
            imgs[(W, H)] = img


def test_cpp_gs_mirror_renders_a_frame(G, tmp_path):
    """The C++ gs:: mirror drives one frame in the reference's call order (preprocess, sort, render)."""
    import subprocess
    from test_host import _build_gs_hpp_check
    exe = _build_gs_hpp_check(tmp_path)
    r = subprocess.run([exe, str(tmp_path), "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gpu ok" in r.stdout


def test_selection_query_rect_brush_matches_oracle(G, O):
    """K1's selection test (rect / brush x Set / Add / Remove): the rewritten selection bitset is
    bit-identical to the oracle's and the same frame is drawn with the new selection highlighted."""
    W, H, n = 640, 360, 50_000
    g = G.gaussian_from_ply(G.synth_scene(0xB2000080, n))
    packed = G.pack_gaussians(2, 1, g)
    cam = G.OrbitCamera.orbit()
    view, proj = cam.view(), cam.projection(np.float32(W) / np.float32(H))
    old = pack_bits(np.arange(n) % 5 == 0)
    cases = [(G.QUERY_RECT, G.SELECT_SET, (100.5, 60.25), (400.0, 300.0), 0.0),
             (G.QUERY_RECT, G.SELECT_ADD, (300, 100), (639, 200), 0.0),
             (G.QUERY_RECT, G.SELECT_REMOVE, (0, 0), (320, 360), 0.0),
             (G.QUERY_BRUSH, G.SELECT_SET, (50, 300), (600, 40), 40.0),
             (G.QUERY_BRUSH, G.SELECT_ADD, (320, 180), (320, 180), 25.0)]
    with G.Viewer(W, H) as v:
        m = v.add_model("m", n)
        m.upload_packed(0, packed)
        v.update_camera(cam)
        v.update_selection_highlight((1.0, 0.0, 1.0, 0.5))
        for kind, op, p0, p1, rad in cases:
            m.upload_selection(old)
            v.update_query(G.query_pod(kind, op, p0, p1, rad))
            f = O.make_frame(view, proj, W, H, highlight=(1.0, 0.0, 1.0, 0.5), query=O.query_pod(kind, op, p0, p1, rad))
            om = O.ModelRef(2, 1, packed, n, selection=old)
            img = v.render_frame_host([m]).copy()
            want = O.query_selection(f, om)
            got = m.download_selection()
            assert np.array_equal(got, want), (kind, op)
            assert 0 < bits_set(want, n).sum() < n
            oi, ok, osp = O.preprocess(f, om)
            ok, oi, osp = O.sort(ok, oi, osp)
            assert np.array_equal(m.indices(), oi)
            assert np.array_equal(np.ascontiguousarray(m.splats()["flags"]), np.ascontiguousarray(osp["flags"]))
            assert_image_close(img, O.composite(f, osp)[0])
        v.update_query(G.query_pod(G.QUERY_NONE))
        m.upload_selection(old)
        v.render_frame_host([m])
        assert np.array_equal(m.download_selection(), old)          # no query: selection untouched


def test_selection_query_texture_mode_and_postprocess(G, O):
    """Non-immediate selection (query_toolset.set_use_texture(true) + query_toolset.render into the query texture,
    reference scene.rs:767-791): strokes are painted into the viewer's query texture, a preprocess with query kind
    TEXTURE selects the Gaussians whose centre falls on a painted texel.  Texture, selection bitset and image match
    the oracle; then postprocess (scene.rs:596-611) commits the selection edit into the edit pods, again checked
    against the oracle's twin."""
    W, H, n = 640, 360, 50_000
    g = G.gaussian_from_ply(G.synth_scene(0xB2000081, n))
    packed = G.pack_gaussians(2, 1, g)
    cam = G.OrbitCamera.orbit()
    view, proj = cam.view(), cam.projection(np.float32(W) / np.float32(H))
    old = pack_bits(np.arange(n) % 7 == 0)
    strokes = [(G.QUERY_BRUSH, (50, 300), (300, 100), 30.0), (G.QUERY_BRUSH, (300, 100), (600, 250), 30.0),
               (G.QUERY_RECT, (10.5, 10.5), (90.25, 70.0), 0.0)]
    tex = np.zeros((H, W), np.uint8)
    for kind, p0, p1, rad in strokes:
        O.query_texture_paint(tex, O.query_pod(kind, 0, p0, p1, rad))
    edit = dict(flag=1, color=(0.5, 1.2, 0.9), contrast=0.2, exposure=0.5, gamma=1.2, alpha=0.8)
    with G.Viewer(W, H) as v:
        m = v.add_model("m", n)
        m.upload_packed(0, packed)
        v.update_camera(cam)
        v.update_selection_highlight((0.0, 1.0, 1.0, 0.4))
        v.query_texture_clear()
        for kind, p0, p1, rad in strokes:
            v.query_texture_paint(G.query_pod(kind, 0, p0, p1, rad))
        assert np.array_equal(v.query_texture_download(), tex)                 # painted texture: byte-exact
        for op in (G.SELECT_SET, G.SELECT_ADD, G.SELECT_REMOVE):
            m.upload_selection(old)
            v.update_query(G.query_pod(G.QUERY_TEXTURE, op))
            f = O.make_frame(view, proj, W, H, highlight=(0.0, 1.0, 1.0, 0.4), query=O.query_pod(4, op), query_texture=tex)
            om = O.ModelRef(2, 1, packed, n, selection=old)
            img = v.render_frame_host([m]).copy()
            want = O.query_selection(f, om)
            assert np.array_equal(m.download_selection(), want), op
            assert 0 < bits_set(want, n).sum() < n
            oi, ok, osp = O.preprocess(f, om)
            ok, oi, osp = O.sort(ok, oi, osp)
            assert np.array_equal(m.indices(), oi)
            assert_image_close(img, O.composite(f, osp)[0])
        # an uploaded texture behaves like a painted one
        v.query_texture_upload(tex[::-1].copy())
        m.upload_selection(old)
        v.update_query(G.query_pod(G.QUERY_TEXTURE, G.SELECT_SET))
        v.render_frame_host([m])
        f = O.make_frame(view, proj, W, H, query=O.query_pod(4, 0), query_texture=tex[::-1].copy())
        sel = O.query_selection(f, O.ModelRef(2, 1, packed, n, selection=old))
        assert np.array_equal(m.download_selection(), sel)
        # postprocess: the selection edit lands in the edit pods of exactly the selected Gaussians
        v.update_query(G.query_pod(G.QUERY_NONE))
        v.update_selection_edit(G.EditPod.new(**edit))
        m.postprocess()
        want_edits = O.postprocess(sel, m_default_edits(O, n), O.edit_pod(**edit))
        got_edits = m.download_edits()
        assert got_edits.tobytes() == want_edits.tobytes()
        # ... and the committed edits draw the same frame as the live selection edit did
        f2 = O.make_frame(view, proj, W, H, highlight=(0.0, 1.0, 1.0, 0.4))
        om2 = O.ModelRef(2, 1, packed, n, selection=sel, edits=want_edits)
        v.update_selection_edit(G.EditPod.default())
        img = v.render_frame_host([m]).copy()
        oi, ok, osp = O.preprocess(f2, om2)
        ok, oi, osp = O.sort(ok, oi, osp)
        assert_image_close(img, O.composite(f2, osp)[0])
        # resizing the viewer clears the texture (update_query_texture_size)
        v.resize(320, 200)
        assert v.query_texture_download().sum() == 0


def m_default_edits(O, n):
    e = np.zeros(n, dtype=O.EDIT)
    e["color"] = (0.0, 1.0, 1.0)
    e["gamma"] = 1.0
    e["alpha"] = 1.0
    return e


def test_viewers_share_resident_records(G, O):
    """b200gs_model_create_shared: a second viewer on the same device renders the records resident in the first
    (ref-counted, like the reference's cloned buffer handles, scene.rs:641): same bytes out, and the records
    survive the destruction of the model that uploaded them."""
    W, H, n = 800, 450, 60_000
    packed = G.pack_gaussians(2, 1, G.gaussian_from_ply(G.synth_scene(0xB2000082, n)))
    cam = G.OrbitCamera.orbit(4.0, 25.0, 80.0)
    v1, v2 = G.Viewer(W, H), G.Viewer(W, H)
    try:
        m1 = v1.add_model("scene", n)
        m1.upload_packed(0, packed)                # returns after enqueue: create_shared orders v2 behind it
        m2 = v2.add_shared_model("scene", m1)
        v1.update_camera(cam)
        v2.update_camera(cam)
        a = v1.render_frame_host([m1]).copy()
        b = v2.render_frame_host([m2]).copy()
        assert np.array_equal(a, b)
        assert m2.download_packed().tobytes() == packed.tobytes()
        v1.remove_model("scene")                   # the records stay alive for v2
        v1.sync()
        c = v2.render_frame_host([m2]).copy()
        assert np.array_equal(a, c)
        with pytest.raises(G.GsError):
            with G.Viewer(W, H, G.SH_HALF, G.COV3D_HALF) as v3:
                v3.add_shared_model("scene", m2)   # different layout
    finally:
        v1.close()
        v2.close()


def _hits_from_oracle(O, f, idx, keys, spl, px, py):
    """Per-pixel hit list restated from the oracle's depth-sorted splats (same alpha rule as its compositor)."""
    W, H = f.size[0], f.size[1]
    out = []
    for k in range(len(idx)):
        s = spl[k]
        r = float(s["radius"])
        if r == 0:
            continue
        x0, x1 = max(np.ceil(s["mx"] - r), 0), min(np.floor(s["mx"] + r), W - 1)
        y0, y1 = max(np.ceil(s["my"] - r), 0), min(np.floor(s["my"] + r), H - 1)
        if not (x0 <= px <= x1 and y0 <= py <= y1):
            continue
        dx, dy = np.float32(px) - s["mx"], np.float32(py) - s["my"]
        power = -0.5 * (s["ca"] * dx * dx + s["cc"] * dy * dy) - s["cb"] * dx * dy
        if power > 0:
            continue
        al = min(0.99, float(s["opacity_h"]) * float(np.exp(power)))
        if al < 1 / 255:
            continue
        out.append((int(idx[k]), al, float(keys[k:k + 1].view(np.float32)[0])))
    return out


def test_hit_query_list_and_positions(G, O):
    """Row N3: the measurement tool's hit query (reference src/tab/scene.rs:617-676)."""
    # three splats stacked on the optical axis + a scene for a denser list
    W = H = 65
    D = 5.0
    g = make_gaussians(G.GAUSSIAN, [[0, 0, 1.0], [0, 0, -1.0], [0, 0, 0.0], [3.0, 0, 0]], scale=0.3, color=(255, 255, 255, 100))
    cam = G.OrbitCamera(pos=(0, 0, D))
    view, proj = cam.view(), cam.projection(np.float32(1.0))
    with G.Viewer(W, H, G.SH_NONE, G.COV3D_SINGLE) as v:
        m = v.add_model("m", 4)
        m.update_range(0, g)
        v.update_camera(cam)
        v.render_frame_host([m])
        hits = v.query_hits([m], W // 2, H // 2)
        assert hits["index"].tolist() == [0, 2, 1] and hits["model"].tolist() == [0, 0, 0]   # nearest first
        assert np.allclose(hits["alpha"], min(0.99, 100 / 255), rtol=2e-3)
        assert np.all(np.diff(hits["depth"]) > 0)
        p = G.hit_pos_by_closest(hits, view, proj, (W, H), W // 2, H // 2)
        assert np.allclose(p, [0, 0, 1.0], atol=2e-3)
        p = G.hit_pos_by_alpha_range(hits, 0.05, view, proj, (W, H), W // 2, H // 2)
        assert abs(p[0]) < 1e-3 and abs(p[1]) < 1e-3 and -1.0 < p[2] < 1.0
        with pytest.raises(G.GsError):
            G.hit_pos_by_alpha_range(hits, 0.9, view, proj, (W, H), W // 2, H // 2)      # nothing that opaque
        assert len(v.query_hits([m], 0, 0)) == 0
    # dense scene: list equals the oracle-derived one
    W, H, n = 320, 180, 30_000
    packed = G.pack_gaussians(2, 1, G.gaussian_from_ply(G.synth_scene(0xB2000081, n)))
    cam = G.OrbitCamera.orbit()
    f = O.make_frame(cam.view(), cam.projection(np.float32(W) / np.float32(H)), W, H)
    oi, ok, osp = O.preprocess(f, O.ModelRef(2, 1, packed, n))
    ok, oi, osp = O.sort(ok, oi, osp)
    with G.Viewer(W, H) as v:
        m = v.add_model("m", n)
        m.upload_packed(0, packed)
        v.update_camera(cam)
        v.render_frame_host([m])
        for (px, py) in [(160, 90), (100, 120), (250, 60)]:
            want = _hits_from_oracle(O, f, oi, ok, osp, px, py)
            got = v.query_hits([m], px, py)
            strong = [w for w in want if w[1] > 1.2 / 255]           # away from the 1/255 cut (exp vs ex2)
            got_idx = got["index"].tolist()
            assert [w[0] for w in strong] == [i for i in got_idx if i in {w[0] for w in strong}]
            assert abs(len(got) - len(want)) <= 2 and len(want) > 0
            d = {w[0]: w for w in want}
            for h in got:
                if int(h["index"]) in d:
                    assert abs(h["alpha"] - d[int(h["index"])][1]) < 2e-3 and h["depth"] == np.float32(d[int(h["index"])][2])


def test_partial_bins_and_quadrants_at_odd_viewports(G, O):
    """Binning is per 32x32-pixel bin, compositing per 16x16 tile (a quadrant of its bin, fed through the quadrant mask of
    every entry): viewports whose last bin column / row is partial — down to a single pixel, with whole quadrants
    outside — must match the oracle like any other, for one and for two layered models."""
    n = 60_000
    cam = G.OrbitCamera.orbit(3.0, 10.0, 40.0)
    packs = [G.pack_gaussians(2, 1, G.gaussian_from_ply(G.synth_scene(seed, n))) for seed in (0xB2000082, 0xB2000083)]
    for W, H in ((333, 217), (353, 97), (64, 33), (17, 48)):
        with G.Viewer(W, H) as v:
            ms = []
            for k in range(2):
                m = v.add_model("m%d" % k, n)
                m.upload_packed(0, packs[k])
                m.set_transform((0.5 * k, 0, 0), G.quat_from_euler_zyx_deg([0, 20 * k, 0]), (1, 1, 1))
                ms.append(m)
            v.update_camera(cam)
            f = O.make_frame(cam.view(), cam.projection(np.float32(W) / np.float32(H)), W, H)
            oms = [O.ModelRef(2, 1, packs[k], n, pos=(0.5 * k, 0, 0), quat=G.quat_from_euler_zyx_deg([0, 20 * k, 0])) for k in range(2)]
            for sel in ([0], [0, 1]):
                img = v.render_frame_host([ms[i] for i in sel]).copy()
                assert v.last_timings().overflow == 0
                ref, _, _ = O.render_frame(f, [oms[i] for i in sel])
                assert_image_close(img, ref)


def test_long_translucent_lists_exercise_the_producer_ring(G, O):
    """120 000 nearly transparent splats over a 2 x 2 block of 32-pixel bins (plus a thin veil over the whole frame): the
    two central bins hold ~55 000 entries each — ~27x the compositor's id ring — and no pixel ever saturates (the oracle's
    alpha stays below 215), so every quadrant CTA walks its whole list: the producer warp runs in steady state (ring
    wrap-around, dozens of rounds) instead of stopping after the cooperative prologue.  Image vs oracle."""
    W, H = 256, 160
    rng = np.random.default_rng(7)
    n_pile, n_veil = 120_000, 3_000
    pile = np.stack([rng.uniform(-0.65, 0.65, n_pile), rng.uniform(-0.45, 0.45, n_pile), rng.uniform(-1.0, 1.0, n_pile)], 1)
    veil = np.stack([rng.uniform(-1.6, 1.6, n_veil), rng.uniform(-1.0, 1.0, n_veil), rng.uniform(-1.0, 1.0, n_veil)], 1)
    g = np.concatenate([make_gaussians(G.GAUSSIAN, pile, scale=0.01, color=(250, 120, 40, 2)),
                        make_gaussians(G.GAUSSIAN, veil, scale=0.03, color=(30, 200, 240, 40))])
    g["color"][:, :3] = rng.integers(0, 256, (len(g), 3), dtype=np.uint8)       # (so that a misordered blend shows)
    cam = G.OrbitCamera(pos=(0, 0, 3.0))
    with G.Viewer(W, H, G.SH_NONE, G.COV3D_SINGLE) as v:
        m = v.add_model("pile", len(g))
        m.update_range(0, g)
        v.update_camera(cam)
        v.enable_timings(True, True)
        img = v.render_frame_host([m]).copy()
        t = v.last_timings()
        assert t.overflow == 0 and t.visible == len(g)
        assert t.staged_entries > n_pile                          # every list was walked to its end (each splat is in >= 1 tile)
    f = O.make_frame(cam.view(), cam.projection(np.float32(W) / np.float32(H)), W, H)
    ref, _, _ = O.render_frame(f, [O.ModelRef(3, 0, G.pack_gaussians(3, 0, g), len(g))])
    assert int(ref[..., 3].max()) < 230                           # (nothing saturates: no early exit anywhere)
    assert_image_close(img, ref)
