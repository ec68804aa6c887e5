"""Small frames for compute-sanitizer that touch every kernel of the library: python tools/sanitize_frame.py
(two layered models with transforms, mask, selection, edits; rect query, texture query (paint + sample), mask evaluation,
postprocess, all three display modes, a viewport with partial bins, a shared model on a second viewer, the hit query, both raw
sorts, a pipelined host frame pair)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200gs as G
N, W, H = 30_000, 333, 187
g = G.gaussian_from_ply(G.synth_scene(0xB2000001, N))
cams = G.view_batch()
with G.Viewer(W, H) as v, G.Viewer(W, H) as v2:
    ms = []
    for k in range(2):
        m = v.add_model("m%d" % k, N)
        m.update_range(0, g)
        m.set_transform((0.6 * k, 0, 0), G.quat_from_euler_zyx_deg([0, 25 * k, 0]), (1, 1, 1))
        ms.append(m)
    ms[0].upload_mask(np.full((N + 31) // 32, 0xF0F0FFFF, np.uint32))
    ms[1].upload_selection(np.full((N + 31) // 32, 0x0000FFFF, np.uint32))
    # mask evaluation (box | ellipsoid - box) on model 1
    shapes = np.zeros(3, dtype=G.MASK_SHAPE)
    shapes["kind"] = (G.MASK_BOX, G.MASK_ELLIPSOID, G.MASK_BOX)
    shapes["pos"] = ((-0.5, 0, 0), (0.5, 0, 0), (0, 0, 0))
    shapes["quat"] = (0, 0, 0, 1)
    shapes["scale"] = ((3, 3, 3), (4, 2, 4), (1, 1, 1))
    ops = np.array([(G.MASKOP_SHAPE, 0), (G.MASKOP_SHAPE, 1), (G.MASKOP_UNION, 0), (G.MASKOP_SHAPE, 2), (G.MASKOP_DIFFERENCE, 0)], dtype=G.MASK_OP)
    ms[1].eval_mask(ops, shapes)
    v.update_selection_highlight((1, 0, 1, 0.5))
    v.update_query(G.query_pod(G.QUERY_RECT, G.SELECT_ADD, (50, 40), (200, 150)))
    for i, cam in enumerate(cams[:3]):
        v.update_camera(cam)
        v.update_gaussian_transform(1.0, (G.DISPLAY_SPLAT, G.DISPLAY_ELLIPSE, G.DISPLAY_POINT)[i], 3 - i, i == 2)
        img = v.render_frame_host(ms)
    v.update_gaussian_transform()
    # texture query: paint two strokes, select with the texture, commit an edit with postprocess
    v.query_texture_clear()
    v.query_texture_paint(G.query_pod(G.QUERY_BRUSH, G.SELECT_SET, (40, 40), (250, 120), 12.0))
    v.query_texture_paint(G.query_pod(G.QUERY_RECT, G.SELECT_SET, (100, 90), (180, 170)))
    tex = v.query_texture_download()
    v.update_query(G.query_pod(G.QUERY_TEXTURE, G.SELECT_SET))
    ms[0].preprocess()
    v.update_query(G.query_pod(G.QUERY_NONE))
    v.update_selection_edit(G.EditPod.new(G.EDIT_ENABLED, (0.5, 1.2, 0.9), 0.2, 0.5, 1.2, 0.8))
    ms[0].postprocess()
    v.update_selection_edit(G.EditPod.default())
    img = v.render_frame_host(ms)
    hits = v.query_hits(ms, W // 2, H // 2)
    # a second viewer on the same records (shared model), pipelined host frames
    sm = v2.add_shared_model("shared", ms[0])
    out = [np.zeros((H, W, 4), np.uint8) for _ in range(2)]
    pinned = [G.PinnedBuffer(W * H * 4) for _ in range(2)]
    v2.render_frame_host_begin([sm], cams[4], pinned[0].array)
    v2.render_frame_host_begin([sm], cams[5], pinned[1].array)
    v2.render_frame_host_end(); v2.render_frame_host_end()
    k, val = v.sort_pairs(np.arange(10000, dtype=np.uint32)[::-1].copy(), np.arange(10000, dtype=np.uint32))
    if os.environ.get("B200GS_SANITIZE_WIDE", "1") == "1":
        # K2w (sort_wide.cu, not on the frame path).  racecheck does not model thread-block-cluster barriers and reports its
        # mbarrier init -> expect_tx sequence (one thread, fenced, then barrier.cluster) as WARNINGS: tools/sanitize.sh leaves
        # this call out of the racecheck runs and keeps it in memcheck / synccheck
        k2, val2 = v.sort_pairs((np.arange(50000, dtype=np.uint32) * 2654435761 % 4096).astype(np.uint32), np.arange(50000, dtype=np.uint32), bits=12, wide=True)
        assert np.all(np.diff(k2.astype(np.int64)) >= 0)
    print("ok", int(img[..., 3].sum()), len(hits), int(k[0]), int(tex.sum() > 0))
