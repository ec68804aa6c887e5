"""Small frame for compute-sanitizer: python tools/sanitize_frame.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200gs as G
N, W, H = 30_000, 320, 180
g = G.gaussian_from_ply(G.synth_scene(0xB2000001, N))
with G.Viewer(W, H) as v:
    ms = []
    for k in range(2):
        m = v.add_model("m%d" % k, N)
        m.update_range(0, g)
        m.set_transform((0.6 * k, 0, 0), G.quat_from_euler_zyx_deg([0, 25 * k, 0]), (1, 1, 1))
        ms.append(m)
    ms[0].upload_mask(np.full((N + 31) // 32, 0xF0F0FFFF, np.uint32))
    ms[1].upload_selection(np.full((N + 31) // 32, 0x0000FFFF, np.uint32))
    v.update_selection_highlight((1, 0, 1, 0.5))
    v.update_query(G.query_pod(G.QUERY_RECT, G.SELECT_ADD, (50, 40), (200, 150)))
    for i, cam in enumerate(G.view_batch()[:3]):
        v.update_camera(cam)
        img = v.render_frame_host(ms)
    hits = v.query_hits(ms, 160, 90)
    k, val = v.sort_pairs(np.arange(10000, dtype=np.uint32)[::-1].copy(), np.arange(10000, dtype=np.uint32))
    print("ok", int(img[..., 3].sum()), len(hits), int(k[0]))
