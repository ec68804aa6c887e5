"""Median per-stage device times for the bench workload: python tools/time_stages.py [N] [frames] [W] [H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
import b200gs as G
N = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 24
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1920
H = int(sys.argv[4]) if len(sys.argv) > 4 else 1080
packed = G.pack_gaussians(G.SH_NORM8, G.COV3D_HALF, G.gaussian_from_ply(G.synth_scene(0xB2000006, N)))
cams = G.view_batch()
if os.environ.get("SORT_CLAIM"):
    G.set_tuning("sort.claim", int(os.environ["SORT_CLAIM"]))
if os.environ.get("SORT_CLUSTER"):
    G.set_tuning("sort.cluster", int(os.environ["SORT_CLUSTER"]))
with G.Viewer(W, H) as v:
    m = v.add_model("scene", N)
    m.upload_packed(0, packed)
    v.enable_timings(True, False)
    rows = []
    for i in range(frames):
        v.update_camera(cams[i % 8])
        v.render_frame([m])
        t = v.last_timings()
        rows.append((t.preprocess_ms, t.sort_ms, t.bin_ms, t.composite_ms, t.total_ms))
    r = np.median(np.array(rows[4:]), 0)
    print("entries", t.tile_entries, "sort.cluster", v.info("sort.cluster"),
          "resident clusters", v.info("sort.resident_clusters"))
    print("median ms: pre %.3f sort %.3f bin %.3f comp %.3f total %.3f" % tuple(r))
