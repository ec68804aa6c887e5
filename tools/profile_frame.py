"""Renders a few frames of the bench workload (for ncu): python tools/profile_frame.py [N] [frames] [W] [H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200gs as G
N = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 4
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
    for i in range(frames):
        v.update_camera(cams[i])
        v.render_frame([m])
    v.sync()
print("done")
