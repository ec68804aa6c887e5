"""Per-view device time of the whole 1024-view batch on one GPU (one stream): python tools/view_costs.py
Writes gpurun_out/view_costs.npy; used to predict the load balance of contiguous per-rank view blocks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200gs as G
N = 6_000_000
packed = G.pack_gaussians(G.SH_NORM8, G.COV3D_HALF, G.gaussian_from_ply(G.synth_scene(0xB2000006, N)))
cams = G.view_batch()
with G.Viewer(1920, 1080) as v:
    m = v.add_model("scene", N)
    m.upload_packed(0, packed)
    v.enable_timings(True, False)
    for i in range(8):
        v.update_camera(cams[i]); v.render_frame([m]); v.last_timings()
    out = np.zeros((2, len(cams)))
    for rep in range(2):
        for i, c in enumerate(cams):
            v.update_camera(c)
            v.render_frame([m])
            out[rep, i] = v.last_timings().total_ms
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/view_costs.npy", out)
t = out.min(0)
print("mean %.4f ms, min %.4f, max %.4f" % (t.mean(), t.min(), t.max()))
for n in (2, 4, 8):
    blocks = t.reshape(n, -1).sum(1)
    print("N=%d contiguous blocks: mean/max = %.4f" % (n, blocks.mean() / blocks.max()))
