"""Throughput of a view batch with 1 vs K viewers (one CUDA stream each) on one GPU:
   python tools/two_streams.py [N] [frames] [K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200gs as G
N = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
W, H = 1920, 1080
packed = G.pack_gaussians(G.SH_NORM8, G.COV3D_HALF, G.gaussian_from_ply(G.synth_scene(0xB2000006, N)))
cams = G.view_batch()
for k in range(1, K + 1):
    vs = [G.Viewer(W, H) for _ in range(k)]
    ms = []
    for v in vs:
        m = v.add_model("scene", N); m.upload_packed(0, packed); ms.append(m)
    def run(nf):
        for i in range(nf):
            v = vs[i % k]
            v.update_camera(cams[i % 64])
            v.render_frame([ms[i % k]])
        for v in vs: v.sync()
    run(8)
    t0 = time.time(); run(frames); dt = time.time() - t0
    print("viewers=%d: %.3f ms/frame, %.1f frames/s" % (k, dt / frames * 1e3, frames / dt), flush=True)
    for v in vs: v.close()
