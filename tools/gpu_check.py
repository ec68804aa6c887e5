"""Developer check on a GPU box: parity of every stage against the oracle + stage timings.
   python tools/gpu_check.py [N] [W] [H] [seed]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200gs as G
from oracle import oracle as O

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
H = int(sys.argv[3]) if len(sys.argv) > 3 else 720
seed = int(sys.argv[4], 0) if len(sys.argv) > 4 else 0xB2000001
check = N <= 2000000

t = time.time()
ply = G.synth_scene(seed, N)
g = G.gaussian_from_ply(ply)
packed = G.pack_gaussians(G.SH_NORM8, G.COV3D_HALF, g)
print("scene built %.2fs" % (time.time() - t), flush=True)
cam = G.OrbitCamera.orbit()
v = G.Viewer(W, H)
m = v.add_model("m", N)
t = time.time(); m.upload_packed(0, packed); print("upload %.3fs" % (time.time() - t), flush=True)
v.update_camera(cam)
v.enable_timings(True, True)
m.preprocess(); v.sync()
V = m.visible_count()
print("V =", V, flush=True)
if check:
    f = O.make_frame(cam.view(), cam.projection(np.float32(W) / np.float32(H)), W, H)
    om = O.ModelRef(2, 1, packed, N)
    oi, ok, osp = O.preprocess(f, om)
    print("oracle V =", len(oi))
    gi, gk, gs = m.indices(), m.depth_keys(), m.splats()
    print("idx equal", np.array_equal(gi, oi), "keys equal", np.array_equal(gk, ok))
    if len(gs) == len(osp):
        for fld in ("mx", "my", "radius", "ca", "cb", "cc", "opacity_h", "r_h", "g_h", "b_h", "flags"):
            a, b = np.ascontiguousarray(gs[fld]), np.ascontiguousarray(osp[fld])
            eq = np.array_equal(a.view(np.uint8), b.view(np.uint8))
            md = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if len(a) else 0.0
            print("  splat.%s bit-equal=%s maxdiff=%g" % (fld, eq, md))
m.sort(); v.sync()
if check:
    ok2, oi2, osp2 = O.sort(ok, oi, osp)
    gi2, gk2 = m.indices(), m.depth_keys()
    print("sorted keys equal", np.array_equal(gk2, ok2), "sorted idx equal", np.array_equal(gi2, oi2))
img = v.render_frame_host([m]).copy()
tm = v.last_timings()
print("timings ms: pre %.3f sort %.3f bin %.3f comp %.3f total %.3f | V=%d entries=%d evals=%d overflow=%d" % (
    tm.preprocess_ms, tm.sort_ms, tm.bin_ms, tm.composite_ms, tm.total_ms, tm.visible, tm.tile_entries, tm.evals, tm.overflow), flush=True)
if check:
    ref, _ = O.composite(f, osp2, False)
    d = np.abs(img.astype(int) - ref.astype(int))
    mse = (d.astype(float) ** 2).mean()
    print("image vs oracle b2f: max|d|=%d psnr=%.2f dB  n(d>1)=%d" % (d.max(), 10 * np.log10(255 ** 2 / max(mse, 1e-12)), int((d > 1).sum())))
    ref2, ev = O.composite(f, osp2, True)
    d2 = np.abs(img.astype(int) - ref2.astype(int))
    print("image vs oracle f2b: max|d|=%d  oracle evals=%d" % (d2.max(), ev))
from PIL import Image
os.makedirs("gpurun_out", exist_ok=True)
Image.fromarray(img[..., :3]).save("gpurun_out/gpu_%d.png" % N)
# timing loop without counters
v.enable_timings(True, False)
ts = []
for i in range(30):
    v.render_frame([m]); 
    tm = v.last_timings()
    ts.append((tm.preprocess_ms, tm.sort_ms, tm.bin_ms, tm.composite_ms, tm.total_ms))
ts = np.array(ts[5:])
print("median ms: pre %.3f sort %.3f bin %.3f comp %.3f total %.3f" % tuple(np.median(ts, 0)), flush=True)
v.enable_timings(False, False)
import ctypes
t0 = time.time()
for i in range(50): v.render_frame([m])
v.sync(); dt = (time.time() - t0) / 50
print("untimed loop: %.3f ms/frame = %.1f fps" % (dt * 1e3, 1 / dt))
t0 = time.time()
for i in range(20): v.render_frame_host([m], cam)
dt = (time.time() - t0) / 20
print("e2e host loop: %.3f ms/frame = %.1f fps" % (dt * 1e3, 1 / dt))
v.close()
