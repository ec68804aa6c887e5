import sys; sys.path.insert(0,"/root/repo")
import numpy as np, b200gs as G
N=6_000_000; W,H=1920,1080
packed=G.pack_gaussians(2,1,G.gaussian_from_ply(G.synth_scene(0xB2000006,N)))
cams=G.view_batch()
with G.Viewer(W,H) as v:
    m=v.add_model("s",N); m.upload_packed(0,packed); v.enable_timings(True,True)
    for i in range(8):
        v.update_camera(cams[i]); v.render_frame([m]); t=v.last_timings()
        print(i, "V",t.visible,"entries",t.tile_entries,"staged",t.staged_entries,"(%.0f%%)"%(100*t.staged_entries/max(1,t.tile_entries)),"evals",t.evals, "comp ms %.3f"%t.composite_ms)
