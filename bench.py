#!/usr/bin/env python
"""bench.py — frames/s of the 3DGS render hot path (preprocess -> depth sort -> compositing).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

Workload (BASELINE.json `metric`, configs[2] + configs[4]): the synthetic 6M-Gaussian SH3 scene
(seed 0xB2000006, Norm8 SH + Half Cov3d, 76 B/record) at 1920x1080, rendered from the 1024-view
orbit batch of SURVEY.md §8d.  One step = every rank renders `--views` consecutive views of ITS
contiguous block of the batch (weak scaling: per-GPU work is fixed; default 64 views per step, so the
default 20 steps render 1280 frames per GPU); the scene is replicated (broadcast once with NCCL), no
collective runs inside a frame, finished images are gathered to rank 0 with NCCL on the side and a sample
of them is verified byte for byte against a local re-render.  Every GPU runs `--viewers` viewer handles
(one CUDA stream each, default 2, sharing ONE resident copy of the scene) and deals the views of a step
round-robin, so independent frames overlap on the device.  `strong_scaling` times the whole 1024-view batch
divided over the ranks.  `value` = frames/s over all ranks with the scene resident in HBM;
`e2e` = the same through b200gs_render_frame_host_begin/_end (host camera in, RGBA8 image out to
pinned host memory, D2H inside the timed region).  `--impl reference` times the CPU restatement of the
reference path (oracle/, all host threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0xB2000006
METRIC = "frames/sec @1080p for 6M-splat SH3 scene"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--views", type=int, default=64, help="views rendered per step per GPU")
    ap.add_argument("--viewers", type=int, default=2,
                    help="viewer handles (one CUDA stream + one scene replica each) per GPU; the views of a step are "
                         "dealt round-robin, so independent frames overlap on the device")
    ap.add_argument("--gaussians", type=int, default=6_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs stage timings (1M, 6M@4K, config 4)")
    ap.add_argument("--no-gather", action="store_true", help="diagnostic: skip the image gather")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="how finished images reach rank 0: p2p = every rank copies its slot straight into rank 0's buffer over "
                         "NVLink with the copy engines (symmetric memory, no SMs); nccl = dist.gather (SM kernels on every rank)")
    return ap.parse_args()


def workload_name(a):
    return ("synthetic %.1fM-Gaussian SH3 scene (seed 0x%X, Norm8 SH + Half Cov3d, 76 B/record) at %dx%d, "
            "1024-view orbit batch" % (a.gaussians / 1e6, SEED, a.width, a.height))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ reference arm
def run_reference(a, rank):
    """The reference's CPU implementation of the path is not buildable here (no rustc/cargo, crate
    source absent; tools/plan_a_probe.py records the probe): this arm times the CPU restatement (oracle/) on
    EVERY host core, one frame of the same workload per step.  It never imports the product package."""
    if rank != 0:
        return
    from oracle import oracle as O
    cores = O.set_num_threads(os.cpu_count())   # torch.distributed.run exports OMP_NUM_THREADS=1: override it
    packed = O.pack(2, 1, O.gaussian_from_ply(O.synth_scene(SEED, a.gaussians)))
    cams = O.view_batch(a.width, a.height)
    model = O.ModelRef(2, 1, packed, a.gaussians)

    def frame(i):
        view, proj = cams[i % len(cams)]
        f = O.make_frame(view, proj, a.width, a.height)
        return O.render_frame(f, [model], front_to_back=False, fp32=True)

    for i in range(a.warmup):
        frame(i)
    t0 = time.perf_counter()
    stages = np.zeros(3)
    for i in range(a.steps):
        _, _, st = frame(a.warmup + i)
        stages += np.array(st)
    dt = time.perf_counter() - t0
    fps = a.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "frames_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "1 frame (1 view of the batch) per step; CPU restatement of the reference path in fp32, "
                                   "back-to-front blending, OpenMP on %d threads; stage seconds/frame pre=%.3f sort=%.3f "
                                   "composite=%.3f" % ((cores,) + tuple(stages / a.steps)),
                         "plan_a_probe": plan_a_probe()},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def plan_a_probe():
    """can the reference's own wgpu pipeline run on this box? (tools/plan_a_probe.py)"""
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("plan_a_probe", os.path.join(ROOT, "tools", "plan_a_probe.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        r = mod.probe()
        return {k: r[k] for k in ("runnable", "which_cargo", "which_rustc", "cargo_registry", "crate_source_found", "vulkan_icd",
                                  "software_adapter_libs", "crates_io_reachable", "nproc", "hostname_kind")}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


# ------------------------------------------------------------------------------ our arm
def stage_times(G, v, m, mats, W, H, frames=14, spread=None):
    """median per-stage device times (CUDA events on the viewer's stream) of `frames` consecutive views, timed with the
    production kernels; the work counters (evaluations) come from a second pass over the same views with the counting
    instantiation of the compositor, whose atomics would otherwise sit inside the timed numbers.  `spread`, if a dict, receives
    the p10 / p90 of the frame time."""
    rows = []
    for counting in (False, True):
        v.enable_timings(True, counting)
        for s in range(frames if not counting else min(frames, 14)):
            view, proj = mats[s % len(mats)]
            v.update_camera_matrices(view, proj, (W, H))
            v.render_frame([m])
            tm = v.last_timings()
            if not counting:
                rows.append([tm.preprocess_ms, tm.sort_ms, tm.bin_ms, tm.composite_ms, tm.total_ms, tm.visible, tm.tile_entries, 0])
            else:
                rows[s][7] = tm.evals
    v.enable_timings(False, False)
    arr = np.array(rows[2:], dtype=np.float64)
    if spread is not None:
        spread["frame_ms_p10"], spread["frame_ms_p90"] = float(np.percentile(arr[:, 4], 10)), float(np.percentile(arr[:, 4], 90))
        spread["frames_timed"] = int(len(arr))
    med = np.median(arr, axis=0)
    med[7] = np.median(arr[:12, 7])      # (evaluation counts exist for the first frames only)
    return med


def ply_stream_rate(G, local_rank, n=1_000_000, chunk=65536):
    """Row N1 of SURVEY.md §8: the loading loop of the reference (scene.rs:341-380 — read a chunk of the PLY, Gaussian::from,
    gaussians_buffer.update_range, every frame while the file streams in) through this library: PLY file -> b200gs_ply_read ->
    b200gs_gaussian_from_ply -> b200gs_model_update_range (host packer + pinned ring + H2D).  1M Gaussians (248 MB of PLY)."""
    import tempfile
    ply = G.synth_scene(0xB2000002, n)
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "scene.ply")
        G.write_ply(path, ply)
        size = os.path.getsize(path)
        del ply
        with G.Viewer(640, 360, G.SH_NORM8, G.COV3D_HALF, device=local_rank) as v:
            m = v.add_model("stream", n)
            t0 = time.perf_counter()
            start = parse_s = conv_s = up_s = 0.0
            start = 0
            ta = time.perf_counter()
            for verts in G.read_ply(path, chunk=chunk):
                tb = time.perf_counter()
                g = G.gaussian_from_ply(verts)
                tc = time.perf_counter()
                m.update_range(start, g)
                td_ = time.perf_counter()
                parse_s += tb - ta; conv_s += tc - tb; up_s += td_ - tc
                start += len(verts)
                ta = time.perf_counter()
            v.sync()
            dt = time.perf_counter() - t0
    return {"gaussians": n, "ply_bytes": size, "chunk": chunk, "seconds": dt, "ply_gb_per_s": size / dt / 1e9,
            "mgaussians_per_s": n / dt / 1e6, "read_s": parse_s, "gaussian_from_ply_s": conv_s, "update_range_s": up_s,
            "note": "file -> b200gs_ply_read -> b200gs_gaussian_from_ply -> b200gs_model_update_range (pack + pinned ring + H2D); the converters and the packer split large ranges over the host threads (host.cpp parallel_for)"}


def extra_config_times(G, local_rank):
    """The other BASELINE.json configs on one stream (CUDA events, median over views of the batch): configs[1]
    1M @ 1080p, configs[2] second leg 6M @ 3840x2160 (reuses nothing of the timed run), configs[3] three 2M models
    with per-model transforms, colour edits on a rect-selected half and a composite mask."""
    out = {}
    cams = G.view_batch()

    def mats_for(W, H):
        asp = np.float32(W) / np.float32(H)
        return [(c.view(), c.projection(asp)) for c in cams[:16]]

    for name, n, seed, W, H in (("1M@1920x1080", 1_000_000, 0xB2000002, 1920, 1080), ("6M@3840x2160", 6_000_000, SEED, 3840, 2160)):
        packed = G.pack_gaussians(G.SH_NORM8, G.COV3D_HALF, G.gaussian_from_ply(G.synth_scene(seed, n)))
        with G.Viewer(W, H, G.SH_NORM8, G.COV3D_HALF, device=local_rank) as v:
            m = v.add_model("scene", n)
            m.upload_packed(0, packed)
            pre, srt, bn, comp, tot, vis, ent, ev = stage_times(G, v, m, mats_for(W, H), W, H)
        out[name] = {"frame_ms": tot, "frames_per_s_one_stream": 1e3 / tot, "preprocess_ms": pre, "sort_ms": srt, "bin_ms": bn,
                     "composite_ms": comp, "visible": int(vis), "tile_entries": int(ent)}
        del packed
    # secondary layouts of SURVEY.md §8(d): the same 6M scene packed Half + Half (120 B) and Single + Single (220 B), 1080p
    g6 = G.gaussian_from_ply(G.synth_scene(SEED, 6_000_000))
    for name, sh, cov in (("6M@1920x1080 Half+Half (120 B/record)", G.SH_HALF, G.COV3D_HALF),
                          ("6M@1920x1080 Single+Single (220 B/record)", G.SH_SINGLE, G.COV3D_SINGLE)):
        W, H, n = 1920, 1080, 6_000_000
        rb2 = G.record_bytes(sh, cov)
        with G.Viewer(W, H, sh, cov, device=local_rank) as v:
            m = v.add_model("scene", n)
            m.upload_packed(0, G.pack_gaussians(sh, cov, g6))
            pre, srt, bn, comp, tot, vis, ent, ev = stage_times(G, v, m, mats_for(W, H), W, H)
        out[name] = {"frame_ms": tot, "frames_per_s_one_stream": 1e3 / tot, "preprocess_ms": pre, "sort_ms": srt, "bin_ms": bn,
                     "composite_ms": comp, "visible": int(vis), "record_bytes": rb2,
                     "preprocess_hbm_frac": ((n * rb2 + 8 * vis) / (pre * 1e-3) / 1e9) / peaks()[0]}
    del g6
    # config 4
    W, H, n = 1920, 1080, 2_000_000
    xf = [((-2, 0, 0), (0, 30, 0), (1, 1, 1)), ((0, 0, 0), (10, 0, 45), (1.2, 0.8, 1)), ((2, 0.5, 0), (0, -60, 0), (0.7, 0.7, 0.7))]
    with G.Viewer(W, H, G.SH_NORM8, G.COV3D_HALF, device=local_rank) as v:
        models, centers = [], []
        for k, seed in enumerate((0xB2000041, 0xB2000042, 0xB2000043)):
            g = G.gaussian_from_ply(G.synth_scene(seed, n))
            mm = v.add_model("m%d" % k, n)
            mm.update_range(0, g)
            pos, rot, scale = xf[k]
            mm.set_transform(pos, G.quat_from_euler_zyx_deg(rot), scale)
            centers.append(g["pos"].mean(axis=0))
            models.append(mm)
            del g
        shapes = np.zeros(3, dtype=G.MASK_SHAPE)
        shapes["kind"] = (G.MASK_BOX, G.MASK_ELLIPSOID, G.MASK_BOX)
        shapes["pos"] = ((-0.5, 0, 0), (0.5, 0, 0), (0, 0, 0))
        shapes["quat"] = (0, 0, 0, 1)
        shapes["scale"] = ((3, 3, 3), (4, 2, 4), (1, 1, 1))
        ops = np.array([(G.MASKOP_SHAPE, 0), (G.MASKOP_SHAPE, 1), (G.MASKOP_UNION, 0), (G.MASKOP_SHAPE, 2), (G.MASKOP_DIFFERENCE, 0)],
                       dtype=G.MASK_OP)
        models[1].eval_mask(ops, shapes)
        # edit on a rect-selected half of model 2: select with a query, commit with postprocess
        v.update_camera(cams[5])
        v.update_query(G.query_pod(G.QUERY_RECT, G.SELECT_SET, (0, 0), (W / 2, H)))
        models[2].preprocess()
        v.update_query(G.query_pod(G.QUERY_NONE))
        v.update_selection_edit(G.EditPod.new(G.EDIT_ENABLED, (0.5, 1.2, 0.9), 0.2, 0.5, 1.2, 0.8))
        models[2].postprocess()
        v.update_selection_edit(G.EditPod.default())
        rows = []
        v.enable_timings(True, False)
        for c in cams[:14]:
            v.update_camera(c)
            v.render_frame(v.order_models(models, np.array(centers, np.float32)))   # (returns the models, farthest first)
            tm = v.last_timings()
            rows.append((tm.total_ms, tm.bin_ms, tm.composite_ms, tm.visible, tm.tile_entries))
        r = np.median(np.array(rows[2:], np.float64), axis=0)
        out["config4 3x2M + transforms + edits + mask @1920x1080"] = {
            "frame_ms": r[0], "frames_per_s_one_stream": 1e3 / r[0], "bin_ms": r[1], "composite_ms": r[2],
            "visible": int(r[3]), "tile_entries": int(r[4])}
    return out


def run_ours(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import b200gs as G

    nccl_init_ms = 0.0
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        t0 = time.perf_counter()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        x = torch.zeros(1, device=torch.device("cuda", local_rank))
        dist.all_reduce(x)                      # communicator set-up happens on the first collective: keep it out of
        torch.cuda.synchronize()                # the broadcast's time
        nccl_init_ms = (time.perf_counter() - t0) * 1e3
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    W, H, B, N = a.width, a.height, a.views, a.gaussians
    rb = G.record_bytes(G.SH_NORM8, G.COV3D_HALF)

    # K viewer handles per GPU: frames of a view batch are independent, so they are dealt round-robin to K
    # viewers (one stream each) and the latency-bound kernels of one frame overlap the issue-bound kernels
    # of another.  The viewers SHARE one resident copy of the packed scene (b200gs_model_create_shared).
    K = max(1, a.viewers)
    viewers = [G.Viewer(W, H, G.SH_NORM8, G.COV3D_HALF, device=local_rank) for _ in range(K)]
    models = [viewers[0].add_model("scene", N)]
    streams = [torch.cuda.ExternalStream(vv.stream(), device=dev) for vv in viewers]
    v, m, stream = viewers[0], models[0], streams[0]

    # ---- scene: generated on rank 0, broadcast once over NCCL/NVLink, one resident replica per GPU
    t0 = time.perf_counter()
    packed = None
    if rank == 0:
        ply = G.synth_scene(SEED, N)
        packed = G.pack_gaussians(G.SH_NORM8, G.COV3D_HALF, G.gaussian_from_ply(ply))
        del ply
    scene_build_s = time.perf_counter() - t0
    bcast_ms = upload_ms = 0.0
    if world > 1:
        buf = torch.empty(N * rb, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf.copy_(torch.from_numpy(packed))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(buf, 0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        m.upload_packed_device(0, buf.data_ptr(), N)
        v.sync()
        del buf
    else:
        t1 = time.perf_counter()
        m.upload_packed(0, packed)              # host -> pinned ring -> HBM
        v.sync()
        upload_ms = (time.perf_counter() - t1) * 1e3
    for vv in viewers[1:]:
        models.append(vv.add_shared_model("scene", m))

    # ---- this rank's contiguous block of the 1024-view batch
    cams = G.view_batch()
    lo, hi = G.partition_views(len(cams), world, rank)
    block = cams[lo:hi]
    asp = np.float32(W) / np.float32(H)
    mats = [(c.view(), c.projection(asp)) for c in block]

    img_bytes = W * H * 4
    ring = [torch.empty((B, H, W, 4), dtype=torch.uint8, device=dev) for _ in range(2)]
    # ---- image gather.  p2p: rank 0's (2, world, B, H, W, 4) buffer is symmetric memory mapped into every rank; a rank's
    # finished slot is one peer-to-peer cudaMemcpyAsync over NVLink on a side stream (copy engines: no SM on either side is
    # taken from the renderers, which NCCL's send / recv kernels do — rank 0 receives 7 x 531 MB per step at N = 8).  Falls
    # back to the NCCL gather if symmetric memory cannot be set up on every rank.
    gather_mode, gather_note, sym_local, sym_rank0, copy_stream = "none", "", None, None, None
    if world > 1 and not a.no_gather:
        gather_mode = "nccl"
        if a.gather == "p2p":
            ok = 1
            try:
                import torch.distributed._symmetric_memory as symm_mem
                shape = (2, world, B, H, W, 4)
                sym_local = symm_mem.empty(shape, dtype=torch.uint8, device=dev)
                hdl = symm_mem.rendezvous(sym_local, dist.group.WORLD)
                sym_rank0 = hdl.get_buffer(0, shape, torch.uint8)
                copy_stream = torch.cuda.Stream(device=dev)
            except Exception as e:  # noqa: BLE001
                ok, gather_note = 0, "symmetric memory unavailable: %r" % (e,)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                gather_mode = "p2p"
            else:
                sym_local = sym_rank0 = copy_stream = None
    gathered = [[torch.empty((B, H, W, 4), dtype=torch.uint8, device=dev) for _ in range(world)] for _ in range(2)] \
        if (gather_mode == "nccl" and rank == 0) else None
    pending = [None, None]
    rendered = [torch.cuda.Event() for _ in range(K)]

    def step(s):
        """render B views of this rank's block into ring[s % 2] (the gather of the slot's previous use is done),
        then hand the slot to an asynchronous NCCL gather that runs behind the next step's rendering"""
        slot = s % 2
        if pending[slot] is not None:
            if gather_mode == "p2p":
                for st in streams:
                    st.wait_event(pending[slot])    # the slot's previous copy has left before it is rendered into again
                pending[slot] = None
            else:
                pending[slot].wait()            # (on `stream`: orders later work of viewer 0 behind the gather)
                pending[slot] = None
                done = torch.cuda.Event()
                done.record(stream)
                for st in streams[1:]:
                    st.wait_event(done)
        base = ring[slot].data_ptr()
        for j in range(B):
            view, proj = mats[(s * B + j) % len(mats)]
            k = j % K
            viewers[k].update_camera_matrices(view, proj, (W, H))
            viewers[k].render_frame([models[k]], base + j * img_bytes, W * 4)
        if gather_mode == "p2p":
            for k in range(K):                  # the slot is complete when every viewer's last frame of it is
                rendered[k].record(streams[k])
                copy_stream.wait_event(rendered[k])
            with torch.cuda.stream(copy_stream):
                sym_rank0[slot, rank].copy_(ring[slot], non_blocking=True)
                pending[slot] = torch.cuda.Event()
                pending[slot].record(copy_stream)
        elif gather_mode == "nccl":
            for k in range(1, K):               # the slot is complete when every viewer's last frame of it is
                rendered[k].record(streams[k])
                stream.wait_event(rendered[k])
            pending[slot] = dist.gather(ring[slot], gathered[slot] if rank == 0 else None, dst=0, async_op=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def drain():
        for p in pending:
            if p is not None:
                if gather_mode == "p2p":
                    stream.wait_event(p)
                else:
                    p.wait()
        pending[0] = pending[1] = None
        for st in streams[1:]:
            stream.wait_stream(st)

    with torch.cuda.stream(stream):
        for s in range(a.warmup):
            step(s)
        drain()
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
            time.sleep(0.15)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        launches_before = sum(vv.launch_count() for vv in viewers)
        e0.record(stream)                       # every stream is idle here (barrier above)
        for s in range(a.warmup, a.warmup + a.steps):
            step(s)
        drain()                                 # e1 is behind the last frame of every viewer and the last gather
        e1.record(stream)
        barrier()
        launches_timed = sum(vv.launch_count() for vv in viewers) - launches_before
        elapsed_ms = e0.elapsed_time(e1)
        clocks = sampler.finish() if sampler else None
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())

        # ---- are the gathered bytes right?  rank 0 re-renders, on its own GPU, a sample of the views every rank sent
        # in the last step and compares them byte for byte with what arrived (same view -> same bytes on every rank)
        gather_verified = None
        if gather_mode != "none":
            ok = True
            if rank == 0:
                s_last = a.warmup + a.steps - 1
                check = torch.empty((H, W, 4), dtype=torch.uint8, device=dev)
                for r in range(world):
                    rlo, rhi = G.partition_views(len(cams), world, r)
                    for j in (0, B // 2, B - 1):
                        c = cams[rlo + (s_last * B + j) % (rhi - rlo)]
                        v.update_camera_matrices(c.view(), c.projection(asp), (W, H))
                        v.render_frame([m], check.data_ptr(), W * 4)
                        v.sync()
                        got = sym_local[s_last % 2, r, j] if gather_mode == "p2p" else gathered[s_last % 2][r][j]
                        ok = ok and bool(torch.equal(check, got))
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.broadcast(flag, 0)
            gather_verified = bool(flag.item())
    frames = a.steps * B * world
    value = frames / (elapsed_ms * 1e-3)

    # ---- strong scaling beside the weak number: the WHOLE 1024-view batch, each rank its block, images gathered
    strong = None
    with torch.cuda.stream(stream):
        nsteps = (len(block) + B - 1) // B
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for s in range(nsteps):
            step(s)
        drain()
        s1.record(stream)
        barrier()
        t = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        strong = {"views": nsteps * B * world, "ms": float(t.item()), "frames_per_s": nsteps * B * world / (float(t.item()) * 1e-3),
                  "note": "fixed 1024-view batch divided over the ranks (total work fixed), gather included"}

    # ---- end to end: host camera in, host image out (pinned), D2H inside the timed region.
    # Two frames in flight (b200gs_render_frame_host_begin/_end): the D2H copy of frame i overlaps the
    # rendering of frame i+1; every frame's image still lands in host memory inside the timed region.
    host_img = [[G.PinnedBuffer(img_bytes) for _ in range(2)] for _ in range(K)]

    def e2e_loop(n_frames, first):
        inflight = [0] * K
        for s in range(n_frames):
            k = s % K
            if inflight[k] == 2:                # at most two frames in flight per viewer: retire the oldest
                viewers[k].render_frame_host_end()
                inflight[k] -= 1
            viewers[k].render_frame_host_begin([models[k]], block[(first + s) % len(block)], host_img[k][(s // K) & 1].array)
            inflight[k] += 1
        for k in range(K):
            while inflight[k]:
                viewers[k].render_frame_host_end()
                inflight[k] -= 1

    e2e_frames = a.steps * B
    e2e_loop(32, 0)
    barrier()
    launches1 = sum(vv.launch_count() for vv in viewers)
    t0 = time.perf_counter()
    e2e_loop(e2e_frames, 0)
    barrier()
    e2e_s = time.perf_counter() - t0
    launches_e2e = sum(vv.launch_count() for vv in viewers) - launches1
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = frames / float(t.item())

    # ---- per-stage device times (CUDA events on the viewer's stream), separate short loop on one stream
    spread = {}
    pre_ms, sort_ms, bin_ms, comp_ms, tot_ms, vis, entries, evals = stage_times(G, v, m, mats, W, H, frames=122, spread=spread)

    if rank == 0:
        peak, peak_src = peaks()
        # ALGORITHMIC bytes: preprocess = N*R + 8*V (key + index); sort = 68*V (histogram read + 4 x (read 8 +
        # write 8)).  The kernel also writes this design's intermediates (32-byte projected splat + 4-byte bin word per
        # visible Gaussian): reported separately as frac_with_intermediates.
        b_pre = N * rb + 8 * vis
        b_pre_all = b_pre + 36 * vis
        b_sort = 68 * vis
        ach_pre = b_pre / (pre_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("preprocess_dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "views_per_step_per_gpu": B, "viewers_per_gpu": K, "gaussians": N,
                       "record_bytes": rb, "scene_replicas_per_gpu": 1,
                       "parallelism": "views partitioned over %d GPU(s), scene replicated (one NCCL broadcast); images gathered to rank 0 "
                                      "off the critical path (%s)" % (world, {"p2p": "peer-to-peer copies over NVLink into rank 0's symmetric-memory "
                                      "buffer, copy engines", "nccl": "async NCCL gather", "none": "no gather"}[gather_mode] + (("; " + gather_note) if gather_note else "")),
                       "l2": "inputs larger than L2: the %.0f MB packed scene is re-streamed every frame (L2 = 126 MB); "
                             "camera changes every frame" % (N * rb / 1e6)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 136, "d2h_bytes_per_step": B * img_bytes,
                    "call": "b200gs_render_frame_host_begin/_end, %d viewer(s) x 2 frames in flight (camera pod in, RGBA8 "
                            "image out to pinned host memory)" % K, "gpu_launches": int(launches_e2e)},
            "gpu_launches": int(launches_timed),
            "gather_verified": gather_verified,
            "strong_scaling": strong,
            "roofline": {"bound": "hbm", "kernel": "k_preprocess", "achieved": ach_pre, "peak": peak, "unit": "GB/s",
                         "frac": ach_pre / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_pre,
                         "frac_with_intermediates": (b_pre_all / (pre_ms * 1e-3) / 1e9) / peak,
                         "frac_of_nominal_8TBs": ach_pre / 8000.0},
            "stages": {
                "preprocess_ms": pre_ms, "sort_ms": sort_ms, "bin_ms": bin_ms, "composite_ms": comp_ms, "frame_ms": tot_ms,
                "frames_per_s_one_stream": 1e3 / tot_ms,   # one viewer, frames back to back on its stream
                "visible": int(vis), "tile_entries": int(entries), "splat_evals": int(evals),
                "sort_gkeys_per_s": vis / (sort_ms * 1e-3) / 1e9,
                "sort_hbm_frac": (b_sort / (sort_ms * 1e-3) / 1e9) / peak,
                "pre_plus_sort_hbm_frac": ((b_pre + b_sort) / ((pre_ms + sort_ms) * 1e-3) / 1e9) / peak,
                "composite_gevals_per_s": evals / (comp_ms * 1e-3) / 1e9,
                # evaluations/s against min(FP32 ceiling 148 SMs x 128 lanes x f / ~10 instr, MUFU ceiling 148 x 16 x f)
                "composite_frac_of_fp32_ceiling": (evals / (comp_ms * 1e-3) / 1e9) / min(148 * 128 * 1.965 / 10.0, 148 * 16 * 1.965),
                "frame_ms_p10": spread.get("frame_ms_p10"), "frame_ms_p90": spread.get("frame_ms_p90"),
                "frames_timed": spread.get("frames_timed"),
                "sort_cluster": v.info("sort.cluster"), "sort_resident_clusters": v.info("sort.resident_clusters"),
            },
            "clocks": clocks,
            "setup": {"scene_build_s": scene_build_s, "nccl_init_ms": nccl_init_ms, "scene_broadcast_ms": bcast_ms,
                      "scene_upload_ms": upload_ms,
                      "scene_upload_gb_per_s": (N * rb / 1e9) / (upload_ms * 1e-3) if upload_ms else None},
        }
    for vv in viewers:
        vv.close()
    del ring, gathered, sym_local, sym_rank0
    torch.cuda.empty_cache()
    if rank == 0:
        if world == 1 and not a.no_extra:
            try:
                out["setup"]["ply_stream"] = ply_stream_rate(G, local_rank)
            except Exception as e:  # noqa: BLE001
                out["setup"]["ply_stream"] = {"error": repr(e)}
            try:
                out["extra_configs"] = extra_config_times(G, local_rank)
            except Exception as e:  # noqa: BLE001  (a side table must never cost the headline line)
                out["extra_configs"] = {"error": repr(e)}
        if world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(a, packed, block)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(a, packed, block):
    """The CPU restatement (oracle/) timed beside the GPU number: 2 frames after 1 warm-up, every host core."""
    from oracle import oracle as O
    cores = O.set_num_threads(os.cpu_count())
    model = O.ModelRef(2, 1, packed, a.gaussians)
    asp = np.float32(a.width) / np.float32(a.height)

    def frame(i):
        c = block[i % len(block)]
        f = O.make_frame(c.view(), c.projection(asp), a.width, a.height)
        return O.render_frame(f, [model], front_to_back=False, fp32=True)

    frame(0)
    t0 = time.perf_counter()
    n = 2
    for i in range(n):
        frame(1 + i)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d frames (views 1..%d of rank 0's block) of the same workload after 1 warm-up frame; CPU "
                      "restatement of the reference path in fp32, OpenMP on %d threads" % (n, n, cores),
            "plan_a_probe": plan_a_probe()}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    if world != a.gpus and world == 1 and a.gpus > 1:
        # launched without torchrun: re-exec under it, one rank per GPU
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
