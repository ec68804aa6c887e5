"""World-size-2 gloo test (CPU) of the multi-GPU host logic: the view batch is partitioned into
contiguous blocks, every rank holds a full scene replica (broadcast once), renders its block with no
collective inside a frame, and the images are gathered to rank 0 in batch order.
The renderer here is the CPU oracle standing in for the CUDA library (this box has no GPU); what is
under test is the partition / broadcast / gather plumbing that bench.py uses with NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, W, H, N_VIEWS = 4000, 96, 54, 6


def _render_views(packed, views):
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    import b200gs as G
    cams = G.view_batch(n_az=3, n_el=2, radii=(4.5,))
    model = O.ModelRef(2, 1, packed, N)
    out = []
    for i in views:
        c = cams[i]
        f = O.make_frame(c.view(), c.projection(np.float32(W) / np.float32(H)), W, H)
        img, _, _ = O.render_frame(f, [model])
        out.append(img)
    return np.stack(out) if out else np.zeros((0, H, W, 4), np.uint8)


def _worker(rank, world, port, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import b200gs as G
    rb = G.record_bytes(2, 1)
    # scene generated on rank 0 only, broadcast once
    buf = torch.zeros(N * rb, dtype=torch.uint8)
    if rank == 0:
        buf.copy_(torch.from_numpy(G.pack_gaussians(2, 1, G.gaussian_from_ply(G.synth_scene(0xB2000001, N)))))
    dist.broadcast(buf, 0)
    lo, hi = G.partition_views(N_VIEWS, world, rank)
    mine = torch.from_numpy(_render_views(buf.numpy(), range(lo, hi)))
    # blocks may be ragged: pad to the largest block for the gather
    per = -(-N_VIEWS // world)
    pad = torch.zeros((per, H, W, 4), dtype=torch.uint8)
    pad[: hi - lo] = mine
    gathered = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, gathered, dst=0)
    if rank == 0:
        parts = []
        for r in range(world):
            l2, h2 = G.partition_views(N_VIEWS, world, r)
            parts.append(gathered[r][: h2 - l2].numpy())
        np.save(result_path, np.concatenate(parts))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_contiguous_and_complete():
    sys.path.insert(0, ROOT)
    import b200gs as G
    for n in (1024, 6, 7, 1):
        for world in (1, 2, 3, 4, 8):
            blocks = [G.partition_views(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert [G.partition_views(1024, 8, r) for r in (0, 7)] == [(0, 128), (896, 1024)]


@pytest.mark.timeout(300)
def test_two_rank_view_batch_matches_single_process(tmp_path, G):
    world = 2
    port = 29600 + os.getpid() % 300
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = np.load(out)
    packed = G.pack_gaussians(2, 1, G.gaussian_from_ply(G.synth_scene(0xB2000001, N)))
    ref = _render_views(packed, range(N_VIEWS))
    assert got.shape == ref.shape == (N_VIEWS, H, W, 4)
    assert np.array_equal(got, ref)          # same view -> same bytes on every rank, batch order kept
    assert len({img.tobytes() for img in ref}) == N_VIEWS   # the views really differ
