"""Times the raw device sort (pairs) for depth-like and tile-like keys: python tools/time_sort.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import b200gs as G
rng = np.random.default_rng(0)
with G.Viewer(16, 16) as v:
    for name, n, bits, keys in [("depth 3.7M x32b", 3_700_000, 32, np.float32(rng.uniform(0.9, 1.0, 3_700_000)).view(np.uint32)),
                                ("tile 10.3M x16b", 10_300_000, 16, rng.integers(0, 8160, 10_300_000).astype(np.uint32)),
                                ("random 16M x32b", 16_000_000, 32, rng.integers(0, 1 << 32, 16_000_000, dtype=np.uint64).astype(np.uint32))]:
        k0 = torch.from_numpy(keys.astype(np.int64)).to(torch.int64).cuda().to(torch.int32)  # bit pattern irrelevant for timing
        k0 = torch.from_numpy(keys.view(np.int32)).cuda()
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        ts = []
        for it in range(8):
            k = k0.clone(); vv = vals.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            v.sort_pairs_device(k.data_ptr(), vv.data_ptr(), n, bits)
            ts.append(time.perf_counter() - t0)
        ks = k.cpu().numpy().view(np.uint32)
        mask = (1 << bits) - 1
        assert np.all(np.diff((ks & mask).astype(np.int64)) >= 0)
        print("%s: %.3f ms wall incl. alloc/sync (best of 8)" % (name, 1e3 * min(ts)))
