"""Raw sort API against numpy's stable sort: python tools/sort_check.py  (GPU box; run under `timeout`)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import b200gs as G
rng = np.random.default_rng(7)
WIDE = os.environ.get("WIDE", "1") == "1"
if len(sys.argv) > 1:
    G.set_tuning("sort.cluster", int(sys.argv[1]))
ok = True
with G.Viewer(64, 64) as v:
    for n in (1, 33, 4096, 4097, 32768, 32769, 100_000, 1_000_003, 5_900_000):
        for bits, hi in ((16, 1 << 16), (32, 1 << 32), (32, 1 << 11), (32, 1)):
            keys = rng.integers(0, hi, n, dtype=np.uint64).astype(np.uint32)
            if hi == 1 << 32:   # depth-like: [0.8, 0.99)
                keys = rng.uniform(0.8, 0.99, n).astype(np.float32).view(np.uint32)
            vals = np.arange(n, dtype=np.uint32)
            t0 = time.time()
            k2, v2 = v.sort_pairs(keys, vals, bits, wide=WIDE)
            order = np.argsort(keys, kind="stable")
            good = np.array_equal(k2, keys[order]) and np.array_equal(v2, vals[order])
            ok &= good
            print("n=%d bits=%d range=%d %s %.3fs" % (n, bits, hi, "ok" if good else "MISMATCH", time.time() - t0), flush=True)
print("ALL OK" if ok else "FAILED")
