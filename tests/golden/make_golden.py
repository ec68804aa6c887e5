"""Writes tests/golden/*.npz from the CPU oracle (run from the repo root:
    python tests/golden/make_golden.py).
The reference holds no golden vectors for this path and cannot be built here, so
these fixtures are self-generated: they guard the oracle against drift and let the GPU tests
compare against committed bytes without running the oracle's compositor at size."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def main():
    seed, n, W, H = 0xB2000001, 3000, 160, 96
    out = dict(seed=np.uint64(seed), n=np.int64(n), W=np.int64(W), H=np.int64(H))
    ply = O.synth_scene(seed, n)
    g = O.gaussian_from_ply(ply)
    for sh, cov in ((2, 1), (1, 0), (0, 0), (3, 1)):
        packed = O.pack(sh, cov, g)
        view, proj = O.orbit_camera(width=W, height=H)
        f = O.make_frame(view, proj, W, H)
        idx, keys, spl = O.preprocess(f, O.ModelRef(sh, cov, packed, n))
        keys, idx, spl = O.sort(keys, idx, spl)
        img, _ = O.composite(f, spl, False)
        tag = "%d%d" % (sh, cov)
        out["idx_" + tag], out["keys_" + tag], out["img_" + tag] = idx, keys, img
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), **out)
    print("wrote oracle_small.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
