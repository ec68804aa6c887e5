"""Shared helpers for the parity tests."""
import numpy as np

SEED_100K, SEED_1M, SEED_6M = 0xB2000001, 0xB2000002, 0xB2000006
SEEDS_CFG4 = (0xB2000041, 0xB2000042, 0xB2000043)

# RGBA tolerance stated by BASELINE.json's north_star
MAX_ABS_DIFF = 2       # per channel, in 1/255 units
MIN_PSNR_DB = 50.0


def psnr(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    mse = float((d * d).mean())
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)


def assert_image_close(img, ref):
    d = np.abs(img.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= MAX_ABS_DIFF, "max |d| = %d > %d at %s" % (d.max(), MAX_ABS_DIFF, np.unravel_index(d.argmax(), d.shape))
    assert psnr(img, ref) >= MIN_PSNR_DB, "PSNR %.2f dB < %.1f" % (psnr(img, ref), MIN_PSNR_DB)


def scene(mod, seed, n, sh=2, cov3d=1):
    """(ply, gaussians, packed) of the synthetic scene from module `mod` (oracle or product)."""
    ply = mod.synth_scene(seed, n)
    g = mod.gaussian_from_ply(ply)
    packed = mod.pack(sh, cov3d, g) if hasattr(mod, "pack") else mod.pack_gaussians(sh, cov3d, g)
    return ply, g, packed


def make_gaussians(dtype, pos, scale=0.05, rot=(0, 0, 0, 1), color=(255, 255, 255, 255), sh=None):
    """Hand-made Gaussians for known-answer tests."""
    pos = np.atleast_2d(np.asarray(pos, np.float32))
    n = len(pos)
    g = np.zeros(n, dtype=dtype)
    g["pos"] = pos
    g["rot"] = np.broadcast_to(np.asarray(rot, np.float32), (n, 4))
    g["scale"] = np.broadcast_to(np.asarray(scale, np.float32), (n, 3)) if np.ndim(scale) <= 1 else scale
    g["color"] = np.broadcast_to(np.asarray(color, np.uint8), (n, 4))
    if sh is not None:
        g["sh"] = sh
    return g


def bits_set(words, n):
    return np.unpackbits(np.asarray(words, dtype="<u4").view(np.uint8), bitorder="little")[:n].astype(bool)


def pack_bits(flags):
    flags = np.asarray(flags, bool)
    n = len(flags)
    pad = np.zeros(((n + 31) // 32) * 32, bool)
    pad[:n] = flags
    return np.packbits(pad, bitorder="little").view("<u4").copy()
