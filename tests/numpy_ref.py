"""Independent float64 NumPy restatement of the hot path, written from the text of SURVEY.md §8(c) — NOT from
oracle/gs_oracle.c — as the slower cross-check SURVEY.md §7 step 2 asks for.  Test infrastructure only.

It shares no code with the C oracle: matrices are applied as whole 4x4 products in float64, the covariance
projection is written with explicit 3x3 matrix algebra (J W Σ Wᵀ Jᵀ), the SH basis comes from the closed-form real
spherical harmonics (with the Inria sign convention stated in §8c.3), and the image is blended back to front with
the "over" operator of §8c.8.  Agreement with the C oracle therefore checks the oracle's reading of §8(c), not its
typing.  Layout covered: Norm8 SH + Half Cov3d (the app default), identity or TRS model transform.
"""
import numpy as np

CULL_XY = 1.3          # §8c.5
CLAMP_XY = 1.3         # §8c.7
LOWPASS = 0.3
EXTENT_SIGMA = 3.0
ALPHA_MAX = 0.99
ALPHA_MIN = 1.0 / 255.0


def decode_norm8_half(packed, n):
    """pos f32x3 | colour u8x4 | 48 unorm8 SH (45 used, value = q/255*2-1) | 6 f16 covariance (xx,xy,xz,yy,yz,zz)"""
    rec = np.dtype([("pos", "<f4", 3), ("col", "u1", 4), ("sh", "u1", 48), ("cov", "<f2", 6)])
    assert rec.itemsize == 76
    r = np.frombuffer(np.ascontiguousarray(packed, np.uint8).tobytes(), dtype=rec, count=n)
    pos = r["pos"].astype(np.float64)
    col = r["col"].astype(np.float64) / 255.0
    sh = r["sh"][:, :45].astype(np.float64) / 255.0 * 2.0 - 1.0
    c = r["cov"].astype(np.float64)
    cov = np.empty((n, 3, 3))
    cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2] = c[:, 0], c[:, 1], c[:, 2]
    cov[:, 1, 0], cov[:, 1, 1], cov[:, 1, 2] = c[:, 1], c[:, 3], c[:, 4]
    cov[:, 2, 0], cov[:, 2, 1], cov[:, 2, 2] = c[:, 2], c[:, 4], c[:, 5]
    return pos, col, sh.reshape(n, 15, 3), cov


def quat_to_mat(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def sh_basis(d):
    """Real SH bands 1..3 at unit directions d (n,3), ordered and signed as Inria's computeColorFromSH (§8c.3)."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    c1 = 0.5 * np.sqrt(3.0 / np.pi)
    c2 = [0.5 * np.sqrt(15.0 / np.pi), -0.5 * np.sqrt(15.0 / np.pi), 0.25 * np.sqrt(5.0 / np.pi),
          -0.5 * np.sqrt(15.0 / np.pi), 0.25 * np.sqrt(15.0 / np.pi)]
    c3 = [-0.25 * np.sqrt(35.0 / (2 * np.pi)), 0.5 * np.sqrt(105.0 / np.pi), -0.25 * np.sqrt(21.0 / (2 * np.pi)),
          0.25 * np.sqrt(7.0 / np.pi), -0.25 * np.sqrt(21.0 / (2 * np.pi)), 0.25 * np.sqrt(105.0 / np.pi),
          -0.25 * np.sqrt(35.0 / (2 * np.pi))]
    xx, yy, zz = x * x, y * y, z * z
    b = np.stack([
        -c1 * y, c1 * z, -c1 * x,
        c2[0] * x * y, c2[1] * y * z, c2[2] * (2 * zz - xx - yy), c2[3] * x * z, c2[4] * (xx - yy),
        c3[0] * y * (3 * xx - yy), c3[1] * x * y * z, c3[2] * y * (4 * zz - xx - yy),
        c3[3] * z * (2 * zz - 3 * xx - 3 * yy), c3[4] * x * (4 * zz - xx - yy), c3[5] * z * (xx - yy),
        c3[6] * x * (xx - 3 * yy)], axis=1)
    return b


def render(packed, n, view, proj, W, H, sh_deg=3, size=1.0, model_pos=(0, 0, 0), model_quat=(0, 0, 0, 1),
           model_scale=(1, 1, 1), background=(0, 0, 0, 0), near_to_far=None):
    """Returns (visible indices in ascending order, float64 ndc.z of those, near->far order of indices, float image).
    `near_to_far` (Gaussian indices) overrides the depth order used for blending: float64 depths round to f32 keys
    differently from an f32 chain, so near-ties can swap; the caller first checks that the order it passes in is
    a sort of THESE depths up to that rounding, then compares images under the same order."""
    pos, col, sh, cov = decode_norm8_half(packed, n)
    V = np.asarray(view, np.float64).reshape(4, 4).T        # column-major (glam) -> row-major
    P = np.asarray(proj, np.float64).reshape(4, 4).T
    R = quat_to_mat(model_quat)
    S = np.diag(np.asarray(model_scale, np.float64))
    world = (R @ (S @ pos.T)).T + np.asarray(model_pos, np.float64)
    pv = (V @ np.c_[world, np.ones(n)].T).T
    clip = (P @ pv.T).T
    w = clip[:, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        ndc = clip[:, :3] / w[:, None]
    vis = (w > 0) & (ndc[:, 2] > 0) & (ndc[:, 2] < 1) & (np.abs(ndc[:, 0]) <= CULL_XY) & (np.abs(ndc[:, 1]) <= CULL_XY)
    idx = np.nonzero(vis)[0]
    pv, ndc, world_v = pv[idx], ndc[idx], world[idx]
    m = len(idx)

    # Σ' = (R S) Σ (R S)ᵀ size²; cov2d = J W Σ' Wᵀ Jᵀ + 0.3 I  (§8c.7, view space looks down -z, pixel rows flipped)
    M = R @ S
    Sw = M @ cov[idx] @ M.T * (size * size)
    fx, fy = P[0, 0] * W / 2.0, P[1, 1] * H / 2.0
    tz = -pv[:, 2]
    limx, limy = CLAMP_XY / P[0, 0], CLAMP_XY / P[1, 1]
    tx = np.clip(pv[:, 0] / tz, -limx, limx) * tz
    ty = np.clip(pv[:, 1] / tz, -limy, limy) * tz
    J = np.zeros((m, 2, 3))
    J[:, 0, 0] = fx / tz
    J[:, 0, 2] = fx * tx / (tz * tz)
    J[:, 1, 1] = -fy / tz
    J[:, 1, 2] = -fy * ty / (tz * tz)
    T = J @ V[:3, :3]
    c2 = T @ Sw @ np.transpose(T, (0, 2, 1))
    a, b, d = c2[:, 0, 0] + LOWPASS, c2[:, 0, 1], c2[:, 1, 1] + LOWPASS
    det = a * d - b * b
    ok = det > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        conic = np.stack([d / det, -b / det, a / det], axis=1)
    mid = 0.5 * (a + d)
    lam = mid + np.sqrt(np.maximum(0.1, mid * mid - det))
    radius = np.where(ok, np.ceil(EXTENT_SIGMA * np.sqrt(lam)), 0.0)

    # colour: baked SH0 + bands 1..deg along the WORLD-space view direction (§8c.3)
    cam = -V[:3, :3].T @ V[:3, 3]
    dirs = world_v - cam
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    ncoef = {0: 0, 1: 3, 2: 8, 3: 15}[sh_deg]
    rgb = col[idx, :3] + np.einsum("nk,nkc->nc", sh_basis(dirs)[:, :ncoef], sh[idx, :ncoef])
    rgb = np.clip(rgb, 0.0, 1.0)
    op = col[idx, 3]
    mx = ((ndc[:, 0] + 1.0) * W - 1.0) * 0.5
    my = ((1.0 - ndc[:, 1]) * H - 1.0) * 0.5

    # depth order: ascending ndc.z (as f32, the key the sort sees), ties by ascending index (§8c.6); back-to-front blend
    z32 = ndc[:, 2].astype(np.float32)
    order = np.argsort(z32, kind="stable")
    if near_to_far is not None:
        where = np.full(n, -1, np.int64)
        where[idx] = np.arange(m)
        order = where[np.asarray(near_to_far, np.int64)]
        assert len(order) == m and (order >= 0).all(), "near_to_far must be a permutation of the visible set"
    img = np.empty((H, W, 4))
    img[:] = np.asarray(background, np.float64)
    for k in order[::-1]:
        r = radius[k]
        if r <= 0:
            continue
        x0, x1 = int(max(0.0, np.ceil(mx[k] - r))), int(min(W - 1.0, np.floor(mx[k] + r)))
        y0, y1 = int(max(0.0, np.ceil(my[k] - r))), int(min(H - 1.0, np.floor(my[k] + r)))
        if x0 > x1 or y0 > y1:
            continue
        dx = np.arange(x0, x1 + 1) - mx[k]
        dy = (np.arange(y0, y1 + 1) - my[k])[:, None]
        power = -0.5 * (conic[k, 0] * dx * dx + conic[k, 2] * dy * dy) - conic[k, 1] * dx * dy
        al = np.minimum(ALPHA_MAX, op[k] * np.exp(power))
        al = np.where((power > 0) | (al < ALPHA_MIN), 0.0, al)[:, :, None]
        px = img[y0:y1 + 1, x0:x1 + 1]
        px[:, :, :3] = rgb[k] * al + px[:, :, :3] * (1.0 - al)
        px[:, :, 3:] = al + px[:, :, 3:] * (1.0 - al)
    return idx, ndc[:, 2], idx[order], img
