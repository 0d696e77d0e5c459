"""Deterministic synthetic 4x4 block / image generators shared by the parity tests.

The families follow SURVEY.md 8(d): smooth low-frequency fields + noise (16 unique colours per block,
the optimiser's worst case), flat tiles (solid and <=4 unique colour shortcuts), plus edge families the
reference's code paths special-case (two/three/four-colour blocks that exercise
try_combinatorial_encoding, near-black blocks, saturated reds/blues that move the perceptual weights,
transparent pixels for DXT1A)."""
import numpy as np


def smooth_image(w, h, seed, alpha=False):
    """Three low-frequency sin / cos fields per channel + N(0, 10) noise.  The fields are separable in x, y and x + y, so
    they are evaluated on 1-D axes and broadcast / gathered -- bit-identical to evaluating them on the full grid, and several times faster."""
    rng = np.random.default_rng(seed)
    xs = np.arange(w, dtype=np.float64); ys = np.arange(h, dtype=np.float64); ss = np.arange(w + h - 1, dtype=np.float64)
    xy = (np.arange(h)[:, None] + np.arange(w)[None, :])
    chans = []
    for c in range(3):
        p1, p2, p3 = rng.uniform(43, 177, 3)
        f = (np.sin(xs / p1 + c)[None, :] + np.cos(ys / p2 - c)[:, None]) + np.sin(ss / p3)[xy]
        f = (f - f.min()) / max(f.max() - f.min(), 1e-9)
        chans.append(27 + 200 * f + rng.normal(0, 10, (h, w)))
    if alpha:
        a = (128 + 64 * np.sin(xs / 17.0))[None, :] + rng.normal(0, 8, (h, w))
    else:
        a = np.full((h, w), 255.0)
    img = np.stack(chans + [a], axis=-1)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def flat_image(w, h, seed, tile=8):
    rng = np.random.default_rng(seed)
    th, tw = (h + tile - 1) // tile, (w + tile - 1) // tile
    t = rng.integers(0, 256, (th, tw, 4), dtype=np.uint8)
    t[..., 3] = 255
    return np.repeat(np.repeat(t, tile, 0), tile, 1)[:h, :w].copy()


def block_family(kind, n, seed):
    """n blocks of 16 RGBA pixels, shape (n,16,4) uint8."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        b = rng.integers(0, 256, (n, 16, 4), dtype=np.uint8)
    elif kind == "smooth":
        base = rng.integers(0, 256, (n, 1, 3)).astype(np.float64)
        d = rng.normal(0, 1, (n, 1, 3)) * rng.uniform(0, 60, (n, 1, 1))
        t = rng.uniform(-1, 1, (n, 16, 1))
        rgb = base + d * t + rng.normal(0, rng.uniform(0, 8, (n, 1, 1)), (n, 16, 3))
        a = rng.integers(0, 256, (n, 1, 1)) + rng.normal(0, 20, (n, 16, 1))
        b = np.clip(np.rint(np.concatenate([rgb, a], -1)), 0, 255).astype(np.uint8)
    elif kind in ("two", "three", "four", "solid"):
        k = {"solid": 1, "two": 2, "three": 3, "four": 4}[kind]
        pal = rng.integers(0, 256, (n, k, 4), dtype=np.uint8)
        idx = rng.integers(0, k, (n, 16))
        b = np.take_along_axis(pal, idx[..., None].repeat(4, -1), 1)
    elif kind == "dxt_like":   # pixels that are exact DXT1 palette entries -> combinatorial recovery path
        e0 = rng.integers(0, 256, (n, 1, 3)); e1 = rng.integers(0, 256, (n, 1, 3))
        e0 = (e0 >> 3 << 3) | (e0 >> 5); e1 = (e1 >> 3 << 3) | (e1 >> 5)
        pal = np.concatenate([e0, e1, (2 * e0 + e1) // 3, (2 * e1 + e0) // 3], 1)
        idx = rng.integers(0, 4, (n, 16))
        rgb = np.take_along_axis(pal, idx[..., None].repeat(3, -1), 1)
        b = np.concatenate([rgb, np.full((n, 16, 1), 255)], -1).astype(np.uint8)
    elif kind == "dark":
        b = rng.integers(0, 12, (n, 16, 4), dtype=np.uint8)
        b[:, ::3, :3] = rng.integers(0, 256, (n, 6, 3), dtype=np.uint8)
    elif kind == "saturated":
        b = rng.integers(0, 40, (n, 16, 4), dtype=np.uint8)
        ch = rng.integers(0, 2, n) * 2
        for i in range(n):
            b[i, :, ch[i]] = rng.integers(150, 256, 16)
    elif kind == "alpha_mix":  # some pixels under the DXT1A threshold
        b = rng.integers(0, 256, (n, 16, 4), dtype=np.uint8)
        b[..., 3] = np.where(rng.random((n, 16)) < 0.3, rng.integers(0, 128, (n, 16)), 255)
    elif kind == "gray":
        g = rng.integers(0, 256, (n, 16, 1), dtype=np.uint8)
        b = np.concatenate([g, g, g, rng.integers(0, 256, (n, 16, 1), dtype=np.uint8)], -1)
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(b, dtype=np.uint8)


FAMILIES = ["noise", "smooth", "two", "three", "four", "solid", "dxt_like", "dark", "saturated", "alpha_mix", "gray"]
