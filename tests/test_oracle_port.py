"""CPU tests of the oracle: the plain-C port (oracle/port) against (1) the committed golden vectors
generated from the unmodified reference and (2), where oracle/_ref exists, the reference itself."""
import ctypes

import numpy as np
import pytest

import blockgen
import helpers
from golden.make_golden import case_image

GOLD = helpers.golden("pack_golden.json")["cases"]


@pytest.mark.parametrize("idx", range(0, len(GOLD), 1))
def test_port_matches_golden(port, idx):
    c = GOLD[idx]
    packed = helpers.port_pack(port, c["fmt"], case_image(c), c["q"], c["perc"], c["both"])
    assert packed[:64].tobytes().hex() == c["head"]
    assert helpers.sha(packed) == c["sha256"]


@pytest.mark.parametrize("family", blockgen.FAMILIES)
def test_port_matches_reference_blocks(port, ref, family):
    blocks = blockgen.block_family(family, 64, 977)
    img = helpers.blocks_to_image(blocks)
    for fmt in (0, 1, 3, 5):
        for q in (4, 3, 2, 1, 0):
            for perc, both in ((1, 1), (0, 0)):
                a = helpers.ref_pack(ref, fmt, img, q, perc, both)
                b = helpers.port_pack(port, fmt, img, q, perc, both)
                bad = helpers.mismatching_blocks(a, b, helpers.bytes_per_block(fmt))
                assert bad.size == 0, (family, fmt, q, perc, both, bad[:8])


def test_port_matches_reference_flags(port, ref):
    """transparent-indices-for-black and multi-threaded reference packing (thread-count independence
    with caching disabled)."""
    img = blockgen.smooth_image(64, 48, 5, alpha=True)
    img[::5, ::3, :3] = 2
    a = helpers.ref_pack(ref, 0, img, 4, 1, 1, tfb=1)
    b = helpers.port_pack(port, 0, img, 4, 1, 1, tfb=1)
    assert (a == b).all()
    c = helpers.ref_pack(ref, 3, img, 4, 1, 1, threads=3)
    d = helpers.port_pack(port, 3, img, 4, 1, 1)
    assert (c == d).all()


def test_port_n_pixel_clusters(port, ref):
    """N-pixel form used by the clustered paths (16*|cluster| pixels): port == reference."""
    class Prm(ctypes.Structure):
        _fields_ = [(n, ctypes.c_uint32) for n in "quality perceptual pixels_have_alpha use_alpha_blocks alpha_threshold grayscale transparent_for_black force_alpha_blocks".split()]

    class Res(ctypes.Structure):
        _fields_ = [("error", ctypes.c_uint64), ("low", ctypes.c_uint16), ("high", ctypes.c_uint16), ("alpha_block", ctypes.c_uint8)]
    P = helpers.P
    for fam in ("smooth", "noise", "four", "dark"):
        for nblk in (2, 5, 40):
            px = np.ascontiguousarray(blockgen.block_family(fam, nblk, 31).reshape(-1, 4))
            n = len(px)
            for uab in (0, 1):
                lo = ctypes.c_uint16(); hi = ctypes.c_uint16(); err = ctypes.c_uint64(); ab = ctypes.c_uint8()
                sel = np.zeros(n, np.uint8)
                ref.ref_dxt1_optimize(P(px), n, 4, 1, 0, uab, 128, 0, 0, 0, ctypes.byref(lo), ctypes.byref(hi), P(sel), ctypes.byref(err), ctypes.byref(ab))
                p = Prm(4, 1, 0, uab, 128, 0, 0, 0); r = Res(); sel2 = np.zeros(n, np.uint8)
                port.op_dxt1_optimize(P(px), n, ctypes.byref(p), ctypes.byref(r), P(sel2))
                assert (lo.value, hi.value, err.value, ab.value) == (r.low, r.high, r.error, r.alpha_block)
                assert (sel == sel2).all()
            for comp in (0, 3):
                f = np.zeros(1, np.uint8); s = np.zeros(1, np.uint8); sel = np.zeros(n, np.uint8); e = np.zeros(1, np.uint64); bt = np.zeros(1, np.uint8)
                ref.ref_dxt5_optimize_batch(P(px), 1, n, comp, 4, 1, P(f), P(s), P(sel), P(e), P(bt))
                f2 = ctypes.c_uint8(); s2 = ctypes.c_uint8(); e2 = ctypes.c_uint64(); bt2 = ctypes.c_uint8(); sel2 = np.zeros(n, np.uint8)
                port.op_dxt5_optimize(P(px), n, comp, 4, 1, ctypes.byref(f2), ctypes.byref(s2), P(sel2), ctypes.byref(e2), ctypes.byref(bt2))
                assert (f[0], s[0], e[0], bt[0]) == (f2.value, s2.value, e2.value, bt2.value)
                assert (sel == sel2).all()
