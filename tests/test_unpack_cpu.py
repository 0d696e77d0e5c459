"""DXTn -> RGBA unpack kernel (crn_gpu_unpack_image) under the SIMT emulator, bit-exact against the reference's
dxt_image::unpack (oracle/_ref) and against the numpy decoder of tests/quality.py (which thereby gets pinned too)."""
import numpy as np
import pytest

import crunch2_b200 as crn
import helpers
import quality

P = helpers.P


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


def ref_unpack(ref, fmt, blocks, w, h):
    out = np.zeros((h, w, 4), np.uint8)
    assert ref.ref_unpack_image(fmt, P(blocks), w, h, P(out))
    return out


def random_blocks(fmt, w, h, seed):
    """Random bit patterns: every selector value, both block types, equal endpoints."""
    rng = np.random.default_rng(seed)
    n = ((w + 3) // 4) * ((h + 3) // 4)
    b = rng.integers(0, 256, (n, helpers.bytes_per_block(fmt)), dtype=np.uint8)
    b[::7, 0:2] = b[::7, 2:4]                      # color0 == color1 / alpha0 == alpha1 cases
    return np.ascontiguousarray(b)


@pytest.mark.parametrize("fmt", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("size", [(32, 16), (13, 9), (4, 4), (1, 1)])
def test_unpack_matches_reference(simctx, ref, fmt, size):
    w, h = size
    blocks = random_blocks(fmt, w, h, 100 * fmt + w)
    got = simctx.unpack_image(fmt, blocks, w, h)
    want = ref_unpack(ref, fmt, blocks, w, h)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("fmt", [0, 3, 4, 5, 6])
def test_numpy_decoder_matches_reference(ref, fmt):
    """tests/quality.py::decode_blocks is what the tolerance tests measure PSNR with: pin it on the channels it fills."""
    w, h = 32, 32
    blocks = random_blocks(fmt, w, h, 7 + fmt)
    want = quality.image_to_blocks(ref_unpack(ref, fmt, blocks, w, h))
    got = quality.decode_blocks(blocks.tobytes(), fmt)
    ch = {0: [0, 1, 2, 3], 3: [0, 1, 2, 3], 4: [3], 5: [0, 1], 6: [0, 1]}[fmt]
    assert np.array_equal(got[..., ch], want[..., ch])


def test_unpack_roundtrip_of_packed_image(simctx, ref):
    import blockgen
    img = blockgen.smooth_image(24, 20, 3, alpha=True)
    for fmt in (0, 3):
        packed = simctx.pack_image(fmt, img)
        got = simctx.unpack_image(fmt, packed, 24, 20)
        assert np.array_equal(got, ref_unpack(ref, fmt, np.frombuffer(packed, np.uint8), 24, 20))
        assert quality.psnr(got, img, [0, 1, 2]) > 20.0


def test_unpack_rejects_bad_arguments(simctx):
    with pytest.raises(ValueError):
        simctx.unpack_image(3, np.zeros(8, np.uint8), 4, 4)
    with pytest.raises(crn.CrnGpuError):
        simctx.unpack_image(9, np.zeros(64, np.uint8), 4, 4)
