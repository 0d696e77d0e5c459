"""GPU parity tests of the block-by-block path (the parity tests proper): the nvcc-built library,
called through the C-ABI, against the oracle port, the committed golden vectors, the unmodified
reference (oracle/_ref, when its .so travelled to the box) and size-independent properties at the
full BASELINE.json size."""
import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
from golden.make_golden import case_image

pytestmark = pytest.mark.gpu
GOLD = helpers.golden("pack_golden.json")["cases"]


def _params(c):
    return crn.PackParams(dxt_quality=c["q"], perceptual=c["perc"], use_both_block_types=c["both"])


def test_native_library_loaded(gpu_ctx):
    lib = crn.load_library()
    assert lib.crn_gpu_is_native() == 1 and lib.crn_gpu_device_count() >= 1


@pytest.mark.parametrize("idx", range(len(GOLD)))
def test_gpu_matches_golden(gpu_ctx, idx):
    c = GOLD[idx]
    packed = gpu_ctx.pack_image(c["fmt"], case_image(c), _params(c))
    assert packed[:64].tobytes().hex() == c["head"]
    assert helpers.sha(packed) == c["sha256"]


@pytest.mark.parametrize("family", blockgen.FAMILIES)
@pytest.mark.parametrize("fmt", [0, 1, 3, 4, 6])
def test_gpu_matches_port_blocks(gpu_ctx, port, family, fmt):
    img = helpers.blocks_to_image(blockgen.block_family(family, 1500, 4242))
    for q, perc, both in ((4, 1, 1), (4, 0, 0), (3, 1, 1)):
        a = gpu_ctx.pack_image(fmt, img, crn.PackParams(dxt_quality=q, perceptual=perc, use_both_block_types=both))
        b = helpers.port_pack(port, fmt, img, q, perc, both)
        bad = helpers.mismatching_blocks(a, b, helpers.bytes_per_block(fmt))
        assert bad.size == 0, (family, fmt, q, perc, both, bad[:8], bad.size)


@pytest.mark.parametrize("fmt", [0, 1, 2, 3, 4, 5, 6])
def test_gpu_matches_reference_image(gpu_ctx, ref, fmt):
    """512x384 smooth+noise image (SURVEY 8(d) generator): 12 288 blocks, bit-exact vs the reference."""
    img = blockgen.smooth_image(512, 384, 2048, alpha=True)
    a = gpu_ctx.pack_image(fmt, img)
    b = helpers.ref_pack(ref, fmt, img, threads=7)
    bad = helpers.mismatching_blocks(a, b, helpers.bytes_per_block(fmt))
    assert bad.size == 0, (fmt, bad.size, bad[:8])


def test_gpu_ragged_and_tiny(gpu_ctx, port):
    for fmt in range(7):
        for (w, h, seed) in ((1, 1, 1), (2, 7, 2), (5, 3, 3), (13, 9, 4), (4096, 4, 5), (4, 1024, 6)):
            img = blockgen.smooth_image(w, h, seed, alpha=True)
            a = gpu_ctx.pack_image(fmt, img)
            b = helpers.port_pack(port, fmt, img)
            assert (a == b).all(), (fmt, w, h)


def test_gpu_flat_tiles(gpu_ctx, port):
    """8x8 constant tiles: the solid-colour and <=4-unique-colour shortcuts (SURVEY 8(d) second variant)."""
    img = blockgen.flat_image(512, 256, 7, tile=8)
    img2 = blockgen.flat_image(256, 256, 8, tile=2)
    for im in (img, img2):
        for fmt in (0, 3):
            assert (gpu_ctx.pack_image(fmt, im) == helpers.port_pack(port, fmt, im)).all()


def test_gpu_full_size_properties(gpu_ctx, port):
    """BASELINE configs[0] size (2048x2048): (1) run-to-run determinism, (2) block independence -- packing
    the whole image equals packing 512x512 tiles separately, (3) a 16 384-block window equals the oracle,
    (4) every emitted DXT1 block decodes (colour0/colour1 ordering consistent with the selectors used)."""
    img = blockgen.smooth_image(2048, 2048, 2048, alpha=False)
    full = gpu_ctx.pack_image(0, img).reshape(512, 512, 8)
    again = gpu_ctx.pack_image(0, img).reshape(512, 512, 8)
    assert (full == again).all()
    for ty in (0, 3):
        for tx in (1, 2):
            tile = np.ascontiguousarray(img[ty * 512:(ty + 1) * 512, tx * 512:(tx + 1) * 512])
            t = gpu_ctx.pack_image(0, tile).reshape(128, 128, 8)
            assert (t == full[ty * 128:(ty + 1) * 128, tx * 128:(tx + 1) * 128]).all()
    win = np.ascontiguousarray(img[1024:1536, 512:1024])
    want = helpers.port_pack(port, 0, win).reshape(128, 128, 8)
    assert (want == full[256:384, 128:256]).all()
    lo = full[..., 0].astype(np.uint16) | (full[..., 1].astype(np.uint16) << 8)
    hi = full[..., 2].astype(np.uint16) | (full[..., 3].astype(np.uint16) << 8)
    sel = np.unpackbits(full[..., 4:8], axis=-1, bitorder="little").reshape(512, 512, 16, 2)
    selv = sel[..., 0] | (sel[..., 1] << 1)
    three = lo <= hi
    assert not (selv[three] == 3).any()      # opaque input: index 3 of a 3-colour block is never selected


@pytest.mark.parametrize("q", [0, 1, 2])
@pytest.mark.parametrize("fmt", [0, 1, 3])
def test_gpu_low_quality_levels(gpu_ctx, port, q, fmt):
    """row a6 on the device: evaluate_solution_fast and the superfast / fast / normal flows, bit-exact vs the port and the reference"""
    from test_sim_kernels import check_low_quality
    check_low_quality(gpu_ctx, port, q, fmt)


@pytest.mark.parametrize("q", [1, 3, 4])
@pytest.mark.parametrize("fmt", [0, 1])
def test_gpu_transparent_indices_for_black(gpu_ctx, port, q, fmt):
    """row a8 on the device: try_alpha_as_black_optimization"""
    from test_sim_kernels import check_transparent_for_black
    check_transparent_for_black(gpu_ctx, port, q, fmt)
