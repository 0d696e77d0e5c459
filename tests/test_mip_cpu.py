"""Resampler / mip-chain generation under the SIMT emulator, bit-exact against the reference's image_utils::resample in
its task-pool form (threaded_resampler), which is what crn_compress_mip_chain runs."""
import ctypes

import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers

P = helpers.P


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


def ref_resample(ref, img, dw, dh, filt="kaiser", scale=0.9, srgb=True, gamma=2.2, wrap=False, comps=4, multithreaded=True):
    h, w = img.shape[:2]
    out = np.zeros((dh, dw, 4), np.uint8)
    ok = ref.ref_resample(P(np.ascontiguousarray(img)), w, h, P(out), dw, dh, filt.encode(), ctypes.c_float(scale), int(srgb), ctypes.c_float(gamma), int(wrap), comps,
                          int(multithreaded))
    assert ok
    return out


def ref_mips(ref, img, **kw):
    h, w = img.shape[:2]
    levels, l = [img], 1
    while (w >> (l - 1)) > 1 or (h >> (l - 1)) > 1:
        levels.append(ref_resample(ref, img, max(1, w >> l), max(1, h >> l), **kw)); l += 1
    return levels


@pytest.mark.parametrize("size", [(64, 64), (40, 24), (37, 19), (16, 1)])
def test_mip_chain_matches_reference(simctx, ref, size):
    w, h = size
    img = blockgen.smooth_image(w, h, 17 + w, alpha=True)
    got = simctx.generate_mipmaps(img)
    want = ref_mips(ref, img)
    assert len(got) == len(want)
    for l, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape and np.array_equal(a, b), "level %d differs (max abs diff %d)" % (l, int(np.abs(a.astype(int) - b.astype(int)).max()))


@pytest.mark.parametrize("filt", ["box", "tent", "lanczos4", "mitchell", "kaiser"])
@pytest.mark.parametrize("srgb,wrap,comps", [(True, False, 4), (False, True, 3)])
def test_mip_filters_and_modes(simctx, ref, filt, srgb, wrap, comps):
    img = blockgen.smooth_image(48, 32, 5, alpha=True)
    got = simctx.generate_mipmaps(img, filter=filt, filter_scale=1.0, srgb=srgb, wrapping=wrap, num_comps=comps, max_levels=4)
    want = ref_mips(ref, img, filt=filt, scale=1.0, srgb=srgb, wrap=wrap, comps=comps)[:4]
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_mip_level_count_and_errors(simctx, sim):
    assert sim.crn_gpu_mip_level_count(4096, 4096, 1, 0) == 13
    assert sim.crn_gpu_mip_level_count(4096, 4096, 1, 5) == 5
    assert sim.crn_gpu_mip_level_count(37, 19, 1, 16) == 6
    assert sim.crn_gpu_mip_level_count(64, 64, 8, 0) == 4
    with pytest.raises(crn.CrnGpuError):
        simctx.generate_mipmaps(np.zeros((8, 8, 4), np.uint8), filter=7)
