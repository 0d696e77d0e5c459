"""crunch2_b200/dropin.py: the ctypes binding of the drop-in's C++ symbols (what bench.py's end-to-end figure calls), over the emulator build."""
import os

import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
from crunch2_b200 import dropin

SIM = os.path.join(helpers.ROOT, "tests", "cusim", "libcrnlib_b200_sim.so")


@pytest.fixture(scope="module")
def lib(sim):
    if not os.path.exists(SIM):
        pytest.skip("libcrnlib_b200_sim.so not built (needs the reference's headers)")
    return dropin.load(SIM)


def test_crn_compress_dds_equals_the_c_abi_call(lib, sim):
    from bench import mip_chain
    levels = mip_chain(blockgen.smooth_image(48, 40, 3, alpha=True))
    ctx = crn.Context(0, lib=sim)
    for fmt in (0, 2):
        got, q, rate = dropin.crn_compress([levels], fmt, file_type=dropin.FILE_DDS, quality_level=255, flags=1 | 2 | 8 | 32, lib=lib)
        assert got == ctx.compress_dds([levels], fmt, quality_level=255) and q == 255
        assert rate > 0.0                                   # LZMA bits per texel (liblzma present in this image)
    ctx.close()


def test_crn_compress_crn_round_trip_and_progress(lib, sim):
    from bench import mip_chain
    levels = mip_chain(blockgen.smooth_image(64, 64, 5, alpha=False))
    calls = []
    crn_bytes, q, rate = dropin.crn_compress([levels], 0, file_type=dropin.FILE_CRN, quality_level=100, progress=lambda a, b, c, d, u: (calls.append((a, b, c, d)) or 1), lib=lib)
    assert q == 100 and abs(rate - 8.0 * len(crn_bytes) / sum(l.shape[0] * l.shape[1] for l in levels)) < 1e-4
    assert calls == [(24, 25, 1, 1)]                        # crn_comp::compress_internal's only progress call (crn_comp.cpp:1600)
    dds = dropin.crn_decompress_crn_to_dds(crn_bytes, lib=lib)
    ctx = crn.Context(0, lib=sim)
    assert dds == ctx.crn_to_dds(crn_bytes)
    ctx.close()
    with pytest.raises(RuntimeError):                       # a cancelling callback fails the call (inc/crnlib.h:224-228)
        dropin.crn_compress([levels], 0, file_type=dropin.FILE_CRN, quality_level=100, progress=lambda a, b, c, d, u: 0, lib=lib)
    with pytest.raises(RuntimeError):                       # check() rejects the quality level
        dropin.crn_compress([levels], 0, file_type=dropin.FILE_CRN, quality_level=300, lib=lib)


def ref_compress_mip_params(ref, levels, fmt, mp, quality=255, flags=1 | 2 | 8 | 32, file_type=1):
    import ctypes
    ref.ref_compress_mip_params.restype = ctypes.c_void_p
    h, w = levels[0].shape[:2]
    arrs = [np.ascontiguousarray(l) for l in levels]
    ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    size = ctypes.c_uint32()
    p = ref.ref_compress_mip_params(file_type, fmt, w, h, len(arrs), ptrs, flags, quality, 0, ctypes.byref(mp), ctypes.byref(size))
    if not p:
        return None
    d = ctypes.string_at(p, size.value)
    ref.ref_free(ctypes.c_void_p(p))
    return d


MIP_CASES = {
    "window": dict(m_window_left=8, m_window_top=4, m_window_right=40, m_window_bottom=36),
    "window past the edge": dict(m_window_left=30, m_window_top=20, m_window_right=70, m_window_bottom=52),
    "clamp by cropping": dict(m_clamp_width=32, m_clamp_height=24),
    "clamp by scaling": dict(m_clamp_width=32, m_clamp_height=24, m_clamp_scale=1),
    "absolute": dict(m_scale_mode=1, m_scale_x=40.0, m_scale_y=28.0),
    "relative": dict(m_scale_mode=2, m_scale_x=0.5, m_scale_y=0.75),
    "lower pow2": dict(m_scale_mode=3),
    "nearest pow2": dict(m_scale_mode=4),
    "next pow2": dict(m_scale_mode=5),
    "renormalise mips": dict(m_renormalize=1),
    "renormalise top mip": dict(m_renormalize=1, m_rtopmip=1),
    "no mips + window": dict(m_mode=3, m_window_left=0, m_window_top=0, m_window_right=16, m_window_bottom=16),
    "source mips dropped by a resize": dict(m_mode=1, m_scale_mode=3),
}


def expected_source(ref, img, opts):
    """What create_texture_mipmaps (crnlib/crn_texture_comp.cpp:392-540) makes of level 0, from the reference's own pieces: numpy crop
    (image::extract_block with clamped reads), the size rules restated, image_utils::resample through the shim."""
    from test_mip_cpu import ref_resample
    h, w = img.shape[:2]
    l, t, r, b = (opts.get(k, 0) for k in ("m_window_left", "m_window_top", "m_window_right", "m_window_bottom"))
    if r > l and b > t and l < w and t < h:
        ys = np.minimum(np.arange(t, b), h - 1); xs = np.minimum(np.arange(l, r), w - 1)
        img = np.ascontiguousarray(img[ys][:, xs])
        h, w = img.shape[:2]
    cw, chh, cs = opts.get("m_clamp_width", 0), opts.get("m_clamp_height", 0), opts.get("m_clamp_scale", 0)
    nw, nh = w, h
    if cw and chh and (nw > cw or nh > chh) and not cs:
        nw, nh = min(cw, nw), min(chh, nh)
        img = np.ascontiguousarray(img[:nh, :nw]); h, w = nh, nw
    lower = lambda v: 1 << (v.bit_length() - 1)
    upper = lambda v: v if v & (v - 1) == 0 else 1 << v.bit_length()
    mode = opts.get("m_scale_mode", 0)
    p2 = (nw & (nw - 1)) == 0 and (nh & (nh - 1)) == 0
    if mode == 1:
        nw, nh = int(opts["m_scale_x"]), int(opts["m_scale_y"])
    elif mode == 2:
        nw, nh = int(np.float32(opts["m_scale_x"]) * np.float32(nw) + np.float32(.5)), int(np.float32(opts["m_scale_y"]) * np.float32(nh) + np.float32(.5))
    elif mode == 3 and not p2:
        nw, nh = lower(nw), lower(nh)
    elif mode == 4 and not p2:
        nw = lower(nw) if abs(nw - lower(nw)) < abs(nw - upper(nw)) else upper(nw)
        nh = lower(nh) if abs(nh - lower(nh)) < abs(nh - upper(nh)) else upper(nh)
    elif mode == 5 and not p2:
        nw, nh = upper(nw), upper(nh)
    if cw and chh and (nw > cw or nh > chh) and cs:
        nw, nh = min(cw, nw), min(chh, nh)
    if (nw, nh) != (w, h):
        img = ref_resample(ref, img, nw, nh, filt="kaiser", scale=1.0, srgb=True, gamma=2.2, wrap=False, comps=4, multithreaded=False)
    return img


@pytest.mark.parametrize("case", sorted(MIP_CASES))
def test_mipmap_source_options_match_reference_file(lib, ref, case):
    """create_texture_mipmaps' crop / clamp / rescale / renormalise options (crnlib/crn_texture_comp.cpp:392-540) through the drop-in's
    crn_compress(comp_params, mipmap_params): whole .dds files, block-by-block packing (bit-exact class).  The reference's own overload
    (crn_texture_comp.cpp:580-616) forgets to update m_width / m_height after a crop or resize -- it then reads the new images with the old size
    (out of bounds when they shrank) -- so where the size changes the expectation is assembled from the reference's pieces (crop, size rules,
    image_utils::resample) and its crn_compress on the result; where it does not (renormalise), the overload itself is the expectation."""
    from bench import mip_chain
    img = blockgen.smooth_image(56, 44, 17, alpha=True)
    levels = mip_chain(img)[:3] if "source mips" in case else [img]
    opts = MIP_CASES[case]
    mp = dropin.CrnMipmapParams().clear()
    for k, v in opts.items():
        setattr(mp, k, v)
    got, _, _ = dropin.crn_compress([levels], 0, file_type=dropin.FILE_DDS, quality_level=255, flags=1 | 2 | 8 | 32, lib=lib, mipmap_params=mp)
    if "renormalise" in case:
        want = ref_compress_mip_params(ref, levels, 0, mp)
    else:
        src = expected_source(ref, img, opts)
        plain = dropin.CrnMipmapParams().clear()
        plain.m_mode = opts.get("m_mode", 0)
        want = ref_compress_mip_params(ref, [src], 0, plain)
    assert want is not None
    assert got[:128] == want[:128]              # size and level count first
    assert got == want


def test_mipmap_source_options_cubemap(lib, ref):
    """Six faces: the window is ignored ("Can't crop cubemap textures"), a scale mode resamples every face; .dds block by block against the
    reference's crn_compress on the faces its own resampler produces."""
    import ctypes
    from test_mip_cpu import ref_resample
    faces = [blockgen.smooth_image(24, 24, 300 + f, alpha=False) for f in range(6)]
    mp = dropin.CrnMipmapParams().clear()
    mp.m_mode = 3                                            # cCRNMipModeNoMips
    mp.m_window_left, mp.m_window_top, mp.m_window_right, mp.m_window_bottom = 2, 2, 10, 10
    mp.m_scale_mode = 2; mp.m_scale_x = 0.5; mp.m_scale_y = 0.5
    got, _, _ = dropin.crn_compress([[f] for f in faces], 0, file_type=dropin.FILE_DDS, quality_level=255, flags=1 | 2 | 8 | 32, lib=lib, mipmap_params=mp)
    small = [ref_resample(ref, f, 12, 12, filt="kaiser", scale=1.0, srgb=True, gamma=2.2, wrap=False, comps=3, multithreaded=False) for f in faces]
    want, _, _ = helpers.ref_compress(ref, [[f] for f in small], 0, file_type=1, quality=255, threads=0, flags=1 | 2 | 8 | 32)
    assert got[:128] == want[:128]
    assert got == want
