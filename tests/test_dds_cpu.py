"""DDS container edge: crn_gpu_dds_header and crn_gpu_crn_to_dds (crn_decompress_crn_to_dds) against the reference's
crn_decompress_crn_to_dds (oracle/_ref), byte for byte; the transcode runs under the SIMT emulator here."""
import ctypes

import numpy as np
import pytest

import crnsynth
import crunch2_b200 as crn
import helpers


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


def ref_to_dds(ref, data):
    buf = np.frombuffer(data, np.uint8)
    size = ctypes.c_uint32()
    p = ref.ref_crn_to_dds(helpers.P(buf), len(data), ctypes.byref(size))
    assert p
    d = ctypes.string_at(p, size.value)
    ref.ref_free(ctypes.c_void_p(p))
    return d


CASES = [(fmt, w, h, lv, faces) for fmt in ("DXT1", "DXT5", "DXN_XY", "DXN_YX", "DXT5A") for (w, h, lv, faces) in ((64, 32, None, 1), (20, 12, 1, 1), (16, 16, 3, 6), (8, 8, 1, 6))]


@pytest.mark.parametrize("fmt,w,h,lv,faces", CASES)
def test_crn_to_dds_matches_reference(simctx, ref, fmt, w, h, lv, faces):
    data = crnsynth.synth_crn(w, h, fmt, levels=lv, faces=faces, seed=w + h, n_color_ep=16, n_color_sel=16, n_alpha_ep=16, n_alpha_sel=16)
    want = ref_to_dds(ref, data)
    got = simctx.crn_to_dds(data)
    assert got[:128] == want[:128]
    assert got == want


def test_header_arguments(sim):
    out = (ctypes.c_uint8 * 128)()
    assert sim.crn_gpu_dds_header(0, 64, 64, 7, 1, out) == 0
    assert bytes(out[:4]) == b"DDS "
    assert sim.crn_gpu_dds_header(1, 64, 64, 7, 1, out) != 0      # DXT3 never comes out of a .crn
    assert sim.crn_gpu_dds_header(0, 0, 64, 7, 1, out) != 0
    assert sim.crn_gpu_dds_header(0, 64, 64, 1, 2, out) != 0


def test_crn_to_dds_rejects_garbage(simctx):
    with pytest.raises(crn.CrnGpuError):
        simctx.crn_to_dds(b"not a crn file at all, just some bytes to get past any minimum size check........................")
