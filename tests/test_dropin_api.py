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
