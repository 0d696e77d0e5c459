"""The .CRN writer's palette orderings on the device (csrc/writer_kernels.cuh: greedy chains and the three weighted chains of
optimize_color_endpoints_task, crnlib/crn_comp.cpp:767-933) against the host loops of crn_writer.h: the files must be identical.
crn_gpu_compress_crn takes the device path; CRN_B200_HOST_ORDER=1 makes the hook decline, which is the host path of crn_gpu_crn_write."""
import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
from bench import mip_chain


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("fmt,size,q", [("DXT1", 64, 255), ("DXT1", 64, 60), ("DXT5", 48, 200)])
def test_device_orderings_equal_host_orderings(simctx, monkeypatch, fmt, size, q):
    img = blockgen.smooth_image(size, size - 16, 21 + size, alpha=True)
    levels = mip_chain(img)[:3]
    dev, _, _ = simctx.compress_crn([levels], helpers.CRN_FMT[fmt], quality_level=q)
    monkeypatch.setenv("CRN_B200_HOST_ORDER", "1")
    host, _, _ = simctx.compress_crn([levels], helpers.CRN_FMT[fmt], quality_level=q)
    monkeypatch.delenv("CRN_B200_HOST_ORDER")
    assert dev == host


@pytest.mark.parametrize("fmt", ["DXT1", "DXT5"])
def test_chunked_transition_lists_equal_single_chunk(simctx, monkeypatch, fmt):
    """Transitions::build cuts the blocks into chunks for host threads at >= 65536 blocks; forced here on a small input: same file."""
    img = blockgen.smooth_image(80, 64, 5, alpha=True)
    levels = mip_chain(img)[:3]
    monkeypatch.setenv("CRN_B200_WRITER_CHUNKS", "1")
    one, _, _ = simctx.compress_crn([levels], helpers.CRN_FMT[fmt], quality_level=200)
    monkeypatch.setenv("CRN_B200_WRITER_CHUNKS", "7")
    many, _, _ = simctx.compress_crn([levels], helpers.CRN_FMT[fmt], quality_level=200)
    monkeypatch.delenv("CRN_B200_WRITER_CHUNKS")
    assert one == many
