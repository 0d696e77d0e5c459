"""Clustered-DDS quantiser building blocks under the SIMT emulator vs the oracle port."""
import ctypes

import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
import quality

P = helpers.P


class MipDesc(ctypes.Structure):
    _fields_ = [("first_block", ctypes.c_uint32), ("block_width", ctypes.c_uint32), ("block_height", ctypes.c_uint32)]


def layout_levels(levels):
    """levels: list of (h,w,4) images -> (blocks (n,16,4), mips [(first, bw, bh)])"""
    blocks, mips, first = [], [], 0
    for img in levels:
        b = quality.image_to_blocks(img)
        bh, bw = (img.shape[0] + 3) // 4, (img.shape[1] + 3) // 4
        mips.append((first, bw, bh)); first += len(b); blocks.append(b)
    return np.ascontiguousarray(np.concatenate(blocks)), mips


def port_training(port, kind, comp, blocks, mips, hierarchical=1):
    n = len(blocks); D = 2 if kind else 6
    vecs = np.zeros((n, D), np.uint8); w = np.zeros(n, np.uint32)
    nch = sum(((bw + 1) // 2) * ((bh + 1) // 2) for _, bw, bh in mips)
    enc = np.zeros(nch, np.uint8)
    m = np.array(mips, np.uint32).ravel()
    port.op_qdxt_training(kind, comp, P(blocks), n, P(m), len(mips), hierarchical, P(vecs), P(w), P(enc))
    return vecs, w, enc


def gpu_training(ctx, kind, comp, blocks, mips):
    n = len(blocks); D = 2 if kind else 6
    vecs = np.zeros((n, D), np.uint8); w = np.zeros(n, np.uint32)
    nch = sum(((bw + 1) // 2) * ((bh + 1) // 2) for _, bw, bh in mips)
    enc = np.zeros(nch, np.uint8)
    arr = (MipDesc * len(mips))(*[MipDesc(*m) for m in mips])
    ctx._check(ctx._lib.crn_gpu_qdxt_training(ctx._ctx, kind, comp, blocks.ctypes.data, n, arr, len(mips), vecs.ctypes.data, w.ctypes.data, enc.ctypes.data))
    ctx.synchronize()
    return vecs, w, enc


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


@pytest.mark.parametrize("kind,comp", [(0, 3), (1, 3), (1, 0)])
def test_training_vectors_match_port(simctx, port, kind, comp):
    from bench import mip_chain
    for (w, h, seed, flat) in ((64, 64, 1, False), (40, 24, 2, False), (13, 9, 3, False), (64, 32, 4, True)):
        base = blockgen.flat_image(w, h, seed, tile=8) if flat else blockgen.smooth_image(w, h, seed, alpha=True)
        if flat:
            base[::3, ::2, :3] //= 2
        blocks, mips = layout_levels(mip_chain(base))
        a = gpu_training(simctx, kind, comp, blocks, mips)
        b = port_training(port, kind, comp, blocks, mips)
        assert (a[2] == b[2]).all(), ("encoding", kind, w, h, np.nonzero(a[2] != b[2])[0][:5])
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all(), (kind, w, h)
