"""GPU parity tests of the CRN -> DXTn transcoder through the C-ABI: golden .crn files from the reference
compressor, synthetic files up to BASELINE.json's 8192x8192 DXT5 size against the reference decoder
(oracle/_ref) or the oracle port, the crnd_unpack_level contract, and the batched launch."""
import os

import numpy as np
import pytest

import crnsynth
import crunch2_b200 as crn
import helpers
from test_transcode_cpu import GOLD, SYNTH, load, shas, split_levels

pytestmark = pytest.mark.gpu


def device_buffer(nbytes):
    import torch
    return torch.zeros(int(nbytes), dtype=torch.uint8, device="cuda:0")


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_gpu_matches_golden_crn(gpu_ctx, case):
    tex = gpu_ctx.unpack_begin(load(case))
    assert shas(split_levels(tex, tex.unpack_all())) == case["sha256"]
    tex.close()


@pytest.mark.parametrize("fmt,w,h,faces,kw", SYNTH + [("DXT1", 1024, 512, 1, {}), ("DXN_XY", 512, 512, 6, {}), ("DXT5A", 2048, 64, 1, {})])
def test_gpu_synthetic_crn(gpu_ctx, port, fmt, w, h, faces, kw):
    data = crnsynth.synth_crn(w, h, fmt, faces=faces, seed=11, **kw)
    tex = gpu_ctx.unpack_begin(data)
    assert split_levels(tex, tex.unpack_all()) == helpers.port_unpack_all(port, data)
    tex.close()


def test_gpu_full_size_8192_dxt5(gpu_ctx, port):
    """BASELINE configs[3]: 8192x8192 DXT5, all 14 levels, bit-exact vs the reference decoder."""
    data = crnsynth.synth_crn(8192, 8192, "DXT5", seed=4, with_crc=False, n_color_ep=4096, n_color_sel=4096, n_alpha_ep=2048, n_alpha_sel=2048)
    ref = helpers.load_ref()
    want = helpers.ref_unpack_all(ref, data) if ref is not None else helpers.port_unpack_all(port, data)
    tex = gpu_ctx.unpack_begin(data)
    assert tex.info["levels"] == 14 and tex.total_size == sum(len(f) for lv in want for f in lv)
    got = tex.unpack_all()
    flat = np.frombuffer(b"".join(b"".join(lv) for lv in want), np.uint8)
    assert got.size == flat.size
    bad = np.nonzero(got != flat)[0]
    assert bad.size == 0, (bad.size, bad[:4])
    tex.close()


def test_gpu_unpack_level_contract(gpu_ctx, port):
    data = load([c for c in GOLD if c["name"] == "dxt1_cube_32_mips"][0])
    want = helpers.port_unpack_all(port, data)
    tex = gpu_ctx.unpack_begin(data)
    bx, by = tex.level_blocks(1)
    pitch = bx * 8 + 16
    bufs = [device_buffer(pitch * by).fill_(0xEE) for _ in range(6)]
    tex.unpack_level_device([b.data_ptr() for b in bufs], pitch * by, pitch, 1)
    gpu_ctx.synchronize()
    for f in range(6):
        rows = bufs[f].cpu().numpy().reshape(by, pitch)
        assert rows[:, :bx * 8].tobytes() == want[1][f]
        assert (rows[:, bx * 8:] == 0xEE).all()
    with pytest.raises(crn.CrnGpuError):
        tex.unpack_level_device([b.data_ptr() for b in bufs], 8, pitch, 1)
    tex.close()


def test_gpu_batch(gpu_ctx, port):
    files = [crnsynth.synth_crn(256, 256, f, seed=20 + i) for i, f in enumerate(["DXT1", "DXT5", "DXN_XY", "DXT5A"] * 8)]
    texs = [gpu_ctx.unpack_begin(d) for d in files]
    outs = [device_buffer(t.total_size) for t in texs]
    gpu_ctx.unpack_batch(texs, [o.data_ptr() for o in outs], [t.total_size for t in texs])
    for t, o, d in zip(texs, outs, files):
        assert split_levels(t, o.cpu().numpy()) == helpers.port_unpack_all(port, d)
        t.close()


@pytest.mark.parametrize("min_blocks", ["1", "4096"])
@pytest.mark.parametrize("fmt,w,h,faces", [("DXT1", 2048, 1024, 1), ("DXT5", 1024, 1024, 1), ("DXN_XY", 512, 512, 6), ("DXT5A", 4096, 256, 1), ("DXT5", 260, 100, 1)])
def test_gpu_wide_path(gpu_ctx, port, monkeypatch, fmt, w, h, faces, min_blocks):
    """Levels through transcode_wide.cuh (all of them with the threshold at 1; the default threshold otherwise),
    bit-exact against the oracle, and the same bytes as the warp-per-level kernel gives."""
    data = crnsynth.synth_crn(w, h, fmt, faces=faces, seed=31, skew=0.1)
    want = helpers.port_unpack_all(port, data)
    monkeypatch.setenv("CRN_B200_WIDE_MIN_BLOCKS", min_blocks)
    tex = gpu_ctx.unpack_begin(data)
    l0 = gpu_ctx.launch_count
    got = tex.unpack_all()
    wide_launches = gpu_ctx.launch_count - l0
    assert split_levels(tex, got) == want
    monkeypatch.setenv("CRN_B200_WIDE_MIN_BLOCKS", "4000000000")
    l0 = gpu_ctx.launch_count
    narrow = tex.unpack_all()
    bx, by = tex.level_blocks(0)
    assert gpu_ctx.launch_count - l0 == 1 and (wide_launches >= 2 or (((bx + 1) & ~1) * ((by + 1) & ~1) * faces < int(min_blocks)))
    assert np.array_equal(got, narrow)
    tex.close()


def test_gpu_wide_batch(gpu_ctx, port):
    """Batch of files whose large levels take the wide path (more levels than SM pairs: walk and resolve as two launches)."""
    fmts = ["DXT1", "DXT5", "DXN_XY", "DXT5A"]
    files = [crnsynth.synth_crn(512, 512, fmts[i % 4], seed=60 + i, skew=0.1) for i in range(48)]
    texs = [gpu_ctx.unpack_begin(d) for d in files]
    outs = [device_buffer(t.total_size) for t in texs]
    gpu_ctx.unpack_batch(texs, [o.data_ptr() for o in outs], [t.total_size for t in texs])
    for t, o, d in zip(texs, outs, files):
        assert split_levels(t, o.cpu().numpy()) == helpers.port_unpack_all(port, d)
        t.close()
