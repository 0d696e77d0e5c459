"""CPU tests of the CRN -> DXTn transcoder: oracle port against the golden .crn files written and decoded
by the unmodified reference, the reference itself (when built), and the CUDA kernels under the SIMT
emulator.  Synthetic .crn files (tests/crnsynth.py) cover what the reference's compressor cannot
produce: 8192-entry palettes, long Huffman codes, ragged sizes."""
import os

import numpy as np
import pytest

import crnsynth
import crunch2_b200 as crn
import helpers

GOLD = helpers.golden("crn_golden.json")["cases"]


def load(case):
    with open(os.path.join(helpers.ROOT, "tests", "golden", "crn", case["name"] + ".crn"), "rb") as f:
        return f.read()


def shas(levels):
    return [[helpers.sha(np.frombuffer(fc, np.uint8)) for fc in lv] for lv in levels]


def split_levels(tex, flat):
    out = []
    for l in range(tex.info["levels"]):
        bx, by = tex.level_blocks(l)
        n = bx * by * tex.info["bytes_per_block"]
        out.append([flat[tex.level_offset(l, f):tex.level_offset(l, f) + n].tobytes() for f in range(tex.info["faces"])])
    return out


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_port_matches_golden_crn(port, case):
    assert shas(helpers.port_unpack_all(port, load(case))) == case["sha256"]


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_sim_matches_golden_crn(sim, case):
    ctx = crn.Context(0, lib=sim)
    data = load(case)
    info = crn.texture_info(data, sim)
    assert (info["width"], info["height"], info["faces"], info["levels"], info["format"]) == (case["width"], case["height"], case["faces"], case["levels"], case["format"])
    tex = ctx.unpack_begin(data)
    assert shas(split_levels(tex, tex.unpack_all())) == case["sha256"]
    tex.close(); ctx.close()


SYNTH = [("DXT1", 64, 64, 1, {}), ("DXT5", 100, 36, 1, {}), ("DXN_XY", 32, 48, 1, {}), ("DXT5A", 20, 12, 1, {}), ("DXT5", 32, 32, 6, {}),
         ("DXT1", 256, 256, 1, dict(n_color_ep=8192, n_color_sel=8192)), ("DXT5", 8, 8, 1, dict(n_color_ep=9, n_color_sel=8, n_alpha_ep=8, n_alpha_sel=8)),
         ("DXN_YX", 1, 1, 1, {}), ("DXT5", 260, 4, 1, dict(n_alpha_ep=8192, n_alpha_sel=8192))]


@pytest.mark.parametrize("fmt,w,h,faces,kw", SYNTH)
def test_synthetic_crn_port_ref_sim_agree(port, sim, fmt, w, h, faces, kw):
    data = crnsynth.synth_crn(w, h, fmt, faces=faces, seed=5, **kw)
    want = helpers.port_unpack_all(port, data)
    ref = helpers.load_ref()
    if ref is not None:
        assert helpers.ref_unpack_all(ref, data) == want
    ctx = crn.Context(0, lib=sim)
    tex = ctx.unpack_begin(data)
    assert split_levels(tex, tex.unpack_all()) == want
    tex.close(); ctx.close()


def test_unpack_level_contract(sim, port):
    """crnd_unpack_level semantics: per-face destination pointers, explicit pitch, size / pitch validation
    (inc/crn_decomp.h:3569-3575)."""
    import ctypes
    data = load([c for c in GOLD if c["name"] == "dxt1_cube_32_mips"][0])
    want = helpers.port_unpack_all(port, data)
    ctx = crn.Context(0, lib=sim)
    tex = ctx.unpack_begin(data)
    bx, by = tex.level_blocks(1)
    pitch = bx * 8 + 16
    bufs = [np.full(pitch * by, 0xEE, np.uint8) for _ in range(6)]
    tex.unpack_level_device([b.ctypes.data for b in bufs], pitch * by, pitch, 1)
    ctx.synchronize()
    for f in range(6):
        rows = bufs[f].reshape(by, pitch)
        assert rows[:, :bx * 8].tobytes() == want[1][f]
        assert (rows[:, bx * 8:] == 0xEE).all()          # padding untouched
    with pytest.raises(crn.CrnGpuError):
        tex.unpack_level_device([b.ctypes.data for b in bufs], 8, pitch, 1)          # destination too small
    with pytest.raises(crn.CrnGpuError):
        tex.unpack_level_device([b.ctypes.data for b in bufs], pitch * by, bx * 8 - 4, 1)   # pitch too small
    with pytest.raises(crn.CrnGpuError):
        tex.unpack_level_device([b.ctypes.data for b in bufs], pitch * by, pitch, 99)      # bad level
    tex.close(); ctx.close()


def test_bad_files_rejected(sim):
    ctx = crn.Context(0, lib=sim)
    with pytest.raises(crn.CrnGpuError):
        ctx.unpack_begin(b"\0" * 200)
    data = bytearray(load(GOLD[0]))
    with pytest.raises(crn.CrnGpuError):
        ctx.unpack_begin(bytes(data[:40]))
    ctx.close()


def test_batch_matches_single(sim, port):
    ctx = crn.Context(0, lib=sim)
    files = [load(c) for c in GOLD[:4]]
    texs = [ctx.unpack_begin(d) for d in files]
    outs = [np.zeros(t.total_size, np.uint8) for t in texs]
    ctx.unpack_batch(texs, [o.ctypes.data for o in outs], [o.size for o in outs])
    for t, o, d in zip(texs, outs, files):
        assert split_levels(t, o) == helpers.port_unpack_all(port, d)
        t.close()
    ctx.close()


WIDE = [("DXT1", 256, 256, 1, {}), ("DXT5", 260, 100, 1, {}), ("DXN_XY", 128, 64, 1, {}), ("DXT5A", 260, 36, 1, {}), ("DXT5", 64, 64, 6, {}),
        ("DXT1", 512, 64, 1, dict(n_color_ep=8192, n_color_sel=8192)), ("DXT5", 1024, 16, 1, dict(n_alpha_ep=8192, n_alpha_sel=8192, skew=0.1))]


@pytest.mark.parametrize("split", ["0", "1"])
@pytest.mark.parametrize("min_blocks", ["1", "64"])
@pytest.mark.parametrize("fmt,w,h,faces,kw", WIDE)
def test_wide_path_matches_port(port, sim, monkeypatch, fmt, w, h, faces, kw, min_blocks, split):
    """transcode_wide.cuh (transition tables -> walk -> resolve) on files small enough for the emulator: every level
    ("1") or only the large ones ("64", the rest through the warp-per-level kernel), bit-exact against the oracle."""
    monkeypatch.setenv("CRN_B200_WIDE_MIN_BLOCKS", min_blocks)
    monkeypatch.setenv("CRN_B200_WIDE_SPLIT", split)          # walker + resolver in one launch / in two
    data = crnsynth.synth_crn(w, h, fmt, faces=faces, seed=11, **kw)
    want = helpers.port_unpack_all(port, data)
    ctx = crn.Context(0, lib=sim)
    tex = ctx.unpack_begin(data)
    l0 = ctx.launch_count
    got = split_levels(tex, tex.unpack_all())
    assert ctx.launch_count - l0 == (2 if min_blocks == "1" else 3) + int(split)
    assert got == want
    # one level at a time with a pitch, as crnd_unpack_level is called
    bx, by = tex.level_blocks(0)
    bpb = tex.info["bytes_per_block"]
    pitch = bx * bpb + 8
    bufs = [np.full(pitch * by, 0xEE, np.uint8) for _ in range(faces)]
    tex.unpack_level_device([b.ctypes.data for b in bufs], pitch * by, pitch, 0)
    ctx.synchronize()
    for f in range(faces):
        rows = bufs[f].reshape(by, pitch)
        assert rows[:, :bx * bpb].tobytes() == want[0][f]
        assert (rows[:, bx * bpb:] == 0xEE).all()
    tex.close(); ctx.close()


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_wide_path_matches_golden_crn(sim, monkeypatch, case):
    monkeypatch.setenv("CRN_B200_WIDE_MIN_BLOCKS", "1")
    monkeypatch.setenv("CRN_B200_WIDE_SPLIT", "0")
    ctx = crn.Context(0, lib=sim)
    tex = ctx.unpack_begin(load(case))
    assert shas(split_levels(tex, tex.unpack_all())) == case["sha256"]
    tex.close(); ctx.close()


def test_wide_batch_matches_single(sim, port, monkeypatch):
    """crnd_unpack_batch with the large levels of every file on the wide path (one table / walk / resolve launch for all)."""
    monkeypatch.setenv("CRN_B200_WIDE_MIN_BLOCKS", "64")
    ctx = crn.Context(0, lib=sim)
    files = [crnsynth.synth_crn(w, h, f, seed=40 + i) for i, (f, w, h) in enumerate([("DXT1", 128, 64), ("DXT5", 64, 64), ("DXN_XY", 36, 260), ("DXT5A", 64, 32), ("DXT5", 8, 8)])]
    texs = [ctx.unpack_begin(d) for d in files]
    outs = [np.zeros(t.total_size, np.uint8) for t in texs]
    l0 = ctx.launch_count
    ctx.unpack_batch(texs, [o.ctypes.data for o in outs], [o.size for o in outs])
    assert ctx.launch_count - l0 == 4                      # warp-per-level kernel + tables + walk + resolve
    for t, o, d in zip(texs, outs, files):
        assert split_levels(t, o) == helpers.port_unpack_all(port, d)
        t.close()
    ctx.close()


def _level_ofs(data, level):
    return int.from_bytes(data[70 + 4 * level:74 + 4 * level], "big")


@pytest.mark.parametrize("wide", [0, 1])
def test_truncated_and_corrupt_files_are_safe(sim, wide, monkeypatch):
    """a truncated level stream must read as zeros past its end (crn_decomp.h:3168-3170), never past the file image;
    corrupt headers / models are rejected at unpack_begin.  The emulator build runs under the host's memory protection, so an
    out-of-bounds walk of hundreds of KB would fault here."""
    if wide:
        monkeypatch.setenv("CRN_B200_WIDE_MIN_BLOCKS", "64")
    ctx = crn.Context(0, lib=sim)
    data = bytearray(crnsynth.synth_crn(256, 128, "DXT5", seed=11))
    # 1. level 0's stream cut to 40 bytes: declared size stays 256x128, the decoder keeps "decoding" zeros
    cut = bytes(data[:_level_ofs(data, 0) + 40])
    hdr = bytearray(cut)
    hdr[6:10] = len(cut).to_bytes(4, "big")          # data_size must fit the buffer
    try:
        tex = ctx.unpack_begin(bytes(hdr))
    except crn.CrnGpuError:
        tex = None                                   # rejecting is fine too (level offsets of the mips now lie past the end)
    if tex is not None:
        out = tex.unpack_all()
        assert out.size == tex.total_size
        tex.close()
    # 2. a 74-byte buffer that claims 16 levels: header_size > size
    tiny = bytearray(data[:74]); tiny[16] = 16; tiny[2:4] = (70 + 64).to_bytes(2, "big"); tiny[6:10] = (74).to_bytes(4, "big")
    with pytest.raises(crn.CrnGpuError):
        ctx.unpack_begin(bytes(tiny))
    # 3. selector palette count zeroed while the format needs it
    bad = bytearray(data); bad[33 + 8 + 6:33 + 8 + 8] = (0).to_bytes(2, "big")
    with pytest.raises(crn.CrnGpuError):
        ctx.unpack_begin(bytes(bad))
    # 4. random garbage after the header: either rejected or decoded without leaving the file image
    rng = np.random.default_rng(3)
    junk = bytearray(data); first = _level_ofs(data, 0)
    junk[first:] = rng.integers(0, 256, len(junk) - first, dtype=np.uint8).tobytes()
    try:
        tex = ctx.unpack_begin(bytes(junk))
        assert tex.unpack_all().size == tex.total_size
        tex.close()
    except crn.CrnGpuError:
        pass
    ctx.close()


def test_lane_per_stream_kernel_matches_port(sim, port, monkeypatch):
    """transcode_streams_kernel (one lane per level stream, the large-batch path): forced on with CRN_B200_STREAMS_MIN=1.
    Single files of every format (golden .crn from the reference + synthetic ones with 8192-entry palettes and long codes),
    then one mixed-format batch through crn_gpu_crnd_unpack_batch."""
    monkeypatch.setenv("CRN_B200_STREAMS_MIN", "1")
    ctx = crn.Context(0, lib=sim)
    l0 = ctx.launch_count
    datas = [load(c) for c in GOLD]
    datas += [crnsynth.synth_crn(w, h, fmt, faces=faces, seed=7, **kw) for fmt, w, h, faces, kw in SYNTH]
    wants = [helpers.port_unpack_all(port, d) for d in datas]
    for d, want in zip(datas, wants):
        tex = ctx.unpack_begin(d)
        assert split_levels(tex, tex.unpack_all()) == want
        tex.close()
    texs = [ctx.unpack_begin(d) for d in datas]
    bufs = [np.full(t.total_size + 16, 0xEE, np.uint8) for t in texs]
    ctx.unpack_batch(texs, [b.ctypes.data for b in bufs], [t.total_size for t in texs])
    for t, b, want in zip(texs, bufs, wants):
        assert split_levels(t, b[:t.total_size]) == want
        assert (b[t.total_size:] == 0xEE).all()
        t.close()
    assert ctx.launch_count - l0 >= len(datas) + 1
    ctx.close()
