"""The .CRN writer back-end (crn_gpu_crn_write, host C++ in csrc/crn_writer.h) against the reference's own writer.

The reference's crn_compress (oracle/_ref, helper threads = 0: the deterministic configuration) writes a .crn from an
image; the reference's dxt_hc::compress through the shim gives the palettes + indices that file was written from; our
writer codes those same arrays.  Gates: (1) the reference's decoder (crnd_unpack_level) turns both files into exactly the
same blocks, level by level and face by face; (2) header fields agree and crnd_validate_file-style CRCs hold (the port
decoder checks them); (3) the file size is within 0.5 % of the reference's (north_star tolerance: bitrate within 1 %);
(4) crn_gpu_crn_hc_params reproduces crn_comp's codebook sizing (otherwise (1) could not hold)."""
import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
import hc_util

CRN_FMT = helpers.CRN_FMT
HC_FMT = {"DXT1": 0, "DXT5": 3, "DXT5A": 4, "DXN_XY": 5, "DXN_YX": 6}


@pytest.fixture(scope="module")
def lib():
    return crn.load_library()     # the product library: the writer is host code and needs no device


def mips(img, n):
    from bench import mip_chain
    return mip_chain(img)[:n]


def write_from_reference_hc(lib, ref, name, face_levels, quality, threads=0):
    faces, levels = len(face_levels), len(face_levels[0])
    h, w = face_levels[0][0].shape[:2]
    p = crn.crn_params(CRN_FMT[name], w, h, levels, faces, quality, perceptual=True, lib=lib)
    hp = crn.crn_hc_params(p, lib=lib)
    blocks, lv = hc_util.hc_layout(face_levels)
    assert hp.num_blocks == len(blocks)
    for i, (first, nb, bw, weight) in enumerate(lv):
        assert (hp.levels[i].first_block, hp.levels[i].num_blocks, hp.levels[i].block_width) == (first, nb, bw)
        assert np.float32(hp.levels[i].weight) == np.float32(weight)
    cbs = (hp.color_endpoint_codebook_size, hp.color_selector_codebook_size, hp.alpha_endpoint_codebook_size, hp.alpha_selector_codebook_size)
    out = hc_util.ref_hc_compress(ref, HC_FMT[name], blocks, lv, num_faces=faces, perceptual=bool(hp.perceptual), codebook_sizes=cbs,
                                  deratings=(hp.adaptive_tile_color_psnr_derating, hp.adaptive_tile_alpha_psnr_derating, hp.adaptive_tile_color_alpha_weighting_ratio),
                                  alpha_components=tuple(hp.alpha_component_indices), threads=threads)
    return crn.crn_write(p, hp, out, lib=lib), out


@pytest.mark.parametrize("name,w,h,nlev,quality", [
    ("DXT1", 64, 64, 3, 128), ("DXT5", 64, 48, 4, 128), ("DXN_XY", 40, 24, 2, 200), ("DXT5A", 32, 32, 6, 64), ("DXN_YX", 32, 16, 1, 255), ("DXT1", 128, 128, 8, 40)])
def test_same_blocks_and_size_as_reference_writer(lib, ref, port, name, w, h, nlev, quality):
    img = blockgen.smooth_image(w, h, 7 + w, alpha=True)
    face_levels = [mips(img, nlev)]
    want, _, _ = helpers.ref_compress(ref, face_levels, CRN_FMT[name], file_type=0, quality=quality, threads=0)
    assert want is not None
    got, _ = write_from_reference_hc(lib, ref, name, face_levels, quality)
    assert helpers.ref_unpack_all(ref, got) == helpers.ref_unpack_all(ref, want)
    assert helpers.port_unpack_all(port, got) == helpers.ref_unpack_all(ref, want)
    assert got[12:19] == want[12:19]                      # width, height, levels, faces, format
    for k in range(4):                                    # palette entry counts
        assert got[39 + 8 * k:41 + 8 * k] == want[39 + 8 * k:41 + 8 * k]
    assert abs(len(got) - len(want)) <= max(8, 0.005 * len(want)), (len(got), len(want))
    # stronger than the contract: on these vectors the two writers agree byte for byte (orderings, code lengths, CRCs)
    assert got == want


def test_cubemap(lib, ref):
    faces = [mips(blockgen.smooth_image(32, 32, 100 + f, alpha=True), 3) for f in range(6)]
    want, _, _ = helpers.ref_compress(ref, faces, CRN_FMT["DXT5"], file_type=0, quality=160, threads=0)
    got, _ = write_from_reference_hc(lib, ref, "DXT5", faces, 160)
    assert helpers.ref_unpack_all(ref, got) == helpers.ref_unpack_all(ref, want)
    assert got == want


def test_reference_validates_and_converts_file(lib, ref):
    """crnd_validate_file (header + data CRC16) inside the reference's crn_decompress_crn_to_dds accepts the file."""
    import ctypes
    img = blockgen.smooth_image(64, 64, 3, alpha=True)
    got, _ = write_from_reference_hc(lib, ref, "DXT5", [mips(img, 7)], 128)
    buf = np.frombuffer(got, np.uint8)
    size = ctypes.c_uint32()
    dds = ref.ref_crn_to_dds(helpers.P(buf), len(got), ctypes.byref(size))
    assert dds and size.value > 128
    ref.ref_free(ctypes.c_void_p(dds))
    info = crn.texture_info(got, lib=lib)
    assert (info["width"], info["height"], info["levels"], info["faces"]) == (64, 64, 7, 1)


def test_bad_arguments(lib):
    p = crn.crn_params(1, 64, 64, lib=lib)               # DXT3: refused like the reference (crn_comp.cpp:600-604)
    with pytest.raises(crn.CrnGpuError):
        crn.crn_hc_params(p, lib=lib)
    p = crn.crn_params(0, 8, 8, lib=lib)
    hp = crn.crn_hc_params(p, lib=lib)
    n = hp.num_blocks
    out = dict(endpoint_indices=np.zeros((n, 4), np.uint16), selector_indices=np.zeros((n, 4), np.uint16), color_endpoints=np.array([0x1234ABCD], np.uint32),
               alpha_endpoints=np.zeros(0, np.uint32), color_selectors=np.array([0x55AA00FF], np.uint32), alpha_selectors=np.zeros(0, np.uint64))
    data = crn.crn_write(p, hp, out, lib=lib)            # one-entry palettes are legal
    assert crn.texture_info(data, lib=lib)["width"] == 8
    out["endpoint_indices"][1, 0] = 5                      # index past the palette
    with pytest.raises(crn.CrnGpuError):
        crn.crn_write(p, hp, out, lib=lib)
