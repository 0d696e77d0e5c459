"""crn_compress to .CRN end to end (crn_gpu_compress_crn: device block gather + dxt_hc pipeline under the SIMT emulator +
host writer) against the reference's crn_compress (oracle/_ref, helper threads = 0).  Tolerance class (north_star): PSNR of
the decoded file within 0.05 dB of the reference's, file size (the CRN bitrate, crn_comp.cpp:1640-1653) within 1 %.  Both
files are decoded by the reference's own crnd_unpack_level."""
import numpy as np
import pytest

import blockgen
import crunch2_b200 as crn
import helpers
import quality

HC_FMT = {"DXT1": 0, "DXT5": 3, "DXT5A": 4, "DXN_XY": 5}
CHANNELS = {"DXT1": [[0, 1, 2]], "DXT5": [[0, 1, 2], [3]], "DXT5A": [[3]], "DXN_XY": [[0, 1]]}


@pytest.fixture(scope="module")
def simctx(sim):
    ctx = crn.Context(0, lib=sim)
    yield ctx
    ctx.close()


def decoded_psnr(ref, data, name, face_levels):
    """PSNR per channel group of a .crn (decoded by the reference) against the source pixels, all faces and levels."""
    lv = helpers.ref_unpack_all(ref, data)
    got, src = [], []
    for l, faces in enumerate(lv):
        for f, blob in enumerate(faces):
            img = face_levels[f][l]
            h, w = img.shape[:2]
            ph, pw = (h + 3) & ~3, (w + 3) & ~3
            ys = np.minimum(np.arange(ph), h - 1); xs = np.minimum(np.arange(pw), w - 1)
            src.append(quality.image_to_blocks(np.ascontiguousarray(img[ys][:, xs])))
            got.append(quality.decode_blocks(blob, HC_FMT[name]))
    got, src = np.concatenate(got), np.concatenate(src)
    return [quality.psnr(got, src, c) for c in CHANNELS[name]]


def check(ref, name, face_levels, got, want, psnr_tol=0.05, size_tol=0.01):
    pg, pr = decoded_psnr(ref, got, name, face_levels), decoded_psnr(ref, want, name, face_levels)
    for a, b in zip(pg, pr):
        assert a >= b - psnr_tol, (pg, pr)
    assert len(got) <= len(want) * (1 + size_tol) + 8, (len(got), len(want))
    return pg, pr


@pytest.mark.parametrize("name,w,h,nlev,q", [("DXT1", 64, 48, 3, 128), ("DXT5", 64, 64, 2, 100), ("DXN_XY", 32, 32, 2, 200), ("DXT5A", 32, 32, 1, 160)])
def test_quality_level(simctx, ref, name, w, h, nlev, q):
    from bench import mip_chain
    face_levels = [mip_chain(blockgen.smooth_image(w, h, 21 + w, alpha=True))[:nlev]]
    want, _, _ = helpers.ref_compress(ref, face_levels, helpers.CRN_FMT[name], file_type=0, quality=q, threads=0)
    got, rate, used_q = simctx.compress_crn(face_levels, helpers.CRN_FMT[name], quality_level=q)
    assert used_q == q
    texels = sum(l.shape[0] * l.shape[1] for l in face_levels[0])
    assert abs(rate - len(got) * 8.0 / texels) < 1e-4
    check(ref, name, face_levels, got, want)
    info = crn.texture_info(got, lib=simctx._lib)
    assert (info["width"], info["height"], info["levels"], info["faces"], info["format"]) == (w, h, nlev, 1, helpers.CRN_FMT[name])


def test_target_bitrate_search(simctx, ref):
    """The quality search lands on the reference's quality level (or a neighbour whose file is as close to the target)."""
    from bench import mip_chain
    face_levels = [mip_chain(blockgen.smooth_image(64, 64, 5, alpha=True))[:2]]
    target = 1.6
    want, ref_q, ref_rate = helpers.ref_compress(ref, face_levels, helpers.CRN_FMT["DXT1"], file_type=0, bitrate=target, threads=0, want_bitrate=True)
    got, rate, q = simctx.compress_crn(face_levels, helpers.CRN_FMT["DXT1"], target_bitrate=target)
    assert abs(rate - target) <= abs(ref_rate - target) + 0.02, (rate, ref_rate, q, ref_q)
    assert abs(int(q) - int(ref_q)) <= 8, (q, ref_q)
    check(ref, "DXT1", face_levels, got, want, psnr_tol=0.15, size_tol=0.02)   # neighbouring quality levels differ by about this much


def test_cubemap_quality_level(simctx, ref):
    from bench import mip_chain
    faces = [mip_chain(blockgen.smooth_image(32, 32, 300 + f, alpha=True))[:2] for f in range(6)]
    want, _, _ = helpers.ref_compress(ref, faces, helpers.CRN_FMT["DXT1"], file_type=0, quality=128, threads=0)
    got, _, _ = simctx.compress_crn(faces, helpers.CRN_FMT["DXT1"], quality_level=128)
    check(ref, "DXT1", faces, got, want)
    assert crn.texture_info(got, lib=simctx._lib)["faces"] == 6
