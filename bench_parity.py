"""Parity gates of the bench line ("in the same run"): per BASELINE config, our output against the unmodified reference
(oracle/_ref) on the same input AT THE CONFIG'S SIZE, each with a boolean `within_tolerance`.  Only bench.py imports this; the reference side is
test infrastructure (tests/helpers.py -> oracle/_ref/liboracle_ref.so) and runs on the host cores, outside every timed region.

  C1  md5 of the full block payload (2048^2 DXT1 + mips, endpoint caching off)           bit-exact class
  C2  RGB / alpha PSNR + LZMA bits of the clustered DXT5 .dds at 4096^2 + mips            0.05 dB / 1 %
  C3  .crn of the 6 x 2048^2 cubemap at quality 128: PSNR + file bits; the 1.25 bpp search beside the reference's own search on a
      6 x 512^2 cubemap (the reference's search at full size takes minutes): quality level, bpp, PSNR
  C4  every level of the 8192^2 DXT5 .crn equal to crnd_unpack_level's bytes (run_transcode)   bit-exact class
  C5  all 1024 textures through our path; PSNR + LZMA bits of a 12-texture sample against the reference

PSNR is the reference's image_utils::error_metrics definition (crnlib/crn_image_utils.cpp:1048-1123: 10 log10(255^2 / MSE) over the selected
channels, all levels and faces pooled); both files are decoded by the same decoder (our unpack kernel, bit-exact against dxt_image::unpack)."""
import hashlib
import time

import numpy as np

PSNR_TOL, BITS_TOL = 0.05, 0.01


def _sq_err(ctx, fmt, payload, levels, channels):
    """sum of squared errors and sample count of a block payload (levels back to back, faces outermost handled by the caller)"""
    se = {c: 0.0 for c in channels}; n = 0; ofs = 0
    import crunch2_b200 as crn
    bpb = crn.bytes_per_block(fmt)
    for img in levels:
        h, w = img.shape[:2]
        nb = ((w + 3) // 4) * ((h + 3) // 4) * bpb
        dec = ctx.unpack_image(fmt, payload[ofs:ofs + nb], w, h)
        ofs += nb
        d = dec.astype(np.int32) - img.astype(np.int32)
        for c in channels:
            se[c] += float((d[..., list(c)].astype(np.float64) ** 2).sum())
        n += h * w
    return se, n, ofs


def psnr_of_payload(ctx, fmt, payload, faces_levels, channels):
    """faces_levels[face][level]; payload laid out faces outermost, levels inside (write_dds order).  Returns {channels: dB}."""
    tot = {c: 0.0 for c in channels}; n = 0; ofs = 0
    payload = np.frombuffer(payload, np.uint8)
    for levels in faces_levels:
        se, k, used = _sq_err(ctx, fmt, payload[ofs:], levels, channels)
        ofs += used; n += k
        for c in channels:
            tot[c] += se[c]
    out = {}
    for c in channels:
        mse = tot[c] / (n * len(c))
        out[c] = 999.0 if mse == 0 else 10.0 * np.log10(255.0 * 255.0 / mse)
    return out


def gate(ps_ours, ps_ref, bits_ours, bits_ref):
    d = {"psnr_ours": [round(float(v), 4) for v in ps_ours.values()], "psnr_reference": [round(float(v), 4) for v in ps_ref.values()],
         "psnr_delta_db": [round(float(a - b), 4) for a, b in zip(ps_ours.values(), ps_ref.values())],
         "bits_ours": int(bits_ours), "bits_reference": int(bits_ref), "bits_ratio": round(bits_ours / max(1, bits_ref), 5)}
    d["within_tolerance"] = bool(all(abs(x) <= PSNR_TOL for x in d["psnr_delta_db"]) and abs(bits_ours - bits_ref) <= BITS_TOL * bits_ref)
    d["tolerance"] = "PSNR within %.2f dB, bits within %.0f %%" % (PSNR_TOL, BITS_TOL * 100)
    return d


def lzma_bits(ctx, data):
    """LZMA-compressed size in bits with the reference's coder parameters (crn_gpu_lzma_size: liblzma, level-5 parameters of lzma_codec::pack)"""
    import ctypes
    buf = np.frombuffer(data, np.uint8)
    n = int(ctx._lib.crn_gpu_lzma_size(buf.ctypes.data_as(ctypes.c_void_p), len(data)))
    if n:
        return 8 * n
    import lzma
    return 8 * len(lzma.compress(bytes(data), format=lzma.FORMAT_ALONE, preset=5))


def c2_gate(ctx, fmt, levels, ours_payload, ref_dds):
    ch = ((0, 1, 2), (3,)) if fmt == 3 else ((0, 1, 2),)
    a = psnr_of_payload(ctx, fmt, ours_payload, [levels], ch)
    b = psnr_of_payload(ctx, fmt, ref_dds[128:], [levels], ch)
    g = gate(a, b, lzma_bits(ctx, ours_payload), lzma_bits(ctx, ref_dds[128:]))
    g["what"] = "clustered DXT5 .dds q128, %dx%d + %d mips: RGB / alpha PSNR and LZMA bits of the block payload, ours vs the reference's crn_compress" % (
        levels[0].shape[1], levels[0].shape[0], len(levels) - 1)
    return g


def c1_gate(ctx, ref, helpers, levels, threads):
    """block-by-block DXT1 of the whole chain: md5 of the payload, ours (crn_gpu_compress_dds at quality 255) vs the reference with endpoint caching off"""
    t0 = time.perf_counter()
    ours = ctx.compress_dds([levels], 0, quality_level=255)
    t_ours = time.perf_counter() - t0
    t0 = time.perf_counter()
    want, _, _ = helpers.ref_compress(ref, [levels], 0, file_type=1, quality=255, threads=threads, flags=1 | 2 | 8 | 32)
    t_ref = time.perf_counter() - t0
    return {"what": "DXT1 uber .dds of %dx%d + %d mips, cCRNCompFlagDisableEndpointCaching: whole file" % (levels[0].shape[1], levels[0].shape[0], len(levels) - 1),
            "md5_ours": hashlib.md5(ours).hexdigest(), "md5_reference": hashlib.md5(want).hexdigest(), "bytes": len(ours),
            "within_tolerance": bool(ours == want), "tolerance": "bit-exact", "ours_s": round(t_ours, 3), "reference_s": round(t_ref, 3), "reference_threads": threads + 1}


def crn_payload(ctx, crn_bytes):
    """every level / face of a .crn through our transcoder, re-laid faces outermost (the .dds order psnr_of_payload expects)"""
    return ctx.crn_to_dds(crn_bytes)[128:]


def c3_gate(ctx, ref, helpers, faces, ours_crn, ref_crn, small_faces, threads, quiet):
    ch = ((0, 1, 2),)
    a = psnr_of_payload(ctx, 0, crn_payload(ctx, ours_crn), faces, ch)
    b = psnr_of_payload(ctx, 0, crn_payload(ctx, ref_crn), faces, ch)
    g = gate(a, b, 8 * len(ours_crn), 8 * len(ref_crn))
    g["what"] = "DXT1 .crn of the 6 x 2048^2 cubemap + mips at quality 128: RGB PSNR and file bits, ours vs the reference's crn_compress"
    # the target-bitrate search, both sides, on a cubemap the reference finishes in seconds
    t0 = time.perf_counter()
    so, srate, sq = ctx.compress_crn(small_faces, 0, target_bitrate=1.25)
    t_ours = time.perf_counter() - t0
    t0 = time.perf_counter()
    with quiet():
        sr, rq, rrate = helpers.ref_compress(ref, small_faces, 0, file_type=0, quality=128, bitrate=1.25, threads=threads, want_bitrate=True)
    t_ref = time.perf_counter() - t0
    ntex = sum(l.shape[0] * l.shape[1] for f in small_faces for l in f)
    pa = psnr_of_payload(ctx, 0, crn_payload(ctx, so), small_faces, ch)
    pb = psnr_of_payload(ctx, 0, crn_payload(ctx, sr), small_faces, ch)
    s = gate(pa, pb, 8 * len(so), 8 * len(sr))
    s.update({"what": "crn_compress with m_target_bitrate = 1.25 on a 6 x %d^2 cubemap + mips (the reference's search at 2048^2 takes minutes)" % small_faces[0][0].shape[0],
              "quality_level_ours": int(sq), "quality_level_reference": int(rq), "bpp_ours": round(8.0 * len(so) / ntex, 4), "bpp_reference": round(8.0 * len(sr) / ntex, 4),
              "ours_s": round(t_ours, 3), "reference_s": round(t_ref, 3)})
    g["bitrate_search"] = s
    g["within_tolerance"] = bool(g["within_tolerance"] and s["within_tolerance"])
    return g
