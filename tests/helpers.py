"""ctypes plumbing shared by the tests: the oracle port, the unmodified reference (oracle/_ref),
the SIMT-emulation build of the product library, and golden-vector helpers."""
import ctypes
import hashlib
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731


def _make(args, cwd):
    subprocess.run(["make"] + args, cwd=cwd, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)


def load_port():
    so = os.path.join(ROOT, "oracle", "liboracle_port.so")
    srcs = [os.path.join(ROOT, "oracle", "port", f) for f in os.listdir(os.path.join(ROOT, "oracle", "port"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        _make(["port"], os.path.join(ROOT, "oracle"))
    lib = ctypes.CDLL(so)
    lib.op_pack_image.restype = ctypes.c_uint32
    return lib


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "liboracle_ref.so")
    if not os.path.exists(so):
        if not os.path.exists("/root/reference/crnlib/crnlib.cpp"):
            return None
        _make(["-j8", "ref"], os.path.join(ROOT, "oracle"))
    lib = ctypes.CDLL(so)
    lib.ref_pack_image.restype = ctypes.c_uint32
    lib.ref_compress.restype = ctypes.c_void_p
    lib.ref_crn_to_dds.restype = ctypes.c_void_p
    lib.ref_transcode_begin.restype = ctypes.c_void_p
    lib.ref_transcode_all.restype = ctypes.c_double
    return lib


def load_sim():
    so = os.path.join(ROOT, "tests", "cusim", "libcrn_b200_sim.so")
    csrc = os.path.join(ROOT, "crunch2_b200", "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "tests", "cusim", f) for f in ("cusim.h", "cuda_runtime.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in deps):
        _make(["sim"], csrc)
    from crunch2_b200 import api
    return api._declare(ctypes.CDLL(so))


def bytes_per_block(fmt):
    return 8 if fmt in (0, 1, 4) else 16


def port_pack(lib, fmt, img, q=4, perc=1, both=1, thresh=128, tfb=0):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    out = np.zeros(((w + 3) // 4) * ((h + 3) // 4) * bytes_per_block(fmt), np.uint8)
    n = lib.op_pack_image(fmt, P(img), w, h, w * 4, q, perc, both, thresh, tfb, P(out))
    assert n == out.size
    return out


def ref_pack(lib, fmt, img, q=4, perc=1, both=1, thresh=128, tfb=0, threads=0):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    out = np.zeros(((w + 3) // 4) * ((h + 3) // 4) * bytes_per_block(fmt), np.uint8)
    n = lib.ref_pack_image(fmt, P(img), w, h, q, perc, both, tfb, thresh, 0, threads, P(out))
    assert n == out.size
    return out


def blocks_to_image(blocks):
    """(n,16,4) blocks -> a 4-pixel wide, 4n-pixel tall image whose block i is blocks[i]."""
    n = blocks.shape[0]
    return np.ascontiguousarray(blocks.reshape(n * 4, 4, 4))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden(name):
    with open(os.path.join(ROOT, "tests", "golden", name)) as f:
        return json.load(f)


def mismatching_blocks(a, b, bpb):
    a = a.reshape(-1, bpb)
    b = b.reshape(-1, bpb)
    return np.nonzero((a != b).any(axis=1))[0]


# ---- CRN files ------------------------------------------------------------------------------------
CRN_FMT = dict(DXT1=0, DXT3=1, DXT5=2, DXT5_CCxY=3, DXT5_xGxR=4, DXT5_xGBR=5, DXT5_AGBR=6, DXN_XY=7, DXN_YX=8, DXT5A=9)


def ref_compress(lib, images, fmt, file_type=0, quality=128, flags=1 | 2 | 8, bitrate=0.0, dxt_quality=4, threads=7, want_bitrate=False):
    """crn_compress through the unmodified reference.  images[face][level] = (h, w, 4) uint8.
    file_type 0 = CRN, 1 = DDS.  Returns (bytes, actual_quality, actual_bitrate)."""
    faces, levels = len(images), len(images[0])
    h, w = images[0][0].shape[:2]
    flat = [np.ascontiguousarray(images[f][l], np.uint8) for f in range(faces) for l in range(levels)]
    arr = (ctypes.c_void_p * len(flat))(*[a.ctypes.data for a in flat])
    size = ctypes.c_uint32(); aq = ctypes.c_uint32(); ab = ctypes.c_float()
    p = lib.ref_compress(file_type, fmt, w, h, faces, levels, arr, flags, quality, ctypes.c_float(bitrate), dxt_quality, threads, 3,
                         ctypes.byref(size), ctypes.byref(aq), ctypes.byref(ab), int(want_bitrate))
    if not p:
        return None, 0, 0.0
    data = ctypes.string_at(p, size.value)
    lib.ref_free(ctypes.c_void_p(p))
    return data, aq.value, ab.value


def level_dims(w, h, level):
    return max(1, w >> level), max(1, h >> level)


def ref_unpack_all(lib, crn):
    """Every level / face through crnd_unpack_level of the reference; returns list[level][face] -> bytes."""
    buf = np.frombuffer(crn, np.uint8)
    info = (ctypes.c_uint32 * 8)()
    assert lib.ref_crn_info(P(buf), len(crn), info)
    w, h, levels, faces, bpb = info[0], info[1], info[2], info[3], info[4]
    ctx = lib.ref_transcode_begin(P(buf), len(crn))
    assert ctx
    out = []
    for l in range(levels):
        lw, lh = level_dims(w, h, l)
        bx, by = (lw + 3) // 4, (lh + 3) // 4
        faces_np = [np.zeros(bx * by * bpb, np.uint8) for _ in range(faces)]
        ptrs = (ctypes.c_void_p * faces)(*[a.ctypes.data for a in faces_np])
        assert lib.ref_transcode_level(ctypes.c_void_p(ctx), ptrs, bx * by * bpb, bx * bpb, l)
        out.append([a.tobytes() for a in faces_np])
    lib.ref_transcode_end(ctypes.c_void_p(ctx))
    return out


def port_unpack_all(lib, crn):
    lib.op_crnd_begin.restype = ctypes.c_void_p
    buf = np.frombuffer(crn, np.uint8)
    info = (ctypes.c_uint32 * 8)()
    assert lib.op_crnd_info(P(buf), len(crn), info)
    w, h, levels, faces, bpb = info[0], info[1], info[2], info[3], info[4]
    ctx = lib.op_crnd_begin(P(buf), len(crn))
    assert ctx
    out = []
    for l in range(levels):
        lw, lh = level_dims(w, h, l)
        bx, by = (lw + 3) // 4, (lh + 3) // 4
        faces_np = [np.zeros(bx * by * bpb, np.uint8) for _ in range(faces)]
        ptrs = (ctypes.c_void_p * faces)(*[a.ctypes.data for a in faces_np])
        assert lib.op_crnd_unpack_level(ctypes.c_void_p(ctx), ptrs, bx * by * bpb, bx * bpb, l)
        out.append([a.tobytes() for a in faces_np])
    lib.op_crnd_end(ctypes.c_void_p(ctx))
    return out
