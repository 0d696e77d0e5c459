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
