"""ctypes binding of libcrnlib_b200.so -- the reference's PUBLIC C++ API (inc/crnlib.h, inc/crn_defs.h) as crunch2_b200/csrc/crnlib_dropin.cpp
exports it (same mangled symbols as the reference's libcrn).  This is the call a crnlib user makes: `crn_compress(const crn_comp_params&,
crn_uint32&, crn_uint32*, float*)`, `crn_decompress_crn_to_dds`, `crn_free_block`, `crnd::crnd_unpack_*`.  bench.py's end-to-end figure and
tests/test_dropin_api.py go through it.  The structs below restate the reference's layouts (inc/crnlib.h:231-414, :471-574) for ctypes only;
C++ callers use the reference's own headers."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_FACES, MAX_LEVELS = 6, 16
FILE_CRN, FILE_DDS = 0, 1
FLAG_PERCEPTUAL, FLAG_HIERARCHICAL, FLAG_QUICK, FLAG_BOTH_BLOCK_TYPES = 1, 2, 4, 8
FLAG_TRANSPARENT_FOR_BLACK, FLAG_NO_ENDPOINT_CACHING, FLAG_MANUAL_PALETTES, FLAG_DXT1A = 16, 32, 64, 128
PROGRESS_FN = ctypes.CFUNCTYPE(ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p)


class CrnCompParams(ctypes.Structure):      # crn_comp_params, inc/crnlib.h:231-414
    _fields_ = [("m_size_of_obj", ctypes.c_uint32), ("m_file_type", ctypes.c_uint32), ("m_faces", ctypes.c_uint32), ("m_width", ctypes.c_uint32),
                ("m_height", ctypes.c_uint32), ("m_levels", ctypes.c_uint32),
                # enum crn_format holds both -1 (cCRNFmtInvalid) and 0xFFFFFFFF (cCRNFmtForceDWORD): its underlying type is 64 bits wide
                ("m_format", ctypes.c_int64), ("m_flags", ctypes.c_uint32),
                ("m_pImages", (ctypes.c_void_p * MAX_LEVELS) * MAX_FACES), ("m_target_bitrate", ctypes.c_float), ("m_quality_level", ctypes.c_uint32),
                ("m_dxt1a_alpha_threshold", ctypes.c_uint32), ("m_dxt_quality", ctypes.c_uint32), ("m_dxt_compressor_type", ctypes.c_uint32),
                ("m_alpha_component", ctypes.c_uint32), ("m_crn_adaptive_tile_color_psnr_derating", ctypes.c_float),
                ("m_crn_adaptive_tile_alpha_psnr_derating", ctypes.c_float), ("m_crn_color_endpoint_palette_size", ctypes.c_uint32),
                ("m_crn_color_selector_palette_size", ctypes.c_uint32), ("m_crn_alpha_endpoint_palette_size", ctypes.c_uint32),
                ("m_crn_alpha_selector_palette_size", ctypes.c_uint32), ("m_num_helper_threads", ctypes.c_uint32), ("m_userdata0", ctypes.c_uint32),
                ("m_userdata1", ctypes.c_uint32), ("m_pProgress_func", ctypes.c_void_p), ("m_pProgress_func_data", ctypes.c_void_p)]

    def clear(self):                        # crn_comp_params::clear()
        ctypes.memset(ctypes.byref(self), 0, ctypes.sizeof(self))
        self.m_size_of_obj = ctypes.sizeof(self)
        self.m_faces = 1; self.m_levels = 1
        self.m_flags = FLAG_PERCEPTUAL | FLAG_HIERARCHICAL | FLAG_BOTH_BLOCK_TYPES
        self.m_quality_level = 255; self.m_dxt1a_alpha_threshold = 128; self.m_dxt_quality = 4; self.m_alpha_component = 3
        self.m_crn_adaptive_tile_color_psnr_derating = 2.0; self.m_crn_adaptive_tile_alpha_psnr_derating = 2.0
        return self


class CrnMipmapParams(ctypes.Structure):    # crn_mipmap_params, inc/crnlib.h:471-574
    _fields_ = [("m_size_of_obj", ctypes.c_uint32), ("m_mode", ctypes.c_uint32), ("m_filter", ctypes.c_uint32), ("m_gamma_filtering", ctypes.c_uint32),
                ("m_gamma", ctypes.c_float), ("m_blurriness", ctypes.c_float), ("m_max_levels", ctypes.c_uint32), ("m_min_mip_size", ctypes.c_uint32),
                ("m_renormalize", ctypes.c_uint32), ("m_rtopmip", ctypes.c_uint32), ("m_tiled", ctypes.c_uint32), ("m_scale_mode", ctypes.c_uint32),
                ("m_scale_x", ctypes.c_float), ("m_scale_y", ctypes.c_float), ("m_window_left", ctypes.c_uint32), ("m_window_top", ctypes.c_uint32),
                ("m_window_right", ctypes.c_uint32), ("m_window_bottom", ctypes.c_uint32), ("m_clamp_scale", ctypes.c_uint32), ("m_clamp_width", ctypes.c_uint32),
                ("m_clamp_height", ctypes.c_uint32)]

    def clear(self):                        # crn_mipmap_params::clear()
        ctypes.memset(ctypes.byref(self), 0, ctypes.sizeof(self))
        self.m_size_of_obj = ctypes.sizeof(self)
        self.m_mode = 0; self.m_filter = 4; self.m_gamma_filtering = 1; self.m_gamma = 2.2; self.m_blurriness = .9
        self.m_max_levels = MAX_LEVELS; self.m_min_mip_size = 1; self.m_scale_x = 1.0; self.m_scale_y = 1.0
        return self


_SYMS = {"crn_compress": "_Z12crn_compressRK15crn_comp_paramsRjPjPf",
         "crn_compress_mip": "_Z12crn_compressRK15crn_comp_paramsRK17crn_mipmap_paramsRjPjPf",
         "crn_free_block": "_Z14crn_free_blockPv",
         "crn_decompress_crn_to_dds": "_Z25crn_decompress_crn_to_ddsPKvRj",
         "crn_get_version_number": "_Z22crn_get_version_numberv"}
_lib = None


def library_path():
    return os.path.join(_HERE, "libcrnlib_b200.so")


def load(path=None):
    """Loads the drop-in (and, through its DT_NEEDED entry, libcrn_b200.so).  Raises when it has not been built: there is no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or library_path()
    if not os.path.exists(p):
        raise RuntimeError("%s is missing: `make -C crunch2_b200/csrc` builds it where the reference's headers (/root/reference/inc) exist" % p)
    lib = ctypes.CDLL(p)
    f = getattr(lib, _SYMS["crn_compress"])
    f.restype = ctypes.c_void_p
    f.argtypes = [ctypes.POINTER(CrnCompParams), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_float)]
    fm = getattr(lib, _SYMS["crn_compress_mip"])
    fm.restype = ctypes.c_void_p
    fm.argtypes = [ctypes.POINTER(CrnCompParams), ctypes.POINTER(CrnMipmapParams), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_float)]
    g = getattr(lib, _SYMS["crn_free_block"]); g.restype = None; g.argtypes = [ctypes.c_void_p]
    h = getattr(lib, _SYMS["crn_decompress_crn_to_dds"]); h.restype = ctypes.c_void_p; h.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32)]
    if path is None:
        _lib = lib
    return lib


def crn_compress(images, crn_format, file_type=FILE_DDS, quality_level=255, flags=None, target_bitrate=0.0, dxt_quality=4, alpha_component=3, progress=None,
                 want_rate=True, lib=None, mipmap_params=None):
    """crn_compress (inc/crnlib.h:609) through the drop-in.  images[face][level]: (h, w, 4) uint8, C-contiguous (pinned host memory is fine).
    Returns (file bytes, actual quality level, actual bitrate) or raises RuntimeError when the call returns NULL."""
    lib = lib or load()
    p = CrnCompParams().clear()
    faces, levels = len(images), len(images[0])
    h, w = images[0][0].shape[:2]
    p.m_file_type = file_type; p.m_faces = faces; p.m_levels = levels; p.m_width = w; p.m_height = h; p.m_format = int(crn_format)
    if flags is not None:
        p.m_flags = int(flags)
    p.m_quality_level = int(quality_level); p.m_target_bitrate = float(target_bitrate); p.m_dxt_quality = int(dxt_quality); p.m_alpha_component = int(alpha_component)
    keep = []
    for f in range(faces):
        for l in range(levels):
            a = images[f][l]
            if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]):
                a = np.ascontiguousarray(a, np.uint8)
            keep.append(a)
            p.m_pImages[f][l] = a.ctypes.data
    cb = None
    if progress is not None:
        cb = PROGRESS_FN(progress)
        p.m_pProgress_func = ctypes.cast(cb, ctypes.c_void_p)
    size = ctypes.c_uint32(); q = ctypes.c_uint32(); rate = ctypes.c_float()
    if mipmap_params is not None:           # the crn_mipmap_params overload (inc/crnlib.h:614)
        ptr = getattr(lib, _SYMS["crn_compress_mip"])(ctypes.byref(p), ctypes.byref(mipmap_params), ctypes.byref(size), ctypes.byref(q), ctypes.byref(rate) if want_rate else None)
    else:
        ptr = getattr(lib, _SYMS["crn_compress"])(ctypes.byref(p), ctypes.byref(size), ctypes.byref(q), ctypes.byref(rate) if want_rate else None)
    if not ptr:
        raise RuntimeError("crn_compress returned NULL")
    try:
        return ctypes.string_at(ptr, size.value), q.value, rate.value
    finally:
        getattr(lib, _SYMS["crn_free_block"])(ptr)


def crn_decompress_crn_to_dds(crn_bytes, lib=None):
    lib = lib or load()
    size = ctypes.c_uint32(len(crn_bytes))
    ptr = getattr(lib, _SYMS["crn_decompress_crn_to_dds"])(ctypes.c_char_p(crn_bytes), ctypes.byref(size))
    if not ptr:
        raise RuntimeError("crn_decompress_crn_to_dds returned NULL")
    try:
        return ctypes.string_at(ptr, size.value)
    finally:
        getattr(lib, _SYMS["crn_free_block"])(ptr)
