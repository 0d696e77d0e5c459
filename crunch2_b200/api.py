"""Host-side binding of the C-ABI library (include/crn_b200.h) -- the call a user makes.

Mirrors the reference's interface for the path:
  * `Context.pack_image`  <->  crnlib::dxt_image::init(fmt, image, pack_params)
    (reference crnlib/crn_dxt_image.cpp:447-493), same formats, same knobs (`PackParams` mirrors
    dxt_image::pack_params, crnlib/crn_dxt_image.h:166-228), same block layout, results equal to
    the reference with endpoint caching disabled.
Errors: the reference returns false/NULL; here a `CrnGpuError` carries the library's status and
message.  The product path never touches oracle/ and refuses the g++ emulation build.
"""
import ctypes
import os

import numpy as np

FMT_DXT1, FMT_DXT1A, FMT_DXT3, FMT_DXT5, FMT_DXT5A, FMT_DXN_XY, FMT_DXN_YX = range(7)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class CrnGpuError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("crn_gpu status %d: %s" % (status, message))
        self.status = status


class _PackParams(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("dxt_quality", ctypes.c_uint32), ("perceptual", ctypes.c_uint32),
                ("use_both_block_types", ctypes.c_uint32), ("dxt1a_alpha_threshold", ctypes.c_uint32),
                ("use_transparent_indices_for_black", ctypes.c_uint32), ("grayscale_sampling", ctypes.c_uint32),
                ("non_hierarchical", ctypes.c_uint32), ("reserved", ctypes.c_uint32 * 4)]


class _TextureInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("struct_size", "width", "height", "levels", "faces", "bytes_per_block",
                                                "userdata0", "userdata1", "format")]


class _HcLevel(ctypes.Structure):
    _fields_ = [("first_block", ctypes.c_uint32), ("num_blocks", ctypes.c_uint32), ("block_width", ctypes.c_uint32), ("weight", ctypes.c_float)]


class _HcParams(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("format", ctypes.c_uint32), ("num_blocks", ctypes.c_uint32), ("num_levels", ctypes.c_uint32),
                ("num_faces", ctypes.c_uint32), ("levels", _HcLevel * 16), ("perceptual", ctypes.c_uint32),
                ("color_endpoint_codebook_size", ctypes.c_uint32), ("color_selector_codebook_size", ctypes.c_uint32),
                ("alpha_endpoint_codebook_size", ctypes.c_uint32), ("alpha_selector_codebook_size", ctypes.c_uint32),
                ("adaptive_tile_color_psnr_derating", ctypes.c_float), ("adaptive_tile_alpha_psnr_derating", ctypes.c_float),
                ("adaptive_tile_color_alpha_weighting_ratio", ctypes.c_float), ("alpha_component_indices", ctypes.c_uint32 * 2),
                ("shard_rank", ctypes.c_uint32), ("shard_count", ctypes.c_uint32), ("exchange", ctypes.c_void_p), ("exchange_user", ctypes.c_void_p)]


class _CrnParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("struct_size", "crn_format", "width", "height", "levels", "faces", "quality_level", "perceptual",
                                                "alpha_component", "userdata0", "userdata1")] + [("palette_sizes", ctypes.c_uint32 * 4),
                ("adaptive_tile_color_psnr_derating", ctypes.c_float), ("adaptive_tile_alpha_psnr_derating", ctypes.c_float),
                ("target_bitrate", ctypes.c_float), ("shard_rank", ctypes.c_uint32), ("shard_count", ctypes.c_uint32),
                ("exchange", ctypes.c_void_p), ("exchange_user", ctypes.c_void_p)]


class _DdsParams(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_uint32) for n in ("struct_size", "crn_format", "width", "height", "levels", "faces", "quality_level", "dxt1a_for_transparency")] + [("pack", _PackParams)]
                + [("target_bitrate", ctypes.c_float), ("hierarchical", ctypes.c_uint32), ("reserved", ctypes.c_uint32 * 2)])


EXCHANGE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32)


class _DdsDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("struct_size", "width", "height", "levels", "faces", "pixel_format", "file_format", "block_format")]


class _HcInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in ("struct_size", "num_blocks", "num_tiles", "n_color_endpoints", "n_alpha_endpoints",
                                                "n_color_selectors", "n_alpha_selectors")] + [("vq_rounds", ctypes.c_uint32 * 4), ("unique_vectors", ctypes.c_uint32 * 4)]


class _ResampleParams(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("filter", ctypes.c_uint32), ("filter_scale", ctypes.c_float), ("srgb", ctypes.c_uint32),
                ("source_gamma", ctypes.c_float), ("wrapping", ctypes.c_uint32), ("num_comps", ctypes.c_uint32), ("reserved", ctypes.c_uint32 * 3)]


MIP_FILTERS = {"box": 0, "tent": 1, "lanczos4": 2, "mitchell": 3, "kaiser": 4}


class PackParams:
    """dxt_image::pack_params for the block-by-block path (defaults of crn_comp_params::clear())."""

    def __init__(self, dxt_quality=4, perceptual=True, use_both_block_types=True, dxt1a_alpha_threshold=128,
                 use_transparent_indices_for_black=False, grayscale_sampling=False):
        self.dxt_quality = dxt_quality
        self.perceptual = perceptual
        self.use_both_block_types = use_both_block_types
        self.dxt1a_alpha_threshold = dxt1a_alpha_threshold
        self.use_transparent_indices_for_black = use_transparent_indices_for_black
        self.grayscale_sampling = grayscale_sampling

    def _c(self):
        p = _PackParams()
        p.struct_size = ctypes.sizeof(_PackParams)
        p.dxt_quality = int(self.dxt_quality)
        p.perceptual = int(bool(self.perceptual))
        p.use_both_block_types = int(bool(self.use_both_block_types))
        p.dxt1a_alpha_threshold = int(self.dxt1a_alpha_threshold)
        p.use_transparent_indices_for_black = int(bool(self.use_transparent_indices_for_black))
        p.grayscale_sampling = int(bool(self.grayscale_sampling))
        return p


def library_path():
    return os.path.join(_HERE, "libcrn_b200.so")


def _declare(lib):
    vp, u32, i32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int
    lib.crn_gpu_abi_version.restype = u32
    lib.crn_gpu_qdxt_output_size.restype = ctypes.c_uint64
    lib.crn_gpu_qdxt_output_size.argtypes = [vp]
    lib.crn_gpu_qdxt_level_offset.restype = ctypes.c_uint64
    lib.crn_gpu_qdxt_level_offset.argtypes = [vp, u32]
    lib.crn_gpu_qdxt_free.argtypes = [vp]
    lib.crn_gpu_qdxt_free.restype = None
    lib.crn_gpu_qdxt_pack.argtypes = [vp, u32, vp, i32]
    lib.crn_gpu_qdxt_get_info.argtypes = [vp, vp]
    lib.crn_gpu_is_native.restype = i32
    lib.crn_gpu_device_count.restype = i32
    lib.crn_gpu_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.crn_gpu_destroy.argtypes = [vp]
    lib.crn_gpu_destroy.restype = None
    lib.crn_gpu_last_error.argtypes = [vp]
    lib.crn_gpu_last_error.restype = ctypes.c_char_p
    lib.crn_gpu_stream.argtypes = [vp]
    lib.crn_gpu_stream.restype = vp
    lib.crn_gpu_synchronize.argtypes = [vp]
    lib.crn_gpu_launch_count.argtypes = [vp]
    lib.crn_gpu_launch_count.restype = ctypes.c_uint64
    lib.crn_gpu_bytes_per_block.argtypes = [u32]
    lib.crn_gpu_bytes_per_block.restype = u32
    lib.crn_gpu_pack_image.argtypes = [vp, u32, ctypes.POINTER(_PackParams), vp, u32, u32, u32, vp]
    lib.crn_gpu_pack_image_host.argtypes = [vp, u32, ctypes.POINTER(_PackParams), vp, u32, u32, u32, vp]
    u64 = ctypes.c_uint64
    lib.crn_gpu_dxt1_optimize_clusters.argtypes = [vp, ctypes.POINTER(_PackParams), i32, vp, u32, vp, vp, u32, u32, vp, u32, u32, vp, vp]
    lib.crn_gpu_dxt5_optimize_clusters.argtypes = [vp, ctypes.POINTER(_PackParams), u32, vp, u32, vp, vp, u32, u32, vp, u32, u32, vp, vp]
    lib.crn_gpu_qdxt_training.argtypes = [vp, u32, u32, vp, u32, vp, u32, vp, vp, vp]
    lib.crn_gpu_optimize_selectors.argtypes = [vp, u32, ctypes.POINTER(_PackParams), u32, vp, u32, vp, vp, u32, vp, u32, u32]
    lib.crn_gpu_refine_endpoints.argtypes = [vp, ctypes.c_int, ctypes.c_int, u32, vp, vp, vp, u32, vp, vp, vp, vp]
    lib.crn_gpu_nearest_codebook.argtypes = [vp, u32, vp, u32, vp, u32, vp]
    lib.crn_gpu_assign_selectors.argtypes = [vp, u32, ctypes.c_int, u32, vp, u32, vp, vp, vp, u32, vp, vp, vp]
    lib.crn_gpu_default_resample_params.argtypes = [ctypes.POINTER(_ResampleParams)]
    lib.crn_gpu_default_resample_params.restype = None
    lib.crn_gpu_resample.argtypes = [vp, ctypes.POINTER(_ResampleParams), vp, u32, u32, u32, vp, u32, u32, u32]
    lib.crn_gpu_mip_level_count.argtypes = [u32, u32, u32, u32]
    lib.crn_gpu_mip_level_count.restype = u32
    lib.crn_gpu_generate_mipmaps.argtypes = [vp, ctypes.POINTER(_ResampleParams), vp, u32, u32, u32, u32, u32, vp, u64, ctypes.POINTER(u32)]
    lib.crn_gpu_generate_mipmaps_host.argtypes = [vp, ctypes.POINTER(_ResampleParams), vp, u32, u32, u32, u32, u32, vp, u64, ctypes.POINTER(u32)]
    lib.crn_gpu_unpack_image.argtypes = [vp, u32, vp, u32, u32, vp, u32]
    lib.crn_gpu_unpack_image_host.argtypes = [vp, u32, vp, u32, u32, vp, u32]
    lib.crn_gpu_blockify.argtypes = [vp, vp, u32, u32, u32, u32, vp, ctypes.POINTER(u32), ctypes.POINTER(u32)]
    lib.crn_gpu_default_hc_params.argtypes = [ctypes.POINTER(_HcParams)]
    lib.crn_gpu_default_hc_params.restype = None
    lib.crn_gpu_hc_compress.argtypes = [vp, ctypes.POINTER(_HcParams), vp, i32, ctypes.POINTER(vp)]
    lib.crn_gpu_hc_get_info.argtypes = [vp, ctypes.POINTER(_HcInfo)]
    for name in ("endpoint_indices", "selector_indices", "color_endpoints", "alpha_endpoints", "color_selectors", "alpha_selectors", "block_encodings", "tile_indices"):
        f = getattr(lib, "crn_gpu_hc_" + name)
        f.argtypes = [vp]
        f.restype = vp
    lib.crn_gpu_hc_free.argtypes = [vp]
    lib.crn_gpu_hc_free.restype = None
    lib.crn_gpu_default_crn_params.argtypes = [ctypes.POINTER(_CrnParams)]
    lib.crn_gpu_default_crn_params.restype = None
    lib.crn_gpu_crn_hc_params.argtypes = [ctypes.POINTER(_CrnParams), ctypes.POINTER(_HcParams)]
    lib.crn_gpu_crn_write.argtypes = [ctypes.POINTER(_CrnParams), ctypes.POINTER(_HcParams), vp, vp, vp, u32, vp, u32, vp, u32, vp, u32, ctypes.POINTER(vp), ctypes.POINTER(u32)]
    lib.crn_gpu_compress_crn.argtypes = [vp, ctypes.POINTER(_CrnParams), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(u32), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(u32)]
    lib.crn_gpu_free_file.argtypes = [vp]
    lib.crn_gpu_free_file.restype = None
    lib.crn_gpu_crnd_get_texture_info.argtypes = [vp, u32, ctypes.POINTER(_TextureInfo)]
    lib.crn_gpu_crnd_unpack_begin.argtypes = [vp, vp, u32, ctypes.POINTER(vp)]
    lib.crn_gpu_crnd_unpack_level.argtypes = [vp, ctypes.POINTER(vp), u32, u32, u32]
    lib.crn_gpu_crnd_total_size.argtypes = [vp]
    lib.crn_gpu_crnd_total_size.restype = u64
    lib.crn_gpu_crnd_level_offset.argtypes = [vp, u32, u32]
    lib.crn_gpu_crnd_level_offset.restype = u64
    lib.crn_gpu_crnd_unpack_all_levels.argtypes = [vp, vp, u64]
    lib.crn_gpu_crnd_unpack_all_levels_host.argtypes = [vp, vp, u64]
    lib.crn_gpu_crnd_unpack_batch.argtypes = [vp, ctypes.POINTER(vp), u32, ctypes.POINTER(vp), ctypes.POINTER(u64)]
    lib.crn_gpu_crnd_unpack_end.argtypes = [vp]
    lib.crn_gpu_dds_header.argtypes = [u32, u32, u32, u32, u32, vp]
    lib.crn_gpu_compress_mip_chain.argtypes = [vp, u32, ctypes.POINTER(_CrnParams), ctypes.POINTER(_DdsParams), vp, u32, u32, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(u32)]
    lib.crn_gpu_default_dds_params.argtypes = [ctypes.POINTER(_DdsParams)]
    lib.crn_gpu_default_dds_params.restype = None
    lib.crn_gpu_compress_dds.argtypes = [vp, ctypes.POINTER(_DdsParams), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(u32)]
    lib.crn_gpu_compress_dds_ex.argtypes = [vp, ctypes.POINTER(_DdsParams), ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(u32), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(u32)]
    lib.crn_gpu_lzma_size.argtypes = [vp, u64]
    lib.crn_gpu_lzma_size.restype = u64
    lib.crn_gpu_crnd_unpack_level_host.argtypes = [vp, ctypes.POINTER(vp), u32, u32, u32]
    lib.crn_gpu_dds_get_desc.argtypes = [vp, u32, ctypes.POINTER(_DdsDesc)]
    lib.crn_gpu_dds_to_images.argtypes = [vp, vp, u32, ctypes.POINTER(vp), u32, ctypes.POINTER(_DdsDesc)]
    lib.crn_gpu_convert_pixels.argtypes = [vp, vp, u32, u32, u32, u32]
    lib.crn_gpu_pool_mallocs.argtypes = [vp]
    lib.crn_gpu_pool_mallocs.restype = u64
    lib.crn_gpu_set_vq_mode.argtypes = [vp, ctypes.c_int]
    lib.crn_gpu_set_vq_mode.restype = None
    lib.crn_gpu_set_progress.argtypes = [vp, vp, vp]
    lib.crn_gpu_set_progress.restype = None
    lib.crn_gpu_crn_to_dds.argtypes = [vp, vp, u32, ctypes.POINTER(vp), ctypes.POINTER(u32)]
    return lib


def load_library(path=None):
    """Loads the nvcc-built libcrn_b200.so.  Fails loudly: no emulation build, no CPU path."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or os.environ.get("CRN_B200_LIB") or library_path()   # CRN_B200_LIB: A/B builds of the same sources
    if not os.path.exists(p):
        raise CrnGpuError(-1, "%s is missing: build it with `make -C crunch2_b200/csrc` (nvcc, sm_100a); "
                              "there is no CPU fallback" % p)
    lib = _declare(ctypes.CDLL(p))
    if not lib.crn_gpu_is_native():
        raise CrnGpuError(-4, "%s is the g++ SIMT-emulation test build, not the product" % p)
    if path is None:
        _LIB = lib
    return lib


def bytes_per_block(fmt):
    return 8 if fmt in (FMT_DXT1, FMT_DXT1A, FMT_DXT5A) else 16


def _devptr(t):
    return ctypes.c_void_p(t.data_ptr())


class Context:
    """One GPU + one stream (crn_gpu_ctx).  `lib` lets the tests drive the emulation build."""

    def __init__(self, device=0, lib=None):
        self._lib = lib if lib is not None else load_library()
        self._ctx = ctypes.c_void_p()
        rc = self._lib.crn_gpu_create(int(device), ctypes.byref(self._ctx))
        if rc != 0:
            self._ctx = None
            raise CrnGpuError(rc, "crn_gpu_create(device=%d) failed (no CUDA device?)" % device)
        self.device = device

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.crn_gpu_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise CrnGpuError(rc, (self._lib.crn_gpu_last_error(self._ctx) or b"").decode())

    @property
    def stream(self):
        return self._lib.crn_gpu_stream(self._ctx)

    @property
    def launch_count(self):
        return int(self._lib.crn_gpu_launch_count(self._ctx))

    @property
    def pool_mallocs(self):
        """cudaMalloc calls of the context's buffer pool so far (stands still in steady state)"""
        return int(self._lib.crn_gpu_pool_mallocs(self._ctx))

    def set_vq_mode(self, exact_member_order):
        """crn_gpu_set_vq_mode: True = the reference's member-order float sums bit for bit (slow, verification); False = single-launch frontier splits."""
        self._lib.crn_gpu_set_vq_mode(self._ctx, 1 if exact_member_order else 0)
        self.vq_exact = bool(exact_member_order)

    def synchronize(self):
        self._check(self._lib.crn_gpu_synchronize(self._ctx))

    # --- block-by-block packing (dxt_image::init) -------------------------------------------------
    def pack_image(self, fmt, rgba, params=None):
        """rgba: (H, W, 4) uint8 numpy array (host; copies are part of the call).  Returns the packed
        blocks as a uint8 numpy array of blocks_y*blocks_x*bytes_per_block bytes."""
        params = params or PackParams()
        a = np.ascontiguousarray(rgba, dtype=np.uint8)
        if a.ndim != 3 or a.shape[2] != 4:
            raise ValueError("expected an (H, W, 4) uint8 image")
        h, w = a.shape[:2]
        out = np.empty(((w + 3) // 4) * ((h + 3) // 4) * bytes_per_block(fmt), np.uint8)
        cp = params._c()
        self._check(self._lib.crn_gpu_pack_image_host(self._ctx, fmt, ctypes.byref(cp), a.ctypes.data_as(ctypes.c_void_p),
                                                       w, h, w * 4, out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def pack_image_device(self, fmt, d_rgba, width, height, pitch, d_out, params=None):
        """Asynchronous, device pointers (ints or objects with data_ptr()) on this context's stream."""
        params = params or PackParams()
        cp = params._c()
        src = _devptr(d_rgba) if hasattr(d_rgba, "data_ptr") else ctypes.c_void_p(int(d_rgba))
        dst = _devptr(d_out) if hasattr(d_out, "data_ptr") else ctypes.c_void_p(int(d_out))
        self._check(self._lib.crn_gpu_pack_image(self._ctx, fmt, ctypes.byref(cp), src, width, height, pitch, dst))

    # --- cluster optimisers (qdxt1::pack_endpoints_task / qdxt5::pack_endpoints_task) -----------------
    def optimize_clusters(self, kind, d_blocks, n_blocks, d_offsets, d_members, n_clusters, total_member_blocks, d_out, out_stride,
                          out_offset, params=None, component=3, dxt1a=False, d_endpoints=None, d_error=None):
        """kind = "color" or "alpha".  All pointers are device pointers (ints or objects with data_ptr())."""
        params = params or PackParams()
        cp = params._c()

        def ptr(x):
            if x is None:
                return ctypes.c_void_p(0)
            return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x))
        if kind == "color":
            rc = self._lib.crn_gpu_dxt1_optimize_clusters(self._ctx, ctypes.byref(cp), int(bool(dxt1a)), ptr(d_blocks), n_blocks, ptr(d_offsets),
                                                          ptr(d_members), n_clusters, total_member_blocks, ptr(d_out), out_stride, out_offset,
                                                          ptr(d_endpoints), ptr(d_error))
        else:
            rc = self._lib.crn_gpu_dxt5_optimize_clusters(self._ctx, ctypes.byref(cp), component, ptr(d_blocks), n_blocks, ptr(d_offsets),
                                                          ptr(d_members), n_clusters, total_member_blocks, ptr(d_out), out_stride, out_offset,
                                                          ptr(d_endpoints), ptr(d_error))
        self._check(rc)

    def optimize_selectors(self, kind, d_blocks, n_blocks, d_offsets, d_members, n_clusters, d_elements, stride, offset, params=None, component=3):
        """qdxt1/qdxt5::optimize_selectors_task: kind = "color" or "alpha"; d_elements is updated in place."""
        params = params or PackParams()
        cp = params._c()

        def ptr(x):
            return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x))
        self._check(self._lib.crn_gpu_optimize_selectors(self._ctx, 0 if kind == "color" else 1, ctypes.byref(cp), component, ptr(d_blocks), n_blocks,
                                                         ptr(d_offsets), ptr(d_members), n_clusters, ptr(d_elements), stride, offset))

    def vq_clusterize(self, d_vectors, d_weights, n, dims, max_codebook_size=65535, retrieve_max_clusters=0, threaded=False):
        """clusterizer<V>::generate_codebook + retrieve_clusters (threaded: threaded_clusterizer<V>::create_clusters).
        d_vectors: u8[n][dims], d_weights: u32[n] on the device.  Returns (cluster_of[n] numpy, num_clusters, codebook_size)."""
        def ptr(x):
            return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x))
        cluster_of = np.zeros(n, np.uint32)
        k, cb = ctypes.c_uint32(), ctypes.c_uint32()
        self._check(self._lib.crn_gpu_vq_clusterize(self._ctx, dims, ptr(d_vectors), ptr(d_weights), n, max_codebook_size, retrieve_max_clusters,
                                                    1 if threaded else 0, cluster_of.ctypes.data_as(ctypes.c_void_p), ctypes.byref(k), ctypes.byref(cb)))
        return cluster_of, k.value, cb.value

    # --- dxt_hc building blocks ---------------------------------------------------------------------------
    def refine_endpoints(self, dxt1_selectors, d_pixels, d_selectors, d_offsets, n_clusters, d_endpoints, d_error, d_ok,
                         perceptual=True, component=0, d_error_to_beat=None):
        """dxt_endpoint_refiner::refine per CSR cluster of pixels; outputs low | high << 16, error, ok (device arrays)."""
        def ptr(x):
            return None if x is None else (ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x)))
        self._check(self._lib.crn_gpu_refine_endpoints(self._ctx, 1 if dxt1_selectors else 0, 1 if perceptual else 0, component, ptr(d_pixels), ptr(d_selectors),
                                                       ptr(d_offsets), n_clusters, ptr(d_error_to_beat), ptr(d_endpoints), ptr(d_error), ptr(d_ok)))

    def nearest_codebook(self, dims, d_vectors, n, d_codebook, k, d_out):
        """dxt_hc::determine_color/alpha_endpoint_clusters_task: first nearest codebook entry per float vector."""
        def ptr(x):
            return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x))
        self._check(self._lib.crn_gpu_nearest_codebook(self._ctx, dims, ptr(d_vectors), n, ptr(d_codebook), k, ptr(d_out)))

    def assign_selectors(self, kind, d_blocks, n_blocks, d_values, d_codebook, k, d_best_index, d_refined, d_used, perceptual=True, component=0, d_values_accum=None):
        """dxt_hc::create_color/alpha_selector_codebook_task + re-vote: kind = "color" or "alpha" (device arrays)."""
        def ptr(x):
            return None if x is None else (ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x)))
        self._check(self._lib.crn_gpu_assign_selectors(self._ctx, 0 if kind == "color" else 1, 1 if perceptual else 0, component, ptr(d_blocks), n_blocks, ptr(d_values),
                                                       ptr(d_values_accum), ptr(d_codebook), k, ptr(d_best_index), ptr(d_refined), ptr(d_used)))

    # --- mip chain (mipmapped_texture::generate_mipmaps, crnlib/crn_mipmapped_texture.cpp:2140-2220) ---------------
    def _resample_params(self, filter, filter_scale, srgb, source_gamma, wrapping, num_comps):
        p = _ResampleParams()
        self._lib.crn_gpu_default_resample_params(ctypes.byref(p))
        p.filter = MIP_FILTERS[filter] if isinstance(filter, str) else int(filter)
        p.filter_scale = float(filter_scale); p.srgb = int(bool(srgb)); p.source_gamma = float(source_gamma)
        p.wrapping = int(bool(wrapping)); p.num_comps = int(num_comps)
        return p

    def generate_mipmaps(self, rgba, filter="kaiser", filter_scale=0.9, srgb=True, source_gamma=2.2, wrapping=False, num_comps=4,
                         min_mip_size=1, max_levels=16):
        """rgba: (h, w, 4) uint8 -> list of levels [level0, level1, ...] (numpy), level l = max(1, w >> l) x max(1, h >> l)."""
        img = np.ascontiguousarray(rgba, np.uint8)
        h, w = img.shape[:2]
        p = self._resample_params(filter, filter_scale, srgb, source_gamma, wrapping, num_comps)
        n = self._lib.crn_gpu_mip_level_count(w, h, min_mip_size, max_levels)
        sizes = [(max(1, h >> l), max(1, w >> l)) for l in range(1, n)]
        out = np.empty(sum(a * b * 4 for a, b in sizes) or 1, np.uint8)
        nl = ctypes.c_uint32(0)
        self._check(self._lib.crn_gpu_generate_mipmaps_host(self._ctx, ctypes.byref(p), img.ctypes.data_as(ctypes.c_void_p), w, h, w * 4, min_mip_size, max_levels,
                                                            out.ctypes.data_as(ctypes.c_void_p), out.size, ctypes.byref(nl)))
        levels, o = [img], 0
        for a, b in sizes:
            levels.append(out[o:o + a * b * 4].reshape(a, b, 4).copy()); o += a * b * 4
        return levels

    def resample_device(self, d_src, sw, sh, spitch, d_dst, dw, dh, dpitch, filter="kaiser", filter_scale=1.0, srgb=True, source_gamma=2.2, wrapping=False, num_comps=4):
        def ptr(x):
            return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x))
        p = self._resample_params(filter, filter_scale, srgb, source_gamma, wrapping, num_comps)
        self._check(self._lib.crn_gpu_resample(self._ctx, ctypes.byref(p), ptr(d_src), sw, sh, spitch, ptr(d_dst), dw, dh, dpitch))

    def generate_mipmaps_device(self, d_level0, w, h, pitch, d_mips, capacity, filter="kaiser", filter_scale=0.9, srgb=True, source_gamma=2.2, wrapping=False,
                                num_comps=4, min_mip_size=1, max_levels=16):
        def ptr(x):
            return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x))
        p = self._resample_params(filter, filter_scale, srgb, source_gamma, wrapping, num_comps)
        nl = ctypes.c_uint32(0)
        self._check(self._lib.crn_gpu_generate_mipmaps(self._ctx, ctypes.byref(p), ptr(d_level0), w, h, pitch, min_mip_size, max_levels, ptr(d_mips), capacity, ctypes.byref(nl)))
        return nl.value

    # --- dxt_image::unpack (crnlib/crn_dxt_image.cpp:495-567): blocks -> RGBA8 pixels ------------------------------
    def unpack_image(self, fmt, blocks, width, height):
        """blocks: bytes / uint8 array as pack_image returns -> (height, width, 4) uint8."""
        b = np.ascontiguousarray(np.frombuffer(blocks, np.uint8) if not isinstance(blocks, np.ndarray) else blocks.view(np.uint8).ravel())
        need = ((width + 3) // 4) * ((height + 3) // 4) * bytes_per_block(fmt)
        if b.size < need:
            raise ValueError("block buffer too small")
        out = np.empty((height, width, 4), np.uint8)
        self._check(self._lib.crn_gpu_unpack_image_host(self._ctx, int(fmt), b.ctypes.data_as(ctypes.c_void_p), width, height,
                                                        out.ctypes.data_as(ctypes.c_void_p), width * 4))
        return out

    def unpack_image_device(self, fmt, d_blocks, width, height, d_rgba, pitch):
        def ptr(x):
            return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x))
        self._check(self._lib.crn_gpu_unpack_image(self._ctx, int(fmt), ptr(d_blocks), width, height, ptr(d_rgba), pitch))

    def blockify(self, d_rgba, width, height, pitch, d_blocks, pad_pixels=8):
        """crn_comp::quantize_images' gather on the device (d_rgba / d_blocks: torch CUDA tensors or device pointers); returns (blocks_x, blocks_y)."""
        def ptr(x):
            return ctypes.c_void_p(x.data_ptr()) if hasattr(x, "data_ptr") else ctypes.c_void_p(int(x))
        bx, by = ctypes.c_uint32(0), ctypes.c_uint32(0)
        self._check(self._lib.crn_gpu_blockify(self._ctx, ptr(d_rgba), width, height, pitch, pad_pixels, ptr(d_blocks), ctypes.byref(bx), ctypes.byref(by)))
        return bx.value, by.value

    # --- dxt_hc::compress (crnlib/crn_dxt_hc.cpp:98-312): blocks of all levels -> palettes + per-block indices ---------
    def hc_compress(self, fmt, blocks, levels, num_faces=1, perceptual=True, codebook_sizes=(3072, 3072, 3072, 3072),
                    deratings=(2.0, 2.0, 3.0), alpha_components=(3, 0), shard=None):
        """blocks: (n, 16, 4) uint8 numpy array (host) or a torch CUDA tensor of the same shape; levels: list of
        (first_block, num_blocks, block_width, weight) as crn_comp::quantize_images lays them out (crnlib/crn_comp.cpp:706-741);
        codebook_sizes: colour endpoints, colour selectors, alpha endpoints, alpha selectors.  shard = (rank, world,
        allgather) spreads the per-cluster optimisation of this ONE texture over `world` ranks that all make the same call;
        allgather(buf_u8, bytes_per_rank) fills every rank's slice of the numpy buffer in place (crunch2_b200.shard.
        allgather_inplace does it over torch.distributed).  Returns a dict of numpy arrays named as dxt_hc::compress's
        outputs plus 'info'."""
        p = _HcParams()
        self._lib.crn_gpu_default_hc_params(ctypes.byref(p))
        on_host = not hasattr(blocks, "data_ptr")
        if on_host:
            blocks = np.ascontiguousarray(blocks, np.uint8)
            n = blocks.shape[0]
            ptr = blocks.ctypes.data_as(ctypes.c_void_p)
        else:
            n = int(blocks.shape[0])
            ptr = ctypes.c_void_p(blocks.data_ptr())
        p.format = int(fmt); p.num_blocks = n; p.num_levels = len(levels); p.num_faces = int(num_faces)
        for i, (first, nb, bw, weight) in enumerate(levels):
            p.levels[i].first_block = int(first); p.levels[i].num_blocks = int(nb); p.levels[i].block_width = int(bw); p.levels[i].weight = float(weight)
        p.perceptual = int(bool(perceptual))
        (p.color_endpoint_codebook_size, p.color_selector_codebook_size, p.alpha_endpoint_codebook_size, p.alpha_selector_codebook_size) = [int(x) for x in codebook_sizes]
        (p.adaptive_tile_color_psnr_derating, p.adaptive_tile_alpha_psnr_derating, p.adaptive_tile_color_alpha_weighting_ratio) = [float(x) for x in deratings]
        p.alpha_component_indices[0], p.alpha_component_indices[1] = int(alpha_components[0]), int(alpha_components[1])
        cb = None
        if shard is not None and int(shard[1]) > 1:
            rank, world, gather = shard
            failure = []

            def _exchange(user, buf, bytes_per_rank, nranks):
                try:
                    arr = np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ctypes.c_uint8)), (int(bytes_per_rank) * int(nranks),))
                    gather(arr, int(bytes_per_rank))
                    return 0
                except Exception as e:  # never let an exception cross the C boundary
                    failure.append(e)
                    return 1
            cb = EXCHANGE_FN(_exchange)
            p.shard_rank, p.shard_count = int(rank), int(world)
            p.exchange = ctypes.cast(cb, ctypes.c_void_p)
        h = ctypes.c_void_p()
        rc = self._lib.crn_gpu_hc_compress(self._ctx, ctypes.byref(p), ptr, 1 if on_host else 0, ctypes.byref(h))
        if rc and cb is not None and failure:
            raise failure[0]                                  # the collective's own exception, not "the exchange callback failed"
        self._check(rc)
        try:
            info = _HcInfo()
            info.struct_size = ctypes.sizeof(_HcInfo)
            self._check(self._lib.crn_gpu_hc_get_info(h, ctypes.byref(info)))

            def arr(name, count, dtype):
                if not count:
                    return np.zeros(0, dtype)
                addr = getattr(self._lib, "crn_gpu_hc_" + name)(h)
                return np.ctypeslib.as_array(ctypes.cast(addr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype))), (count,)).copy()
            out = {
                "endpoint_indices": arr("endpoint_indices", n * 4, np.uint16).reshape(n, 4),
                "selector_indices": arr("selector_indices", n * 4, np.uint16).reshape(n, 4),
                "color_endpoints": arr("color_endpoints", info.n_color_endpoints, np.uint32),
                "alpha_endpoints": arr("alpha_endpoints", info.n_alpha_endpoints, np.uint32),
                "color_selectors": arr("color_selectors", info.n_color_selectors, np.uint32),
                "alpha_selectors": arr("alpha_selectors", info.n_alpha_selectors, np.uint64),
                "block_encodings": arr("block_encodings", n, np.uint8),
                "tile_indices": arr("tile_indices", n, np.uint32),
                "info": {"num_tiles": info.num_tiles, "vq_rounds": list(info.vq_rounds), "unique_vectors": list(info.unique_vectors)},
            }
        finally:
            self._lib.crn_gpu_hc_free(h)
        return out

    def compress_crn(self, images, crn_format, quality_level=128, perceptual=True, alpha_component=3, palette_sizes=None, userdata=(0, 0), target_bitrate=0.0,
                     shard=None):
        """crn_compress to a .CRN (crn_comp::compress_pass, crnlib/crn_comp.cpp:1613; with target_bitrate > 0 the quality search
        of crnlib/crn_texture_comp.cpp:120-262): images[face][level] = (h, w, 4) uint8 host arrays, level l being
        max(1, w >> l) x max(1, h >> l).  Returns (file bytes, bits per texel, quality level)."""
        faces, levels = len(images), len(images[0])
        h, w = images[0][0].shape[:2]
        p = crn_params(crn_format, w, h, levels, faces, quality_level, perceptual, alpha_component, palette_sizes, userdata, lib=self._lib)
        p.target_bitrate = float(target_bitrate)
        cb = None
        if shard is not None and int(shard[1]) > 1:          # (rank, world, allgather): as in hc_compress, every rank makes this same call
            rank, world, gather = shard
            failure = []

            def _exchange(user, buf, bytes_per_rank, nranks):
                try:
                    arr = np.ctypeslib.as_array(ctypes.cast(buf, ctypes.POINTER(ctypes.c_uint8)), (int(bytes_per_rank) * int(nranks),))
                    gather(arr, int(bytes_per_rank))
                    return 0
                except Exception as e:  # never let an exception cross the C boundary
                    failure.append(e)
                    return 1
            cb = EXCHANGE_FN(_exchange)
            p.shard_rank, p.shard_count = int(rank), int(world)
            p.exchange = ctypes.cast(cb, ctypes.c_void_p)
        flat = [np.ascontiguousarray(images[f][l], np.uint8) for f in range(faces) for l in range(levels)]
        for i, a in enumerate(flat):
            lw, lh = max(1, w >> (i % levels)), max(1, h >> (i % levels))
            if a.shape != (lh, lw, 4):
                raise ValueError("image %d has shape %s, expected %s" % (i, a.shape, (lh, lw, 4)))
        ptrs = (ctypes.c_void_p * len(flat))(*[a.ctypes.data for a in flat])
        out = ctypes.c_void_p(); size = ctypes.c_uint32(); rate = ctypes.c_float(); q = ctypes.c_uint32()
        rc = self._lib.crn_gpu_compress_crn(self._ctx, ctypes.byref(p), ptrs, ctypes.byref(out), ctypes.byref(size), ctypes.byref(rate), ctypes.byref(q))
        if rc and cb is not None and failure:
            raise failure[0]
        self._check(rc)
        try:
            return ctypes.string_at(out, size.value), rate.value, q.value
        finally:
            self._lib.crn_gpu_free_file(out)

    # --- clustered DDS compression (mipmapped_texture::qdxt_pack_init / qdxt_pack) ----------------------
    def qdxt_init(self, fmt, levels, params=None):
        """levels: list of (h, w, 4) uint8 arrays -- numpy (host pixels) or torch CUDA tensors -- faces x mips in
        the reference's order.  Returns a Qdxt whose pack(quality_level) gives the packed blocks of all levels."""
        return Qdxt(self, fmt, levels, params or PackParams())

    # --- CRN -> DXTn transcoding (crnd_unpack_begin / crnd_unpack_level / crnd_unpack_end) -----------
    def compress_dds(self, images, crn_format, quality_level=255, params=None, dxt1a_for_transparency=False, target_bitrate=0.0, with_rate=False, hierarchical=True):
        """crn_compress to a .DDS (dds_comp, crnlib/crn_dds_comp.cpp:148-289): images[face][level] = (h, w, 4) uint8 host arrays.
        quality_level 255 packs block by block, lower values take the clustered path.  Returns the file bytes."""
        faces, levels = len(images), len(images[0])
        h, w = images[0][0].shape[:2]
        p = _DdsParams()
        self._lib.crn_gpu_default_dds_params(ctypes.byref(p))
        p.crn_format, p.width, p.height, p.levels, p.faces = int(crn_format), int(w), int(h), int(levels), int(faces)
        p.quality_level, p.dxt1a_for_transparency = int(quality_level), int(bool(dxt1a_for_transparency))
        p.hierarchical = int(bool(hierarchical))
        if params is not None:
            p.pack = params._c()
        flat = [np.ascontiguousarray(images[f][l], np.uint8) for f in range(faces) for l in range(levels)]
        ptrs = (ctypes.c_void_p * len(flat))(*[a.ctypes.data for a in flat])
        out = ctypes.c_void_p(); size = ctypes.c_uint32()
        if not with_rate and not target_bitrate:
            self._check(self._lib.crn_gpu_compress_dds(self._ctx, ctypes.byref(p), ptrs, ctypes.byref(out), ctypes.byref(size)))
            try:
                return ctypes.string_at(out, size.value)
            finally:
                self._lib.crn_gpu_free_file(out)
        # crn_compress's optional outputs: (file, LZMA bits per texel, quality level); target_bitrate runs the reference's search
        p.target_bitrate = float(target_bitrate)
        rate = ctypes.c_float(); q = ctypes.c_uint32()
        self._check(self._lib.crn_gpu_compress_dds_ex(self._ctx, ctypes.byref(p), ptrs, ctypes.byref(out), ctypes.byref(size), ctypes.byref(rate), ctypes.byref(q)))
        try:
            return ctypes.string_at(out, size.value), rate.value, q.value
        finally:
            self._lib.crn_gpu_free_file(out)

    def compress_mip_chain(self, faces_level0, crn_format, file_type="dds", quality_level=255, params=None, perceptual=True):
        """crn_compress with default crn_mipmap_params (inc/crnlib.h:614): faces_level0 = list of (h, w, 4) uint8 level-0 images;
        the mip chain is generated on the device, then compressed to a .dds (file_type "dds") or .crn ("crn").  Returns bytes."""
        imgs = [np.ascontiguousarray(a, np.uint8) for a in faces_level0]
        h, w = imgs[0].shape[:2]
        ptrs = (ctypes.c_void_p * len(imgs))(*[a.ctypes.data for a in imgs])
        out = ctypes.c_void_p(); size = ctypes.c_uint32()
        if file_type == "crn":
            cp = crn_params(crn_format, w, h, 1, len(imgs), quality_level, perceptual, lib=self._lib)
            rc = self._lib.crn_gpu_compress_mip_chain(self._ctx, 0, ctypes.byref(cp), None, None, 0, 0, ptrs, ctypes.byref(out), ctypes.byref(size))
        else:
            dp = _DdsParams()
            self._lib.crn_gpu_default_dds_params(ctypes.byref(dp))
            dp.crn_format, dp.width, dp.height, dp.levels, dp.faces, dp.quality_level = int(crn_format), int(w), int(h), 1, len(imgs), int(quality_level)
            if params is not None:
                dp.pack = params._c()
            rc = self._lib.crn_gpu_compress_mip_chain(self._ctx, 1, None, ctypes.byref(dp), None, 0, 0, ptrs, ctypes.byref(out), ctypes.byref(size))
        self._check(rc)
        try:
            return ctypes.string_at(out, size.value)
        finally:
            self._lib.crn_gpu_free_file(out)

    def dds_to_images(self, dds_bytes):
        """crn_decompress_dds_to_images (inc/crnlib.h:634): .dds bytes -> (list of (h, w, 4) uint8 images indexed level + levels * face, desc dict).
        Block formats are decoded and uncooked on the device; the reference's uncompressed layouts go through the mask-extraction kernel."""
        buf = np.frombuffer(dds_bytes, np.uint8)
        d = _DdsDesc(); d.struct_size = ctypes.sizeof(_DdsDesc)
        rc = self._lib.crn_gpu_dds_get_desc(buf.ctypes.data_as(ctypes.c_void_p), len(dds_bytes), ctypes.byref(d))
        if rc:
            raise CrnGpuError(rc, "not a .dds file this path reads")
        imgs = []
        for f in range(d.faces):
            for l in range(d.levels):
                imgs.append(np.empty((max(1, d.height >> l), max(1, d.width >> l), 4), np.uint8))
        # index level + levels * face
        ptrs = (ctypes.c_void_p * len(imgs))(*[a.ctypes.data for a in imgs])
        self._check(self._lib.crn_gpu_dds_to_images(self._ctx, buf.ctypes.data_as(ctypes.c_void_p), len(dds_bytes), ptrs, len(imgs), ctypes.byref(d)))
        return imgs, {n: getattr(d, n) for n in ("width", "height", "levels", "faces", "pixel_format", "file_format", "block_format")}

    def convert_pixels_device(self, d_rgba, width, height, pitch, conversion):
        """image_utils::convert_image on a device image in place (1 To_CCxY ... 9 XY_to_XYZ, include/crn_b200.h)."""
        self._check(self._lib.crn_gpu_convert_pixels(self._ctx, ctypes.c_void_p(int(d_rgba)), int(width), int(height), int(pitch), int(conversion)))

    def crn_to_dds(self, crn_bytes):
        """crn_decompress_crn_to_dds (inc/crnlib.h:620): .crn bytes -> .dds bytes, transcoded on the device."""
        buf = np.frombuffer(crn_bytes, np.uint8)
        out = ctypes.c_void_p(); size = ctypes.c_uint32()
        self._check(self._lib.crn_gpu_crn_to_dds(self._ctx, buf.ctypes.data_as(ctypes.c_void_p), len(crn_bytes), ctypes.byref(out), ctypes.byref(size)))
        try:
            return ctypes.string_at(out, size.value)
        finally:
            self._lib.crn_gpu_free_file(out)

    def unpack_begin(self, crn_bytes):
        """crnd_unpack_begin: returns a Texture bound to this context."""
        return Texture(self, crn_bytes)

    def unpack_batch(self, textures, d_dst_ptrs, capacities):
        """All levels of several textures in ONE launch; d_dst_ptrs are device pointers (ints)."""
        n = len(textures)
        tp = (ctypes.c_void_p * n)(*[t._tex for t in textures])
        dp = (ctypes.c_void_p * n)(*[ctypes.c_void_p(int(p)) for p in d_dst_ptrs])
        cp = (ctypes.c_uint64 * n)(*[int(c) for c in capacities])
        self._check(self._lib.crn_gpu_crnd_unpack_batch(self._ctx, tp, n, dp, cp))


def crn_params(crn_format, width, height, levels=1, faces=1, quality_level=128, perceptual=True, alpha_component=3, palette_sizes=None, userdata=(0, 0), lib=None):
    lib = lib if lib is not None else load_library()
    p = _CrnParams()
    lib.crn_gpu_default_crn_params(ctypes.byref(p))
    p.crn_format, p.width, p.height, p.levels, p.faces = int(crn_format), int(width), int(height), int(levels), int(faces)
    p.quality_level, p.perceptual, p.alpha_component = int(quality_level), int(bool(perceptual)), int(alpha_component)
    p.userdata0, p.userdata1 = int(userdata[0]), int(userdata[1])
    if palette_sizes is not None:
        for i in range(4):
            p.palette_sizes[i] = int(palette_sizes[i])
    return p


def crn_hc_params(params, lib=None):
    """crn_comp's level table and dxt_hc parameters for a .CRN of these crn_params (host only)."""
    lib = lib if lib is not None else load_library()
    hp = _HcParams()
    rc = lib.crn_gpu_crn_hc_params(ctypes.byref(params), ctypes.byref(hp))
    if rc != 0:
        raise CrnGpuError(rc, "crn_gpu_crn_hc_params failed (%d)" % rc)
    return hp


def crn_write(params, hc_params, out, lib=None):
    """Palettes + indices (dict named as Context.hc_compress's result) -> .crn bytes; the writer back-end of crn_comp
    (crnlib/crn_comp.cpp:767-1496).  Host only."""
    lib = lib if lib is not None else load_library()
    keep = {k: np.ascontiguousarray(out[k], dt) for k, dt in (("endpoint_indices", np.uint16), ("selector_indices", np.uint16), ("color_endpoints", np.uint32),
                                                              ("alpha_endpoints", np.uint32), ("color_selectors", np.uint32), ("alpha_selectors", np.uint64))}

    def ptr(k):
        return keep[k].ctypes.data_as(ctypes.c_void_p) if keep[k].size else None
    f = ctypes.c_void_p(); size = ctypes.c_uint32()
    rc = lib.crn_gpu_crn_write(ctypes.byref(params), ctypes.byref(hc_params), ptr("endpoint_indices"), ptr("selector_indices"),
                               ptr("color_endpoints"), keep["color_endpoints"].size, ptr("alpha_endpoints"), keep["alpha_endpoints"].size,
                               ptr("color_selectors"), keep["color_selectors"].size, ptr("alpha_selectors"), keep["alpha_selectors"].size,
                               ctypes.byref(f), ctypes.byref(size))
    if rc != 0:
        raise CrnGpuError(rc, "crn_gpu_crn_write failed (%d)" % rc)
    try:
        return ctypes.string_at(f, size.value)
    finally:
        lib.crn_gpu_free_file(f)


def texture_info(crn_bytes, lib=None):
    """crnd_get_texture_info: host-only header crack."""
    lib = lib if lib is not None else load_library()
    info = _TextureInfo()
    info.struct_size = ctypes.sizeof(_TextureInfo)
    buf = np.frombuffer(crn_bytes, np.uint8)
    rc = lib.crn_gpu_crnd_get_texture_info(buf.ctypes.data_as(ctypes.c_void_p), len(crn_bytes), ctypes.byref(info))
    if rc != 0:
        raise CrnGpuError(rc, "not a CRN file")
    return {k: getattr(info, k) for k, _ in _TextureInfo._fields_ if k != "struct_size"}


class _LevelDesc(ctypes.Structure):
    _fields_ = [("rgba", ctypes.c_void_p), ("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("pitch_bytes", ctypes.c_uint32)]


class _QdxtInfo(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_uint32), ("n_blocks", ctypes.c_uint32), ("num_elements", ctypes.c_uint32),
                ("endpoint_codebook_size", ctypes.c_uint32 * 3), ("max_selector_clusters", ctypes.c_uint32 * 3),
                ("endpoint_clusters", ctypes.c_uint32 * 3), ("selector_clusters", ctypes.c_uint32 * 3), ("endpoint_opt_ms", ctypes.c_float * 3),
                ("opt_candidates", ctypes.c_uint64 * 3), ("opt_colour_evals", ctypes.c_uint64 * 3), ("opt_palette_entries", ctypes.c_uint32 * 3), ("reserved", ctypes.c_uint32)]


class Qdxt:
    """mipmapped_texture::qdxt_state: pixel blocks, endpoint trees and selector bounds of one texture on the device."""

    def __init__(self, ctx, fmt, levels, params):
        self._c = ctx
        self._q = ctypes.c_void_p()
        self.fmt = fmt
        on_host = not hasattr(levels[0], "data_ptr")
        keep, descs = [], (_LevelDesc * len(levels))()
        for i, lv in enumerate(levels):
            if on_host:
                lv = np.ascontiguousarray(lv, np.uint8)
                ptr, h, w = lv.ctypes.data, lv.shape[0], lv.shape[1]
            else:
                lv = lv.contiguous()
                ptr, h, w = lv.data_ptr(), lv.shape[0], lv.shape[1]
            keep.append(lv)
            descs[i] = _LevelDesc(ptr, w, h, w * 4)
        cp = params._c()
        ctx._check(ctx._lib.crn_gpu_qdxt_init(ctx._ctx, fmt, ctypes.byref(cp), descs, len(levels), 1 if on_host else 0, ctypes.byref(self._q)))
        self.size = ctx._lib.crn_gpu_qdxt_output_size(self._q)
        self.level_offsets = [ctx._lib.crn_gpu_qdxt_level_offset(self._q, i) for i in range(len(levels))]

    def pack(self, quality_level, out=None):
        """qdxt_pack at crn_comp_params::m_quality_level (0..255).  out: optional torch CUDA uint8 tensor; default returns numpy."""
        if out is not None:
            self._c._check(self._c._lib.crn_gpu_qdxt_pack(self._q, int(quality_level), ctypes.c_void_p(out.data_ptr()), 0))
            return out
        buf = np.zeros(self.size, np.uint8)
        self._c._check(self._c._lib.crn_gpu_qdxt_pack(self._q, int(quality_level), buf.ctypes.data_as(ctypes.c_void_p), 1))
        return buf

    def info(self):
        inf = _QdxtInfo()
        inf.struct_size = ctypes.sizeof(_QdxtInfo)
        self._c._check(self._c._lib.crn_gpu_qdxt_get_info(self._q, ctypes.byref(inf)))
        ne = inf.num_elements
        return dict(n_blocks=inf.n_blocks, num_elements=ne, endpoint_codebook_size=list(inf.endpoint_codebook_size)[:ne],
                    max_selector_clusters=list(inf.max_selector_clusters)[:ne], endpoint_clusters=list(inf.endpoint_clusters)[:ne],
                    selector_clusters=list(inf.selector_clusters)[:ne], endpoint_opt_ms=list(inf.endpoint_opt_ms)[:ne],
                    opt_candidates=list(inf.opt_candidates)[:ne], opt_colour_evals=list(inf.opt_colour_evals)[:ne], opt_palette_entries=list(inf.opt_palette_entries)[:ne])

    def close(self):
        if getattr(self, "_q", None):
            self._c._lib.crn_gpu_qdxt_free(self._q)
            self._q = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Texture:
    """crnd_unpack_context: header + Huffman tables + palettes of one .crn, resident on the device."""

    def __init__(self, ctx, crn_bytes):
        self._c = ctx
        self._lib = ctx._lib
        self.info = texture_info(crn_bytes, self._lib)
        buf = np.frombuffer(crn_bytes, np.uint8)
        self._tex = ctypes.c_void_p()
        ctx._check(self._lib.crn_gpu_crnd_unpack_begin(ctx._ctx, buf.ctypes.data_as(ctypes.c_void_p), len(crn_bytes), ctypes.byref(self._tex)))
        self.total_size = int(self._lib.crn_gpu_crnd_total_size(self._tex))

    def level_offset(self, level, face=0):
        return int(self._lib.crn_gpu_crnd_level_offset(self._tex, level, face))

    def level_blocks(self, level):
        w, h = max(1, self.info["width"] >> level), max(1, self.info["height"] >> level)
        return (w + 3) // 4, (h + 3) // 4

    def unpack_all(self):
        """Every level and face (tight, level-major / face-major) as one uint8 array on the host."""
        out = np.empty(self.total_size, np.uint8)
        self._c._check(self._lib.crn_gpu_crnd_unpack_all_levels_host(self._tex, out.ctypes.data_as(ctypes.c_void_p), out.size))
        return out

    def unpack_all_device(self, d_dst, capacity):
        p = ctypes.c_void_p(d_dst.data_ptr()) if hasattr(d_dst, "data_ptr") else ctypes.c_void_p(int(d_dst))
        self._c._check(self._lib.crn_gpu_crnd_unpack_all_levels(self._tex, p, int(capacity)))

    def unpack_level_device(self, face_ptrs, dst_size, row_pitch, level):
        """crnd_unpack_level with device destination pointers (ints), one per face."""
        n = len(face_ptrs)
        arr = (ctypes.c_void_p * n)(*[ctypes.c_void_p(int(p)) for p in face_ptrs])
        self._c._check(self._lib.crn_gpu_crnd_unpack_level(self._tex, arr, int(dst_size), int(row_pitch), int(level)))

    def close(self):
        if getattr(self, "_tex", None):
            self._lib.crn_gpu_crnd_unpack_end(self._tex)
            self._tex = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
