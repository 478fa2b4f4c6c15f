"""dct_carver_b200 — B200-native DCT-Carver energy hot path.

Thin ctypes mirror of the C ABI in include/dctc.h (the product is the CUDA library, not this file).
Names follow the reference: EnergyParameters (src/render.h:9-18), blocksize / edges / textures
(src/main.h:12-22).  There is no CPU fallback: if libdctc.so is missing or no CUDA device is present,
loading / context creation raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCTC_LIB") or os.path.join(_HERE, "libdctc.so")   # DCTC_LIB: experimental builds (tools/build_exp.sh)

OK = 0
ERR_INVALID, ERR_BLOCKSIZE, ERR_NOMEM, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6, -7
KERNEL_AUTO, KERNEL_FP32_TILE, KERNEL_FP32_MARCH, KERNEL_TC_SPLIT, KERNEL_FP32_STREAM = 0, 1, 2, 3, 4

# every symbol include/dctc.h declares (tests check that the library exports exactly these)
ABI_SYMBOLS = [
    "dctc_create", "dctc_destroy", "dctc_set_params", "dctc_set_kernel", "dctc_last_cuda_error", "dctc_strerror",
    "dctc_version", "dctc_device_count", "dctc_stream", "dctc_launch_count",
    "dctc_energy_full", "dctc_energy_full_dev", "dctc_energy_batch_dev", "dctc_energy_band_dev", "dctc_energy_band_dev_at", "dctc_energy_batch",
    "dctc_carver_load", "dctc_carver_width", "dctc_carver_height", "dctc_carver_energy", "dctc_carve_and_update",
    "dctc_carver_image", "dctc_carver_resize_width", "dctc_carver_enlarge_width", "dctc_carver_set_incremental", "dctc_carver_rebuild_count", "dctc_pixel_energy",
    "dctc_energy_minmax_dev", "dctc_energy_image_dev", "dctc_carver_energy_image", "dctc_preview_energy",
    "dctc_carver_set_dump_vmaps", "dctc_carver_vmap", "dctc_carver_paint_seams",
    "dctc_synth_fill_dev", "dctc_synth_byte", "dctc_ipc_export", "dctc_ipc_open", "dctc_ipc_close",
    "dctc_dev_alloc", "dctc_dev_free", "dctc_host_alloc_pinned", "dctc_host_alloc_pinned_wc", "dctc_host_free_pinned", "dctc_memcpy_h2d",
    "dctc_memcpy_d2h", "dctc_memset_dev", "dctc_sync", "dctc_timer_begin", "dctc_timer_end",
    # multi-GPU host layer (csrc/dctc_multi.cu)
    "dctc_band_plan", "dctc_multi_create", "dctc_multi_destroy", "dctc_multi_device_count", "dctc_multi_context",
    "dctc_multi_set_params", "dctc_multi_set_kernel", "dctc_multi_launch_count", "dctc_multi_energy_batch",
    "dctc_multi_energy_bands", "dctc_multi_bands_create", "dctc_multi_bands_destroy", "dctc_multi_bands_geometry",
    "dctc_multi_bands_upload", "dctc_multi_bands_synth", "dctc_multi_bands_energy", "dctc_multi_bands_download",
    "dctc_multi_bands_energy_image", "dctc_rendezvous_allgather", "dctc_band_runner_create", "dctc_band_runner_geometry",
    "dctc_band_runner_image_dev", "dctc_band_runner_energy_dev", "dctc_band_runner_synth", "dctc_band_runner_upload",
    "dctc_band_runner_connect", "dctc_band_runner_barrier", "dctc_band_runner_energy", "dctc_band_runner_download",
    "dctc_band_runner_energy_image", "dctc_band_runner_destroy", "dctc_pcie_probe",
]


class EnergyParameters(C.Structure):
    """Bit-for-bit the reference's EnergyParameters (src/render.h:9-18)."""
    _fields_ = [("edges", C.c_float), ("textures", C.c_float), ("blocksize", C.c_int),
                ("ip", C.POINTER(C.c_int)), ("w", C.POINTER(C.c_double)), ("data", C.POINTER(C.POINTER(C.c_double)))]


class CarverEnergyParams(C.Structure):
    _fields_ = [("base", EnergyParameters), ("gpu", C.c_void_p)]


class DctcError(RuntimeError):
    def __init__(self, status, what=""):
        self.status = status
        msg = "%s failed: %d" % (what, status)
        try:
            msg += " (%s)" % lib().dctc_strerror(status).decode()
        except Exception:
            pass
        super().__init__(msg)


_lib = None


def lib():
    """Loads libdctc.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("dct_carver_b200: %s is missing — run `python -m dct_carver_b200.build` "
                          "(needs nvcc). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, u32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_float
    sig = {
        "dctc_create": (i32, [C.POINTER(vp), i32]),
        "dctc_destroy": (None, [vp]),
        "dctc_set_params": (i32, [vp, C.POINTER(EnergyParameters)]),
        "dctc_set_kernel": (i32, [vp, i32]),
        "dctc_last_cuda_error": (i32, [vp]),
        "dctc_strerror": (C.c_char_p, [i32]),
        "dctc_version": (i32, []),
        "dctc_device_count": (i32, []),
        "dctc_stream": (vp, [vp]),
        "dctc_launch_count": (C.c_ulonglong, [vp]),
        "dctc_energy_full": (i32, [vp, vp, i32, i32, i32, sz, vp]),
        "dctc_energy_full_dev": (i32, [vp, vp, i32, i32, i32, sz, vp, sz, i32]),
        "dctc_energy_batch_dev": (i32, [vp, vp, i32, sz, i32, i32, i32, sz, vp, sz, sz, i32]),
        "dctc_energy_band_dev": (i32, [vp, vp, i32, i32, i32, sz, vp, i32, sz, vp, i32, sz, vp, sz, i32]),
        "dctc_energy_band_dev_at": (i32, [vp, vp, i32, i32, i32, i32, sz, vp, i32, sz, vp, i32, sz, vp, sz, i32]),
        "dctc_energy_batch": (i32, [vp, vp, i32, sz, i32, i32, i32, sz, vp, sz]),
        "dctc_carver_load": (i32, [vp, vp, i32, i32, i32, sz]),
        "dctc_carver_width": (i32, [vp]),
        "dctc_carver_height": (i32, [vp]),
        "dctc_carver_energy": (i32, [vp, vp]),
        "dctc_carve_and_update": (i32, [vp, vp, vp, vp, vp]),
        "dctc_carver_image": (i32, [vp, vp]),
        "dctc_carver_resize_width": (i32, [vp, i32, vp]),
        "dctc_carver_enlarge_width": (i32, [vp, i32, vp]),
        "dctc_carver_rebuild_count": (i32, [vp]),
        "dctc_carver_set_incremental": (i32, [vp, i32]),
        "dctc_energy_minmax_dev": (i32, [vp, vp, C.c_size_t, i32, i32, vp]),
        "dctc_energy_image_dev": (i32, [vp, vp, C.c_size_t, i32, i32, vp, vp, C.c_size_t, i32]),
        "dctc_carver_energy_image": (i32, [vp, vp]),
        "dctc_preview_energy": (i32, [vp, vp, i32, i32, i32, C.c_size_t, vp, vp]),
        "dctc_carver_set_dump_vmaps": (i32, [vp, i32]),
        "dctc_carver_vmap": (i32, [vp, vp, vp]),
        "dctc_carver_paint_seams": (i32, [vp, vp, i32, C.c_size_t]),
        "dctc_pixel_energy": (f32, [i32, i32, i32, i32, vp, vp]),
        "dctc_synth_fill_dev": (i32, [vp, vp, i32, sz, i32, i32, i32, sz, u32, i32, i32, i32]),
        "dctc_synth_byte": (C.c_uint8, [u32, u32, u32, u32, u32, i32]),
        "dctc_ipc_export": (i32, [vp, vp, vp]),
        "dctc_ipc_open": (i32, [vp, vp, C.POINTER(vp)]),
        "dctc_ipc_close": (i32, [vp, vp]),
        "dctc_dev_alloc": (i32, [vp, C.POINTER(vp), sz]),
        "dctc_dev_free": (i32, [vp, vp]),
        "dctc_host_alloc_pinned": (i32, [C.POINTER(vp), sz]),
        "dctc_host_alloc_pinned_wc": (i32, [C.POINTER(vp), sz]),
        "dctc_host_free_pinned": (i32, [vp]),
        "dctc_memcpy_h2d": (i32, [vp, vp, vp, sz]),
        "dctc_memcpy_d2h": (i32, [vp, vp, vp, sz]),
        "dctc_memset_dev": (i32, [vp, vp, i32, sz]),
        "dctc_sync": (i32, [vp]),
        "dctc_timer_begin": (i32, [vp]),
        "dctc_timer_end": (i32, [vp, C.POINTER(f32)]),
        "dctc_band_plan": (i32, [i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "dctc_multi_create": (i32, [C.POINTER(vp), C.POINTER(i32), i32]),
        "dctc_multi_destroy": (None, [vp]),
        "dctc_multi_device_count": (i32, [vp]),
        "dctc_multi_context": (vp, [vp, i32]),
        "dctc_multi_set_params": (i32, [vp, C.POINTER(EnergyParameters)]),
        "dctc_multi_set_kernel": (i32, [vp, i32]),
        "dctc_multi_launch_count": (C.c_ulonglong, [vp]),
        "dctc_multi_energy_batch": (i32, [vp, vp, i32, sz, i32, i32, i32, sz, vp, sz]),
        "dctc_multi_energy_bands": (i32, [vp, vp, i32, i32, i32, sz, vp, vp]),
        "dctc_multi_bands_create": (i32, [vp, i32, i32, i32, C.POINTER(vp)]),
        "dctc_multi_bands_destroy": (None, [vp]),
        "dctc_multi_bands_geometry": (i32, [vp, i32, C.POINTER(i32), C.POINTER(i32)]),
        "dctc_multi_bands_upload": (i32, [vp, vp, sz]),
        "dctc_multi_bands_synth": (i32, [vp, u32, i32]),
        "dctc_multi_bands_energy": (i32, [vp, i32]),
        "dctc_multi_bands_download": (i32, [vp, vp]),
        "dctc_multi_bands_energy_image": (i32, [vp, vp]),
        "dctc_rendezvous_allgather": (i32, [C.c_char_p, i32, i32, vp, sz, vp, i32]),
        "dctc_band_runner_create": (i32, [vp, C.c_char_p, i32, i32, i32, i32, i32, C.POINTER(vp)]),
        "dctc_band_runner_geometry": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(sz)]),
        "dctc_band_runner_image_dev": (vp, [vp]),
        "dctc_band_runner_energy_dev": (vp, [vp]),
        "dctc_band_runner_synth": (i32, [vp, u32, i32]),
        "dctc_band_runner_upload": (i32, [vp, vp, sz]),
        "dctc_band_runner_connect": (i32, [vp]),
        "dctc_band_runner_barrier": (i32, [vp]),
        "dctc_band_runner_energy": (i32, [vp, i32]),
        "dctc_band_runner_download": (i32, [vp, vp]),
        "dctc_band_runner_energy_image": (i32, [vp, vp]),
        "dctc_band_runner_destroy": (None, [vp]),
        "dctc_pcie_probe": (i32, [vp, sz, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(status, what):
    if status != OK:
        raise DctcError(status, what)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def pinned_array(shape, dtype, write_combined=False):
    """numpy array over cudaMallocHost memory (kept alive by the returned array's base); write_combined: for input frames
    the host only writes."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    if write_combined:
        _check(lib().dctc_host_alloc_pinned_wc(C.byref(p), n), "dctc_host_alloc_pinned_wc")
    else:
        _check(lib().dctc_host_alloc_pinned(C.byref(p), n), "dctc_host_alloc_pinned")
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[id(buf)] = (buf, p.value)
    return arr


_PINNED = {}


class Context:
    """One GPU context (one caller at a time, like the reference's single-threaded callback)."""

    def __init__(self, device=0, blocksize=8, edges=0.5, textures=0.5, kernel=KERNEL_AUTO):
        self._h = C.c_void_p()
        _check(lib().dctc_create(C.byref(self._h), device), "dctc_create")
        self.device = device
        self.set_params(blocksize, edges, textures)
        if kernel != KERNEL_AUTO:
            self.set_kernel(kernel)

    # -- parameters ------------------------------------------------------------------------------------
    def set_params(self, blocksize, edges, textures):
        p = EnergyParameters(edges=edges, textures=textures, blocksize=blocksize)
        _check(lib().dctc_set_params(self._h, C.byref(p)), "dctc_set_params")
        self.blocksize, self.edges, self.textures = blocksize, edges, textures

    def set_kernel(self, kernel):
        _check(lib().dctc_set_kernel(self._h, kernel), "dctc_set_kernel")

    @property
    def handle(self):
        return self._h

    @property
    def launches(self):
        return int(lib().dctc_launch_count(self._h))

    def close(self):
        if self._h:
            lib().dctc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- K1 ---------------------------------------------------------------------------------------------
    def energy_full(self, img):
        """img: uint8 array [h, w] or [h, w, channels] (C-contiguous rows) -> float32 [h, w]."""
        img = np.asarray(img)
        if img.dtype != np.uint8:
            raise TypeError("image must be uint8")
        if img.ndim == 2:
            img = img[:, :, None]
        img = np.ascontiguousarray(img)
        h, w, ch = img.shape
        out = np.empty((h, w), np.float32)
        _check(lib().dctc_energy_full(self._h, _ptr(img), w, h, ch, w * ch, _ptr(out)), "dctc_energy_full")
        return out

    def energy_batch(self, imgs, out=None):
        """imgs: uint8 [n, h, w, ch] host array (pinned for full speed) -> float32 [n, h, w]."""
        n, h, w, ch = imgs.shape
        if out is None:
            out = np.empty((n, h, w), np.float32)
        _check(lib().dctc_energy_batch(self._h, _ptr(imgs), n, h * w * ch, w, h, ch, w * ch, _ptr(out), h * w),
               "dctc_energy_batch")
        return out

    # -- device memory helpers ----------------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        _check(lib().dctc_dev_alloc(self._h, C.byref(p), nbytes), "dctc_dev_alloc")
        return p.value

    def dev_free(self, p):
        _check(lib().dctc_dev_free(self._h, C.c_void_p(p)), "dctc_dev_free")

    def h2d(self, d_ptr, arr):
        arr = np.ascontiguousarray(arr)
        _check(lib().dctc_memcpy_h2d(self._h, C.c_void_p(d_ptr), _ptr(arr), arr.nbytes), "dctc_memcpy_h2d")

    def d2h(self, arr, d_ptr):
        _check(lib().dctc_memcpy_d2h(self._h, _ptr(arr), C.c_void_p(d_ptr), arr.nbytes), "dctc_memcpy_d2h")

    def sync(self):
        _check(lib().dctc_sync(self._h), "dctc_sync")

    def ipc_export(self, d_ptr):
        buf = (C.c_ubyte * 64)()
        _check(lib().dctc_ipc_export(self._h, C.c_void_p(d_ptr), buf), "dctc_ipc_export")
        return bytes(buf)

    def ipc_open(self, handle):
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _check(lib().dctc_ipc_open(self._h, buf, C.byref(p)), "dctc_ipc_open")
        return p.value

    def ipc_close(self, d_ptr):
        _check(lib().dctc_ipc_close(self._h, C.c_void_p(d_ptr)), "dctc_ipc_close")

    def timer_begin(self):
        _check(lib().dctc_timer_begin(self._h), "dctc_timer_begin")

    def timer_end(self):
        ms = C.c_float()
        _check(lib().dctc_timer_end(self._h, C.byref(ms)), "dctc_timer_end")
        return ms.value

    def synth_fill_dev(self, d_img, n_frames, frame_stride, w, h, ch, pitch, seed, pattern=0, first_frame=0, y_offset=0):
        _check(lib().dctc_synth_fill_dev(self._h, C.c_void_p(d_img), n_frames, frame_stride, w, h, ch, pitch, seed,
                                         pattern, first_frame, y_offset), "dctc_synth_fill_dev")

    def energy_batch_dev(self, d_imgs, n, frame_stride, w, h, ch, pitch, d_out, out_frame_stride, out_pitch, sync=False):
        _check(lib().dctc_energy_batch_dev(self._h, C.c_void_p(d_imgs), n, frame_stride, w, h, ch, pitch,
                                           C.c_void_p(d_out), out_frame_stride, out_pitch, int(sync)),
               "dctc_energy_batch_dev")

    def energy_band_dev(self, d_band, w, rows, ch, pitch, d_top, top_rows, top_pitch, d_bot, bot_rows, bot_pitch, d_out,
                        out_pitch, sync=False, band_y0=0):
        """band_y0: image row of the band's first row (dctc_energy_band_dev_at); 0 = dctc_energy_band_dev."""
        _check(lib().dctc_energy_band_dev_at(self._h, C.c_void_p(d_band), w, rows, band_y0, ch, pitch,
                                             C.c_void_p(d_top) if d_top else None, top_rows, top_pitch,
                                             C.c_void_p(d_bot) if d_bot else None, bot_rows, bot_pitch,
                                             C.c_void_p(d_out), out_pitch, int(sync)), "dctc_energy_band_dev_at")

    # -- K2 carver session ----------------------------------------------------------------------------------
    def carver_load(self, img):
        img = np.asarray(img)
        if img.ndim == 2:
            img = img[:, :, None]
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, ch = img.shape
        _check(lib().dctc_carver_load(self._h, _ptr(img), w, h, ch, w * ch), "dctc_carver_load")
        self._carver_ch = ch

    def carver_size(self):
        return lib().dctc_carver_width(self._h), lib().dctc_carver_height(self._h)

    def carver_energy(self):
        w, h = self.carver_size()
        out = np.empty((h, w), np.float32)
        _check(lib().dctc_carver_energy(self._h, _ptr(out)), "dctc_carver_energy")
        return out

    def carver_image(self):
        w, h = self.carver_size()
        out = np.empty((h, w, self._carver_ch), np.uint8)
        _check(lib().dctc_carver_image(self._h, _ptr(out)), "dctc_carver_image")
        return out

    def carver_resize_width(self, n_seams):
        """Whole retarget loop on the device (seam DP, back-track, carve, band energy update per seam); returns the
        removed columns, shape (n_seams, h), in the coordinates of the image at the time of removal."""
        w, h = self.carver_size()
        seams = np.empty((n_seams, h), np.int32)
        _check(lib().dctc_carver_resize_width(self._h, int(n_seams), _ptr(seams) if n_seams else None),
               "dctc_carver_resize_width")
        return seams

    def carver_enlarge_width(self, n_seams):
        """lqr_carver_resize to a larger width on the device: returns the n_seams seams (as carver_resize_width does);
        the session continues on the enlarged image."""
        w, h = self.carver_size()
        seams = np.empty((n_seams, h), np.int32)
        _check(lib().dctc_carver_enlarge_width(self._h, int(n_seams), _ptr(seams) if n_seams else None),
               "dctc_carver_enlarge_width")
        return seams

    def carver_set_incremental(self, on):
        """Device seam loop: update the cumulative map incrementally (liblqr's update_mmap) instead of rebuilding it."""
        _check(lib().dctc_carver_set_incremental(self._h, int(bool(on))), "dctc_carver_set_incremental")

    def carver_rebuild_count(self):
        """Full cumulative-map rebuilds the device seam loop needed since carver_load (the other seams were served by
        the incremental update)."""
        return int(lib().dctc_carver_rebuild_count(self._h))

    def preview_energy(self, img, want_image=True):
        """Preview-path operator (dct_energy_preview, src/render.c:421-501): returns (energy float map, normalised
        8-bit image with the input's channel count)."""
        img = np.asarray(img)
        if img.ndim == 2:
            img = img[:, :, None]
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w, ch = img.shape
        en = np.empty((h, w), np.float32)
        out = np.empty((h, w, ch), np.uint8) if want_image else None
        _check(lib().dctc_preview_energy(self._h, _ptr(img), w, h, ch, w * ch, _ptr(en), _ptr(out) if want_image else None),
               "dctc_preview_energy")
        return en, out

    def carver_set_dump_vmaps(self, on=True):
        _check(lib().dctc_carver_set_dump_vmaps(self._h, int(on)), "dctc_carver_set_dump_vmaps")

    def carver_vmap(self, w0, h):
        """(visibility map of the original w0 x h pixels, depth): removal order per pixel, 0 = never removed."""
        out = np.empty((h, w0), np.int32)
        depth = C.c_int(0)
        _check(lib().dctc_carver_vmap(self._h, _ptr(out), C.byref(depth)), "dctc_carver_vmap")
        return out, depth.value

    def carver_paint_seams(self, img):
        """display_carver_seams (src/render.c:204-240) on a copy of the original image."""
        img = np.array(img, dtype=np.uint8, order="C")
        if img.ndim == 2:
            img = img[:, :, None]
        h, w, ch = img.shape
        _check(lib().dctc_carver_paint_seams(self._h, _ptr(img), ch, w * ch), "dctc_carver_paint_seams")
        return img

    def carver_energy_image(self):
        """8-bit grey energy image of the session's current map (liblqr get_energy_image semantics)."""
        w, h = self.carver_size()
        out = np.empty((h, w), np.uint8)
        _check(lib().dctc_carver_energy_image(self._h, _ptr(out)), "dctc_carver_energy_image")
        return out

    def energy_minmax_dev(self, d_en, en_pitch, w, h):
        lo_hi = np.empty(2, np.float32)
        _check(lib().dctc_energy_minmax_dev(self._h, C.c_void_p(d_en), en_pitch, w, h, _ptr(lo_hi)), "dctc_energy_minmax_dev")
        return lo_hi

    def energy_image_dev(self, d_en, en_pitch, w, h, d_out, out_pitch, lo_hi=None, sync=True):
        p = None
        if lo_hi is not None:
            lo_hi = np.ascontiguousarray(lo_hi, dtype=np.float32)
            p = _ptr(lo_hi)
        _check(lib().dctc_energy_image_dev(self._h, C.c_void_p(d_en), en_pitch, w, h, p, C.c_void_p(d_out), out_pitch,
                                           int(sync)), "dctc_energy_image_dev")

    def carve_and_update(self, seam_x, want_band=True):
        """Removes one vertical seam; returns (band_values, xmin, xmax) like liblqr's update_emap would visit."""
        w, h = self.carver_size()
        seam = np.ascontiguousarray(seam_x, dtype=np.int32)
        if seam.shape != (h,):
            raise ValueError("seam must have one column per row")
        xmin = np.empty(h, np.int32)
        xmax = np.empty(h, np.int32)
        band = np.empty(h * 4 * (self.blocksize // 2), np.float32) if want_band else None
        _check(lib().dctc_carve_and_update(self._h, _ptr(seam), _ptr(band) if want_band else None, _ptr(xmin),
                                           _ptr(xmax)), "dctc_carve_and_update")
        if want_band:
            n = int(np.maximum(xmax - xmin + 1, 0).sum())
            band = band[:n]
        return band, xmin, xmax


    def pcie_probe(self, nbytes=256 << 20, iters=8):
        """Pinned host <-> device copy bandwidth (GB/s): H2D alone, D2H alone, per direction with both running."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _check(lib().dctc_pcie_probe(self._h, nbytes, iters, C.byref(a), C.byref(b), C.byref(c)), "dctc_pcie_probe")
        return {"h2d_gbs": a.value, "d2h_gbs": b.value, "bidir_gbs_per_dir": c.value}


def band_plan(h, world, rank, blocksize):
    """(first row, rows, halo rows needed above, below) of `rank`'s band: dctc_band_plan."""
    y0, rows, tn, bn = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _check(lib().dctc_band_plan(h, world, rank, blocksize, C.byref(y0), C.byref(rows), C.byref(tn), C.byref(bn)), "dctc_band_plan")
    return y0.value, rows.value, tn.value, bn.value


def rendezvous_allgather(name, rank, world, blob, timeout_ms=60000):
    """All-gather of equal-sized byte strings over the ranks of one box through POSIX shared memory (no framework)."""
    blob = bytes(blob)
    mine = (C.c_ubyte * max(len(blob), 1)).from_buffer_copy(blob or b"\0")
    out = (C.c_ubyte * max(len(blob) * world, 1))()
    _check(lib().dctc_rendezvous_allgather(name.encode(), rank, world, mine, len(blob), out, timeout_ms), "dctc_rendezvous_allgather")
    raw = bytes(out)
    return [raw[i * len(blob):(i + 1) * len(blob)] for i in range(world)]


class BandRunner:
    """One rank's row band of a w x h image (one process per GPU): dctc_band_runner_*.  The halo rows are read from the
    neighbour ranks' HBM by the energy kernel (CUDA IPC peer mappings exchanged through the C rendezvous)."""

    def __init__(self, ctx, rendezvous, rank, world, w, h, ch):
        self.ctx, self.rank, self.world, self.w, self.h, self.ch = ctx, rank, world, w, h, ch
        self._r = C.c_void_p()
        _check(lib().dctc_band_runner_create(ctx.handle, rendezvous.encode(), rank, world, w, h, ch, C.byref(self._r)),
               "dctc_band_runner_create")
        y0, rows, pitch = C.c_int(), C.c_int(), C.c_size_t()
        _check(lib().dctc_band_runner_geometry(self._r, C.byref(y0), C.byref(rows), C.byref(pitch)), "dctc_band_runner_geometry")
        self.y0, self.band_rows, self.pitch = y0.value, rows.value, pitch.value

    def synth(self, seed, pattern=0):
        _check(lib().dctc_band_runner_synth(self._r, seed, pattern), "dctc_band_runner_synth")

    def upload(self, band_rows):
        a = np.ascontiguousarray(band_rows, dtype=np.uint8)
        _check(lib().dctc_band_runner_upload(self._r, _ptr(a), self.w * self.ch), "dctc_band_runner_upload")

    def connect(self):
        _check(lib().dctc_band_runner_connect(self._r), "dctc_band_runner_connect")

    def barrier(self):
        _check(lib().dctc_band_runner_barrier(self._r), "dctc_band_runner_barrier")

    def step(self, sync=False):
        _check(lib().dctc_band_runner_energy(self._r, int(sync)), "dctc_band_runner_energy")

    def fetch(self):
        out = np.empty((self.band_rows, self.w), np.float32)
        _check(lib().dctc_band_runner_download(self._r, _ptr(out)), "dctc_band_runner_download")
        return out

    def energy_image(self):
        out = np.empty((self.band_rows, self.w), np.uint8)
        _check(lib().dctc_band_runner_energy_image(self._r, _ptr(out)), "dctc_band_runner_energy_image")
        return out

    def close(self):
        if self._r:
            lib().dctc_band_runner_destroy(self._r)
            self._r = C.c_void_p()


class Multi:
    """All GPUs of the box from one process: dctc_multi_* (frames round-robin, row bands with peer halo reads)."""

    def __init__(self, devices=None, blocksize=8, edges=0.5, textures=0.5, kernel=KERNEL_AUTO):
        self._m = C.c_void_p()
        if devices is None:
            _check(lib().dctc_multi_create(C.byref(self._m), None, 0), "dctc_multi_create")
        else:
            arr = (C.c_int * len(devices))(*devices)
            _check(lib().dctc_multi_create(C.byref(self._m), arr, len(devices)), "dctc_multi_create")
        self.n = lib().dctc_multi_device_count(self._m)
        self.set_params(blocksize, edges, textures)
        if kernel != KERNEL_AUTO:
            _check(lib().dctc_multi_set_kernel(self._m, kernel), "dctc_multi_set_kernel")
        self._bands = C.c_void_p()

    def set_params(self, blocksize, edges, textures):
        p = EnergyParameters(edges=edges, textures=textures, blocksize=blocksize)
        _check(lib().dctc_multi_set_params(self._m, C.byref(p)), "dctc_multi_set_params")

    @property
    def launches(self):
        return int(lib().dctc_multi_launch_count(self._m))

    def energy_batch(self, imgs, out=None):
        n, h, w, ch = imgs.shape
        if out is None:
            out = np.empty((n, h, w), np.float32)
        _check(lib().dctc_multi_energy_batch(self._m, _ptr(imgs), n, h * w * ch, w, h, ch, w * ch, _ptr(out), h * w),
               "dctc_multi_energy_batch")
        return out

    def energy_bands(self, img, want_image=False):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        if img.ndim == 2:
            img = img[:, :, None]
        h, w, ch = img.shape
        out = np.empty((h, w), np.float32)
        im8 = np.empty((h, w), np.uint8) if want_image else None
        _check(lib().dctc_multi_energy_bands(self._m, _ptr(img), w, h, ch, w * ch, _ptr(out), _ptr(im8) if want_image else None),
               "dctc_multi_energy_bands")
        return (out, im8) if want_image else out

    # device-resident bands (bench)
    def bands_create(self, w, h, ch):
        self.bands_destroy()
        _check(lib().dctc_multi_bands_create(self._m, w, h, ch, C.byref(self._bands)), "dctc_multi_bands_create")
        self._bw, self._bh = w, h

    def bands_synth(self, seed, pattern=0):
        _check(lib().dctc_multi_bands_synth(self._bands, seed, pattern), "dctc_multi_bands_synth")

    def bands_energy(self, sync=False):
        _check(lib().dctc_multi_bands_energy(self._bands, int(sync)), "dctc_multi_bands_energy")

    def bands_download(self):
        out = np.empty((self._bh, self._bw), np.float32)
        _check(lib().dctc_multi_bands_download(self._bands, _ptr(out)), "dctc_multi_bands_download")
        return out

    def bands_destroy(self):
        if self._bands:
            lib().dctc_multi_bands_destroy(self._bands)
            self._bands = C.c_void_p()

    def close(self):
        self.bands_destroy()
        if self._m:
            lib().dctc_multi_destroy(self._m)
            self._m = C.c_void_p()
