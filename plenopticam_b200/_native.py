"""ctypes binding of libpcb200.so (C-ABI declared in include/plenopticam_b200.h).

There is no CPU fallback: if the library cannot be loaded, every operator raises.  `build()` compiles the
library in-tree with nvcc for sm_100a (the built .so is git-ignored but travels with the snapshot).
"""
import ctypes as C
import fcntl
import hashlib
import os
import shutil
import subprocess
import threading

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libpcb200.so")
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "plenopticam_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

PCB_U16, PCB_F32, PCB_U8, PCB_F64 = 0, 1, 2, 3
PCB_REC, PCB_HEX = 0, 1
PCB_LENS_INVALID = -2 ** 31

# numpy mirrors of the C structs
PCB_MAX_PEERS = 16


class Fanout(C.Structure):
    """pcb_fanout_t"""
    _fields_ = [("out", C.c_void_p * PCB_MAX_PEERS), ("mm", C.c_void_p * PCB_MAX_PEERS), ("n", C.c_int32),
                ("reserved", C.c_int32)]


class RefiRows(C.Structure):
    """pcb_refi_rows_t"""
    _fields_ = [("out_lo", C.c_int32), ("out_hi", C.c_int32), ("h_lo", C.c_int32), ("h_hi", C.c_int32),
                ("tmp_lo", C.c_int32), ("tmp_hi", C.c_int32)]


class RefiQuads(C.Structure):
    """pcb_refi_quads_t"""
    _fields_ = [("wq", C.c_void_p), ("fu", C.c_void_p), ("G", C.c_int32), ("reserved", C.c_int32),
                ("woff", C.c_int32 * 32), ("W", C.c_int32 * 32), ("lo", C.c_int32 * 32), ("hi", C.c_int32 * 32)]


LENS_DTYPE = np.dtype([("oy", "<i4"), ("ox", "<i4"), ("wy", "<f4"), ("wx", "<f4")])
HEXCOL_DTYPE = np.dtype([("x0", "<i4"), ("x1", "<i4"), ("w", "<f4"), ("reserved", "<i4")])

_lock = threading.Lock()
_lib = None


class NativeLibraryError(RuntimeError):
    pass


def _sources():
    return sorted(os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith((".cu", ".cuh")))


#: ABI the binding below was written against; `pcb_version()` of the loaded library must return it.  Bump together
#: with PCB_ABI_VERSION in include/plenopticam_b200.h whenever an entry point is added or its arguments change.
ABI_VERSION = 202

STAMP_PATH = LIB_PATH + ".srchash"


def source_hash():
    """sha256 over the CUDA sources, the header and the compiler flags: what the .so must have been built from.
    (File times do not survive the snapshot to the GPU box, a content hash does.)"""
    h = hashlib.sha256()
    for path in _sources() + [HEADER]:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _stale():
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return True
    with open(STAMP_PATH) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile csrc/pcb200.cu -> libpcb200.so for sm_100a (nvcc cross-compiles without a GPU).  Serialised over
    processes with a file lock: of N ranks starting together one compiles, the others wait and find a fresh library."""
    if not force and not _stale():
        return LIB_PATH
    with open(LIB_PATH + ".lock", "w") as lockf:
        fcntl.flock(lockf, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return LIB_PATH
            nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
            if not os.path.exists(nvcc):
                raise NativeLibraryError("nvcc not found; cannot build %s" % LIB_PATH)
            want = source_hash()
            tmp = LIB_PATH + ".tmp.%d" % os.getpid()
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp, os.path.join(CSRC_DIR, "pcb200.cu")]
            if verbose:
                print(" ".join(cmd))
            res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise NativeLibraryError("nvcc failed:\n" + res.stdout)
            os.replace(tmp, LIB_PATH)
            with open(STAMP_PATH, "w") as f:
                f.write(want + "\n")
        finally:
            fcntl.flock(lockf, fcntl.LOCK_UN)
    return LIB_PATH


_vp = C.c_void_p
_i = C.c_int
_i64 = C.c_int64

_SIGNATURES = {
    "pcb_version": (C.c_int, []),
    "pcb_last_error": (C.c_char_p, []),
    "pcb_key_to_float": (C.c_float, [C.c_uint32]),
    "pcb_debug_kernel_events": (_i, [_vp, _vp]),
    "pcb_minmax_reset": (_i, [_vp, _vp]),
    "pcb_minmax_f32": (_i, [_vp, _i64, _vp, _vp]),
    "pcb_resample_minmax": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "pcb_resample_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "pcb_resample_u16": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "pcb_quantize_u16": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "pcb_quantize_u8": (_i, [_vp, _i64, _vp, _vp, _vp]),
    "pcb_spline_prefilter": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pcb_rotate_sample": (_i, [_vp, _i, _i, _i, C.c_double, _vp, _vp]),
    "pcb_viewpoints": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "pcb_viewpoints_inverse": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "pcb_crop_micro_images": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "pcb_crop_micro_images_yx": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "pcb_refocus_int_scratch_bytes": (_i64, [_i, _i, _i, _i, _i, _vp, _i, _vp]),
    "pcb_refocus_int": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i64, _vp, _vp, _vp]),
    "pcb_refocus_int_srgb": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i64, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcb_refocus_int_aligned": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i64, _vp, _vp, _i, _vp]),
    "pcb_refocus_int_srgb_aligned": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i64, C.c_double, C.c_double, _vp,
                                          _vp, _vp, _vp, _vp, _i, _vp]),
    "pcb_refocus_int_direct": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "pcb_refocus_int_zero": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "pcb_refi_prefilter": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "pcb_refi_prefilter_toeplitz": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp,
                                         _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp]),
    "pcb_refocus_refi": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                              _vp, _vp, _vp]),
    "pcb_refocus_refi_fanout": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _vp, _vp, _vp, _vp]),
    "pcb_refi_upscaled_slice": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pcb_peer_alloc": (_i, [_i64, C.POINTER(C.c_void_p)]),
    "pcb_peer_free": (_i, [_vp]),
    "pcb_peer_export": (_i, [_vp, C.c_char_p]),
    "pcb_peer_open": (_i, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "pcb_peer_close": (_i, [_vp]),
    "pcb_percentile_scratch_bytes": (_i64, []),
    "pcb_percentile_f32": (_i, [_vp, _i64, _i, C.POINTER(C.c_double), _i, _vp, _vp, _vp]),
    "pcb_contrast_bounds": (_i, [_vp, _i64, C.c_double, C.c_double, _vp, _vp, _vp]),
    "pcb_percentile_f64": (_i, [_vp, _i64, _i, C.POINTER(C.c_double), _i, _vp, _vp, _vp]),
    "pcb_contrast_srgb": (_i, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "pcb_contrast_srgb_typed": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp]),
    "pcb_image_minmax": (_i, [_vp, _i, _i64, _vp, _vp]),
    "pcb_warp_projective": (_i, [_vp, _i, _i, _i, _i, C.POINTER(C.c_double), _i, _i, _vp, _vp, _vp]),
    "pcb_hex_align": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp,
                           _vp, _vp, _vp]),
    "pcb_unpack_bayer": (_i, [_vp, _i64, _i, _vp, _vp]),
    "pcb_cfa_safe_awb": (_i, [_vp, _vp, _i, _i, C.POINTER(C.c_double), _i, C.c_double, C.c_double, C.c_double, _vp, _vp]),
    "pcb_scheimpflug": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "pcb_devig_white": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "pcb_devig_normalize": (_i, [_vp, _i64, _vp, _vp]),
    "pcb_devig_noise_level": (_i, [_vp, _i, _i, C.POINTER(C.c_double), _i, _vp, _vp, _vp]),
    "pcb_devig_divide": (_i, [_vp, _i, _i64, _i, _vp, _vp, _i, _vp]),
}


def exported_symbols():
    """Names the header declares and the binding expects."""
    return sorted(_SIGNATURES)


def lib():
    """Load (building first if the sources are newer and nvcc is present) and return the CDLL."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            # never load a library that was not built from the sources in this tree: a failed rebuild raises
            build()
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:
            raise NativeLibraryError("cannot load %s (%s); run `python -c 'import __graft_entry__ as g; g.build()'`"
                                     % (LIB_PATH, e))
        try:
            handle.pcb_version.restype = C.c_int
            got = int(handle.pcb_version())
        except AttributeError:
            raise NativeLibraryError("%s does not export pcb_version" % LIB_PATH)
        if got != ABI_VERSION:
            raise NativeLibraryError("%s has ABI version %d, the binding expects %d (header and binding out of step)"
                                     % (LIB_PATH, got, ABI_VERSION))
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError:
                raise NativeLibraryError("%s does not export %s" % (LIB_PATH, name))
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().pcb_last_error().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError(msg)
        if rc == 3:
            raise NotImplementedError(msg)
        raise RuntimeError(msg)
