"""PFM in / out through libmpt.so (include/mpt.h "PFM in / out"; format of src/merian/io/image_io.cpp:53-168). Host-only: these
calls work without a GPU. Images are float32 arrays with top-down rows."""
import ctypes as C
import os

import numpy as np

from . import abi
from ._lib import MptError, lib


def _check(st):
    if st != abi.MPT_OK:
        raise MptError(st, lib().mpt_last_error(None).decode())


def pfm_info(path):
    w, h, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    _check(lib().mpt_pfm_info(os.fsencode(path), C.byref(w), C.byref(h), C.byref(c)))
    return w.value, h.value, c.value


def load_pfm(path, desired_channels=0):
    w, h, c = pfm_info(path)
    ch = desired_channels or c
    out = np.empty((h, w, ch), np.float32)
    _check(lib().mpt_pfm_load(os.fsencode(path), desired_channels, C.c_void_p(out.ctypes.data), out.size))
    return out


def save_pfm(path, image):
    img = np.ascontiguousarray(image, dtype=np.float32)
    if img.ndim == 2:
        img = img[..., None]
    h, w, c = img.shape
    _check(lib().mpt_pfm_save(os.fsencode(path), C.c_void_p(img.ctypes.data), w, h, c))


def load_hdr(path, desired_channels=3):
    """Radiance RGBE, decoded as stb_image does for image_load_f32."""
    w, h = C.c_uint32(), C.c_uint32()
    _check(lib().mpt_hdr_load(os.fsencode(path), desired_channels, C.byref(w), C.byref(h), None, 0))
    out = np.empty((h.value, w.value, desired_channels), np.float32)
    _check(lib().mpt_hdr_load(os.fsencode(path), desired_channels, C.byref(w), C.byref(h), C.c_void_p(out.ctypes.data), out.size))
    return out
