"""ORACLE — TEST INFRASTRUCTURE ONLY. Loads the real reference pieces built by oracle/Makefile
into oracle/_ref/ (from /root/reference sources, unmodified). Nothing here reads /root/reference
at run time, so it also works on the GPU box where only the prebuilt files exist."""
import ctypes
import importlib.util
import os
import sysconfig

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def _load_ext(name):
    path = os.path.join(REF_DIR, name + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        return None
    if not hasattr(np, "float"):
        np.float = float  # bbox.pyx:12 uses the alias numpy removed; shim, do not edit the source
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_bbox():
    """The reference's lib/fpn/box_intersections_cpu/bbox.pyx, compiled as is (or None)."""
    return _load_ext("bbox")


def ref_draw_rectangles():
    """The reference's lib/draw_rectangles/draw_rectangles.pyx, compiled as is (or None)."""
    return _load_ext("draw_rectangles")


def ref_kernels():
    """ctypes handle on the reference's three .cu files compiled unmodified for sm_100a
    (needs a GPU to call), or None when not built."""
    path = os.path.join(REF_DIR, "libref_kernels.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.ROIAlignForwardLaucher.restype = I
    lib.ROIAlignForwardLaucher.argtypes = [P, P, I, I, I, I, I, I, I, F, P, P]
    lib.ROIAlignBackwardLaucher.restype = I
    lib.ROIAlignBackwardLaucher.argtypes = [P, P, I, I, I, I, I, I, I, P, P]
    lib.ApplyNMSGPU.restype = I
    lib.ApplyNMSGPU.argtypes = [P, P, I, F, I]
    lib.highway_lstm_forward_ongpu.restype = None
    lib.highway_lstm_forward_ongpu.argtypes = [I] * 5 + [P] * 10 + [I, P, P]
    lib.highway_lstm_backward_ongpu.restype = None
    lib.highway_lstm_backward_ongpu.argtypes = [I] * 5 + [P] * 16 + [I, I, P, P]
    return lib
