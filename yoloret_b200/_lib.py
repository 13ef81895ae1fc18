"""ctypes binding of the C-ABI in include/yoloret_b200.h.

The product path has no CPU fallback: if the CUDA library is missing or a call
fails, this module raises.  (``yoloret_b200.build.build_library`` /
``__graft_entry__.build()`` produce the library in-tree.)
"""
from __future__ import annotations

import ctypes as C
import functools
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libyoloret_b200.so")

# enums (include/yoloret_b200.h)
ACT_NONE, ACT_RELU6, ACT_SWISH = 0, 1, 2
OP_STEM, OP_PW, OP_DW, OP_RESAMPLE, OP_RFCR, OP_SE, OP_SE_FC, OP_DWPW = 0, 1, 2, 3, 4, 5, 6, 8
UP2, POOL2, POOL4 = 0, 1, 2
PW_AUTO, PW_SIMT, PW_TC, PW_TS, PW_TS2 = 0, 1, 2, 3, 4


class YrOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("act", C.c_int32), ("mode", C.c_int32), ("in_is_u8", C.c_int32),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
        ("Ho", C.c_int32), ("Wo", C.c_int32), ("N", C.c_int32),
        ("k", C.c_int32), ("stride", C.c_int32), ("pad_t", C.c_int32), ("pad_l", C.c_int32),
        ("ld_in", C.c_int32), ("ld_in2", C.c_int32), ("ld_in3", C.c_int32), ("ld_in4", C.c_int32),
        ("ld_out", C.c_int32), ("ld_res", C.c_int32),
        ("K2", C.c_int32), ("K3", C.c_int32), ("K4", C.c_int32),
        ("variant", C.c_int32),
        ("in_", C.c_void_p), ("in2", C.c_void_p), ("in3", C.c_void_p), ("in4", C.c_void_p),
        ("out", C.c_void_p),
        ("w", C.c_void_p), ("bias", C.c_void_p), ("res", C.c_void_p), ("scale", C.c_void_p),
        ("w_tc", C.c_void_p), ("aux", C.c_void_p),
    ]


class YrDecodeParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("num_classes", C.c_int32), ("num_scales", C.c_int32),
        ("grid_h", C.c_int32 * 3), ("grid_w", C.c_int32 * 3), ("ld", C.c_int32 * 3),
        ("anchors", ((C.c_float * 2) * 3) * 3),
        ("input_h", C.c_int32), ("input_w", C.c_int32),
        ("score_threshold", C.c_float), ("cand_cap", C.c_int32),
    ]


class YrLossParams(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("gh", C.c_int32), ("gw", C.c_int32), ("A", C.c_int32), ("C", C.c_int32),
        ("ld_logits", C.c_int32), ("ld_true", C.c_int32),
        ("anchors", (C.c_float * 2) * 3),
        ("input_h", C.c_int32), ("input_w", C.c_int32),
        ("ignore_thresh", C.c_float), ("max_true", C.c_int32),
    ]


class YrLoss3Params(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("num_scales", C.c_int32),
        ("gh", C.c_int32 * 3), ("gw", C.c_int32 * 3), ("ld_logits", C.c_int32 * 3),
        ("anchors", ((C.c_float * 2) * 3) * 3),
        ("input_h", C.c_int32), ("input_w", C.c_int32),
        ("ignore_thresh", C.c_float), ("max_records", C.c_int32),
    ]


# every symbol the header declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "yr_version": (C.c_int, []),
    "yr_last_error": (C.c_char_p, []),
    "yr_sizeof_op": (C.c_int, []),
    "yr_dw_se_slots": (C.c_int, [C.POINTER(YrOp)]),
    "yr_run_ops": (C.c_int, [C.POINTER(YrOp), C.c_int, _P]),
    "yr_pw_tc_packed_floats": (C.c_int64, [C.c_int, C.c_int]),
    "yr_pw_tc_pack": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "yr_pw_ts_packed_floats": (C.c_int64, [C.c_int, C.c_int]),
    "yr_pw_ts_pack": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "yr_pw_ts2_supported": (C.c_int, [C.c_int, C.c_int]),
    "yr_dwpw_packed_floats": (C.c_int64, [C.c_int, C.c_int]),
    "yr_dwpw_pack": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "yr_dwpw_supported": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "yr_dwpw_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "yr_decode_filter": (C.c_int, [C.POINTER(_P), _P, C.POINTER(YrDecodeParams), _P, _P, _P, _P, _P]),
    "yr_yolo_head": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int,
                               _P, _P, _P, _P, _P, _P]),
    "yr_nms_classwise": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                   _P, _P, _P, _P]),
    "yr_pack_detections": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "yr_letterbox_u8": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, _P]),
    "yr_yolo_loss_workspace": (C.c_int64, [C.POINTER(YrLossParams)]),
    "yr_yolo_loss_gather_true": (C.c_int, [_P, C.POINTER(YrLossParams), _P, _P, _P]),
    "yr_yolo_loss": (C.c_int, [_P, _P, _P, _P, C.POINTER(YrLossParams), _P, _P, _P, C.c_int64, _P]),
    "yr_encode_true_boxes_sparse": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.POINTER(_P), _P, _P, C.c_int, _P]),
    "yr_yolo_loss3_workspace": (C.c_int64, [C.POINTER(YrLoss3Params)]),
    "yr_yolo_loss3": (C.c_int, [C.POINTER(_P), C.POINTER(_P), _P, _P, C.POINTER(YrLoss3Params), _P, C.POINTER(_P), _P,
                                C.c_int64, _P]),
    "yr_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, _P]),
    "yr_encode_true_boxes": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(_P), _P]),
}


class YrError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libyoloret_b200.so (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise YrError("CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.yr_sizeof_op() != C.sizeof(YrOp):
            raise YrError("yr_op layout mismatch: C %d vs ctypes %d" % (l.yr_sizeof_op(), C.sizeof(YrOp)))
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise YrError("%s failed (status %d): %s" % (what or "yoloret_b200 call", rc,
                                                     lib().yr_last_error().decode(errors="replace")))


def on_device(fn):
    """Method decorator: runs ``fn`` with ``self.device`` as the current CUDA device.  Function attributes, the SM
    count and streams are per device, so an Engine / PostProcess built for ``cuda:1`` must make that device current
    around its library calls even when the caller's current device is ``cuda:0``."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        import torch
        dev = getattr(self, "device", None)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(self, *args, **kwargs)
        with torch.cuda.device(dev):
            return fn(self, *args, **kwargs)
    return wrapper
