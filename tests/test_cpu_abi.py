"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU,
exports every symbol include/yoloret_b200.h declares, and validates arguments before touching CUDA."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "yoloret_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(yr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    from yoloret_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 11
    for n in names:
        assert hasattr(built_lib, n), "library lacks %s" % n
        assert n in _lib.SYMBOLS, "ctypes binding lacks %s" % n
    assert set(_lib.SYMBOLS) == set(names)


def test_header_compiles_as_plain_c(tmp_path):
    """The boundary is a C ABI: the header must be consumable by a C compiler (no C++/torch types)."""
    c = tmp_path / "t.c"
    c.write_text('#include "yoloret_b200.h"\nint main(void){ yr_op op; (void)op; return (int)sizeof(yr_decode_params) == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c),
                           "-o", str(tmp_path / "t.o")])


def test_struct_layouts_match(built_lib, tmp_path):
    from yoloret_b200._lib import YrOp, YrDecodeParams, YrLossParams
    assert built_lib.yr_sizeof_op() == C.sizeof(YrOp)
    c = tmp_path / "s.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "yoloret_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                 'sizeof(yr_op),sizeof(yr_decode_params),sizeof(yr_loss_params),offsetof(yr_op,in),offsetof(yr_decode_params,cand_cap));return 0;}\n')
    exe = tmp_path / "s"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    a, b, c_, d, e = (int(v) for v in subprocess.check_output([str(exe)]).split())
    assert (a, b, c_) == (C.sizeof(YrOp), C.sizeof(YrDecodeParams), C.sizeof(YrLossParams))
    assert d == YrOp.in_.offset and e == YrDecodeParams.cand_cap.offset


def test_library_is_sm100a_with_lineinfo(built_lib):
    from yoloret_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:400]
    assert built_lib.yr_version() >= 100


def test_argument_validation_needs_no_gpu(built_lib):
    from yoloret_b200 import _lib
    from yoloret_b200._lib import YrOp, YrDecodeParams, YrLossParams
    op = YrOp()
    op.kind = 99
    assert built_lib.yr_run_ops((YrOp * 1)(op), 1, None) == -1
    assert b"unknown kind" in built_lib.yr_last_error()
    for kind, word in ((_lib.OP_PW, b"pw"), (_lib.OP_DW, b"dw"), (_lib.OP_STEM, b"stem"), (_lib.OP_SE, b"se"),
                       (_lib.OP_RFCR, b"rfcr"), (_lib.OP_RESAMPLE, b"resample")):
        op = YrOp()
        op.kind = kind
        assert built_lib.yr_run_ops((YrOp * 1)(op), 1, None) == -1  # null pointers
        assert word in built_lib.yr_last_error()
    assert built_lib.yr_run_ops(None, 0, None) == 0
    p = YrDecodeParams()
    assert built_lib.yr_decode_filter(None, None, C.byref(p), None, None, None, None, None) == -1
    assert built_lib.yr_nms_classwise(None, 0, None, None, None, 1, 1, 1, 1, 0.5, None, None, None, None) == -1
    assert built_lib.yr_pack_detections(None, None, 1, 1, 1, None, None, None, None, None, None) == -1
    assert built_lib.yr_letterbox_u8(None, 1, 1, None, 1, 1, 1, 1, 0, 0, None) == -1
    lp = YrLossParams()
    lp.B, lp.gh, lp.gw, lp.A, lp.C = 2, 13, 13, 3, 80
    assert built_lib.yr_yolo_loss_workspace(C.byref(lp)) > 0
    assert built_lib.yr_yolo_loss(None, None, None, None, C.byref(lp), None, None, None, 0, None) == -1
    with pytest.raises(_lib.YrError):
        _lib.check(-1, "demo")


def test_product_fails_loudly_without_cuda(built_lib):
    """No CPU fallback: constructing the engine / post-process without a GPU raises."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from yoloret_b200 import _lib
    from yoloret_b200.engine import Engine
    from yoloret_b200.postprocess import PostProcess
    with pytest.raises(_lib.YrError):
        Engine("mobilenetv2x75", 20, (96, 96), 1, {}, np.zeros((9, 2), np.float32))
    with pytest.raises(_lib.YrError):
        PostProcess(1, [(3, 3), (6, 6), (12, 12)], 20, np.zeros((9, 2), np.float32))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure; nothing under yoloret_b200/ may import it."""
    pkg = os.path.join(ROOT, "yoloret_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)
                assert "/root/reference" not in src, os.path.join(dp, f)
