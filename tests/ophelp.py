"""Helpers for the GPU parity tests: run single yr_ops through the C-ABI on torch CUDA tensors."""
import ctypes as C

import numpy as np
import torch

from yoloret_b200 import _lib
from yoloret_b200._lib import YrOp

ACT = {"none": 0, "relu6": 1, "swish": 2}


def run_op(op: YrOp):
    ops = (YrOp * 1)(op)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().yr_run_ops(ops, 1, st), "yr_run_ops")
    torch.cuda.synchronize()


def act_ref(x, act):
    if act == "relu6":
        return torch.clamp(x, 0.0, 6.0)
    if act == "swish":
        return x * torch.sigmoid(x)
    return x


def pw_op(a, w, bias, act="none", res=None, scale=None, ld_out=None, variant=0, up2=False):
    """a [B,H,W,ld_in] (uses first K=w.shape[0] channels), w [K,N], returns [B,H,W,N] ([B,2H,2W,N] with up2)."""
    B, H, W, ld = a.shape
    K, N = w.shape
    ldo = ld_out or N
    Ho, Wo = (2 * H, 2 * W) if up2 else (H, W)
    out = torch.full((B, Ho, Wo, ldo), float("nan"), device="cuda")
    op = YrOp()
    op.kind, op.act, op.variant = _lib.OP_PW, ACT[act], variant
    op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, K, Ho, Wo, N
    op.ld_in, op.ld_out = ld, ldo
    op.in_, op.out, op.w, op.bias = a.data_ptr(), out.data_ptr(), w.data_ptr(), bias.data_ptr()
    if res is not None:
        op.res, op.ld_res = res.data_ptr(), res.shape[-1]
    if scale is not None:
        op.scale = scale.data_ptr()
    if variant != _lib.PW_SIMT:
        packed = pack_tc(w, variant)
        op.w_tc = packed.data_ptr()
    run_op(op)
    return out


def pack_tc(w, variant=_lib.PW_TC):
    """W [K,N] (CUDA) -> the tensor-core weight image (yr_pw_tc_pack; yr_pw_ts_pack for variant 3)."""
    K, N = w.shape
    lib = _lib.lib()
    sizer, packer = ((lib.yr_pw_ts_packed_floats, lib.yr_pw_ts_pack) if variant in (_lib.PW_TS, _lib.PW_TS2)
                     else (lib.yr_pw_tc_packed_floats, lib.yr_pw_tc_pack))
    n = int(sizer(K, N))
    assert n > 0, "no tensor-core tiling for K=%d N=%d" % (K, N)
    packed = torch.full((n,), float("nan"), device="cuda")
    _lib.check(packer(w.contiguous().data_ptr(), K, N, packed.data_ptr(),
                      torch.cuda.current_stream().cuda_stream), "yr_pw_pack")
    torch.cuda.synchronize()
    assert not torch.isnan(packed).any()
    return packed


# ---- end-to-end detection comparison with a tie margin (shared with bench.py's verify leg) ----------
from oracle.verify import assert_detections_match  # noqa: E402,F401
