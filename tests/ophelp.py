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


def pw_op(a, w, bias, act="none", res=None, scale=None, ld_out=None, variant=0):
    """a [B,H,W,ld_in] (uses first K=w.shape[0] channels), w [K,N], returns [B,H,W,N]."""
    B, H, W, ld = a.shape
    K, N = w.shape
    ldo = ld_out or N
    out = torch.full((B, H, W, ldo), float("nan"), device="cuda")
    op = YrOp()
    op.kind, op.act, op.variant = _lib.OP_PW, ACT[act], variant
    op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, K, H, W, N
    op.ld_in, op.ld_out = ld, ldo
    op.in_, op.out, op.w, op.bias = a.data_ptr(), out.data_ptr(), w.data_ptr(), bias.data_ptr()
    if res is not None:
        op.res, op.ld_res = res.data_ptr(), res.shape[-1]
    if scale is not None:
        op.scale = scale.data_ptr()
    run_op(op)
    return out
