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
    sizer, packer = ((lib.yr_pw_ts_packed_floats, lib.yr_pw_ts_pack) if variant == _lib.PW_TS
                     else (lib.yr_pw_tc_packed_floats, lib.yr_pw_tc_pack))
    n = int(sizer(K, N))
    assert n > 0, "no tensor-core tiling for K=%d N=%d" % (K, N)
    packed = torch.full((n,), float("nan"), device="cuda")
    _lib.check(packer(w.contiguous().data_ptr(), K, N, packed.data_ptr(),
                      torch.cuda.current_stream().cuda_stream), "yr_pw_pack")
    torch.cuda.synchronize()
    assert not torch.isnan(packed).any()
    return packed


# ---- end-to-end detection comparison with a tie margin -------------------------------------------
def _iou(a, b):
    iy0, ix0 = max(a[0], b[0]), max(a[1], b[1])
    iy1, ix1 = min(a[2], b[2]), min(a[3], b[3])
    inter = max(iy1 - iy0, 0.0) * max(ix1 - ix0, 0.0)
    ua = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return inter / ua if ua > 0 else 0.0


def assert_detections_match(got, ref, score_thr, iou_thr, tol=1e-3, box_tol_px=1.0, margin=5e-3):
    """got / ref: (boxes [n,4], scores [n], classes [n]).  Two fp32 implementations of the network differ
    by rounding, and yolo_eval is discontinuous (score > thr, IoU > thr), so a detection may legally
    appear on one side only when it sits within ``margin`` of a decision boundary.  Everything else must
    pair up one-to-one with the same class, |score diff| <= tol and |box diff| <= box_tol_px.
    Returns (matched, marginal)."""
    gb, gs, gc = (np.asarray(v) for v in got)
    rb, rs, rc = (np.asarray(v) for v in ref)
    used = np.zeros(len(rs), bool)
    unmatched = []
    matched = 0
    for i in range(len(gs)):
        cand = [j for j in range(len(rs)) if not used[j] and rc[j] == gc[i] and abs(float(rs[j]) - float(gs[i])) <= tol
                and np.abs(rb[j].astype(np.float64) - gb[i].astype(np.float64)).max() <= box_tol_px]
        if cand:
            used[cand[0]] = True
            matched += 1
        else:
            unmatched.append(("got", gb[i], float(gs[i]), int(gc[i])))
    unmatched += [("ref", rb[j], float(rs[j]), int(rc[j])) for j in range(len(rs)) if not used[j]]
    for side, box, score, cls in unmatched:
        near_thr = abs(score - score_thr) <= margin
        ob, os_, oc = (gb, gs, gc) if side == "ref" else (rb, rs, rc)  # the side that dropped it
        near_iou = any(oc[k] == cls and os_[k] >= score - tol and abs(_iou(box.astype(np.float64), ob[k].astype(np.float64))
                                                                     - iou_thr) <= 10 * margin for k in range(len(os_)))
        capped = (np.sum(gc == cls) >= 20) or (np.sum(rc == cls) >= 20)  # max_boxes cut shifts the tail
        assert near_thr or near_iou or capped, "unexplained %s-only detection: class %d score %.5f box %s" % (
            side, cls, score, box)
    return matched, len(unmatched)
