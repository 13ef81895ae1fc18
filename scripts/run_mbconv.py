"""Runs one fused inverted-residual block through the C-ABI (timing / ncu / YR_MBCONV_DEBUG timeline).
usage: run_mbconv.py B H W Cin Ce Cout stride [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yoloret_b200 import _lib  # noqa: E402
from yoloret_b200._lib import YrOp  # noqa: E402

B, H, W, Cin, Ce, Cout, stride = (int(v) for v in sys.argv[1:8])
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
lib = _lib.lib()
x = torch.randn(B, H, W, Cin, device="cuda")
w1, b1 = torch.randn(Cin, Ce, device="cuda") * Cin ** -0.5, torch.randn(Ce, device="cuda")
wd, b2 = torch.randn(9, Ce, device="cuda") * 0.3, torch.randn(Ce, device="cuda")
w2, b3 = torch.randn(Ce, Cout, device="cuda") * Ce ** -0.5, torch.randn(Cout, device="cuda")
Ho, Wo = -(-H // stride), -(-W // stride)
pt = max((Ho - 1) * stride + 3 - H, 0) // 2
pl = max((Wo - 1) * stride + 3 - W, 0) // 2
blob = torch.zeros(int(lib.yr_mbconv_packed_floats(Cin, Ce, Cout)), device="cuda")
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.yr_mbconv_pack(w1.data_ptr(), Ce, b1.data_ptr(), wd.data_ptr(), Ce, b2.data_ptr(), w2.data_ptr(), Cout,
                              b3.data_ptr(), Cin, Ce, Cout, blob.data_ptr(), st), "pack")
out = torch.empty(B, Ho, Wo, Cout, device="cuda")
op = YrOp()
op.kind = _lib.OP_MBCONV
op.B, op.H, op.W, op.C, op.K2, op.Ho, op.Wo, op.N = B, H, W, Cin, Ce, Ho, Wo, Cout
op.k, op.stride, op.pad_t, op.pad_l, op.ld_in, op.ld_out = 3, stride, pt, pl, Cin, Cout
op.in_, op.out, op.w_tc = x.data_ptr(), out.data_ptr(), blob.data_ptr()
if stride == 1 and Cin == Cout:
    op.res, op.ld_res = x.data_ptr(), Cin
ops = (YrOp * 1)(op)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.yr_run_ops(ops, 1, st), "yr_run_ops")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
byts = (B * H * W * Cin + B * Ho * Wo * Cout * (2 if op.res else 1)) * 4
print("mbconv %dx%dx%d Cin%d Ce%d Cout%d s%d: best %.4f ms (%.1f GB/s of fused-minimum traffic)" % (
    B, H, W, Cin, Ce, Cout, stride, min(ts), byts / min(ts) / 1e6))
