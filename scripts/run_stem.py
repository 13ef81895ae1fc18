"""Times the stem op (3x3 stride-2 conv, Cin=3) through the C-ABI.  usage: run_stem.py B H W N [u8]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yoloret_b200 import _lib  # noqa: E402
from yoloret_b200._lib import YrOp  # noqa: E402

B, H, W, N = (int(v) for v in sys.argv[1:5])
u8 = len(sys.argv) > 5 and sys.argv[5] == "u8"
x = (torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8) if u8 else torch.rand(B, H, W, 3)).cuda()
w = torch.randn(27, N, device="cuda") * 0.2
b = torch.randn(N, device="cuda")
out = torch.empty(B, H // 2, W // 2, N, device="cuda")
op = YrOp()
op.kind, op.act, op.in_is_u8 = _lib.OP_STEM, 1, int(u8)
op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, 3, H // 2, W // 2, N
op.k, op.stride, op.pad_t, op.pad_l, op.ld_in, op.ld_out = 3, 2, 0, 0, 3, N
op.in_, op.out, op.w, op.bias = x.data_ptr(), out.data_ptr(), w.data_ptr(), b.data_ptr()
ops = (YrOp * 1)(op)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(_lib.lib().yr_run_ops(ops, 1, st), "yr_run_ops")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
byts = x.numel() * x.element_size() + out.numel() * 4
print("stem %dx%dx%d N%d %s CPT=%s: best %.4f ms  %.0f GB/s  checksum %.6f" % (
    B, H, W, N, "u8" if u8 else "f32", os.environ.get("YR_STEM_CPT", "4"), min(ts), byts / min(ts) / 1e6, float(out.double().sum())))
