"""Runs one pointwise layer through the C-ABI a few times (for ncu captures and quick timing).
usage: run_pw_layer.py B H W K N [variant] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from yoloret_b200 import _lib  # noqa: E402
from yoloret_b200._lib import YrOp  # noqa: E402
from ophelp import pack_tc  # noqa: E402

B, H, W, K, N = (int(v) for v in sys.argv[1:6])
variant = int(sys.argv[6]) if len(sys.argv) > 6 else 2
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
a = torch.randn(B, H, W, K, device="cuda")
w = torch.randn(K, N, device="cuda") * K ** -0.5
bias = torch.randn(N, device="cuda")
out = torch.empty(B, H, W, N, device="cuda")
op = YrOp()
op.kind, op.act, op.variant = _lib.OP_PW, 1, variant
op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, K, H, W, N
op.ld_in, op.ld_out = K, N
op.in_, op.out, op.w, op.bias = a.data_ptr(), out.data_ptr(), w.data_ptr(), bias.data_ptr()
if variant != 1:
    packed = pack_tc(w, variant)  # variant 4 (CTA pair) uses the variant-3 image
    op.w_tc = packed.data_ptr()
ops = (YrOp * 1)(op)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(_lib.lib().yr_run_ops(ops, 1, st), "yr_run_ops")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
M = B * H * W
byts = M * (K + N) * 4
ref = torch.clamp(a.view(M, K) @ w + bias, 0, 6)
err = float((out.view(M, N) - ref).abs().max())
print("pw %dx%dx%d K%d N%d variant %d: best %.4f ms  %.1f GB/s  %.2f TF  (max err vs torch tf32/fp32 matmul %.2e)" % (
    B, H, W, K, N, variant, min(ts), byts / min(ts) / 1e6, 2.0 * M * K * N / min(ts) / 1e9, err))
