"""Runs one fused depthwise->pointwise layer pair (YR_OP_DWPW) through the C-ABI, beside the two separate ops.
usage: run_dwpw_layer.py B H W C N stride [reps]     (YR_PW_TC_DEBUG=1 prints the CTA-0 timeline of the fused kernel)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from yoloret_b200 import _lib  # noqa: E402
from yoloret_b200._lib import YrOp  # noqa: E402
from ophelp import pack_tc  # noqa: E402

B, H, W, C, N, S = (int(v) for v in sys.argv[1:7])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
lib = _lib.lib()
Ho, Wo = -(-H // S), -(-W // S)
pt = max((Ho - 1) * S + 3 - H, 0) // 2
pl = max((Wo - 1) * S + 3 - W, 0) // 2
x = torch.randn(B, H, W, C, device="cuda")
wd = torch.randn(9, C, device="cuda") * 0.3
bd = torch.randn(C, device="cuda")
wp = torch.randn(C, N, device="cuda") * C ** -0.5
bp = torch.randn(N, device="cuda")
mid = torch.empty(B, Ho, Wo, C, device="cuda")
out_s = torch.empty(B, Ho, Wo, N, device="cuda")
out_f = torch.empty(B, Ho, Wo, N, device="cuda")
d = YrOp()
d.kind, d.act = _lib.OP_DW, 1
d.B, d.H, d.W, d.C, d.Ho, d.Wo, d.N = B, H, W, C, Ho, Wo, C
d.k, d.stride, d.pad_t, d.pad_l, d.ld_in, d.ld_out = 3, S, pt, pl, C, C
d.in_, d.out, d.w, d.bias = x.data_ptr(), mid.data_ptr(), wd.data_ptr(), bd.data_ptr()
q = YrOp()
q.kind, q.act, q.variant = _lib.OP_PW, 0, 3
q.B, q.H, q.W, q.C, q.Ho, q.Wo, q.N = B, Ho, Wo, C, Ho, Wo, N
q.ld_in, q.ld_out = C, N
packed = pack_tc(wp, 3)
q.in_, q.out, q.w, q.bias, q.w_tc = mid.data_ptr(), out_s.data_ptr(), wp.data_ptr(), bp.data_ptr(), packed.data_ptr()
n = int(lib.yr_dwpw_packed_floats(C, N))
blob = torch.zeros(n, device="cuda")
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.yr_dwpw_pack(wp.data_ptr(), C, N, wd.data_ptr(), bd.data_ptr(), blob.data_ptr(), st), "pack")
f = YrOp()
f.kind, f.act, f.mode = _lib.OP_DWPW, 0, 1
f.B, f.H, f.W, f.C, f.Ho, f.Wo, f.N = B, H, W, C, Ho, Wo, N
f.k, f.stride, f.pad_t, f.pad_l, f.ld_in, f.ld_out = 3, S, pt, pl, C, N
f.in_, f.out, f.w_tc, f.bias = x.data_ptr(), out_f.data_ptr(), blob.data_ptr(), bp.data_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(ops):
    arr = (YrOp * len(ops))(*ops)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.yr_run_ops(arr, len(ops), st), "yr_run_ops")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


only_fused = os.environ.get("YR_ONLY_FUSED")
t_sep = 0.0 if only_fused else timed([d, q])
t_fus = timed([f])
byts = (B * H * W * C + B * Ho * Wo * N) * 4
print("dwpw %dx%dx%dx%d -> N%d s%d: separate %.4f ms, fused %.4f ms (%.0f GB/s of in+out), identical=%s" % (
    B, H, W, C, N, S, t_sep, t_fus, byts / t_fus / 1e6, bool(only_fused) or torch.equal(out_s, out_f)))
