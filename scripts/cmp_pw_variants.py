"""Times every distinct pointwise layer shape of a network with the SS (2) and TS (3) tcgen05 kernels.
usage: cmp_pw_variants.py [model] [size] [classes] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from yoloret_b200 import _lib  # noqa: E402
from yoloret_b200._lib import YrOp  # noqa: E402
from yoloret_b200.netdef import NetDef  # noqa: E402
from ophelp import pack_tc  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "mobilenetv2x75"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 416
ncls = int(sys.argv[3]) if len(sys.argv) > 3 else 80
B = int(sys.argv[4]) if len(sys.argv) > 4 else 64
nd = NetDef(model, ncls, (size, size))
seen = {}
for L in nd.layers:
    if L.kind == "pw":
        key = (L.inp[0].H, L.inp[0].W, L.inp[0].C, L.out.C, L.res is not None)
        seen.setdefault(key, []).append(L.name)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
tot = {2: 0.0, 3: 0.0, "best": 0.0, "ideal": 0.0}
for (H, W, K, N, has_res), names in seen.items():
    a = torch.randn(B, H, W, K, device="cuda")
    w = torch.randn(K, N, device="cuda") * K ** -0.5
    bias = torch.randn(N, device="cuda")
    res = torch.randn(B, H, W, N, device="cuda") if has_res else None
    out = torch.empty(B, H, W, N, device="cuda")
    best = {}
    for v in (2, 3):
        op = YrOp()
        op.kind, op.act, op.variant = _lib.OP_PW, 1, v
        op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, K, H, W, N
        op.ld_in, op.ld_out = K, N
        op.in_, op.out, op.w, op.bias = a.data_ptr(), out.data_ptr(), w.data_ptr(), bias.data_ptr()
        if has_res:
            op.res, op.ld_res = res.data_ptr(), N
        packed = pack_tc(w, v)
        op.w_tc = packed.data_ptr()
        ops = (YrOp * 1)(op)
        ts = []
        for i in range(4):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(_lib.lib().yr_run_ops(ops, 1, st), "yr_run_ops")
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        best[v] = min(ts) * 1e3
    byts = B * H * W * (K + N * (2 if has_res else 1)) * 4
    ideal = byts / 6.5e6
    n = len(names)
    tot[2] += n * best[2]
    tot[3] += n * best[3]
    tot["best"] += n * min(best.values())
    tot["ideal"] += n * ideal
    print("%3dx%-3d K%-4d N%-4d res%d x%d | v2 %7.1f us | v3 %7.1f us | ideal %6.1f us" % (
        H, W, K, N, has_res, n, best[2], best[3], ideal), flush=True)
print("sum over layers (us): v2 %.0f  v3 %.0f  best-of %.0f  ideal(6.5TB/s) %.0f" % (
    tot[2], tot[3], tot["best"], tot["ideal"]))
