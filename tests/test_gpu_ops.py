"""GPU parity of the single-layer kernels against plain PyTorch fp32/fp64 references (through the C-ABI)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from yoloret_b200 import _lib  # noqa: E402
from yoloret_b200._lib import YrOp  # noqa: E402
from ophelp import run_op, act_ref, pw_op, pack_tc, ACT  # noqa: E402

RTOL, ATOL = 2e-5, 2e-5  # fp32 kernels vs an fp64 reference of the same op


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


@pytest.mark.parametrize("B,H,W,K,N,ld_in,act,use_res,use_scale", [
    (2, 7, 9, 16, 8, 16, "none", False, False),
    (2, 7, 9, 24, 24, 24, "relu6", True, False),
    (3, 16, 16, 144, 24, 144, "none", True, False),
    (1, 13, 13, 216, 512, 216, "relu6", False, False),
    (2, 13, 13, 512, 80, 512, "none", False, True),
    (2, 26, 26, 424, 256, 424, "relu6", False, False),
    (1, 52, 52, 256, 256, 256, "none", False, False),
    (2, 5, 5, 72, 72, 168, "swish", False, False),      # strided A (concat slice), N=72 tile
    (2, 5, 5, 96, 144, 96, "swish", True, True),
    (1, 3, 3, 720, 120, 720, "none", True, False),
    (2, 9, 9, 48, 96, 48, "relu6", False, False),
    (2, 40, 40, 16, 96, 16, "relu6", False, False),     # K < one k-block, many M tiles (weights stay resident)
    (1, 30, 30, 120, 720, 120, "relu6", False, False),  # N > 256: three n tiles
    (3, 13, 13, 512, 512, 512, "relu6", False, False),  # streamed weights, two n tiles
    (2, 64, 64, 256, 256, 256, "none", True, False),    # widest single n tile, > 1 tile per CTA? (64 tiles)
    (4, 100, 100, 24, 144, 24, "relu6", False, False),  # 313 tiles > 148 CTAs: persistent loop, both accumulators
    (8, 52, 52, 128, 256, 128, "none", False, False),   # 169 row blocks x 2 n tiles: ring wrap-around of every role
    (8, 40, 40, 432, 72, 432, "none", True, False),     # streamed weights over many k-blocks and tiles
    (6, 48, 48, 48, 32, 48, "swish", False, True),      # 4 narrow accumulators, SE gate, 108 tiles
])
@pytest.mark.parametrize("variant", [1, 2, 3, 4])   # 4 = the CTA-pair (cta_group::2) form of 3
def test_pw_parity(built_lib, B, H, W, K, N, ld_in, act, use_res, use_scale, variant):
    a = _rand(B, H, W, ld_in, seed=1)
    w = _rand(K, N, seed=2, scale=K ** -0.5)
    bias = _rand(N, seed=3)
    res = _rand(B, H, W, N, seed=4) if use_res else None
    scale = torch.rand(B, K, generator=torch.Generator().manual_seed(5)) if use_scale else None
    ad = a[..., :K].double()
    if use_scale:
        ad = ad * scale.double()[:, None, None, :]
    ref = act_ref(ad @ w.double() + bias.double(), act)
    if use_res:
        ref = ref + res.double()
    out = pw_op(a.cuda(), w.cuda(), bias.cuda(), act, res.cuda() if use_res else None,
                scale.cuda() if use_scale else None, ld_out=N + 8, variant=variant)
    got = out[..., :N].cpu().double()
    assert torch.isnan(out[..., N:]).all(), "kernel wrote outside its channel slice"
    torch.testing.assert_close(got, ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("B,H,W,K,N,act,use_res,use_scale", [
    (3, 16, 16, 144, 24, "none", True, False),
    (2, 26, 26, 424, 256, "relu6", False, False),
    (2, 13, 13, 512, 80, "none", False, True),
    (1, 30, 30, 120, 720, "swish", False, False),
    (4, 100, 100, 24, 144, "relu6", False, False),
])
def test_pw_tensor_core_variants_bit_identical(built_lib, B, H, W, K, N, act, use_res, use_scale):
    """The shared-memory-A (2) and tensor-memory-A (3) tcgen05 kernels issue the same three MMAs per K step in the
    same order on the same split operands, so the engine's per-layer autotuner may pick either without changing
    a single output bit."""
    a = _rand(B, H, W, K, seed=11).cuda()
    w = _rand(K, N, seed=12, scale=K ** -0.5).cuda()
    bias = _rand(N, seed=13).cuda()
    res = _rand(B, H, W, N, seed=14).cuda() if use_res else None
    scale = torch.rand(B, K, generator=torch.Generator().manual_seed(15)).cuda() if use_scale else None
    o2 = pw_op(a, w, bias, act, res, scale, variant=2)
    o3 = pw_op(a, w, bias, act, res, scale, variant=3)
    assert torch.equal(o2, o3), float((o2 - o3).abs().max())
    o4 = pw_op(a, w, bias, act, res, scale, variant=4)   # the CTA pair issues the same MMAs on 256-row tiles
    assert torch.equal(o3, o4), float((o3 - o4).abs().max())


@pytest.mark.parametrize("variant", [2, 3, 4])
@pytest.mark.parametrize("B,H,W,K,N,act", [(3, 13, 13, 256, 256, "relu6"), (2, 26, 26, 256, 128, "relu6"), (2, 7, 5, 48, 48, "swish"),
                                           (5, 26, 26, 48, 96, "none")])
def test_pw_fused_upsampling(built_lib, variant, B, H, W, K, N, act):
    """Conv1x1+BN+act followed by UpSampling2D (nearest x2, reference code/yolo3/model.py:254,274) as ONE op: the
    epilogue stores every output row to its 2x2 upsampled pixels of a wider concat buffer.  Must equal the plain op
    followed by a nearest upsample bit for bit, and leave the rest of the buffer alone."""
    a = _rand(B, H, W, K, seed=31).cuda()
    w = _rand(K, N, seed=32, scale=K ** -0.5).cuda()
    bias = _rand(N, seed=33).cuda()
    plain = pw_op(a, w, bias, act, variant=variant)
    fused = pw_op(a, w, bias, act, variant=variant, ld_out=N + 16, up2=True)
    want = plain.repeat_interleave(2, 1).repeat_interleave(2, 2)
    assert fused.shape == (B, 2 * H, 2 * W, N + 16)
    assert torch.equal(fused[..., :N], want)
    assert torch.isnan(fused[..., N:]).all()


@pytest.mark.parametrize("variant", [3, 4])
@pytest.mark.parametrize("B,H,W,K,N1,N2,act2,gate", [
    (8, 52, 52, 128, 256, 128, "relu6", True),    # conv2d_23*{conv2d_24 (y3, linear), conv2d_25} of the benchmarked net
    (16, 26, 26, 256, 256, 256, "relu6", True),   # conv2d_29*{conv2d_30 (y2), conv2d_31}
    (3, 13, 13, 72, 24, 40, "swish", False),      # split inside an n tile and inside a 32-column chunk (24 % 32 != 0)
    (2, 9, 7, 48, 200, 120, "none", False),       # split in the second n tile (N = 320: two tiles of 160)
    (1, 6, 8, 256, 256, 256, "relu6", True),      # 48 rows: one partial row block (a CTA pair's second block is empty)
    (2, 12, 16, 128, 256, 128, "relu6", True),    # three row blocks: an odd number for the pair kernel
    (7, 6, 8, 256, 256, 256, "relu6", True),
])
def test_pw_stacked_outputs(built_lib, variant, B, H, W, K, N1, N2, act2, gate):
    """Two 1x1 convs that read the same tensor (reference code/yolo3/model.py:296-305: a head stage's y conv and the next
    bottom-up conv) as ONE GEMM over [W1 | W2] with two destinations: every output bit equals the two separate ops
    (a column's accumulation does not depend on its neighbours), the first block stays linear, and nothing is written
    outside either channel slice."""
    a = _rand(B, H, W, K, seed=41).cuda()
    w1 = _rand(K, N1, seed=42, scale=K ** -0.5).cuda()
    w2 = _rand(K, N2, seed=43, scale=K ** -0.5).cuda()
    b1, b2 = _rand(N1, seed=44).cuda(), _rand(N2, seed=45).cuda()
    scale = torch.rand(B, K, generator=torch.Generator().manual_seed(46)).cuda() if gate else None
    sep1 = pw_op(a, w1, b1, "none", None, scale, variant=variant)
    sep2 = pw_op(a, w2, b2, act2, None, scale, variant=variant)
    w = torch.cat([w1, w2], 1).contiguous()
    bias = torch.cat([b1, b2]).contiguous()
    out1 = torch.full((B, H, W, N1 + 8), float("nan"), device="cuda")
    out2 = torch.full((B, H, W, N2 + 16), float("nan"), device="cuda")
    op = YrOp()
    op.kind, op.act, op.variant = _lib.OP_PW, ACT[act2], variant
    op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, K, H, W, N1 + N2
    op.ld_in, op.ld_out, op.ld_in2, op.K2, op.K3 = K, N1 + 8, N2 + 16, N1, 1
    op.in_, op.out, op.aux = a.data_ptr(), out1.data_ptr(), out2.data_ptr()
    op.w, op.bias = w.data_ptr(), bias.data_ptr()
    if gate:
        op.scale = scale.data_ptr()
    packed = pack_tc(w, variant)
    op.w_tc = packed.data_ptr()
    run_op(op)
    assert torch.equal(out1[..., :N1], sep1), float((out1[..., :N1] - sep1).abs().max())
    assert torch.equal(out2[..., :N2], sep2), float((out2[..., :N2] - sep2).abs().max())
    assert torch.isnan(out1[..., N1:]).all() and torch.isnan(out2[..., N2:]).all()
    bad = YrOp.from_buffer_copy(op)
    bad.variant = 2                                   # the shared-memory-A kernel has no second destination
    assert _lib.lib().yr_run_ops((YrOp * 1)(bad), 1, torch.cuda.current_stream().cuda_stream) < 0


@pytest.mark.parametrize("variant", [2, 3, 4])
@pytest.mark.parametrize("B,H,W,K,N", [(32, 26, 26, 256, 256), (64, 13, 13, 512, 256), (16, 52, 52, 128, 256)])
def test_pw_streamed_weights_repeatable(built_lib, variant, B, H, W, K, N):
    """Wide layers stream their weight tiles through a shared-memory ring while the activation ring, the TMEM
    stages and the accumulators all wrap many times: 25 back-to-back launches must give the same bits every time
    (a ring-protocol race shows up as a changed output or a trapped launch, not as a tolerance miss)."""
    a = _rand(B, H, W, K, seed=21).cuda()
    w = _rand(K, N, seed=22, scale=K ** -0.5).cuda()
    bias = _rand(N, seed=23).cuda()
    first = pw_op(a, w, bias, "relu6", variant=variant)
    ref = act_ref(a.double().cpu() @ w.double().cpu() + bias.double().cpu(), "relu6")
    torch.testing.assert_close(first.cpu().double(), ref, rtol=RTOL, atol=ATOL)
    for _ in range(24):
        again = pw_op(a, w, bias, "relu6", variant=variant)
        assert torch.equal(first, again)


def _same_pad_lead(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def _dw_ref(x, w, bias, k, s, act):
    B, H, W, C = x.shape
    Ho, pt, pb = _same_pad_lead(H, k, s)
    Wo, pl, pr = _same_pad_lead(W, k, s)
    xn = F.pad(x.permute(0, 3, 1, 2).double(), (pl, pr, pt, pb))
    wn = w.reshape(k, k, C).permute(2, 0, 1)[:, None].double()
    y = F.conv2d(xn, wn, bias.double(), stride=s, groups=C)
    return act_ref(y, act).permute(0, 2, 3, 1), (Ho, Wo, pt, pl)


@pytest.mark.parametrize("B,H,W,C,k,s,act", [
    (2, 16, 16, 24, 3, 1, "relu6"),
    (2, 16, 16, 96, 3, 2, "relu6"),    # even input: TF SAME pads (0,1)
    (1, 13, 13, 720, 3, 1, "relu6"),   # odd input
    (2, 13, 13, 432, 3, 2, "relu6"),   # odd input stride 2: pads (1,1)
    (2, 26, 26, 48, 5, 1, "relu6"),
    (1, 13, 13, 512, 3, 1, "swish"),
    (2, 17, 11, 40, 5, 2, "swish"),
    (1, 5, 3, 8, 3, 1, "none"),
    (3, 104, 104, 144, 3, 1, "relu6"),  # many spatial tiles per image, last channel tile half empty
    (3, 104, 104, 144, 3, 2, "relu6"),
    (2, 208, 208, 24, 3, 1, "relu6"),   # narrow channel box (24 < 32)
    (5, 52, 52, 128, 3, 1, "swish"),    # more work items than resident CTAs: both pipeline stages wrap
    (2, 61, 45, 64, 3, 2, "none"),      # odd sizes, ragged tiles in both directions
    (2, 26, 26, 288, 3, 1, "relu6"),
])
@pytest.mark.parametrize("pad_ld", [0, 24])
def test_dw_parity(built_lib, B, H, W, C, k, s, act, pad_ld):
    """pad_ld > 0: input and output are channel slices of wider buffers (the concat layout)."""
    x = _rand(B, H, W, C, seed=1)
    w = _rand(k * k, C, seed=2, scale=0.3)
    bias = _rand(C, seed=3)
    ref, (Ho, Wo, pt, pl) = _dw_ref(x, w, bias, k, s, act)
    wd, bd = w.cuda(), bias.cuda()
    xw = torch.full((B, H, W, C + pad_ld), float("nan"), device="cuda")
    xw[..., 8 * (pad_ld > 0):8 * (pad_ld > 0) + C] = x.cuda()
    out = torch.full((B, Ho, Wo, C + pad_ld), float("nan"), device="cuda")
    off = 8 * (pad_ld > 0)
    op = YrOp()
    op.kind, op.act = _lib.OP_DW, ACT[act]
    op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, C, Ho, Wo, C
    op.k, op.stride, op.pad_t, op.pad_l, op.ld_in, op.ld_out = k, s, pt, pl, C + pad_ld, C + pad_ld
    op.in_, op.out, op.w, op.bias = xw.data_ptr() + 4 * off, out.data_ptr() + 4 * off, wd.data_ptr(), bd.data_ptr()
    run_op(op)
    torch.testing.assert_close(out[..., off:off + C].cpu().double(), ref, rtol=RTOL, atol=ATOL)
    if pad_ld:
        assert torch.isnan(out[..., :off]).all() and torch.isnan(out[..., off + C:]).all()


@pytest.mark.parametrize("B,H,W,C,N,s,dw_act,pw_act,use_res", [
    (2, 16, 16, 96, 24, 1, "relu6", "none", True),      # one tile per image, resident weights
    (3, 104, 104, 144, 24, 1, "relu6", "none", True),   # block_2: many tiles, C not a multiple of 32
    (2, 104, 104, 144, 24, 2, "relu6", "none", False),  # block_3: stride 2, TF SAME pads (0,1)
    (2, 208, 208, 96, 24, 2, "relu6", "none", False),   # block_1
    (2, 208, 208, 24, 16, 1, "relu6", "none", False),   # expanded_conv: narrow K
    (3, 52, 52, 144, 48, 1, "relu6", "none", True),
    (2, 26, 26, 288, 72, 1, "relu6", "none", True),
    (3, 26, 26, 432, 72, 1, "relu6", "none", False),    # streamed weight ring
    (2, 26, 26, 432, 120, 2, "relu6", "none", False),   # block_13: odd output 13x13, pads (0,1), streamed
    (5, 13, 13, 720, 120, 1, "relu6", "none", True),    # 13x13: two ragged tiles per image
    (2, 61, 45, 64, 40, 2, "none", "relu6", False),     # odd sizes, other activations
    (2, 17, 23, 40, 192, 1, "swish", "swish", True),    # widest single n tile
    (1, 5, 3, 8, 8, 1, "relu6", "none", False),
])
def test_dwpw_bit_identical_to_separate_ops(built_lib, B, H, W, C, N, s, dw_act, pw_act, use_res):
    """YR_OP_DWPW (3x3 depthwise + BN + act computed by the converter warps of the tcgen05 pointwise kernel) gives the
    same BITS as the depthwise op followed by the pointwise op (variant 3), and matches an fp64 reference; concat-slice
    leading dimensions on both sides; repeated launches are bit-stable (ring protocol)."""
    x = _rand(B, H, W, C, seed=31)
    wd = _rand(9, C, seed=32, scale=0.3).cuda()
    bd = _rand(C, seed=33).cuda()
    wp = _rand(C, N, seed=34, scale=C ** -0.5).cuda()
    bp = _rand(N, seed=35).cuda()
    ref_dw, (Ho, Wo, pt, pl) = _dw_ref(x, wd.cpu(), bd.cpu(), 3, s, dw_act)
    res = _rand(B, Ho, Wo, N, seed=36).cuda() if use_res else None
    ref = act_ref(ref_dw @ wp.double().cpu() + bp.double().cpu(), pw_act)
    if use_res:
        ref = ref + res.double().cpu()
    ld_in, off = C + 16, 8
    xw = torch.full((B, H, W, ld_in), float("nan"), device="cuda")
    xw[..., off:off + C] = x.cuda()
    # separate ops
    mid = torch.full((B, Ho, Wo, C), float("nan"), device="cuda")
    op = YrOp()
    op.kind, op.act = _lib.OP_DW, ACT[dw_act]
    op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, C, Ho, Wo, C
    op.k, op.stride, op.pad_t, op.pad_l, op.ld_in, op.ld_out = 3, s, pt, pl, ld_in, C
    op.in_, op.out, op.w, op.bias = xw.data_ptr() + 4 * off, mid.data_ptr(), wd.data_ptr(), bd.data_ptr()
    run_op(op)
    sep = pw_op(mid, wp, bp, pw_act, res=res, variant=3, ld_out=N + 8)
    # fused
    lib = _lib.lib()
    assert lib.yr_dwpw_supported(C, N, s, Ho, Wo) == 1
    n = int(lib.yr_dwpw_packed_floats(C, N))
    blob = torch.full((n,), float("nan"), device="cuda")
    _lib.check(lib.yr_dwpw_pack(wp.data_ptr(), C, N, wd.data_ptr(), bd.data_ptr(), blob.data_ptr(),
                                torch.cuda.current_stream().cuda_stream), "yr_dwpw_pack")
    torch.cuda.synchronize()
    assert not torch.isnan(blob).any()
    out = torch.full((B, Ho, Wo, N + 8), float("nan"), device="cuda")
    f = YrOp()
    f.kind, f.act, f.mode = _lib.OP_DWPW, ACT[pw_act], ACT[dw_act]
    f.B, f.H, f.W, f.C, f.Ho, f.Wo, f.N = B, H, W, C, Ho, Wo, N
    f.k, f.stride, f.pad_t, f.pad_l, f.ld_in, f.ld_out = 3, s, pt, pl, ld_in, N + 8
    f.in_, f.out, f.w_tc, f.bias = xw.data_ptr() + 4 * off, out.data_ptr(), blob.data_ptr(), bp.data_ptr()
    if use_res:
        f.res, f.ld_res = res.data_ptr(), N
    run_op(f)
    assert torch.equal(out[..., :N], sep[..., :N]), float((out[..., :N] - sep[..., :N]).abs().max())
    assert torch.isnan(out[..., N:]).all()
    torch.testing.assert_close(out[..., :N].cpu().double(), ref, rtol=1e-4, atol=1e-4)
    first = out.clone()
    for _ in range(5):
        run_op(f)
        assert torch.equal(first[..., :N], out[..., :N])


def test_dwpw_rejects_unfusable(built_lib):
    lib = _lib.lib()
    assert lib.yr_dwpw_supported(1344, 224, 1, 19, 19) == 0      # two n tiles: the depthwise would be recomputed
    assert int(lib.yr_dwpw_packed_floats(1344, 224)) == 0
    f = YrOp()
    f.kind = _lib.OP_DWPW
    with pytest.raises(_lib.YrError):
        run_op(f)


@pytest.mark.parametrize("u8", [False, True])
@pytest.mark.parametrize("H,W,N,act", [(32, 32, 24, "relu6"), (33, 47, 40, "swish"),  # odd size: the generic kernel
                                       (96, 160, 24, "relu6"), (70, 52, 48, "relu6"), (37, 44, 32, "swish"),
                                       (416, 416, 24, "relu6")])
def test_stem_parity(built_lib, u8, H, W, N, act):
    B = 2
    g = torch.Generator().manual_seed(0)
    if u8:
        xi = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
        x = xi.float() * np.float32(1.0 / 255.0)
    else:
        x = xi = torch.rand(B, H, W, 3, generator=g)
    w = _rand(27, N, seed=2, scale=0.2)
    bias = _rand(N, seed=3)
    Ho, pt, pb = _same_pad_lead(H, 3, 2)
    Wo, pl, pr = _same_pad_lead(W, 3, 2)
    xn = F.pad(x.permute(0, 3, 1, 2).double(), (pl, pr, pt, pb))
    wn = w.reshape(3, 3, 3, N).permute(3, 2, 0, 1).double()
    ref = act_ref(F.conv2d(xn, wn, bias.double(), stride=2), act).permute(0, 2, 3, 1)
    xd, wd, bd = xi.cuda(), w.cuda(), bias.cuda()
    out = torch.full((B, Ho, Wo, N), float("nan"), device="cuda")
    op = YrOp()
    op.kind, op.act, op.in_is_u8 = _lib.OP_STEM, ACT[act], int(u8)
    op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, 3, Ho, Wo, N
    op.k, op.stride, op.pad_t, op.pad_l, op.ld_in, op.ld_out = 3, 2, pt, pl, 3, N
    op.in_, op.out, op.w, op.bias = xd.data_ptr(), out.data_ptr(), wd.data_ptr(), bd.data_ptr()
    run_op(op)
    torch.testing.assert_close(out.cpu().double(), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("mode", ["up2", "pool2", "pool4"])
def test_resample_exact(built_lib, mode):
    B, H, W, C = 2, 8, 12, 24
    x = _rand(B, H, W, C, seed=1)
    xn = x.permute(0, 3, 1, 2)
    if mode == "up2":
        ref = xn.repeat_interleave(2, 2).repeat_interleave(2, 3)
    else:
        p = 2 if mode == "pool2" else 4
        ref = F.max_pool2d(xn, p, p)
    ref = ref.permute(0, 2, 3, 1)
    Ho, Wo = ref.shape[1:3]
    out = torch.full((B, Ho, Wo, C + 16), float("nan"), device="cuda")
    xd = x.cuda()
    op = YrOp()
    op.kind, op.mode = _lib.OP_RESAMPLE, {"up2": 0, "pool2": 1, "pool4": 2}[mode]
    op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, C, Ho, Wo, C
    op.ld_in, op.ld_out = C, C + 16
    op.in_, op.out = xd.data_ptr(), out.data_ptr() + 8 * 4  # write into channel slice [8, 8+C)
    run_op(op)
    assert torch.equal(out[..., 8:8 + C].cpu(), ref)
    assert torch.isnan(out[..., :8]).all() and torch.isnan(out[..., 8 + C:]).all()


@pytest.mark.parametrize("H,W", [(6, 10), (8, 6), (26, 26)])  # 8 rows: ragged last row group of the kernel
def test_rfcr_parity(built_lib, H, W):
    B = 2  # H x W = the stride-16 grid
    K1, K2, K3, K4, N = 120, 72, 24, 24, 48
    b1, b2 = _rand(B, H // 2, W // 2, K1, seed=1), _rand(B, H, W, K2, seed=2)
    b3, b4 = _rand(B, 2 * H, 2 * W, K3, seed=3), _rand(B, 4 * H, 4 * W, K4, seed=4)
    ws = [_rand(k, N, seed=10 + i, scale=k ** -0.5) for i, k in enumerate((K1, K2, K3, K4))]
    alpha = torch.tensor([1.007, 1.663, -0.924, 0.589])  # a2 < 0 exercises the a2*max() order

    def c(x, w):
        return x.double() @ w.double()

    def nchw(x):
        return x.permute(0, 3, 1, 2)

    up = nchw(c(b1, ws[0])).repeat_interleave(2, 2).repeat_interleave(2, 3)
    p3 = F.max_pool2d(nchw(c(b3, ws[2])), 2, 2)
    p4 = nchw(c(F.max_pool2d(nchw(b4), 4, 4).permute(0, 2, 3, 1), ws[3]))
    a = alpha.double()
    ref = (a[0] * up + a[1] * nchw(c(b2, ws[1])) + a[2] * p3 + a[3] * p4).permute(0, 2, 3, 1)
    wcat = torch.cat(ws, 0).cuda()
    d = [t.cuda() for t in (b1, b2, b3, b4)]
    al = alpha.cuda()
    out = torch.full((B, H, W, N), float("nan"), device="cuda")
    op = YrOp()
    op.kind = _lib.OP_RFCR
    op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H // 2, W // 2, K1, H, W, N
    op.K2, op.K3, op.K4 = K2, K3, K4
    op.ld_in, op.ld_in2, op.ld_in3, op.ld_in4, op.ld_out = K1, K2, K3, K4, N
    op.in_, op.in2, op.in3, op.in4 = (t.data_ptr() for t in d)
    op.out, op.w, op.bias = out.data_ptr(), wcat.data_ptr(), al.data_ptr()
    run_op(op)
    torch.testing.assert_close(out.cpu().double(), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("HW,F_,R", [((13, 13), 512, 128), ((7, 5), 128, 32), ((4, 4), 144, 6), ((3, 3), 1392, 58)])
def test_se_parity(built_lib, HW, F_, R):
    B = 3
    x = _rand(B, HW[0], HW[1], F_, seed=1)
    w1, b1 = _rand(F_, R, seed=2, scale=F_ ** -0.5), _rand(R, seed=3, scale=0.1)
    w2, b2 = _rand(R, F_, seed=4, scale=R ** -0.5), _rand(F_, seed=5, scale=0.1)
    m = x.double().mean(dim=(1, 2))
    h = m @ w1.double() + b1.double()
    h = h * torch.sigmoid(h)
    ref = torch.sigmoid(h @ w2.double() + b2.double())
    xd = x.cuda()
    wcat = torch.cat([w1.reshape(-1), w2.reshape(-1)]).cuda()
    bcat = torch.cat([b1, b2]).cuda()
    out = torch.full((B, F_), float("nan"), device="cuda")
    op = YrOp()
    op.kind = _lib.OP_SE
    op.B, op.H, op.W, op.C, op.N = B, HW[0], HW[1], F_, R
    op.ld_in = F_
    op.in_, op.out, op.w, op.bias = xd.data_ptr(), out.data_ptr(), wcat.data_ptr(), bcat.data_ptr()
    run_op(op)
    torch.testing.assert_close(out.cpu().double(), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("B,H,W,F_,R,k,s", [(3, 13, 13, 512, 128, 3, 1), (2, 26, 26, 256, 64, 3, 1), (2, 20, 12, 128, 32, 3, 1),
                                             (2, 9, 7, 40, 10, 5, 2), (1, 8, 8, 1392, 58, 5, 1)])
def test_fused_squeeze_excite(built_lib, B, H, W, F_, R, k, s):
    """Depthwise op with the fused squeeze (aux partial sums) + SE_FC == mean/FC/swish/FC/sigmoid of its output;
    bit-identical run to run (no atomics)."""
    import ctypes as C
    x = _rand(B, H, W, F_, seed=1)
    w = _rand(k * k, F_, seed=2, scale=0.3)
    bias = _rand(F_, seed=3)
    dwref, (Ho, Wo, pt, pl) = _dw_ref(x, w, bias, k, s, "swish")
    w1, b1 = _rand(F_, R, seed=4, scale=F_ ** -0.5), _rand(R, seed=5, scale=0.1)
    w2, b2 = _rand(R, F_, seed=6, scale=R ** -0.5), _rand(F_, seed=7, scale=0.1)
    m = dwref.mean(dim=(1, 2))
    h = m @ w1.double() + b1.double()
    h = h * torch.sigmoid(h)
    ref = torch.sigmoid(h @ w2.double() + b2.double())
    xd, wd, bd = x.cuda(), w.cuda(), bias.cuda()
    out = torch.full((B, Ho, Wo, F_), float("nan"), device="cuda")
    op = YrOp()
    op.kind, op.act = _lib.OP_DW, ACT["swish"]
    op.B, op.H, op.W, op.C, op.Ho, op.Wo, op.N = B, H, W, F_, Ho, Wo, F_
    op.k, op.stride, op.pad_t, op.pad_l, op.ld_in, op.ld_out = k, s, pt, pl, F_, F_
    op.in_, op.out, op.w, op.bias = xd.data_ptr(), out.data_ptr(), wd.data_ptr(), bd.data_ptr()
    slots = built_lib.yr_dw_se_slots(C.byref(op))
    assert slots > 0
    gates = []
    for _ in range(2):
        part = torch.full((B, slots, F_), float("nan"), device="cuda")
        op.aux = part.data_ptr()
        run_op(op)
        torch.testing.assert_close(out.cpu().double(), dwref, rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(part.sum(1).cpu().double() / (Ho * Wo), m, rtol=1e-5, atol=1e-5)
        wcat = torch.cat([w1.t().reshape(-1), w2.reshape(-1)]).cuda()
        bcat = torch.cat([b1, b2]).cuda()
        gate = torch.full((B, F_), float("nan"), device="cuda")
        fc = YrOp()
        fc.kind = _lib.OP_SE_FC
        fc.B, fc.H, fc.W, fc.C, fc.N, fc.K2 = B, Ho, Wo, F_, R, slots
        fc.in_, fc.out, fc.w, fc.bias = part.data_ptr(), gate.data_ptr(), wcat.data_ptr(), bcat.data_ptr()
        run_op(fc)
        torch.testing.assert_close(gate.cpu().double(), ref, rtol=RTOL, atol=ATOL)
        gates.append(gate.clone())
    assert torch.equal(gates[0], gates[1])


def test_bad_arguments_fail_loudly(built_lib):
    op = YrOp()
    op.kind = _lib.OP_PW
    op.B, op.H, op.W, op.C, op.N = 1, 1, 1, 12, 8  # K not a multiple of 8, null pointers
    ops = (YrOp * 1)(op)
    rc = built_lib.yr_run_ops(ops, 1, None)
    assert rc == -1 and b"pw" in built_lib.yr_last_error()
    op.kind = 99
    assert built_lib.yr_run_ops((YrOp * 1)(op), 1, None) == -1
