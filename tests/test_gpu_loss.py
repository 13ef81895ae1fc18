"""GPU parity of the fused YoloLoss forward/backward against the torch-autograd oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from yoloret_b200.yolo3.model import YoloLoss, yolo_loss  # noqa: E402
from oracle import loss as oloss  # noqa: E402


def _make(B, hw, ncls, anchors, n_boxes, seed):
    rng = np.random.default_rng(seed)
    yts = [[] for _ in range(3)]
    for _ in range(B):
        wh = rng.uniform(0.05, 0.6, (n_boxes, 2)) * np.array(hw[::-1])
        cxy = rng.uniform(0.15, 0.85, (n_boxes, 2)) * np.array(hw[::-1])
        lim = np.array([hw[1] - 1, hw[0] - 1])
        tb = np.concatenate([np.clip(cxy - wh / 2, 0, lim), np.clip(cxy + wh / 2, 0, lim),
                             rng.integers(0, ncls, (n_boxes, 1))], 1)
        yt = oloss.preprocess_true_boxes(tb, hw, anchors, ncls) if n_boxes else \
            [np.zeros((hw[0] // s, hw[1] // s, 3, 5 + ncls), np.float32) for s in (32, 16, 8)]
        for s in range(3):
            yts[s].append(yt[s])
    yts = [torch.tensor(np.stack(y)) for y in yts]
    g = torch.Generator().manual_seed(seed)
    yos = [torch.randn(y.shape, generator=g) * 1.5 for y in yts]
    return yts, yos


@pytest.mark.parametrize("B,hw,ncls,n_boxes", [(2, (128, 128), 20, 6), (3, (96, 160), 80, 8), (2, (64, 64), 4, 0)])
def test_loss_and_grad_match_oracle(built_lib, anchors, B, hw, ncls, n_boxes):
    yts, yos = _make(B, hw, ncls, anchors, n_boxes, seed=B)
    exercised = 0
    for idx in range(3):
        ref_in = yos[idx].double().requires_grad_(True)
        ref, parts = oloss.yolo_loss_scale(yts[idx].double(), ref_in, idx, anchors)
        ref.backward()
        out = yos[idx].cuda().requires_grad_(True)
        L = YoloLoss(idx, anchors, 3, print_loss=False)
        loss = L(yts[idx].cuda(), out)
        loss.backward()
        got_parts = L.last_parts.cpu().numpy()
        np.testing.assert_allclose(got_parts[:3], [float(p) for p in parts[:3]], rtol=2e-4, atol=1e-5)
        assert got_parts[3] == float(parts[3])  # sum(ignore_mask)
        np.testing.assert_allclose(float(loss.detach()), float(ref.detach()), rtol=2e-4)
        gref = ref_in.grad.float().numpy()
        ggot = out.grad.cpu().numpy()
        np.testing.assert_allclose(ggot, gref, rtol=2e-3, atol=2e-6)
        if float(yts[idx][..., 4].sum()) > 0:  # this scale holds objects: the GIoU branch is exercised
            assert np.abs(gref[..., :4]).max() > 0
            exercised += 1
    assert exercised > 0 or n_boxes == 0


def test_functional_yolo_loss_sums_scales(built_lib, anchors):
    yts, yos = _make(2, (96, 96), 20, anchors, 5, seed=9)
    ref = oloss.yolo_loss([y.double() for y in yts], [y.double() for y in yos], anchors)
    got = yolo_loss([y.cuda() for y in yts], [y.cuda() for y in yos], anchors)
    np.testing.assert_allclose(float(got), float(ref), rtol=2e-4)


@pytest.mark.parametrize("B,T,hw,ncls,seed", [(4, 8, (416, 416), 80, 0), (3, 20, (320, 480), 20, 1), (5, 6, (96, 96), 4, 2),
                                              (2, 0, (64, 64), 4, 3)])
def test_preprocess_true_boxes_gpu_equals_host_and_oracle(built_lib, anchors, B, T, hw, ncls, seed):
    """The device y_true encoder (yr_encode_true_boxes) == the oracle's restatement of preprocess_true_boxes
    (reference code/yolo3/utils.py:298-376), bit for bit: float32 floor-divided centres, float64 divisions, best
    anchor by shape IoU, later boxes overwriting earlier ones in a shared slot (both class bits stay set), zero-width
    padding rows, and the reference's valid-box counter indexing the unfiltered rows."""
    from yoloret_b200.yolo3.utils import preprocess_true_boxes_gpu, preprocess_true_boxes
    rng = np.random.default_rng(seed)
    boxes = np.zeros((B, T, 5), np.float32)
    for b in range(B):
        n = T if b == 0 else int(rng.integers(0, T + 1))        # valid boxes first, zero padding behind (data.py)
        wh = rng.uniform(4, 0.7 * min(hw), (n, 2))
        c = rng.uniform(0, 1, (n, 2)) * np.array(hw[::-1])
        lo = np.clip(np.floor(c - wh / 2), 0, np.array(hw[::-1]) - 2)
        hi = np.clip(np.ceil(c + wh / 2), lo + 1, np.array(hw[::-1]) - 1)
        boxes[b, :n, 0:2], boxes[b, :n, 2:4] = lo, hi
        boxes[b, :n, 4] = rng.integers(0, ncls, n)
        if n >= 3:                                              # two boxes in one slot, different classes
            boxes[b, 1] = boxes[b, 0]
            boxes[b, 1, 4] = (boxes[b, 0, 4] + 1) % ncls
    if T >= 4:
        boxes[B - 1, 1, 2] = boxes[B - 1, 1, 0]                 # a zero-width row BEFORE valid ones: the index quirk
    got = preprocess_true_boxes_gpu(torch.from_numpy(boxes).cuda(), hw, anchors, ncls)
    torch.cuda.synchronize()
    for b in range(B):
        ref = oloss.preprocess_true_boxes(boxes[b], hw, anchors, ncls)
        host = preprocess_true_boxes(boxes[b], hw, anchors, ncls)
        for l in range(3):
            assert got[l].shape[1:] == ref[l].shape
            assert np.array_equal(host[l], ref[l])
            assert np.array_equal(got[l][b].cpu().numpy(), ref[l]), (b, l)
    with pytest.raises(ValueError):
        preprocess_true_boxes_gpu(torch.from_numpy(boxes), hw, anchors, ncls)   # host tensor: no CPU fallback


# ---- sparse y_true + the all-scales loss kernel ----------------------------------------------------------------------
def _boxes(B, T, hw, ncls, seed):
    rng = np.random.default_rng(seed)
    boxes = np.zeros((B, T, 5), np.float32)
    for b in range(B):
        n = T if b == 0 else int(rng.integers(0, T + 1))
        wh = rng.uniform(4, 0.7 * min(hw), (n, 2))
        c = rng.uniform(0, 1, (n, 2)) * np.array(hw[::-1])
        lo = np.clip(np.floor(c - wh / 2), 0, np.array(hw[::-1]) - 2)
        hi = np.clip(np.ceil(c + wh / 2), lo + 1, np.array(hw[::-1]) - 1)
        boxes[b, :n, 0:2], boxes[b, :n, 2:4] = lo, hi
        boxes[b, :n, 4] = rng.integers(0, ncls, n)
        if n >= 3:                                              # two boxes in one slot, different classes
            boxes[b, 1] = boxes[b, 0]
            boxes[b, 1, 4] = (boxes[b, 0, 4] + 1) % ncls
    if T >= 4:
        boxes[B - 1, 1, 2] = boxes[B - 1, 1, 0]                 # a zero-width row before valid ones: the index quirk
    return boxes


@pytest.mark.parametrize("B,T,hw,ncls,seed", [(4, 8, (416, 416), 80, 0), (3, 20, (320, 480), 20, 1), (5, 6, (96, 96), 4, 2),
                                              (2, 0, (64, 64), 4, 3), (3, 12, (128, 128), 128, 4)])
def test_sparse_y_true_equals_dense(built_lib, anchors, B, T, hw, ncls, seed):
    """yr_encode_true_boxes_sparse (slot maps + records) expands to exactly the dense tensors of the reference's
    preprocess_true_boxes (code/yolo3/utils.py:298-376), slot collisions and the row-index quirk included."""
    from yoloret_b200.yolo3.utils import encode_true_boxes_sparse, preprocess_true_boxes_gpu
    boxes = _boxes(B, T, hw, ncls, seed)
    tb = torch.from_numpy(boxes).cuda()
    sp = encode_true_boxes_sparse(tb, hw, anchors, ncls)
    dense = preprocess_true_boxes_gpu(tb, hw, anchors, ncls)
    for l, (a, b) in enumerate(zip(sp.to_dense(), dense)):
        assert torch.equal(a, b), l
        ref = np.stack([oloss.preprocess_true_boxes(boxes[i], hw, anchors, ncls)[l] for i in range(B)])
        assert np.array_equal(a.cpu().numpy(), ref)
    assert int(sp.counts.sum()) == int(sum((d[..., 4] > 0).sum() for d in dense))
    sp2 = encode_true_boxes_sparse(tb, hw, anchors, ncls, out=sp)   # refill in place: same result
    assert sp2 is sp and all(torch.equal(a, b) for a, b in zip(sp.to_dense(), dense))
    with pytest.raises(ValueError):
        encode_true_boxes_sparse(tb.cpu(), hw, anchors, ncls)
    with pytest.raises(ValueError):
        encode_true_boxes_sparse(tb, hw, anchors, 200)


@pytest.mark.parametrize("B,T,hw,ncls", [(3, 8, (128, 160), 20), (32, 8, (416, 416), 80), (2, 0, (64, 64), 4), (5, 30, (96, 96), 3)])
def test_fused_loss_matches_per_scale_loss_and_oracle(built_lib, anchors, B, T, hw, ncls):
    """FusedYoloLoss (one launch, sparse y_true) == the sum of the per-scale YoloLoss on the dense y_true of the same
    boxes (values and gradients), and both match the fp64 autograd oracle; (32, 8, 416, 80) is the cfg5 per-GPU shape."""
    from yoloret_b200.yolo3.utils import encode_true_boxes_sparse
    from yoloret_b200.yolo3.model import FusedYoloLoss
    boxes = _boxes(B, T, hw, ncls, seed=B + T)
    tb = torch.from_numpy(boxes).cuda()
    sp = encode_true_boxes_sparse(tb, hw, anchors, ncls)
    dense = sp.to_dense()
    g = torch.Generator().manual_seed(5)
    outs_h = [torch.randn(d.shape, generator=g) * 1.5 for d in dense]
    outs = [o.cuda().requires_grad_(True) for o in outs_h]
    fused = FusedYoloLoss(anchors, 3)
    total = fused(sp, outs)
    total.backward()
    parts = fused.last_parts.cpu().numpy()
    ref_total = 0.0
    for idx in range(3):
        o2 = outs_h[idx].cuda().requires_grad_(True)
        L = YoloLoss(idx, anchors, 3, print_loss=False)
        loss = L(dense[idx], o2)
        loss.backward()
        np.testing.assert_allclose(parts[idx, :3], L.last_parts.cpu().numpy()[:3], rtol=2e-5, atol=1e-6)
        assert parts[idx, 3] == float(L.last_parts[3])
        np.testing.assert_allclose(outs[idx].grad.cpu().numpy(), o2.grad.cpu().numpy(), rtol=1e-5, atol=1e-7)
        ref_in = outs_h[idx].double().requires_grad_(True)
        ref, rparts = oloss.yolo_loss_scale(dense[idx].cpu().double(), ref_in, idx, anchors)
        ref.backward()
        np.testing.assert_allclose(parts[idx, :3], [float(p) for p in rparts[:3]], rtol=2e-4, atol=1e-5)
        np.testing.assert_allclose(outs[idx].grad.cpu().numpy(), ref_in.grad.float().numpy(), rtol=2e-3, atol=2e-6)
        ref_total += float(ref.detach())
    np.testing.assert_allclose(float(total.detach()), ref_total, rtol=2e-4)


def test_fused_loss_graph_replay(built_lib, anchors):
    """encode (sparse) + loss forward/backward captured in ONE CUDA graph: replays follow new boxes / logits written into
    the captured buffers and repeat bit-exactly (the kernel re-arms its own arrival counter)."""
    from yoloret_b200.yolo3.utils import encode_true_boxes_sparse
    from yoloret_b200.yolo3.model import FusedYoloLoss
    B, T, hw, ncls = 4, 8, (128, 128), 20
    tb = torch.from_numpy(_boxes(B, T, hw, ncls, 1)).cuda()
    sp = encode_true_boxes_sparse(tb, hw, anchors, ncls)
    outs = [torch.randn(B, hw[0] // s, hw[1] // s, 3, 5 + ncls, device="cuda") for s in (32, 16, 8)]
    fused = FusedYoloLoss(anchors, 3)
    fused._run(sp, outs, True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        encode_true_boxes_sparse(tb, hw, anchors, ncls, out=sp)
        fused._run(sp, outs, True)
    seen = []
    for seed in (1, 2, 1):
        tb.copy_(torch.from_numpy(_boxes(B, T, hw, ncls, seed)))
        for o in outs:
            o.copy_(torch.randn(o.shape, generator=torch.Generator().manual_seed(seed)).cuda())
        g.replay()
        torch.cuda.synchronize()
        seen.append((fused.last_parts.clone(), [d.clone() for d in fused._dl]))
        eager = FusedYoloLoss(anchors, 3)
        ep, ed = eager._run(encode_true_boxes_sparse(tb, hw, anchors, ncls), outs, True)
        assert torch.equal(ep, seen[-1][0]) and all(torch.equal(a, b) for a, b in zip(ed, seen[-1][1]))
    assert torch.equal(seen[0][0], seen[2][0]) and not torch.equal(seen[0][0], seen[1][0])
