"""GPU parity of yolo_eval (decode, class-wise NMS, packing) and letterbox against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from yoloret_b200 import _lib  # noqa: E402
from yoloret_b200.postprocess import PostProcess  # noqa: E402
from yoloret_b200.yolo3.model import yolo_eval, YoloEval  # noqa: E402
from yoloret_b200.yolo3.utils import letterbox_image  # noqa: E402
from oracle import postprocess as opp, letterbox as olb  # noqa: E402


def _synthetic_heads(B, grids, ncls, seed=1234):
    """SURVEY.md §8d synthetic head statistics."""
    rng = np.random.default_rng(seed)
    out = []
    for gh, gw in grids:
        t = np.empty((B, gh, gw, 3, 5 + ncls), np.float32)
        t[..., 0:2] = rng.normal(0, 1, t[..., 0:2].shape)
        t[..., 2:4] = rng.normal(0, 0.5, t[..., 2:4].shape)
        t[..., 4] = rng.normal(-4, 2, t[..., 4].shape)
        t[..., 5:] = rng.normal(-3, 2, t[..., 5:].shape)
        out.append(t)
    return out


def _run_nms(boxes, scores_per_class, max_boxes, iou_thr, score_thr):
    """boxes [T,4], scores_per_class [C,T] -> list of selected indices per class via the C-ABI."""
    lib = _lib.lib()
    Cn, T = scores_per_class.shape
    cap = T
    bd = torch.from_numpy(boxes[None].copy()).cuda()
    cs = torch.zeros(1, Cn, cap, device="cuda")
    ci = torch.zeros(1, Cn, cap, dtype=torch.int32, device="cuda")
    cc = torch.zeros(1, Cn, dtype=torch.int32, device="cuda")
    rng = np.random.default_rng(0)
    for c in range(Cn):
        idx = np.nonzero(scores_per_class[c] > np.float32(score_thr))[0]
        idx = rng.permutation(idx)  # candidate order inside a list is unspecified
        cs[0, c, :len(idx)] = torch.from_numpy(scores_per_class[c][idx])
        ci[0, c, :len(idx)] = torch.from_numpy(idx.astype(np.int32))
        cc[0, c] = len(idx)
    det = torch.zeros(1, Cn, max_boxes, 6, device="cuda")
    dc = torch.zeros(1, Cn, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.yr_nms_classwise(bd.data_ptr(), T, cs.data_ptr(), ci.data_ptr(), cc.data_ptr(), 1, Cn, cap,
                                    max_boxes, float(iou_thr), det.data_ptr(), dc.data_ptr(), status.data_ptr(), st))
    torch.cuda.synchronize()
    det, dc = det.cpu().numpy(), dc.cpu().numpy()
    return [det[0, c, :dc[0, c], 5].view(np.int32).copy() for c in range(Cn)], det, dc


@pytest.mark.parametrize("seed,T,quant", [(0, 300, 0), (1, 2000, 64), (2, 5000, 8), (3, 37, 4),
                                          (4, 10647, 0),      # MAP-mode sized list: shared-memory subset path
                                          (5, 12000, 4096),   # ... with score ties inside the subset threshold bin
                                          (6, 6000, 16)])     # ... ties too wide for the subset: plain loop
def test_nms_bit_exact(built_lib, seed, T, quant):
    """Selected indices identical to the oracle, including score ties (quantised scores),
    degenerate/zero-area boxes and swapped corners."""
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(0, 100, (T, 2)).astype(np.float32)
    wh = rng.uniform(0, 40, (T, 2)).astype(np.float32)
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
    boxes[::17, 2:] = boxes[::17, :2]                 # zero area
    boxes[5::23] = boxes[5::23][:, [2, 3, 0, 1]]      # swapped corners
    boxes[1::29] = boxes[0::29][:len(boxes[1::29])]   # exact duplicates
    Cn = 6
    scores = rng.uniform(0, 1, (Cn, T)).astype(np.float32)
    if quant:
        scores = (np.floor(scores * quant) / quant).astype(np.float32)
    for max_boxes, iou_thr, score_thr in [(20, 0.5, 0.2), (20, 0.3, 0.0), (7, 0.9, 0.5), (50, 0.0, 0.1)]:
        got, _, _ = _run_nms(boxes, scores, max_boxes, iou_thr, score_thr)
        for c in range(Cn):
            ref = opp.nms_c(boxes, scores[c], max_boxes, iou_thr, score_thr)
            assert np.array_equal(got[c], ref), (seed, c, max_boxes, iou_thr, score_thr, got[c], ref)


def test_nms_empty_and_single(built_lib):
    boxes = np.array([[0, 0, 10, 10]], np.float32)
    got, _, dc = _run_nms(boxes, np.array([[0.9], [0.1]], np.float32), 20, 0.5, 0.2)
    assert list(got[0]) == [0] and len(got[1]) == 0 and dc[0, 1] == 0


@pytest.mark.parametrize("ncls,hw,pad,thr", [(80, (416, 416), True, 0.2), (20, (320, 320), False, 0.3),
                                             (20, (96, 160), True, 0.0), (3, (64, 64), False, 0.05)])
def test_yolo_eval_parity(built_lib, anchors, ncls, hw, pad, thr):
    """Whole post-process on identical logits: same detections in the same (class-major, NMS)
    order; scores/boxes within 1e-3 relative (expf/sigmoid ULP differences GPU vs numpy);
    int boxes equal or +-1 at truncation boundaries."""
    B = 2
    grids = [(hw[0] // s, hw[1] // s) for s in (32, 16, 8)]
    heads = _synthetic_heads(B, grids, ncls)
    E = 3 * (ncls + 5)
    shapes = np.array([[375, 500], [600, 420]], np.float32)
    dev = []
    for t in heads:
        flat = torch.from_numpy(t.reshape(B, t.shape[1], t.shape[2], E))
        if pad:
            ld = (E + 7) // 8 * 8 + 8
            buf = torch.zeros(B, t.shape[1], t.shape[2], ld)
            buf[..., :E] = flat
            buf = buf.cuda()
            dev.append(buf.as_strided((B, t.shape[1], t.shape[2], 3, ncls + 5),
                                      (t.shape[1] * t.shape[2] * ld, t.shape[2] * ld, ld, ncls + 5, 1)))
        else:
            dev.append(flat.cuda().view(B, t.shape[1], t.shape[2], 3, ncls + 5))
    got = YoloEval(anchors, 3, ncls, score_threshold=thr, iou_threshold=0.5)(dev, shapes)
    for b in range(B):
        rb, rs, rc, rf = opp.yolo_eval([t[b:b + 1] for t in heads], anchors, 3, ncls, shapes[b],
                                       score_threshold=thr, iou_threshold=0.5, return_float_boxes=True)
        gb, gs, gc = (x.cpu().numpy() for x in got[b])
        assert len(gs) == len(rs) and len(rs) > 0
        assert np.array_equal(gc, rc)
        np.testing.assert_allclose(gs, rs, rtol=1e-3, atol=1e-6)
        assert np.abs(gb.astype(np.int64) - rb.astype(np.int64)).max() <= 1


def test_decode_boxes_and_candidates(built_lib, anchors):
    """Decoded boxes for ALL anchors and the per-class candidate sets match the oracle."""
    B, ncls, hw = 2, 20, (128, 96)
    grids = [(hw[0] // s, hw[1] // s) for s in (32, 16, 8)]
    heads = _synthetic_heads(B, grids, ncls, seed=5)
    pp = PostProcess(B, grids, ncls, anchors)
    shapes = np.array([[300, 500], [128, 96]], np.float32)
    pp.set_image_shapes(shapes)
    dev = [torch.from_numpy(t).cuda() for t in heads]
    pp.run([t.data_ptr() for t in dev], [3 * (ncls + 5)] * 3, 0.1, 0.5)
    torch.cuda.synchronize()
    boxes = pp.boxes.cpu().numpy()
    for b in range(B):
        rb, rs = opp.decode_all([t[b:b + 1] for t in heads], anchors, 3, ncls, shapes[b])
        np.testing.assert_allclose(boxes[b], rb, rtol=1e-5, atol=1e-4)
    # candidate sets: rerun decode only (NMS consumes the scores)
    p = pp.params(0.1, [3 * (ncls + 5)] * 3)
    fp = (C.c_void_p * 3)(*[t.data_ptr() for t in dev])
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(pp.lib.yr_decode_filter(fp, pp.image_shapes.data_ptr(), C.byref(p), pp.boxes.data_ptr(),
                                       pp.cand_score.data_ptr(), pp.cand_index.data_ptr(), pp.cand_count.data_ptr(), st))
    torch.cuda.synchronize()
    cnt, ci, cs = pp.cand_count.cpu().numpy(), pp.cand_index.cpu().numpy(), pp.cand_score.cpu().numpy()
    for b in range(B):
        _, rs = opp.decode_all([t[b:b + 1] for t in heads], anchors, 3, ncls, shapes[b])
        for c in range(ncls):
            got = dict(zip(ci[b, c, :cnt[b, c]].tolist(), cs[b, c, :cnt[b, c]].tolist()))
            sure = set(np.nonzero(rs[:, c] > 0.1 + 1e-5)[0].tolist())
            maybe = set(np.nonzero(rs[:, c] > 0.1 - 1e-5)[0].tolist())
            assert sure <= set(got) <= maybe
            for i, s in got.items():
                assert abs(s - rs[i, c]) <= 1e-5 + 1e-4 * abs(rs[i, c])


def test_candidate_overflow_is_reported(built_lib, anchors):
    B, ncls, hw = 1, 4, (64, 64)
    grids = [(2, 2), (4, 4), (8, 8)]
    heads = [np.full((B, gh, gw, 3, 5 + ncls), 5.0, np.float32) for gh, gw in grids]  # everything passes
    pp = PostProcess(B, grids, ncls, anchors, cand_cap=16)
    pp.set_image_shapes((64, 64))
    dev = [torch.from_numpy(t).cuda() for t in heads]
    pp.run([t.data_ptr() for t in dev], [3 * (ncls + 5)] * 3, 0.2, 0.5)
    with pytest.raises(_lib.YrError):
        pp.results()


@pytest.mark.parametrize("ih,iw,size", [(375, 500, (320, 320)), (600, 600, (416, 416)), (567, 850, (416, 416)),
                                        (100, 37, (96, 160))])
def test_letterbox_parity(built_lib, ih, iw, size):
    rng = np.random.default_rng(ih)
    img = rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
    ref = olb.letterbox_image(olb.u8_to_float(img), size)
    got = letterbox_image(torch.from_numpy(img).cuda(), size).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-6)


def test_yolo_head_matches_oracle(built_lib, anchors):
    """Public yolo_head (reference model.py:344): all four outputs + the calc_loss=True form."""
    from yoloret_b200.yolo3.model import yolo_head
    rng = np.random.default_rng(3)
    feats = rng.standard_normal((2, 5, 7, 3, 11)).astype(np.float32) * 2
    anc = anchors[[3, 4, 5]]
    ref = opp.yolo_head(feats, anc, (160, 224))
    got = yolo_head(torch.from_numpy(feats).cuda(), anc, (160, 224))
    for g, r in zip(got, ref):
        np.testing.assert_allclose(g.cpu().numpy(), np.asarray(r).reshape(g.shape), rtol=2e-6, atol=1e-7)
    grid, xy, wh, conf = yolo_head(torch.from_numpy(feats).cuda(), anc, (160, 224), calc_loss=True)
    assert grid.shape == (5, 7, 1, 2) and float(grid[4, 6, 0, 0]) == 6.0 and float(grid[4, 6, 0, 1]) == 4.0
    np.testing.assert_allclose(xy.cpu().numpy(), np.asarray(ref[0]).reshape(xy.shape), rtol=2e-6, atol=1e-7)
