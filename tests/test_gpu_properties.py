"""Property tests (hypothesis, derandomised so every run sees the same examples) of the post-process and loss kernels
against the CPU oracle: extreme logits (exp overflow to +inf, saturated sigmoids), score ties, degenerate boxes,
ragged candidate counts, thresholds on both sides of the data.  SURVEY.md section 4 asks for these on top of the
seeded cases in test_gpu_postprocess.py / test_gpu_loss.py."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st, HealthCheck

pytestmark = pytest.mark.gpu

from yoloret_b200.yolo3.model import YoloEval, YoloLoss, yolo_head  # noqa: E402
from oracle import postprocess as opp, loss as oloss  # noqa: E402
from test_gpu_postprocess import _run_nms  # noqa: E402

ANCHORS = np.array([10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326], np.float32).reshape(-1, 2)
SET = dict(deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


def _heads(rng, B, grids, ncls, wh_scale, obj_mu, extreme):
    out = []
    for gh, gw in grids:
        t = np.empty((B, gh, gw, 3, 5 + ncls), np.float32)
        t[..., 0:2] = rng.normal(0, 2, t[..., 0:2].shape)
        t[..., 2:4] = rng.normal(0, wh_scale, t[..., 2:4].shape)
        t[..., 4] = rng.normal(obj_mu, 2, t[..., 4].shape)
        t[..., 5:] = rng.normal(-2, 2, t[..., 5:].shape)
        if extreme:  # saturate some cells: exp(t_wh) -> +inf, sigmoid -> exactly 0 / 1
            m = rng.random(t.shape[:4]) < 0.02
            t[..., 2][m] = 100.0
            t[..., 3][rng.random(t.shape[:4]) < 0.02] = -120.0
            t[..., 0][rng.random(t.shape[:4]) < 0.02] = 95.0
            t[..., 4][rng.random(t.shape[:4]) < 0.02] = 60.0
            t[..., 5][rng.random(t.shape[:4]) < 0.02] = -110.0
        out.append(t)
    return out


@settings(max_examples=20, **SET)
@given(seed=st.integers(0, 2 ** 20), ncls=st.sampled_from([1, 3, 20, 80]), gh=st.integers(1, 4), gw=st.integers(1, 5),
       thr=st.sampled_from([0.0, 0.05, 0.2, 0.6, 0.999]), iou=st.sampled_from([0.0, 0.3, 0.5, 0.9]),
       wh_scale=st.sampled_from([0.3, 1.0, 3.0]), extreme=st.booleans())
def test_yolo_eval_property(built_lib, seed, ncls, gh, gw, thr, iou, wh_scale, extreme):
    """yolo_eval (reference code/yolo3/model.py:431-491) on arbitrary logits: same detections in the same order."""
    rng = np.random.default_rng(seed)
    B = 2
    grids = [(gh, gw), (2 * gh, 2 * gw), (4 * gh, 4 * gw)]
    heads = _heads(rng, B, grids, ncls, wh_scale, obj_mu=-1.0, extreme=extreme)
    shapes = np.array([[rng.integers(50, 700), rng.integers(50, 700)] for _ in range(B)], np.float32)
    dev = [torch.from_numpy(t).cuda() for t in heads]
    got = YoloEval(ANCHORS, 3, ncls, score_threshold=thr, iou_threshold=iou)(dev, shapes)
    for b in range(B):
        rb, rs, rc = opp.yolo_eval([t[b:b + 1] for t in heads], ANCHORS, 3, ncls, shapes[b], score_threshold=thr,
                                   iou_threshold=iou)
        gb, gs, gc = (x.cpu().numpy() for x in got[b])
        assert len(gs) == len(rs), (len(gs), len(rs))
        assert np.array_equal(gc, rc)
        np.testing.assert_allclose(gs, rs, rtol=1e-3, atol=1e-6)
        assert np.abs(gb.astype(np.int64) - rb.astype(np.int64)).max(initial=0) <= 1


@settings(max_examples=25, **SET)
@given(seed=st.integers(0, 2 ** 20), T=st.sampled_from([1, 2, 31, 32, 33, 257, 1025, 3000]),
       quant=st.sampled_from([0, 2, 7, 64]), max_boxes=st.sampled_from([1, 20, 50]),
       iou=st.sampled_from([0.0, 0.5, 0.75, 1.0]), thr=st.sampled_from([0.0, 0.3, 0.9]), span=st.sampled_from([5.0, 100.0]))
def test_nms_property(built_lib, seed, T, quant, max_boxes, iou, thr, span):
    """Class-wise greedy NMS (tf.image.non_max_suppression semantics): selected indices bit-exact for any mix of
    score ties, duplicates, zero-area / inverted / huge boxes and thresholds."""
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(0, span, (T, 2)).astype(np.float32)
    wh = rng.uniform(0, 0.4 * span, (T, 2)).astype(np.float32)
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], 1).astype(np.float32)
    boxes[::7, 2:] = boxes[::7, :2]
    boxes[3::11] = boxes[3::11][:, [2, 3, 0, 1]]
    boxes[1::5] = boxes[0::5][:len(boxes[1::5])]
    boxes[2::13] *= np.float32(1e6)
    Cn = 3
    scores = rng.uniform(0, 1, (Cn, T)).astype(np.float32)
    if quant:
        scores = (np.floor(scores * quant) / quant).astype(np.float32)
    got, _, _ = _run_nms(boxes, scores, max_boxes, iou, thr)
    for c in range(Cn):
        ref = opp.nms_c(boxes, scores[c], max_boxes, iou, thr)
        assert np.array_equal(got[c], ref), (c, got[c], ref)
        assert len(ref) <= max_boxes and len(set(ref.tolist())) == len(ref)


@settings(max_examples=12, **SET)
@given(seed=st.integers(0, 2 ** 20), ncls=st.sampled_from([1, 4, 20]), gh=st.integers(1, 3), gw=st.integers(1, 3),
       wh_scale=st.sampled_from([0.5, 2.0]), extreme=st.booleans())
def test_yolo_head_property(built_lib, seed, ncls, gh, gw, wh_scale, extreme):
    """yolo_head (reference code/yolo3/model.py:344-371) element-wise, incl. +inf box sizes from exp overflow."""
    rng = np.random.default_rng(seed)
    t = _heads(rng, 2, [(gh, gw)], ncls, wh_scale, -1.0, extreme)[0]
    anc = ANCHORS[[6, 7, 8]]
    got = yolo_head(torch.from_numpy(t).cuda(), anc, (gh * 32, gw * 32))
    ref = opp.yolo_head(t, anc, (gh * 32, gw * 32))
    for g, r in zip(got, ref):
        g = g.cpu().numpy()
        assert np.array_equal(np.isinf(g), np.isinf(r))
        fin = np.isfinite(r)
        np.testing.assert_allclose(g[fin], r[fin], rtol=2e-6, atol=1e-7)


@settings(max_examples=10, **SET)
@given(seed=st.integers(0, 2 ** 20), ncls=st.sampled_from([1, 4, 20]), n_boxes=st.integers(0, 6), idx=st.integers(0, 2),
       scale=st.sampled_from([0.5, 1.5, 6.0]), B=st.integers(1, 3))
def test_loss_property(built_lib, seed, ncls, n_boxes, idx, scale, B):
    """YoloLoss (reference code/yolo3/model.py:585-671) value and analytic gradient vs fp64 autograd for logits from
    timid to saturated (|logit| up to ~25: BCE in its stable form, GIoU of huge / tiny predicted boxes)."""
    from test_gpu_loss import _make
    hw = (96, 128)
    yts, yos = _make(B, hw, ncls, ANCHORS, n_boxes, seed=seed % 1000)
    yo = yos[idx] * (scale / 1.5)
    yo[..., 2:4] = torch.clamp(yo[..., 2:4], -8, 8)  # exp(t_wh) stays finite: the reference would yield NaN losses beyond
    ref_in = yo.double().requires_grad_(True)
    ref, parts = oloss.yolo_loss_scale(yts[idx].double(), ref_in, idx, ANCHORS)
    ref.backward()
    out = yo.cuda().requires_grad_(True)
    L = YoloLoss(idx, ANCHORS, 3, print_loss=False)
    loss = L(yts[idx].cuda(), out)
    loss.backward()
    got = L.last_parts.cpu().numpy()
    np.testing.assert_allclose(got[:3], [float(p) for p in parts[:3]], rtol=3e-4, atol=1e-5)
    assert got[3] == float(parts[3])
    np.testing.assert_allclose(out.grad.cpu().numpy(), ref_in.grad.float().numpy(), rtol=3e-3, atol=3e-6)
