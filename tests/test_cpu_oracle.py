"""CPU tests of the oracle itself: against the committed golden fixtures (generated from the
reference's shipped checkpoint + demo images, tests/golden/make_golden.py), against the survey's
independently probed detections, and restatement-vs-restatement (C vs Python NMS)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import graph as ograph, postprocess as opp, letterbox as olb, loss as oloss

GOLD = os.path.join(os.path.dirname(__file__), "golden")
VOC = ["aeroplane", "bicycle", "bird", "boat", "bottle", "bus", "car", "cat", "chair", "cow", "diningtable",
       "dog", "horse", "motorbike", "person", "pottedplant", "sheep", "sofa", "train", "tvmonitor"]
# detections of an independent throw-away probe at 320x320, score 0.3, IoU 0.5
SURVEY_PINS = {
    "2008_003205.jpg": [("bicycle", .990, [98, 120, 336, 339]), ("car", .954, [53, 128, 75, 162]),
                        ("person", .950, [70, 160, 338, 345])],
    "2011_006155.jpg": [("person", 1.000, [58, 183, 292, 262]), ("bicycle", .998, [177, 178, 305, 258])],
    "2011_001694.jpg": [("bird", .951, None)],
    "2011_002558.jpg": [("boat", .987, None)],
}


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "demo_golden.npz"))


@pytest.fixture(scope="module")
def voc_weights():
    z = np.load(os.path.join(GOLD, "voc_mbv2x75_weights.npz"))
    return {k.replace("__", "/"): z[k] for k in z.files}


def test_checkpoint_matches_graph_spec(voc_weights):
    """The shipped checkpoint loads by Keras layer name into the restated graph: every shape matches
    and the parameter count is the survey's 1 887 687."""
    spec = ograph.weight_spec("mobilenetv2x75", 20)
    assert set(spec) == set(voc_weights)
    for k, shp in spec.items():
        assert tuple(voc_weights[k].shape) == tuple(shp), k
    assert sum(int(v.size) for v in voc_weights.values()) == 1887687


def test_oracle_reproduces_golden_and_survey_pins(gold, voc_weights):
    names = [str(n) for n in gold["names"]]
    readable = json.load(open(os.path.join(GOLD, "demo_detections.json")))
    for i, name in enumerate(names[:3]):  # 3 images keep the CPU suite short; all 7 run on the GPU path
        img = olb.decode_image_u8(gold["jpeg_%d" % i].tobytes())
        x = olb.letterbox_image(olb.u8_to_float(img), (320, 320))
        if i == 0:
            np.testing.assert_array_equal(x, gold["letterbox_0"])
        ys = [y.numpy() for y in ograph.forward(voc_weights, x[None], "mobilenetv2x75", 20)]
        if i < 2:
            for s in range(3):
                np.testing.assert_allclose(ys[s], gold["y%d_%d" % (s + 1, i)], atol=2e-5)
        b, sc, cl = opp.yolo_eval(ys, gold["anchors"], 3, 20, img.shape[:2], score_threshold=0.3, iou_threshold=0.5)
        np.testing.assert_array_equal(cl, gold["det_classes_%d" % i])
        np.testing.assert_allclose(sc, gold["det_scores_%d" % i], atol=1e-5)
        assert np.abs(b - gold["det_boxes_i_%d" % i]).max(initial=0) <= 1
        got = {(VOC[c]): (float(s), bb.tolist()) for bb, s, c in zip(b, sc, cl)}
        assert [d["cls"] for d in readable[name]] == [VOC[c] for c in cl]
        for cls, score, box in SURVEY_PINS.get(name, []):
            assert cls in got and abs(got[cls][0] - score) < 2e-3, (name, cls, got)
            if box is not None:
                assert np.abs(np.array(got[cls][1]) - np.array(box)).max() <= 1


def test_all_golden_detections_match_survey(gold):
    """Every survey-listed detection is in the committed golden file (cheap: no network run)."""
    readable = json.load(open(os.path.join(GOLD, "demo_detections.json")))
    for name, pins in SURVEY_PINS.items():
        got = {d["cls"]: d for d in readable[name]}
        for cls, score, box in pins:
            assert abs(got[cls]["score"] - score) < 2e-3
            if box is not None:
                assert np.abs(np.array(got[cls]["box"]) - np.array(box)).max() <= 1


def _rand_boxes(rng, n, degenerate=False):
    c = rng.uniform(0, 100, (n, 2))
    wh = rng.uniform(1, 40, (n, 2))
    b = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    if degenerate and n > 4:
        b[1] = b[0]            # exact duplicate
        b[2, 2:] = b[2, :2]    # zero area
        b[3] = b[3][[2, 3, 0, 1]]  # swapped corners
    return b


@pytest.mark.parametrize("n,seed,ties", [(0, 0, False), (1, 1, False), (50, 2, False), (300, 3, True), (700, 4, True)])
def test_nms_c_equals_python(n, seed, ties):
    rng = np.random.default_rng(seed)
    boxes = _rand_boxes(rng, n, degenerate=True)
    scores = rng.uniform(0, 1, n).astype(np.float32)
    if ties and n:
        scores = np.round(scores, 1)  # many exact score ties -> lower index first
    for thr, iou, k in ((0.2, 0.5, 20), (0.0, 0.3, 5), (0.9, 0.5, 20), (0.2, 0.0, 1000)):
        a = opp.nms_c(boxes, scores, k, iou, thr)
        b = opp.nms_python(boxes, scores, k, iou, thr)
        np.testing.assert_array_equal(a, b)
        assert len(a) <= k and np.all(scores[a] > thr)
        assert np.all(np.diff(scores[a]) <= 0)  # selection order is score-descending


def test_nms_semantics_known_answers():
    # two heavy overlaps + one disjoint; strict '>' on both thresholds
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [0, 0, 10, 5], [20, 20, 30, 30]], np.float32)
    scores = np.array([0.9, 0.9, 0.8, 0.5], np.float32)
    np.testing.assert_array_equal(opp.nms_c(boxes, scores, 20, 0.5, 0.2), [0, 2, 3])   # IoU(0,2)=0.5 is NOT > 0.5: kept
    np.testing.assert_array_equal(opp.nms_c(boxes, scores, 20, 0.5, 0.2), opp.nms_python(boxes, scores, 20, 0.5, 0.2))
    np.testing.assert_array_equal(opp.nms_c(boxes, scores, 20, 0.49, 0.2), [0, 3])
    np.testing.assert_array_equal(opp.nms_c(boxes, scores, 20, 0.5, 0.5), [0, 2])      # score 0.5 is not > 0.5
    np.testing.assert_array_equal(opp.nms_c(boxes, scores, 1, 0.5, 0.2), [0])          # tie -> lower index


def test_decode_identity_letterbox():
    """image_shape == input_shape makes yolo_correct_boxes the identity mapping."""
    rng = np.random.default_rng(0)
    anchors = np.array([10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326], np.float32).reshape(-1, 2)
    ys = [rng.standard_normal((1, 64 // s, 96 // s, 3, 9)).astype(np.float32) for s in (32, 16, 8)]
    boxes, scores = opp.decode_all(ys, anchors, 3, 4, (64, 96))
    assert boxes.shape == (3 * (2 * 3 + 4 * 6 + 8 * 12), 4) and scores.shape == (boxes.shape[0], 4)
    assert boxes.min() >= 0 and boxes[:, [0, 2]].max() <= 64 and boxes[:, [1, 3]].max() <= 96
    xy, wh, conf, cls = opp.yolo_head(ys[0], anchors[[6, 7, 8]], (64, 96))
    cy = (boxes[:18, 0] + boxes[:18, 2]) / 2
    inside = (boxes[:18, 0] > 0) & (boxes[:18, 2] < 64)
    np.testing.assert_allclose(cy[inside], (xy[..., 1].reshape(-1) * 64)[inside], rtol=1e-5, atol=1e-4)


def test_loss_oracle_properties():
    """YoloLoss restatement: zero-object batches give ignore-mask==1 everywhere (max over an empty
    axis is -inf, reference model.py:647-649) and only the confidence term; gradients are finite."""
    anchors = np.array([10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326], np.float32).reshape(-1, 2)
    g = torch.Generator().manual_seed(0)
    yo = torch.randn(2, 4, 4, 3, 9, generator=g, dtype=torch.float64, requires_grad=True)
    yt = torch.zeros(2, 4, 4, 3, 9, dtype=torch.float64)
    loss, parts = oloss.yolo_loss_scale(yt, yo, 0, anchors)
    assert float(parts[0]) == 0 and float(parts[2]) == 0 and float(parts[3]) == 2 * 4 * 4 * 3
    expect = torch.nn.functional.softplus(yo[..., 4]).sum() / 2  # BCE(0, x) = softplus(x)
    np.testing.assert_allclose(float(loss.detach()), float(expect.detach()), rtol=1e-12)
    loss.backward()
    assert torch.isfinite(yo.grad).all()
    # do_giou_calculate: identical boxes -> 1, disjoint unit boxes -> -(enclose-union)/enclose
    b = torch.tensor([[0.1, 0.1, 0.5, 0.5]], dtype=torch.float64)
    assert abs(float(oloss.do_giou_calculate(b, b)) - 1.0) < 1e-12
    b2 = torch.tensor([[0.6, 0.6, 0.9, 0.9]], dtype=torch.float64)
    enclose, union = 0.8 * 0.8, 0.16 + 0.09
    assert abs(float(oloss.do_giou_calculate(b, b2)) + (enclose - union) / enclose) < 1e-12


# ---- EfficientNet-B3 (SE blocks, 5x5 depthwise, Swish): the shipped COCO checkpoint ---------------------------------
def test_b3_checkpoint_realigns_and_oracle_reproduces_golden(gold):
    """reference code/checkpoints/efficientnetb3_416_coco.h5 as stored (tests/golden/b3_coco_weights.npz keeps the
    checkpoint's own Keras names).  The reference builds the EfficientNet backbone twice (code/yolo3/model.py:205-217),
    which shifts Keras' auto-numbering: 141 of the 603 names differ from a single graph build, so loading goes through
    ``weights.align_weights`` (per layer kind, in creation order).  Pins: every array is consumed exactly once, shapes
    and the parameter count (8 868 749) match, and the oracle on demo image 0 reproduces the committed logits and the
    confident COCO detections (person / bicycle / car) of tests/golden/demo_detections_b3.json."""
    from yoloret_b200.netdef import NetDef
    from yoloret_b200.weights import align_weights
    z = np.load(os.path.join(GOLD, "b3_coco_weights.npz"))
    have = {k.replace("__", "/"): z[k] for k in z.files}
    spec = NetDef("efficientnetb3", 80, (416, 416)).weight_shapes
    assert dict(spec) == dict(ograph.weight_spec("efficientnetb3", 80))   # product and oracle agree on the graph
    assert len(have) == len(spec) == 603 and sum(1 for k in spec if k not in have) == 141
    w = align_weights(have, spec)
    assert list(w) == list(spec) and sum(int(v.size) for v in w.values()) == 8868749
    used = {id(v) for v in w.values()}
    assert len(used) == 603 and used == {id(v) for v in have.values()}    # a bijection: nothing reused or dropped
    # names that exist on both sides but hold DIFFERENT layers must have been remapped, not taken by name
    assert w["batch_normalization_78/gamma"] is have["batch_normalization_156/gamma"]
    with pytest.raises(KeyError):
        align_weights({k: v for k, v in have.items() if k != "conv2d_10/kernel"}, spec)

    g = np.load(os.path.join(GOLD, "demo_golden_b3.npz"))
    img = olb.decode_image_u8(gold["jpeg_0"].tobytes())
    x = olb.letterbox_image(olb.u8_to_float(img), (416, 416))
    ys = [y.numpy() for y in ograph.forward(w, x[None], "efficientnetb3", 80)]
    for s in range(3):
        np.testing.assert_allclose(ys[s], g["y%d_0" % (s + 1)], atol=5e-5)
    b, sc, cl = opp.yolo_eval(ys, g["anchors"], 3, 80, img.shape[:2], score_threshold=0.3, iou_threshold=0.5)
    np.testing.assert_array_equal(cl, g["det_classes_0"])
    np.testing.assert_allclose(sc, g["det_scores_0"], atol=1e-5)
    assert np.abs(b.astype(np.int64) - g["det_boxes_i_0"]).max() <= 1
    names = [str(c) for c in g["classes"]]
    best = {}
    for c, s_ in zip(cl, sc):
        best[names[c]] = max(best.get(names[c], 0.0), float(s_))
    assert best["person"] > 0.99 and best["bicycle"] > 0.8 and best["car"] > 0.9
