"""GPU parity at the BENCHMARKED shapes (BASELINE.json configs[1..4]) and on the shipped EfficientNet-B3 checkpoint.

The op / network tests elsewhere run small shapes; tile counts, the persistent-CTA item split, resident-vs-streamed
weights and the depthwise item ring all take different branches at batch 64 / 416x416 (cfg2), batch 32 / 320x320
EfficientNet-lite0 (cfg3) and batch 16 / 608x608 MobileNetV2-1.4 (cfg4), so those exact workloads - same seeded
weights and images as ``bench.py`` - are checked here against the CPU oracle on sampled images of the batch.
"""
import argparse
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bench  # noqa: E402  (repo root: the workload definitions of the benchmark itself)
from yoloret_b200.netdef import NetDef  # noqa: E402
from yoloret_b200.weights import align_weights  # noqa: E402
from yoloret_b200.yolo import YOLO  # noqa: E402
from yoloret_b200.yolo3.model import YoloLoss  # noqa: E402
from oracle import verify as overify, loss as oloss  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _flags(tmp_path, a, weights, **extra):
    (tmp_path / "anchors.txt").write_text(", ".join("%d" % v for v in bench.ANCHORS))
    (tmp_path / "classes.txt").write_text("\n".join("class%d" % i for i in range(a.classes)) + "\n")
    f = {"backbone": a.model, "classes_path": str(tmp_path / "classes.txt"), "anchors_path": str(tmp_path / "anchors.txt"),
         "input_size": (a.size, a.size), "score": bench.SCORE, "nms": bench.IOU, "weights": weights, "batch": a.batch,
         "quiet": True}
    f.update(extra)
    return f


@pytest.mark.parametrize("workload,model,size,batch,sample", [
    ("cfg2", "mobilenetv2x75", 416, 64, [0, 21, 42, 63]),       # the headline: BASELINE.json configs[1]
    ("cfg3", "efficientnetlite0", 320, 32, [0, 10, 21, 31]),    # configs[2], per-GPU shard of batch 256 / 8
    ("cfg4", "mobilenetv2x14", 608, 16, [0, 5, 10, 15]),        # configs[3], per-GPU shard of batch 128 / 8
])
@pytest.mark.parametrize("u8", [False, True])
def test_benchmarked_workload_matches_oracle(built_lib, anchors, tmp_path, workload, model, size, batch, sample, u8):
    """The exact bench.py workload (seeded weights, seeded images, score 0.2 / IoU 0.5) through ``YOLO.detect_batch``:
    head logits, post-process on identical inputs and end-to-end detections of four images spread over the batch
    (first, last and two across the CTA work split) against the oracle.  ``u8``: the same batch quantised to uint8 and
    uploaded as bytes (what the end-to-end bench leg sends), against the oracle on ``u8 / 255``."""
    a = argparse.Namespace(model=model, size=size, classes=80, batch=batch)
    nd, weights = bench.make_weights(a)
    x = bench.make_inputs(a, batch, 1234)
    if u8:
        xu = bench.quantise_u8(x)
        x = xu.float() * np.float32(1.0 / 255.0)
    yolo = YOLO(_flags(tmp_path, a, weights, input_u8=u8))
    dets = yolo.detect_batch((xu if u8 else x).pin_memory())
    logits = [y.cpu().numpy() for y in yolo.engine.raw_outputs()]
    rep = overify.verify_batch(weights, x, model, 80, anchors, logits, dets, sample, bench.SCORE, bench.IOU)
    assert rep["detections"] > 0, "the sampled images produced no detections: the check would be vacuous"
    # a second pass through the captured graph and through the streaming API returns the same bits
    again = yolo.detect_batch((xu if u8 else x).pin_memory())
    stream = list(yolo.detect_stream(iter([(xu if u8 else x).pin_memory()] * 2)))
    for other in (again, stream[0], stream[1]):
        for p, q in zip(dets, other):
            assert all(np.array_equal(u, v) for u, v in zip(p, q))


def _b3_weights():
    z = np.load(os.path.join(GOLD, "b3_coco_weights.npz"))
    have = {k.replace("__", "/"): z[k] for k in z.files}
    return align_weights(have, NetDef("efficientnetb3", 80, (416, 416)).weight_shapes)


def test_b3_checkpoint_detect_image_golden(built_lib, tmp_path):
    """The shipped EfficientNet-B3 COCO checkpoint (reference code/checkpoints/efficientnetb3_416_coco.h5, loaded through
    the name re-alignment of code/yolo3/model.py:205-217's double build) on the 7 demo JPEGs at 416x416:
    ``YOLO.detect_image`` == the committed oracle detections; head logits of image 0 within 3e-4 x max."""
    g = np.load(os.path.join(GOLD, "demo_golden_b3.npz"))
    jp = np.load(os.path.join(GOLD, "demo_golden.npz"))
    (tmp_path / "anchors.txt").write_text(",  ".join("%g,%g" % (a, b) for a, b in g["anchors"]))
    (tmp_path / "classes.txt").write_text("\n".join(str(c) for c in g["classes"]) + "\n")
    from yoloret_b200.yolo3.enums import BACKBONE
    yolo = YOLO({"backbone": BACKBONE.EFFICIENTNETB3, "classes_path": str(tmp_path / "classes.txt"),
                 "anchors_path": str(tmp_path / "anchors.txt"), "input_size": (416, 416), "score": 0.3, "nms": 0.5,
                 "weights": _b3_weights(), "model": "golden-b3", "quiet": True})
    total = 0
    for i in range(len(g["names"])):
        data = jp["jpeg_%d" % i].tobytes()
        boxes, scores, cls = yolo.detect_image(io.BytesIO(data), draw=False)
        if i == 0:
            for s, y in enumerate(yolo.engine.raw_outputs()):
                r = g["y%d_0" % (s + 1)]
                err = float(np.abs(y.cpu().numpy() - r).max())
                assert err <= 3e-4 * max(1.0, float(np.abs(r).max())), (s, err)
        fb = yolo.engine.results(with_float_boxes=True)[0][3]
        overify.assert_detections_match((fb, scores, cls), (g["det_boxes_f_%d" % i], g["det_scores_%d" % i],
                                                            g["det_classes_%d" % i]), 0.3, 0.5, tol=1e-3,
                                        box_tol_px=1e-3 * float(max(g["shape_%d" % i])))
        assert np.array_equal(cls, g["det_classes_%d" % i]), (i, cls, g["det_classes_%d" % i])
        np.testing.assert_allclose(scores, g["det_scores_%d" % i], atol=1e-3)
        assert np.abs(boxes.astype(np.int64) - g["det_boxes_i_%d" % i]).max(initial=0) <= 1
        total += len(scores)
    assert total >= 20


def test_loss_at_cfg5_shape(built_lib, anchors):
    """YoloLoss forward + analytic gradient at the cfg5 per-GPU shape (BASELINE.json configs[4]: 416x416, COCO-80,
    batch 32, 8 boxes per image = 256 true boxes in the batch-wide ignore mask) against the fp64 autograd oracle."""
    from test_gpu_loss import _make
    yts, yos = _make(32, (416, 416), 80, anchors, 8, seed=5)
    for idx in range(3):
        ref_in = yos[idx].double().requires_grad_(True)
        ref, parts = oloss.yolo_loss_scale(yts[idx].double(), ref_in, idx, anchors)
        ref.backward()
        out = yos[idx].cuda().requires_grad_(True)
        L = YoloLoss(idx, anchors, 3, print_loss=False)
        loss = L(yts[idx].cuda(), out)
        loss.backward()
        got = L.last_parts.cpu().numpy()
        np.testing.assert_allclose(got[:3], [float(p) for p in parts[:3]], rtol=2e-4, atol=1e-5)
        assert got[3] == float(parts[3])
        np.testing.assert_allclose(out.grad.cpu().numpy(), ref_in.grad.float().numpy(), rtol=2e-3, atol=2e-6)


def test_engine_on_second_device_context(built_lib, anchors):
    """An engine constructed for the current device keeps working when called under another current-device context
    manager (function attributes / SM count / streams are per device)."""
    hw, ncls, B = (96, 96), 20, 2
    nd = NetDef("mobilenetv2x75", ncls, hw)
    from yoloret_b200.weights import synthetic_weights
    from yoloret_b200.yolo3.model import yolov3_body
    w = synthetic_weights(nd.weight_shapes, ncls, seed=7)
    x = torch.rand(B, hw[0], hw[1], 3, generator=torch.Generator().manual_seed(1))
    m0 = yolov3_body((B, hw[0], hw[1], 3), "mobilenetv2x75", 3, num_classes=ncls, device="cuda:0").set_weights(w, anchors)
    y0 = m0(x.cuda(0))
    if torch.cuda.device_count() > 1:
        m1 = yolov3_body((B, hw[0], hw[1], 3), "mobilenetv2x75", 3, num_classes=ncls, device="cuda:1").set_weights(w, anchors)
        y1 = m1(x.cuda(1))  # current device is still cuda:0
        for a, b in zip(y0, y1):
            assert torch.equal(a.cpu(), b.cpu())
    with pytest.raises(ValueError):
        yolov3_body((B, hw[0], hw[1], 3), "mobilenetv2x75", 4, num_classes=ncls)
