"""numpy restatement of the reference post-process (TEST INFRASTRUCTURE).

Follows:
  * ``yolo_head``               reference code/yolo3/model.py:344-371
  * ``yolo_correct_boxes``      code/yolo3/model.py:374-399
  * ``yolo_boxes_and_scores``   code/yolo3/model.py:402-428
  * ``yolo_eval``               code/yolo3/model.py:431-491
  * ``tf.image.non_max_suppression`` (NonMaxSuppressionV3 CPU kernel; third-party,
    un-vendored, unpinned - published algorithm restated):
    candidates score > score_threshold (strict); processed by score desc, ties
    -> lower box index; a candidate is kept unless IoU > iou_threshold (strict)
    with an already kept box (checked newest -> oldest); stop at max_output_size.
    IoU: corners normalised with min/max, 0 if either area <= 0,
    inter / (a_i + a_j - inter) in fp32.

All arithmetic is float32, one numpy op per TF op, in the reference's order.
Parity is unpinned by the reference (see oracle/__init__.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

f32 = np.float32
ANCHOR_MASK = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]


def sigmoid(x):
    x = np.asarray(x, dtype=f32)
    return (f32(1.0) / (f32(1.0) + np.exp(-x, dtype=f32))).astype(f32)


def yolo_head(feats, anchors, input_shape, calc_loss=False):
    """feats [B,H,W,A,5+C] f32; anchors [A,2] (w,h) px; input_shape (h,w)."""
    feats = np.asarray(feats, dtype=f32)
    anchors = np.asarray(anchors, dtype=f32)
    num_anchors = len(anchors)
    anchors_tensor = anchors.reshape(1, 1, 1, num_anchors, 2)
    gh, gw = feats.shape[1:3]
    grid_y = np.tile(np.arange(gh).reshape(-1, 1, 1, 1), [1, gw, 1, 1])
    grid_x = np.tile(np.arange(gw).reshape(1, -1, 1, 1), [gh, 1, 1, 1])
    grid = np.concatenate([grid_x, grid_y], -1).astype(f32)
    grid_wh = np.array([gw, gh], dtype=f32)
    in_wh = np.array([input_shape[1], input_shape[0]], dtype=f32)
    box_xy = ((sigmoid(feats[..., :2]) + grid) / grid_wh).astype(f32)
    box_wh = (np.exp(feats[..., 2:4], dtype=f32) * anchors_tensor / in_wh).astype(f32)
    box_confidence = sigmoid(feats[..., 4:5])
    if calc_loss:
        return grid, box_xy, box_wh, box_confidence
    box_class_probs = sigmoid(feats[..., 5:])
    return box_xy, box_wh, box_confidence, box_class_probs


def yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape):
    box_yx = box_xy[..., ::-1]
    box_hw = box_wh[..., ::-1]
    input_shape = np.asarray(input_shape, dtype=f32)
    image_shape = np.asarray(image_shape, dtype=f32)
    max_shape = np.maximum(image_shape[0], image_shape[1])
    ratio = (image_shape / max_shape).astype(f32)
    boxed_shape = (input_shape * ratio).astype(f32)
    offset = ((input_shape - boxed_shape) / f32(2.0)).astype(f32)
    scale = (image_shape / boxed_shape).astype(f32)
    box_yx = ((box_yx * input_shape - offset) * scale).astype(f32)
    box_hw = (box_hw * (input_shape * scale).astype(f32)).astype(f32)
    box_mins = (box_yx - (box_hw / f32(2.0))).astype(f32)
    box_maxes = (box_yx + (box_hw / f32(2.0))).astype(f32)
    boxes = np.concatenate([
        np.clip(box_mins[..., 0:1], f32(0), image_shape[0]),
        np.clip(box_mins[..., 1:2], f32(0), image_shape[1]),
        np.clip(box_maxes[..., 0:1], f32(0), image_shape[0]),
        np.clip(box_maxes[..., 1:2], f32(0), image_shape[1]),
    ], -1).astype(f32)
    return boxes


def yolo_boxes_and_scores(feats, anchors, num_classes, input_shape, image_shape):
    box_xy, box_wh, box_confidence, box_class_probs = yolo_head(feats, anchors, input_shape)
    boxes = yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape).reshape(-1, 4)
    box_scores = (box_confidence * box_class_probs).astype(f32).reshape(-1, num_classes)
    return boxes, box_scores


def decode_all(yolo_outputs, anchors, num_scales, num_classes, image_shape):
    """The part of yolo_eval before NMS (model.py:443-469): all boxes / scores,
    scales concatenated in the order s32, s16, s8."""
    anchors = np.asarray(anchors, dtype=f32)
    mask = ANCHOR_MASK[-num_scales:]
    input_shape = np.array(yolo_outputs[0].shape[1:3]) * 32
    bs, ss = [], []
    for l in range(num_scales):
        b, s = yolo_boxes_and_scores(yolo_outputs[l], anchors[mask[l]], num_classes, input_shape, image_shape)
        bs.append(b)
        ss.append(s)
    return np.concatenate(bs, 0), np.concatenate(ss, 0)


# --------------------------------------------------------------------------
# NMS: pure-Python restatement (small cases) and plain-C restatement (oracle/nms_ref.c)
# --------------------------------------------------------------------------
def _iou_tf(boxes, i, j):
    bi, bj = boxes[i], boxes[j]
    ymin_i, xmin_i = min(bi[0], bi[2]), min(bi[1], bi[3])
    ymax_i, xmax_i = max(bi[0], bi[2]), max(bi[1], bi[3])
    ymin_j, xmin_j = min(bj[0], bj[2]), min(bj[1], bj[3])
    ymax_j, xmax_j = max(bj[0], bj[2]), max(bj[1], bj[3])
    area_i = f32(f32(ymax_i - ymin_i) * f32(xmax_i - xmin_i))
    area_j = f32(f32(ymax_j - ymin_j) * f32(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return f32(0.0)
    iy0, ix0 = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iy1, ix1 = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = f32(max(f32(iy1 - iy0), f32(0.0)) * max(f32(ix1 - ix0), f32(0.0)))
    return f32(inter / f32(f32(area_i + area_j) - inter))


def nms_python(boxes, scores, max_output_size, iou_threshold, score_threshold) -> np.ndarray:
    boxes = np.asarray(boxes, dtype=f32)
    scores = np.asarray(scores, dtype=f32)
    cand = [i for i in range(len(scores)) if scores[i] > f32(score_threshold)]
    cand.sort(key=lambda i: (-float(scores[i]), i))
    sel: List[int] = []
    thr = f32(iou_threshold)
    for i in cand:
        if len(sel) >= max_output_size:
            break
        keep = True
        for j in reversed(sel):
            if _iou_tf(boxes, i, j) > thr:
                keep = False
                break
        if keep:
            sel.append(i)
    return np.asarray(sel, dtype=np.int32)


_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force: bool = False) -> str:
    """Compile oracle/nms_ref.c -> oracle/_build/libnms_ref.so with gcc."""
    src = os.path.join(_HERE, "nms_ref.c")
    out_dir = os.path.join(_HERE, "_build")
    out = os.path.join(out_dir, "libnms_ref.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-o", out, src, "-lm"])
    return out


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
        _LIB.nms_ref.restype = ctypes.c_int
        _LIB.nms_ref.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_float, ctypes.c_float, ctypes.c_void_p]
    return _LIB


def nms_c(boxes, scores, max_output_size, iou_threshold, score_threshold, score_stride: int = 1) -> np.ndarray:
    boxes = np.ascontiguousarray(boxes, dtype=f32)
    scores = np.asarray(scores, dtype=f32)
    n = boxes.shape[0]
    if score_stride == 1:
        scores = np.ascontiguousarray(scores)
    out = np.empty(max(1, max_output_size), dtype=np.int32)
    k = _lib().nms_ref(boxes.ctypes.data, scores.ctypes.data, n, score_stride, max_output_size,
                       ctypes.c_float(iou_threshold), ctypes.c_float(score_threshold), out.ctypes.data)
    return out[:k].copy()


def yolo_eval(yolo_outputs, anchors, num_scales, num_classes, image_shape, max_boxes=20,
              score_threshold=.6, iou_threshold=.5, nms=nms_c, return_float_boxes=False):
    """Reference ``yolo_eval`` for ONE image (the reference is batch-1).
    Returns (boxes int32 [N,4] (ymin,xmin,ymax,xmax), scores f32 [N], classes int32 [N])."""
    boxes, box_scores = decode_all(yolo_outputs, anchors, num_scales, num_classes, image_shape)
    boxes_, scores_, classes_ = [], [], []
    for c in range(num_classes):
        col = np.ascontiguousarray(box_scores[:, c])
        idx = nms(boxes, col, max_boxes, iou_threshold, score_threshold)
        boxes_.append(boxes[idx])
        scores_.append(col[idx])
        classes_.append(np.full(len(idx), c, dtype=np.int32))
    fb = np.concatenate(boxes_, 0).astype(f32).reshape(-1, 4)
    sc = np.concatenate(scores_, 0).astype(f32)
    cl = np.concatenate(classes_, 0).astype(np.int32)
    ib = fb.astype(np.int32)  # tf.cast(float->int32) truncates toward zero
    if return_float_boxes:
        return ib, sc, cl, fb
    return ib, sc, cl
