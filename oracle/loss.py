"""torch-CPU restatement of the reference loss (TEST INFRASTRUCTURE).

Follows ``YoloLoss.call`` (reference code/yolo3/model.py:607-671), ``yolo_head(calc_loss=True)``
(model.py:344-369), ``do_giou_calculate`` (code/yolo3/utils.py:9-53) and the per-scale sum of
``AdvLossModel._compute_total_loss`` (code/yolo3/train.py:11-16).  Gradients come from torch
autograd (tf.maximum/minimum tie conventions differ only on exact ties, which seeded random
tests do not produce).  Parity is unpinned by the reference (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np
import torch

ANCHOR_MASK = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]


def divide_no_nan(a, b):
    return torch.where(b == 0, torch.zeros_like(a), a / torch.where(b == 0, torch.ones_like(b), b))


def do_giou_calculate(b1, b2, mode="giou"):
    zero = torch.zeros((), dtype=b1.dtype)
    b1_ymin, b1_xmin, b1_ymax, b1_xmax = torch.unbind(b1, -1)
    b2_ymin, b2_xmin, b2_ymax, b2_xmax = torch.unbind(b2, -1)
    b1_width = torch.maximum(zero, b1_xmax - b1_xmin)
    b1_height = torch.maximum(zero, b1_ymax - b1_ymin)
    b2_width = torch.maximum(zero, b2_xmax - b2_xmin)
    b2_height = torch.maximum(zero, b2_ymax - b2_ymin)
    b1_area = b1_width * b1_height
    b2_area = b2_width * b2_height
    intersect_ymin = torch.maximum(b1_ymin, b2_ymin)
    intersect_xmin = torch.maximum(b1_xmin, b2_xmin)
    intersect_ymax = torch.minimum(b1_ymax, b2_ymax)
    intersect_xmax = torch.minimum(b1_xmax, b2_xmax)
    intersect_width = torch.maximum(zero, intersect_xmax - intersect_xmin)
    intersect_height = torch.maximum(zero, intersect_ymax - intersect_ymin)
    intersect_area = intersect_width * intersect_height
    union_area = b1_area + b2_area - intersect_area
    iou = divide_no_nan(intersect_area, union_area)
    if mode == "iou":
        return iou
    enclose_ymin = torch.minimum(b1_ymin, b2_ymin)
    enclose_xmin = torch.minimum(b1_xmin, b2_xmin)
    enclose_ymax = torch.maximum(b1_ymax, b2_ymax)
    enclose_xmax = torch.maximum(b1_xmax, b2_xmax)
    enclose_width = torch.maximum(zero, enclose_xmax - enclose_xmin)
    enclose_height = torch.maximum(zero, enclose_ymax - enclose_ymin)
    enclose_area = enclose_width * enclose_height
    return iou - divide_no_nan(enclose_area - union_area, enclose_area)


def yolo_head_loss(feats, anchors, input_shape):
    """yolo_head(..., calc_loss=True): returns grid, box_xy, box_wh, box_confidence."""
    dt = feats.dtype
    A = len(anchors)
    anchors_tensor = torch.as_tensor(np.asarray(anchors), dtype=dt).reshape(1, 1, 1, A, 2)
    gh, gw = feats.shape[1:3]
    grid_y = torch.arange(gh).reshape(-1, 1, 1, 1).expand(gh, gw, 1, 1)
    grid_x = torch.arange(gw).reshape(1, -1, 1, 1).expand(gh, gw, 1, 1)
    grid = torch.cat([grid_x, grid_y], -1).to(dt)
    box_xy = (torch.sigmoid(feats[..., :2]) + grid) / torch.tensor([gw, gh], dtype=dt)
    box_wh = torch.exp(feats[..., 2:4]) * anchors_tensor / torch.tensor([input_shape[1], input_shape[0]], dtype=dt)
    return grid, box_xy, box_wh, torch.sigmoid(feats[..., 4:5])


def bce_with_logits(labels, logits):  # tf.nn.sigmoid_cross_entropy_with_logits
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-torch.abs(logits)))


def yolo_loss_scale(y_true, yolo_output, idx, anchors, num_scales=3, ignore_thresh=0.5):
    """One ``YoloLoss(idx, ...)`` call.  Returns (loss, (giou, conf, class, sum_ignore))."""
    grid_step = [32, 16, 8][idx]
    anchor = np.asarray(anchors, np.float32).reshape(-1, 2)[ANCHOR_MASK[-num_scales:][idx]]
    m = yolo_output.shape[0]
    object_mask = y_true[..., 4:5]
    true_class_probs = y_true[..., 5:]
    input_shape = (yolo_output.shape[1] * grid_step, yolo_output.shape[2] * grid_step)
    _grid, pred_xy, pred_wh, _conf = yolo_head_loss(yolo_output, anchor, input_shape)
    pred_max = torch.flip(pred_xy + pred_wh / 2.0, [-1])
    pred_min = torch.flip(pred_xy - pred_wh / 2.0, [-1])
    pred_box = torch.cat([pred_min, pred_max], -1)
    true_xy, true_wh = y_true[..., :2], y_true[..., 2:4]
    true_max = torch.flip(true_xy + true_wh / 2.0, [-1])
    true_min = torch.flip(true_xy - true_wh / 2.0, [-1])
    true_box = torch.clamp(torch.cat([true_min, true_max], -1), 0, 1)
    masked_true_box = true_box[object_mask[..., 0] != 0]          # whole batch, model.py:643
    if masked_true_box.shape[0] > 0:
        iou = do_giou_calculate(pred_box.unsqueeze(-2).detach(), masked_true_box.unsqueeze(0), mode="iou")
        best_iou = iou.max(dim=-1).values
    else:
        best_iou = torch.full(pred_box.shape[:-1], -float("inf"), dtype=pred_box.dtype)
    ignore_mask = (best_iou < ignore_thresh).to(y_true.dtype).unsqueeze(-1)
    ce = bce_with_logits(object_mask, yolo_output[..., 4:5])
    confidence_loss = object_mask * ce + (1 - object_mask) * ce * ignore_mask
    class_loss = object_mask * bce_with_logits(true_class_probs, yolo_output[..., 5:])
    class_loss = class_loss.sum() / m
    confidence_loss = confidence_loss.sum() / m
    giou = do_giou_calculate(pred_box, true_box)
    giou_loss = (object_mask * (1 - giou.unsqueeze(-1))).sum() / m
    loss = giou_loss + confidence_loss + class_loss
    return loss, (giou_loss.detach(), confidence_loss.detach(), class_loss.detach(), ignore_mask.sum())


def yolo_loss(y_trues, yolo_outputs, anchors, num_scales=3, ignore_thresh=0.5):
    total = 0
    for idx, (yt, yo) in enumerate(zip(y_trues, yolo_outputs)):
        total = total + yolo_loss_scale(yt, yo, idx, anchors, num_scales, ignore_thresh)[0]
    return total


def preprocess_true_boxes(true_boxes, input_shape, anchors, num_classes, num_scales=3):
    """y_true encoder, reference code/yolo3/utils.py:298-376, for ONE image.
    true_boxes [T,5] (xmin,ymin,xmax,ymax,class) in input pixels; returns list of [gh,gw,3,5+C]."""
    mask = ANCHOR_MASK[-num_scales:]
    true_boxes = np.array(true_boxes, dtype="float32")
    input_shape = np.array(input_shape, dtype="int32")
    anchors = np.asarray(anchors, np.float32).reshape(-1, 2)
    boxes_xy = (true_boxes[..., 0:2] + true_boxes[..., 2:4]) // 2
    boxes_wh = true_boxes[..., 2:4] - true_boxes[..., 0:2]
    true_boxes[..., 0:2] = boxes_xy / input_shape[::-1]
    true_boxes[..., 2:4] = boxes_wh / input_shape[::-1]
    grid_shapes = [np.round(input_shape / [32, 16, 8][l]).astype(np.int32) for l in range(num_scales)]
    y_true = [np.zeros((grid_shapes[l][0], grid_shapes[l][1], 3, 5 + num_classes), dtype="float32")
              for l in range(num_scales)]
    valid = boxes_wh[..., 0] > 0
    wh = boxes_wh[valid][:, None, :]
    amax = anchors[None] / 2.0
    inter = np.maximum(0, np.minimum(wh / 2, amax) - np.maximum(-wh / 2, -amax))
    inter = inter[..., 0] * inter[..., 1]
    union = wh[..., 0] * wh[..., 1] + anchors[None, :, 0] * anchors[None, :, 1] - inter
    best_anchor = np.argmax(inter / union, axis=-1)
    for t, n in enumerate(best_anchor):
        for l in range(num_scales):
            if n in mask[l]:
                i = np.floor(true_boxes[t, 0] * grid_shapes[l][1]).astype("int32")
                j = np.floor(true_boxes[t, 1] * grid_shapes[l][0]).astype("int32")
                k = mask[l].index(n)
                c = true_boxes[t, 4].astype("int32")
                y_true[l][j, i, k, 0:4] = true_boxes[t, 0:4]
                y_true[l][j, i, k, 4] = 1.0
                y_true[l][j, i, k, 5 + c] = 1.0
    return y_true
