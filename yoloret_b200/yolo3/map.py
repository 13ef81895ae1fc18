# -*- coding: utf-8 -*-
"""Host-side mirror of reference code/yolo3/map.py: VOC mAP of a ``YOLO`` engine over an annotation list.

  MAPCallback(glob_path, input_shape, class_names, iou=.5, batch_size=1)   reference map.py:225-235
      .set_model(yolo_model) / .calculate_aps() -> {class: AP}              reference map.py:75-222
  calculate_map(yolo, glob)                                                 reference code/yolo.py:397-405

The reference feeds one image at a time through ``YoloModel.call`` (map.py:107-111).  Here the list is read
with the reference's text format (``path xmin ymin xmax ymax label ...``, map.py:55-73; data.py:71-83) and
every image goes through the same ``yolo_model([bytes])`` call; the AP arithmetic (sort by score, greedy IoU
match with the VOC "+1 pixel" convention, monotone precision envelope) is vectorised numpy on the host - it is
O(detections x ground truths) of float64 work the reference also does on the CPU, not part of the GPU hot path.
TFRecord inputs (map.py:34-53) need TensorFlow's reader and are not supported.
"""
from __future__ import annotations

import glob as _glob
import os
from timeit import default_timer as timer
from typing import Dict, List, Tuple

import numpy as np


def voc_ap(rec: np.ndarray, prec: np.ndarray) -> float:
    """reference map.py:16-32."""
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]           # precision envelope
    i = np.nonzero(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def parse_text_line(line: str) -> Tuple[str, np.ndarray]:
    """``path xmin ymin xmax ymax label ...`` -> (path, float32 [n,5] xmin,ymin,xmax,ymax,label); map.py:55-73."""
    values = line.strip().split(' ')
    if (len(values) - 1) % 5:
        raise ValueError("annotation line of %s does not hold groups of 5 numbers" % values[0])
    return values[0], np.asarray(values[1:], dtype=np.float32).reshape(-1, 5)


def class_aps(pred: np.ndarray, true_res: Dict[int, np.ndarray], num_classes: int, iou_thr: float = 0.5) -> Dict[int, float]:
    """``pred``: [n,7] rows (image idx, class, score, left, top, right, bottom) in detection order;
    ``true_res``: {image idx: [m,5] xmin,ymin,xmax,ymax,label}.  reference map.py:157-222."""
    APs: Dict[int, float] = {}
    pred = np.asarray(pred, dtype=np.float64).reshape(-1, 7)
    for cls in range(num_classes):
        pc = pred[pred[:, 1] == cls]
        if len(pc) == 0:
            APs[cls] = 0
            continue
        gts = {idx: np.asarray(t, dtype=np.float64).reshape(-1, 5) for idx, t in true_res.items()}
        gts = {idx: t[t[:, 4] == cls][:, :4] for idx, t in gts.items()}
        npos = sum(len(t) for t in gts.values())
        det = {idx: np.zeros(len(t), bool) for idx, t in gts.items()}
        order = np.argsort(-pc[:, 2])                       # same (unstable quicksort) call as the reference
        pc = pc[order]
        tp = np.zeros(len(pc))
        fp = np.zeros(len(pc))
        for j, row in enumerate(pc):
            idx = int(row[0])
            bb = row[3:7]
            g = gts[idx]
            ovmax, jmax = -np.inf, -1
            if g.size > 0:
                iw = np.maximum(np.minimum(g[:, 2], bb[2]) - np.maximum(g[:, 0], bb[0]) + 1., 0.)
                ih = np.maximum(np.minimum(g[:, 3], bb[3]) - np.maximum(g[:, 1], bb[1]) + 1., 0.)
                inters = iw * ih
                uni = ((bb[2] - bb[0] + 1.) * (bb[3] - bb[1] + 1.) +
                       (g[:, 2] - g[:, 0] + 1.) * (g[:, 3] - g[:, 1] + 1.) - inters)
                ov = inters / uni
                jmax = int(np.argmax(ov))
                ovmax = ov[jmax]
            if ovmax > iou_thr and not det[idx][jmax]:
                tp[j] = 1.
                det[idx][jmax] = True
            else:
                fp[j] = 1.
        fp, tp = np.cumsum(fp), np.cumsum(tp)
        rec = tp / np.maximum(float(npos), np.finfo(np.float64).eps)
        prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
        APs[cls] = voc_ap(rec, prec)
    return APs


class MAPCallback(object):
    """reference code/yolo3/map.py:10 (a ``tf.keras.callbacks.Callback``; here a plain object with the same
    constructor, ``set_model``, ``calculate_aps`` and ``on_train_end``)."""

    def __init__(self, glob_path, input_shape, class_names, iou=.5, batch_size=1, image_root=None):
        self.input_shape = input_shape
        self.class_names = class_names
        self.num_classes = len(class_names)
        self.glob_path = glob_path
        self.iou = iou
        self.batch_size = batch_size
        self.image_root = image_root  # engine extension: directory the list's relative image paths live under
        self.model = None

    def set_model(self, model):
        self.model = model

    def _lines(self) -> List[str]:
        files = sorted(_glob.glob(self.glob_path)) if isinstance(self.glob_path, str) else list(self.glob_path)
        if not files:
            raise FileNotFoundError("no annotation list matches %r" % (self.glob_path,))
        lines: List[str] = []
        for f in files:
            if f.endswith(".tfrecord") or f.endswith(".tfrecords"):
                raise ValueError("TFRecord inputs need TensorFlow's reader; give the text list (reference map.py:55-73)")
            with open(f) as fh:
                lines += [ln for ln in fh.read().splitlines() if ln.strip()]
        return lines

    def calculate_aps(self) -> Dict[int, float]:
        if self.model is None:
            raise RuntimeError("set_model(yolo.yolo_model) first (reference code/yolo.py:399)")
        true_res: Dict[int, np.ndarray] = {}
        pred_res: List[List[float]] = []
        start = timer()
        lines = self._lines()
        for idx, line in enumerate(lines):
            if idx % 100 == 0:
                print(idx)
            path, bbox = parse_text_line(line)
            if self.image_root is not None and not os.path.isabs(path):
                path = os.path.join(self.image_root, path)
            with open(path, "rb") as fh:
                image = fh.read()
            out_boxes, out_scores, out_classes = self.model([image])          # map.py:111
            for out_box, out_score, out_class in zip(out_boxes, out_scores, out_classes):
                top, left, bottom, right = out_box
                pred_res.append([idx, out_class, out_score, left, top, right, bottom])  # map.py:126-131
            true_res[idx] = bbox
        end = timer()
        print((end - start) / max(1, len(lines)))
        return class_aps(np.asarray(pred_res, dtype=np.float64).reshape(-1, 7), true_res, self.num_classes, self.iou)

    def on_train_end(self, logs=None):
        logs = {} if logs is None else logs
        APs = self.calculate_aps()
        for cls in range(self.num_classes):
            if cls in APs:
                print(self.class_names[cls] + ' ap: ', APs[cls])
        mAP = float(np.mean([APs[cls] for cls in APs]))
        print('mAP: ', mAP)
        logs['mAP'] = mAP
        return mAP


def calculate_map(yolo, glob, image_root=None):
    """reference code/yolo.py:397-405."""
    m = MAPCallback(glob, yolo.input_shape, yolo.class_names, image_root=image_root)
    m.set_model(yolo.yolo_model)
    APs = m.calculate_aps()
    for cls in range(len(yolo.class_names)):
        if cls in APs:
            print(yolo.class_names[cls] + ' ap: ', APs[cls])
    mAP = float(np.mean([APs[cls] for cls in APs]))
    print('mAP: ', mAP)
    return mAP, APs
