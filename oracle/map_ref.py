"""CPU restatement of the reference's VOC mAP evaluation (TEST INFRASTRUCTURE ONLY - see oracle/__init__.py).

Follows reference code/yolo3/map.py line by line, with plain Python loops:
  * ``voc_ap``            map.py:16-32   (area under the monotone precision envelope)
  * ``parse_text_line``   map.py:55-73   (``path xmin ymin xmax ymax label ...``, boxes as [xmin,ymin,xmax,ymax,label])
  * ``class_aps``         map.py:157-222 (per class: sort by -score, greedy match at IoU > thr with the VOC
                                           "+1 pixel" box sizes, first match of a ground truth is the TP)
Parity is unpinned by the reference (it has no tests or expected mAP values); the pins are
hand-computed cases in tests/test_cpu_map.py.
"""
import numpy as np


def voc_ap(rec, prec):
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])


def parse_text_line(line):
    values = line.strip().split(' ')
    nums = np.array([float(v) for v in values[1:]], dtype=np.float32).reshape(-1, 5)
    return values[0], nums  # columns xmin, ymin, xmax, ymax, label (map.py:59-73)


def class_aps(pred_res, true_res, num_classes, iou_thr=0.5):
    """pred_res: list of [image idx, class, score, left, top, right, bottom] (map.py:126-131);
    true_res: {image idx: array [n,5] xmin,ymin,xmax,ymax,label} (map.py:132)."""
    APs = {}
    for cls in range(num_classes):
        pred_res_cls = [x for x in pred_res if x[1] == cls]
        if len(pred_res_cls) == 0:
            APs[cls] = 0
            continue
        true_res_cls = {}
        npos = 0
        for index in true_res:
            objs = [obj for obj in true_res[index] if obj[4] == cls]
            npos += len(objs)
            BBGT = np.array([x[:4] for x in objs])
            true_res_cls[index] = {'bbox': BBGT, 'det': [False] * len(objs)}
        ids = [x[0] for x in pred_res_cls]
        scores = np.array([x[2] for x in pred_res_cls])
        bboxs = np.array([x[3:] for x in pred_res_cls])
        sorted_ind = np.argsort(-scores)
        bboxs = bboxs[sorted_ind, :]
        ids = [ids[x] for x in sorted_ind]
        nd = len(ids)
        tp = np.zeros(nd)
        fp = np.zeros(nd)
        for j in range(nd):
            res = true_res_cls[ids[j]]
            bbox = bboxs[j, :].astype(float)
            ovmax = -np.inf
            BBGT = res['bbox'].astype(float)
            if BBGT.size > 0:
                ixmin = np.maximum(BBGT[:, 0], bbox[0])
                iymin = np.maximum(BBGT[:, 1], bbox[1])
                ixmax = np.minimum(BBGT[:, 2], bbox[2])
                iymax = np.minimum(BBGT[:, 3], bbox[3])
                iw = np.maximum(ixmax - ixmin + 1., 0.)
                ih = np.maximum(iymax - iymin + 1., 0.)
                inters = iw * ih
                uni = ((bbox[2] - bbox[0] + 1.) * (bbox[3] - bbox[1] + 1.) +
                       (BBGT[:, 2] - BBGT[:, 0] + 1.) * (BBGT[:, 3] - BBGT[:, 1] + 1.) - inters)
                overlaps = inters / uni
                ovmax = np.max(overlaps)
                jmax = np.argmax(overlaps)
            if ovmax > iou_thr:
                if not res['det'][jmax]:
                    tp[j] = 1.
                    res['det'][jmax] = 1
                else:
                    fp[j] = 1.
            else:
                fp[j] = 1.
        fp = np.cumsum(fp)
        tp = np.cumsum(tp)
        rec = tp / np.maximum(float(npos), np.finfo(np.float64).eps)
        prec = tp / np.maximum(tp + fp, np.finfo(np.float64).eps)
        APs[cls] = voc_ap(rec, prec)
    return APs
