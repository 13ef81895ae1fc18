"""Host-side mirror of reference code/yolo3/utils.py (the functions on the hot path)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def get_anchors(anchors_path):
    """reference code/yolo3/utils.py:100-104: first line, comma separated, (N,2) float32 (w,h)."""
    with open(anchors_path) as f:
        anchors = f.readline()
    anchors = [float(x) for x in anchors.split(',')]
    return np.array(anchors, np.float32).reshape(-1, 2)


def get_classes(classes_path):
    """reference code/yolo3/utils.py:115-120."""
    with open(classes_path) as f:
        class_names = f.readlines()
    return [c.strip() for c in class_names]


def letterbox_geometry(ih: int, iw: int, size):
    """nh, nw, dy, dx exactly as reference code/yolo3/utils.py:75-79 (float64 scale, int truncation)."""
    h, w = int(size[0]), int(size[1])
    r = min(w / iw, h / ih)
    nh, nw = int(float(ih) * r), int(float(iw) * r)
    return nh, nw, (h - nh) // 2, (w - nw) // 2


def letterbox_image(image: torch.Tensor, size, out: torch.Tensor = None) -> torch.Tensor:
    """reference code/yolo3/utils.py:67-83 on the GPU.

    ``image``: uint8 CUDA tensor [ih, iw, 3] (a decoded image).  Returns float32
    [h, w, 3] in [0,1]: (1/255) scaling as tf.io.decode_image(dtype=float32), bilinear
    half-pixel resize keeping aspect ratio, zero padding.  ``out``: a contiguous float32 CUDA
    [h, w, 3] tensor to write into (e.g. one image of an engine's input batch) instead of a new one."""
    if not (image.is_cuda and image.dtype == torch.uint8 and image.dim() == 3 and image.shape[2] == 3):
        raise ValueError("letterbox_image expects a uint8 CUDA tensor [H,W,3]")
    image = image.contiguous()
    ih, iw = int(image.shape[0]), int(image.shape[1])
    h, w = int(size[0]), int(size[1])
    nh, nw, dy, dx = letterbox_geometry(ih, iw, size)
    if out is None:
        out = torch.empty(h, w, 3, dtype=torch.float32, device=image.device)
    elif not (out.is_cuda and out.dtype == torch.float32 and tuple(out.shape) == (h, w, 3) and out.is_contiguous()):
        raise ValueError("letterbox_image: out must be a contiguous float32 CUDA tensor [%d,%d,3]" % (h, w))
    st = torch.cuda.current_stream(image.device).cuda_stream
    _lib.check(_lib.lib().yr_letterbox_u8(image.data_ptr(), ih, iw, out.data_ptr(), h, w, nh, nw, dy, dx, st),
               "yr_letterbox_u8")
    return out


_ANCHOR_MASK = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]


def preprocess_true_boxes(true_boxes, input_shape, anchors, num_classes, num_scales=3):
    """y_true encoder, reference code/yolo3/utils.py:298-376 (host-side numpy there too; it runs inside the
    tf.data pipeline, code/yolo3/data.py:84-121).  ONE image: ``true_boxes`` [T,5] = (xmin, ymin, xmax, ymax,
    class) in input pixels, rows with zero width are padding.  Returns one array per scale,
    [gh, gw, 3, 5+num_classes]: (cx, cy, w, h) normalised, objectness, one-hot class - written at the cell
    holding the box centre, for the anchor (of all 9) whose shape has the best IoU with the box; a later box
    overwrites an earlier one in the same slot, as the reference's loop does."""
    mask = _ANCHOR_MASK[-num_scales:]
    tb = np.array(true_boxes, dtype=np.float32).reshape(-1, 5)
    hw = np.array(input_shape, dtype=np.int32)
    wh_in = hw[::-1]
    centre = (tb[:, 0:2] + tb[:, 2:4]) // 2            # floor-divided centre, utils.py:321
    size = tb[:, 2:4] - tb[:, 0:2]
    rel = np.concatenate([centre / wh_in, size / wh_in], 1).astype(np.float32)
    grids = [np.round(hw / s).astype(np.int32) for s in (32, 16, 8)[:num_scales]]
    y_true = [np.zeros((g[0], g[1], 3, 5 + num_classes), np.float32) for g in grids]
    keep = size[:, 0] > 0
    if not keep.any():
        return y_true
    anc = np.asarray(anchors, np.float32).reshape(-1, 2)
    bw = size[keep]
    inter = np.minimum(bw[:, None, 0], anc[None, :, 0]) * np.minimum(bw[:, None, 1], anc[None, :, 1])
    union = bw[:, None, 0] * bw[:, None, 1] + anc[None, :, 0] * anc[None, :, 1] - inter
    best = np.argmax(inter / union, axis=1)             # centred boxes: IoU of shapes only
    # the reference enumerates best_anchor (valid boxes only) but indexes true_boxes[t] with that counter
    # (utils.py:357-368), i.e. valid boxes are assumed to come first; mirrored here
    for t, n in enumerate(best):
        for l, m in enumerate(mask):
            if n in m:
                i = int(np.floor(rel[t, 0] * grids[l][1]))
                j = int(np.floor(rel[t, 1] * grids[l][0]))
                y_true[l][j, i, m.index(n), 0:4] = rel[t]
                y_true[l][j, i, m.index(n), 4] = 1.0
                y_true[l][j, i, m.index(n), 5 + int(tb[t, 4])] = 1.0
    return y_true


def encode_true_boxes_batch(boxes_batch, input_shape, anchors, num_classes, num_scales=3):
    """Stacks ``preprocess_true_boxes`` over a batch: list of [T,5] -> list (per scale) of [B,gh,gw,3,5+C]."""
    per_image = [preprocess_true_boxes(b, input_shape, anchors, num_classes, num_scales) for b in boxes_batch]
    return [np.stack([y[l] for y in per_image]) for l in range(num_scales)]


def preprocess_true_boxes_gpu(true_boxes: torch.Tensor, input_shape, anchors, num_classes, num_scales=3):
    """``preprocess_true_boxes`` for a whole batch on the GPU (reference code/yolo3/utils.py:298-376, which the
    reference runs per sample as a numpy py_function, code/yolo3/data.py:84-121).  ``true_boxes``: float32 CUDA
    tensor [B,T,5] = (xmin, ymin, xmax, ymax, class) in input pixels, zero-width rows are padding.  Returns the
    per-scale dense y_true tensors [B,gh,gw,3,5+num_classes] (CUDA), bit-identical to stacking the host encoder."""
    if not (true_boxes.is_cuda and true_boxes.dtype == torch.float32 and true_boxes.dim() == 3 and true_boxes.shape[2] == 5):
        raise ValueError("preprocess_true_boxes_gpu expects a float32 CUDA tensor [B,T,5]")
    import ctypes as C
    tb = true_boxes.contiguous()
    B, T = int(tb.shape[0]), int(tb.shape[1])
    h, w = int(input_shape[0]), int(input_shape[1])
    anc = np.ascontiguousarray(np.asarray(anchors, np.float32).reshape(-1))
    if anc.size != 18:
        raise ValueError("preprocess_true_boxes needs the 9 anchors (18 numbers), got %d" % anc.size)
    grids = [np.round(np.array([h, w], np.int32) / s).astype(np.int32) for s in (32, 16, 8)[:num_scales]]
    ys = [torch.empty(B, int(g[0]), int(g[1]), 3, 5 + num_classes, dtype=torch.float32, device=tb.device) for g in grids]
    ptrs = (C.c_void_p * num_scales)(*[y.data_ptr() for y in ys])
    st = torch.cuda.current_stream(tb.device).cuda_stream
    _lib.check(_lib.lib().yr_encode_true_boxes(tb.data_ptr(), B, T, anc.ctypes.data_as(C.POINTER(C.c_float)), h, w,
                                               num_classes, num_scales, ptrs, st), "yr_encode_true_boxes")
    return ys


class SparseYTrue:
    """Sparse y_true of a batch: what ``preprocess_true_boxes`` (reference
    code/yolo3/utils.py:298-376) encodes, without the > 99.9 % zeros of the dense tensors.  Per scale ``l`` an int32 slot
    map ``maps[l]`` [B,gh,gw,3] (record index or -1) and a record list ``records[l]`` [cap,8] = (x, y, w, h normalised
    centre form, 4 words of class bits); ``counts`` [3] records per scale.  ``to_dense()`` rebuilds the reference's
    tensors (the adapter for code that wants the reference signature)."""

    def __init__(self, B, grids, num_classes, cap, device):
        self.B, self.grids, self.num_classes, self.cap = int(B), [tuple(int(v) for v in g) for g in grids], int(num_classes), int(cap)
        self.maps = [torch.full((self.B, g[0], g[1], 3), -1, dtype=torch.int32, device=device) for g in self.grids]
        self.records = torch.zeros(3, self.cap, 8, dtype=torch.float32, device=device)
        self.counts = torch.zeros(3, dtype=torch.int32, device=device)

    def to_dense(self):
        out = []
        E = 5 + self.num_classes
        counts = self.counts.cpu().numpy()
        for l, (m, g) in enumerate(zip(self.maps, self.grids)):
            y = torch.zeros(self.B, g[0], g[1], 3, E, dtype=torch.float32, device=m.device)
            idx = torch.nonzero(m >= 0, as_tuple=False)
            if len(idx):
                if int(counts[l]) > self.cap:
                    raise _lib.YrError("sparse y_true overflow: %d records > capacity %d" % (int(counts[l]), self.cap))
                r = m[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]].long()
                rec = self.records[l][r]
                y[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3], 0:4] = rec[:, 0:4]
                y[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3], 4] = 1.0
                bits = rec[:, 4:8].contiguous().view(torch.int32)
                for c in range(self.num_classes):
                    on = ((bits[:, c >> 5] >> (c & 31)) & 1).float()
                    y[idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3], 5 + c] = on
            out.append(y)
        return out


def encode_true_boxes_sparse(true_boxes: torch.Tensor, input_shape, anchors, num_classes, num_scales=3,
                             out: "SparseYTrue" = None) -> "SparseYTrue":
    """``preprocess_true_boxes`` for a whole batch on the GPU, emitting the SPARSE form (slot maps + records) instead of
    the dense tensors: same box -> (scale, cell, anchor) assignment, same slot-collision behaviour (reference
    code/yolo3/utils.py:298-376).  ``true_boxes``: float32 CUDA [B,T,5] = (xmin, ymin, xmax, ymax, class) in input
    pixels, zero-width rows are padding.  ``out``: a ``SparseYTrue`` to refill (no allocation; graph-capturable)."""
    if not (true_boxes.is_cuda and true_boxes.dtype == torch.float32 and true_boxes.dim() == 3 and true_boxes.shape[2] == 5):
        raise ValueError("encode_true_boxes_sparse expects a float32 CUDA tensor [B,T,5]")
    if not (1 <= int(num_classes) <= 128):
        raise ValueError("the sparse y_true holds up to 128 classes, got %d" % num_classes)
    import ctypes as C
    tb = true_boxes.contiguous()
    B, T = int(tb.shape[0]), int(tb.shape[1])
    h, w = int(input_shape[0]), int(input_shape[1])
    anc = np.ascontiguousarray(np.asarray(anchors, np.float32).reshape(-1))
    if anc.size != 18:
        raise ValueError("preprocess_true_boxes needs the 9 anchors (18 numbers), got %d" % anc.size)
    grids = [tuple(int(v) for v in np.round(np.array([h, w], np.int32) / s).astype(np.int32)) for s in (32, 16, 8)[:num_scales]]
    if out is None:
        out = SparseYTrue(B, grids, num_classes, max(1, B * T), tb.device)
    elif out.B != B or out.grids != grids or out.num_classes != num_classes or out.cap < B * T:
        raise ValueError("the SparseYTrue to refill does not match this batch")
    with torch.cuda.device(tb.device):
        ptrs = (C.c_void_p * 3)(*[m.data_ptr() for m in out.maps] + [None] * (3 - len(out.maps)))
        st = torch.cuda.current_stream(tb.device).cuda_stream
        _lib.check(_lib.lib().yr_encode_true_boxes_sparse(tb.data_ptr(), B, T, anc.ctypes.data_as(C.POINTER(C.c_float)), h, w,
                                                          num_classes, num_scales, ptrs, out.records.data_ptr(),
                                                          out.counts.data_ptr(), out.cap, st), "yr_encode_true_boxes_sparse")
    return out
