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


def letterbox_image(image: torch.Tensor, size) -> torch.Tensor:
    """reference code/yolo3/utils.py:67-83 on the GPU.

    ``image``: uint8 CUDA tensor [ih, iw, 3] (a decoded image).  Returns float32
    [h, w, 3] in [0,1]: (1/255) scaling as tf.io.decode_image(dtype=float32), bilinear
    half-pixel resize keeping aspect ratio, zero padding."""
    if not (image.is_cuda and image.dtype == torch.uint8 and image.dim() == 3 and image.shape[2] == 3):
        raise ValueError("letterbox_image expects a uint8 CUDA tensor [H,W,3]")
    image = image.contiguous()
    ih, iw = int(image.shape[0]), int(image.shape[1])
    h, w = int(size[0]), int(size[1])
    nh, nw, dy, dx = letterbox_geometry(ih, iw, size)
    out = torch.empty(h, w, 3, dtype=torch.float32, device=image.device)
    st = torch.cuda.current_stream(image.device).cuda_stream
    _lib.check(_lib.lib().yr_letterbox_u8(image.data_ptr(), ih, iw, out.data_ptr(), h, w, nh, nw, dy, dx, st),
               "yr_letterbox_u8")
    return out
