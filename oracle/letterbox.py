"""torch-CPU restatement of the reference pre-process (TEST INFRASTRUCTURE).

Follows ``YoloModel.parse_image`` (reference code/yolo.py:105-112) and
``letterbox_image`` (code/yolo3/utils.py:67-83):
  decode_image(dtype=float32) == uint8 * (1/255)  (tf.image.convert_image_dtype
  multiplies by the fp32 reciprocal); nh/nw = int(ih * min(w/iw, h/ih)) in
  float64; tf.image.resize bilinear with half-pixel centres, no antialias;
  pad_to_bounding_box with zeros at ((h-nh)//2, (w-nw)//2).
JPEG decoding itself (PIL here, libjpeg-turbo in TF) is outside the parity
contract.  Parity is unpinned by the reference.
"""
from __future__ import annotations

import io

import numpy as np
import torch
import torch.nn.functional as F


def decode_image_u8(data: bytes) -> np.ndarray:
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"), dtype=np.uint8)


def u8_to_float(img_u8: np.ndarray) -> np.ndarray:
    return img_u8.astype(np.float32) * np.float32(1.0 / 255.0)


def letterbox_dims(ih: int, iw: int, h: int, w: int):
    r = min(w / iw, h / ih)  # python float == float64, utils.py:76-77
    nh, nw = int(float(ih) * r), int(float(iw) * r)
    return nh, nw, (h - nh) // 2, (w - nw) // 2


def letterbox_image(image_hwc: np.ndarray, size) -> np.ndarray:
    """image_hwc float32 [ih,iw,3] in [0,1]; size (h,w).  Returns [h,w,3] f32."""
    h, w = int(size[0]), int(size[1])
    ih, iw = image_hwc.shape[:2]
    nh, nw, dy, dx = letterbox_dims(ih, iw, h, w)
    t = torch.from_numpy(np.ascontiguousarray(image_hwc)).permute(2, 0, 1)[None]
    r = F.interpolate(t, size=(nh, nw), mode="bilinear", align_corners=False, antialias=False)
    out = torch.zeros(1, 3, h, w, dtype=torch.float32)
    out[:, :, dy:dy + nh, dx:dx + nw] = r
    return out[0].permute(1, 2, 0).contiguous().numpy()
