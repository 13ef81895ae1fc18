"""yolo_eval on the GPU: workspace + calls into the C-ABI post-process kernels.

Mirrors reference ``yolo_eval`` (code/yolo3/model.py:431-491) for a batch of independent
images (the reference is batch-1, SURVEY.md F6: batch = per-image application).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import YrDecodeParams

ANCHOR_MASK = [[6, 7, 8], [3, 4, 5], [0, 1, 2]]  # reference code/yolo3/model.py:444


class PostProcess:
    def __init__(self, batch: int, grids: Sequence[Sequence[int]], num_classes: int, anchors, num_scales: int = 3,
                 max_boxes: int = 20, device=None, cand_cap: Optional[int] = None):
        if not torch.cuda.is_available():
            raise _lib.YrError("yoloret_b200 post-process needs a CUDA device (no CPU fallback exists)")
        self.lib = _lib.lib()
        dev = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.device = dev
        self.batch, self.num_classes, self.num_scales, self.max_boxes = batch, num_classes, num_scales, max_boxes
        self.grids = [tuple(g) for g in grids[:num_scales]]
        self.anchors = np.asarray(anchors, dtype=np.float32).reshape(-1, 2)
        if len(self.anchors) < 3 * num_scales:
            raise ValueError("need %d anchors, got %d" % (3 * num_scales, len(self.anchors)))
        self.total_boxes = sum(3 * h * w for h, w in self.grids)
        self.cand_cap = int(cand_cap) if cand_cap else self.total_boxes
        B, Cn = batch, num_classes
        f32, i32 = torch.float32, torch.int32
        self.boxes = torch.zeros(B, self.total_boxes, 4, dtype=f32, device=dev)
        self.cand_score = torch.empty(B, Cn, self.cand_cap, dtype=f32, device=dev)
        self.cand_index = torch.empty(B, Cn, self.cand_cap, dtype=i32, device=dev)
        self.cand_count = torch.zeros(B, Cn, dtype=i32, device=dev)
        self.det = torch.zeros(B, Cn, max_boxes, 6, dtype=f32, device=dev)
        self.det_count = torch.zeros(B, Cn, dtype=i32, device=dev)
        self.status = torch.zeros(1, dtype=i32, device=dev)
        slots = Cn * max_boxes
        self.out_boxes_f = torch.zeros(B, slots, 4, dtype=f32, device=dev)
        self.out_boxes_i = torch.zeros(B, slots, 4, dtype=i32, device=dev)
        self.out_scores = torch.zeros(B, slots, dtype=f32, device=dev)
        self.out_classes = torch.zeros(B, slots, dtype=i32, device=dev)
        self.out_count = torch.zeros(B, dtype=i32, device=dev)
        self.image_shapes = torch.zeros(B, 2, dtype=f32, device=dev)
        self.workspace_bytes = sum(t.numel() * t.element_size() for t in (
            self.boxes, self.cand_score, self.cand_index, self.cand_count, self.det, self.det_count,
            self.out_boxes_f, self.out_boxes_i, self.out_scores, self.out_classes, self.out_count))

    def set_image_shapes(self, shapes):
        """shapes: [B,2] (h,w) of the original images (yolo_eval's image_shape), or one (h,w) for all."""
        s = torch.as_tensor(np.asarray(shapes, dtype=np.float32))
        if s.dim() == 1:
            s = s[None].expand(self.batch, 2)
        self.image_shapes.copy_(s.contiguous(), non_blocking=True)

    def params(self, score_threshold: float, ld: Sequence[int]) -> YrDecodeParams:
        p = YrDecodeParams()
        p.B, p.num_classes, p.num_scales = self.batch, self.num_classes, self.num_scales
        mask = ANCHOR_MASK[-self.num_scales:]
        for s in range(self.num_scales):
            p.grid_h[s], p.grid_w[s], p.ld[s] = self.grids[s][0], self.grids[s][1], int(ld[s])
            for k in range(3):
                p.anchors[s][k][0] = float(self.anchors[mask[s][k]][0])
                p.anchors[s][k][1] = float(self.anchors[mask[s][k]][1])
        p.input_h, p.input_w = self.grids[0][0] * 32, self.grids[0][1] * 32  # model.py:449
        p.score_threshold = float(score_threshold)
        p.cand_cap = self.cand_cap
        return p

    def run(self, feat_ptrs: Sequence[int], ld: Sequence[int], score_threshold: float, iou_threshold: float,
            stream: Optional[int] = None) -> int:
        """decode+filter -> class-wise NMS -> pack.  Returns the number of kernels launched."""
        st = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        p = self.params(score_threshold, ld)
        fp = (C.c_void_p * 3)()
        for s in range(self.num_scales):
            fp[s] = feat_ptrs[s]
        lib = self.lib
        _lib.check(lib.yr_decode_filter(fp, self.image_shapes.data_ptr(), C.byref(p), self.boxes.data_ptr(),
                                        self.cand_score.data_ptr(), self.cand_index.data_ptr(),
                                        self.cand_count.data_ptr(), st), "yr_decode_filter")
        _lib.check(lib.yr_nms_classwise(self.boxes.data_ptr(), self.total_boxes, self.cand_score.data_ptr(),
                                        self.cand_index.data_ptr(), self.cand_count.data_ptr(), self.batch,
                                        self.num_classes, self.cand_cap, self.max_boxes, float(iou_threshold),
                                        self.det.data_ptr(), self.det_count.data_ptr(), self.status.data_ptr(), st),
                   "yr_nms_classwise")
        _lib.check(lib.yr_pack_detections(self.det.data_ptr(), self.det_count.data_ptr(), self.batch, self.num_classes,
                                          self.max_boxes, self.out_boxes_f.data_ptr(), self.out_boxes_i.data_ptr(),
                                          self.out_scores.data_ptr(), self.out_classes.data_ptr(),
                                          self.out_count.data_ptr(), st), "yr_pack_detections")
        return 3

    def d2h_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.out_boxes_i, self.out_scores, self.out_classes,
                                                          self.out_count, self.status))

    def results(self, with_float_boxes: bool = False):
        """Device->host read.  Per image: (boxes int32 [n,4] (ymin,xmin,ymax,xmax), scores f32 [n], classes int32 [n])."""
        cnt = self.out_count.cpu().numpy()
        bi, sc, cl = self.out_boxes_i.cpu().numpy(), self.out_scores.cpu().numpy(), self.out_classes.cpu().numpy()
        bf = self.out_boxes_f.cpu().numpy() if with_float_boxes else None
        if int(self.status.item()) != 0:
            raise _lib.YrError("candidate list overflow (cand_cap=%d): raise cand_cap" % self.cand_cap)
        out = []
        for b, n in enumerate(cnt):
            r = (bi[b, :n].copy(), sc[b, :n].copy(), cl[b, :n].copy())
            out.append(r + (bf[b, :n].copy(),) if with_float_boxes else r)
        return out
