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
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
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
        slots = Cn * max_boxes
        # One flat int32 "wire" buffer holds everything a caller reads back, so a step's result is ONE
        # device->host copy and (multi-GPU) ONE all-gather:  [count B | status 1 | pad | boxes_i B*slots*4 |
        # scores B*slots | classes B*slots] and, behind the wire part, the un-truncated float boxes.
        hdr = (B + 1 + 3) // 4 * 4
        self.wire_words = hdr + B * slots * 6
        self.flat = torch.zeros(self.wire_words + B * slots * 4, dtype=i32, device=dev)
        o = 0
        self.out_count = self.flat[o:o + B]
        self.status = self.flat[B:B + 1]
        o = hdr
        self.out_boxes_i = self.flat[o:o + B * slots * 4].view(B, slots, 4)
        o += B * slots * 4
        self.out_scores = self.flat[o:o + B * slots].view(torch.float32).view(B, slots)
        o += B * slots
        self.out_classes = self.flat[o:o + B * slots].view(B, slots)
        o += B * slots
        self.out_boxes_f = self.flat[o:o + B * slots * 4].view(torch.float32).view(B, slots, 4)
        self.wire = self.flat[:self.wire_words]
        self._hdr, self._slots = hdr, slots
        self.host_flat = torch.zeros(self.flat.numel(), dtype=i32).pin_memory()
        self.host_flat2 = None  # second pinned landing buffer, allocated by streaming callers
        self._snap = None       # device snapshots of the wire + side stream for enqueue_read
        self.image_shapes = torch.zeros(B, 2, dtype=f32, device=dev)
        self._shapes_host = None
        self.workspace_bytes = sum(t.numel() * t.element_size() for t in (
            self.boxes, self.cand_score, self.cand_index, self.cand_count, self.det, self.det_count, self.flat))

    @_lib.on_device
    def set_image_shapes(self, shapes):
        """shapes: [B,2] (h,w) of the original images (yolo_eval's image_shape), or one (h,w) for all."""
        a = np.asarray(shapes, dtype=np.float32)
        if a.ndim == 1:
            a = np.broadcast_to(a[None], (self.batch, 2))
        if a.shape != (self.batch, 2):
            raise ValueError("image_shapes must be [%d,2] or one (h,w), got %s" % (self.batch, a.shape))
        if self._shapes_host is not None and np.array_equal(a, self._shapes_host):
            return  # unchanged since the last call: nothing to upload
        self._shapes_host = a.copy()
        self.image_shapes.copy_(torch.from_numpy(self._shapes_host))

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

    @_lib.on_device
    def run(self, feat_ptrs: Sequence[int], ld: Sequence[int], score_threshold: float, iou_threshold: float,
            stream: Optional[int] = None, events=None) -> int:
        """decode+filter -> class-wise NMS -> pack.  Returns the number of kernels launched."""
        st = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        p = self.params(score_threshold, ld)
        fp = (C.c_void_p * 3)()
        for s in range(self.num_scales):
            fp[s] = feat_ptrs[s]
        lib = self.lib
        if events is not None:  # bench.py per-kernel timing: 4 events around the 3 launches
            events[0].record()
        _lib.check(lib.yr_decode_filter(fp, self.image_shapes.data_ptr(), C.byref(p), self.boxes.data_ptr(),
                                        self.cand_score.data_ptr(), self.cand_index.data_ptr(),
                                        self.cand_count.data_ptr(), st), "yr_decode_filter")
        if events is not None:
            events[1].record()
        _lib.check(lib.yr_nms_classwise(self.boxes.data_ptr(), self.total_boxes, self.cand_score.data_ptr(),
                                        self.cand_index.data_ptr(), self.cand_count.data_ptr(), self.batch,
                                        self.num_classes, self.cand_cap, self.max_boxes, float(iou_threshold),
                                        self.det.data_ptr(), self.det_count.data_ptr(), self.status.data_ptr(), st),
                   "yr_nms_classwise")
        if events is not None:
            events[2].record()
        _lib.check(lib.yr_pack_detections(self.det.data_ptr(), self.det_count.data_ptr(), self.batch, self.num_classes,
                                          self.max_boxes, self.out_boxes_f.data_ptr(), self.out_boxes_i.data_ptr(),
                                          self.out_scores.data_ptr(), self.out_classes.data_ptr(),
                                          self.out_count.data_ptr(), st), "yr_pack_detections")
        if events is not None:
            events[3].record()
        return 3

    def d2h_bytes(self, with_float_boxes: bool = False) -> int:
        return (self.flat.numel() if with_float_boxes else self.wire_words) * 4

    @_lib.on_device
    def enqueue_read(self, slot: int = 0):
        """Asynchronous read of the wire words into pinned landing buffer ``slot`` (0/1) that does NOT hold up the
        compute stream: the current stream only snapshots the wire (device-to-device, a few microseconds), the
        device->host copy runs on a side stream and overlaps the next step.  Returns ``(host tensor, event)``; the
        caller synchronises the event before touching the tensor.  A slot may be reused once its event has fired."""
        if self._snap is None:
            self._snap = [torch.zeros(self.wire_words, dtype=torch.int32, device=self.device) for _ in range(2)]
            self._read_stream = torch.cuda.Stream(self.device)
            self._snapped = [torch.cuda.Event(), torch.cuda.Event()]
            self._read_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._read_used = [False, False]
        if slot == 1 and self.host_flat2 is None:
            self.host_flat2 = torch.zeros(self.wire_words, dtype=torch.int32).pin_memory()
        dst = (self.host_flat if slot == 0 else self.host_flat2)[:self.wire_words]
        main = torch.cuda.current_stream(self.device)
        if self._read_used[slot]:
            main.wait_event(self._read_done[slot])  # the previous read out of this snapshot (two steps ago) is over
        self._snap[slot].copy_(self.wire, non_blocking=True)
        self._snapped[slot].record(main)
        with torch.cuda.stream(self._read_stream):
            self._read_stream.wait_event(self._snapped[slot])
            dst.copy_(self._snap[slot], non_blocking=True)
            self._read_done[slot].record(self._read_stream)
        self._read_used[slot] = True
        return dst, self._read_done[slot]

    @_lib.on_device
    def read_wire(self, with_float_boxes: bool = False) -> np.ndarray:
        """ONE device->host copy of the wire buffer into pinned host memory (synchronises the stream)."""
        n = self.flat.numel() if with_float_boxes else self.wire_words
        self.host_flat[:n].copy_(self.flat[:n], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.host_flat[:n].numpy()

    def unpack_wire(self, w: np.ndarray, with_float_boxes: bool = False):
        """wire words (one rank's) -> per image (boxes int32 [n,4] (ymin,xmin,ymax,xmax), scores f32 [n],
        classes int32 [n]) in yolo_eval order (class-major, NMS selection order within a class)."""
        B, slots, hdr = self.batch, self._slots, self._hdr
        if int(w[B]) != 0:
            raise _lib.YrError("candidate list overflow (cand_cap=%d): raise cand_cap" % self.cand_cap)
        cnt = w[:B]
        o = hdr
        bi = w[o:o + B * slots * 4].reshape(B, slots, 4)
        o += B * slots * 4
        sc = w[o:o + B * slots].view(np.float32).reshape(B, slots)
        o += B * slots
        cl = w[o:o + B * slots].reshape(B, slots)
        o += B * slots
        bf = w[o:o + B * slots * 4].view(np.float32).reshape(B, slots, 4) if with_float_boxes else None
        out = []
        for b in range(B):
            n = int(cnt[b])
            r = (bi[b, :n].copy(), sc[b, :n].copy(), cl[b, :n].copy())
            out.append(r + (bf[b, :n].copy(),) if with_float_boxes else r)
        return out

    def padded_views(self, w: np.ndarray):
        """wire words -> (counts [B], boxes int32 [B,slots,4], scores f32 [B,slots], classes int32 [B,slots])
        numpy views (no copies); rows >= counts[b] are zero / class -1."""
        B, slots, hdr = self.batch, self._slots, self._hdr
        if int(w[B]) != 0:
            raise _lib.YrError("candidate list overflow (cand_cap=%d): raise cand_cap" % self.cand_cap)
        o = hdr
        bi = w[o:o + B * slots * 4].reshape(B, slots, 4)
        o += B * slots * 4
        sc = w[o:o + B * slots].view(np.float32).reshape(B, slots)
        o += B * slots
        return w[:B], bi, sc, w[o:o + B * slots].reshape(B, slots)

    def results(self, with_float_boxes: bool = False):
        """Device->host read.  Per image: (boxes int32 [n,4] (ymin,xmin,ymax,xmax), scores f32 [n], classes int32 [n])."""
        return self.unpack_wire(self.read_wire(with_float_boxes), with_float_boxes)
