# -*- coding: utf-8 -*-
"""Host-side mirror of reference code/yolo.py: ``YOLO(FLAGS).detect_image`` on the B200 engine.

  YoloModel   reference code/yolo.py:51-165  (decode -> letterbox -> body -> YoloEval)
  YOLO        reference code/yolo.py:168-315 (flag dict, generate(), detect_image())
plus ``detect_batch`` - the batched entry point the reference lacks (its graph is batch-1,
SURVEY.md F6): batch = independent per-image application of the batch-1 path.
"""
from __future__ import annotations

import colorsys
import io
import os
from timeit import default_timer as timer
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .yolo3.enums import BACKBONE
from .yolo3.model import yolov3_body, YoloBody
from .yolo3.utils import get_anchors, get_classes, letterbox_image

_BACKBONE_NAME = {BACKBONE.MOBILENETV2x75: "mobilenetv2x75", BACKBONE.MOBILENETV2x14: "mobilenetv2x14",
                  BACKBONE.EFFICIENTNETB3: "efficientnetb3", BACKBONE.EFFICIENTNETLITE0: "efficientnetlite0"}


class YoloModel:
    """reference code/yolo.py:51.  ``model([bytes]) -> (boxes, scores, classes)``."""

    def __init__(self, model_body, num_anchors, num_scales, classes, model_path, anchors, input_shape, score=0.2,
                 nms=0.5, with_classes=False, name=None, batch=1, weights=None, input_u8=False, quiet=False,
                 **kwargs):
        self.num_anchors, self.num_scales, self.classes = num_anchors, num_scales, classes
        self.with_classes, self.num_classes = with_classes, len(classes)
        self.model_path, self.anchors, self.score, self.nms = model_path, anchors, score, nms
        self.input_shapes = tuple(input_shape)
        self.batch = batch
        self.model: YoloBody = model_body((batch, input_shape[0], input_shape[1], 3),
                                          num_anchors=self.num_anchors // self.num_scales,
                                          num_classes=self.num_classes, input_u8=input_u8, num_scales=num_scales,
                                          **kwargs)
        if weights is not None:
            self.model.set_weights(weights, anchors)
        else:
            self.model.load_weights(self.model_path, anchors=anchors)
        self.engine = self.model.engine
        self._graph = None
        if not quiet:
            print(self.model_path)

    def parse_image(self, image: bytes):
        """reference code/yolo.py:105-112: decode (host, PIL) -> letterbox on the GPU."""
        from PIL import Image
        arr = np.array(Image.open(io.BytesIO(image)).convert("RGB"), dtype=np.uint8)
        dev_img = torch.from_numpy(arr).to(self.engine.device, non_blocking=True)
        return arr.shape[:2], letterbox_image(dev_img, self.input_shapes)

    def __call__(self, input):
        """``input``: list with one encoded image (reference: ``self.yolo_model([image_data])``)."""
        if self.batch != 1:
            raise ValueError("YoloModel.__call__ is the reference's batch-1 path; use YOLO.detect_batch")
        e = self.engine
        shape, lb = self.parse_image(input[0])
        e.input_slot(0, False)[0].copy_(lb)  # the letterboxed image is float32 in [0,1] (reference code/yolo.py:105-112)
        e.pp.set_image_shapes(shape)
        if self._graph is None:  # the ~90 launches of a step replay as one CUDA graph from the second image on
            self._graph = e.capture(self.score, self.nms, 0, False)
        self._graph.replay()
        boxes, scores, classes = e.results()[0]
        if self.with_classes:
            classes = np.asarray([self.classes[c].encode() for c in classes])
        return boxes, scores, classes


class YOLO(object):
    """reference code/yolo.py:168.  Same FLAGS keys and defaults."""

    def __init__(self, FLAGS):
        self.backbone = FLAGS.get('backbone', BACKBONE.MOBILENETV2x75)
        self.class_names = get_classes(FLAGS.get('classes_path', 'model_data/voc_classes.txt'))
        self.anchors = get_anchors(FLAGS.get('anchors_path', 'model_data/yolo_anchors'))
        self.input_shape = FLAGS.get('input_size', (416, 416))
        self.score = FLAGS.get('score', 0.2)
        self.nms = FLAGS.get('nms', 0.5)
        self.with_classes = FLAGS.get('with_classes', False)
        self.num_scales = FLAGS.get('num_scales', 3)
        self.quiet = bool(FLAGS.get('quiet', False))  # engine extension: silence the reference's prints
        self.generate(FLAGS)

    def generate(self, FLAGS):
        # engine extensions (not in the reference): batch size, in-memory weights, u8 batches
        batch = int(FLAGS.get('batch', 1))
        weights = FLAGS.get('weights', None)
        model_path = os.path.expanduser(FLAGS['model']) if weights is None else FLAGS.get('model', '<in-memory>')
        num_anchors = len(self.anchors)
        backbone = self.backbone
        if isinstance(backbone, str):
            backbone_name = backbone
        else:
            backbone_name = _BACKBONE_NAME[backbone]

        def model_body(inputs, **kw):
            return yolov3_body(inputs, model_name=backbone_name, drop_rate=0.2, data_format="channels_last", **kw)

        extra = {k: FLAGS[k] for k in ('micro_batch', 'device', 'pw_variant', 'fuse_se', 'lanes', 'autotune', 'fuse_up2', 'fuse_dwpw', 'fold_linear', 'stack_pw') if k in FLAGS}
        self.yolo_model = YoloModel(model_body, num_anchors, self.num_scales, self.class_names, model_path,
                                    self.anchors, self.input_shape, self.score, self.nms, self.with_classes,
                                    batch=batch, weights=weights, input_u8=bool(FLAGS.get('input_u8', False)),
                                    quiet=self.quiet, **extra)
        self.engine = self.yolo_model.engine
        if not self.quiet:
            print('{} model, anchors, and classes loaded.'.format(model_path))
        hsv_tuples = [(x / len(self.class_names), 1., 1.) for x in range(len(self.class_names))]
        self.colors = list(map(lambda x: colorsys.hsv_to_rgb(*x), hsv_tuples))
        self.colors = list(map(lambda x: (int(x[0] * 255), int(x[1] * 255), int(x[2] * 255)), self.colors))
        np.random.seed(10101)
        np.random.shuffle(self.colors)
        np.random.seed(None)
        self._graph = None

    def detect_image(self, image, draw=True):
        """reference code/yolo.py:235.  ``image``: bytes of an encoded image or a binary file object.
        draw=False -> (out_boxes int32 [N,4] (top,left,bottom,right), out_scores f32 [N], out_classes int32 [N])."""
        image_data = image
        if isinstance(image, bytes) is False:
            image_data = image.read()
        start = timer()
        out_boxes, out_scores, out_classes = self.yolo_model([image_data])
        end = timer()
        if not self.quiet:
            print('Found {} boxes for {}'.format(len(out_boxes), 'img'))
        if not draw:
            return out_boxes, out_scores, out_classes
        from PIL import Image, ImageDraw, ImageFont
        img = Image.open(io.BytesIO(image_data)).convert("RGB")
        try:
            font = ImageFont.truetype(font='font/FiraMono-Medium.otf',
                                      size=np.floor(3e-2 * img.size[1] + 0.5).astype('int32'))
        except OSError:  # the reference's font file is not shipped with it either
            font = ImageFont.load_default()
        thickness = max(1, (img.size[1] + img.size[0]) // 300)
        d = ImageDraw.Draw(img)
        for i, c in reversed(list(enumerate(out_classes))):
            if self.with_classes:
                c = self.class_names.index(str(c, encoding="utf-8"))
            label = '{} {:.2f}'.format(self.class_names[c], out_scores[i])
            top, left, bottom, right = (int(v) for v in out_boxes[i])
            for t in range(thickness):
                d.rectangle([left + t, top + t, right - t, bottom - t], outline=self.colors[c])
            d.text((left, max(0, top - 10)), label, fill=self.colors[c], font=font)
        if not self.quiet:
            print(end - start)
        return img

    # ---- batched extension -----------------------------------------------------------
    def detect_batch(self, images, image_shapes=None, use_graph: bool = True, unpack: bool = True):
        """``images``: [B,H,W,3] tensor already at the network input size - uint8 (scaled by 1/255
        on the GPU like tf.io.decode_image(dtype=float32)) or float32 in [0,1], either accepted by any engine - on the host
        (pinned memory recommended) or on the device.  ``image_shapes``: [B,2] original (h,w)
        per image (default: the input size).  Returns per-image (boxes, scores, classes) numpy
        arrays (``unpack=False``: the padded form ``(counts [B], boxes [B,20*C,4] int32, scores
        [B,20*C], classes [B,20*C])`` as views of the pinned host buffer, valid until the next call);
        includes the host->device copy of the batch and ONE device->host read of the results."""
        e = self.engine
        dst, u8 = e.slot_for(images)
        dst.copy_(images, non_blocking=True)
        e.pp.set_image_shapes(image_shapes if image_shapes is not None else self.input_shape)
        if use_graph:
            if getattr(self, "_batch_graphs", None) is None:
                self._batch_graphs = {}
            if u8 not in self._batch_graphs:
                self._batch_graphs[u8] = e.capture(self.score, self.nms, 0, u8)
            self._batch_graphs[u8].replay()
        else:
            e.step(self.score, self.nms, 0, u8)
        if unpack:
            return e.results()
        return e.pp.padded_views(e.pp.read_wire())

    def detect_images(self, images):
        """Batched ``detect_image`` (the reference handles one image per call, code/yolo.py:235): ``images`` is a list of
        up to ``batch`` encoded images (bytes or binary file objects) of ANY sizes.  Each is decoded on the host (PIL, as
        ``parse_image`` does, code/yolo.py:105-112), letterboxed on the GPU straight into its row of the engine's input
        batch (``letterbox_image``, code/yolo3/utils.py:67-83) and mapped back to its own image shape by ``yolo_eval``;
        ONE network pass serves the whole list.  Returns per image what ``detect_image(image, draw=False)`` returns."""
        from PIL import Image
        e = self.engine
        if not (1 <= len(images) <= e.batch):
            raise ValueError("detect_images takes 1..%d images (the engine's batch), got %d" % (e.batch, len(images)))
        dst = e.input_slot(0, False)
        shapes = np.tile(np.asarray(self.input_shape, np.float32), (e.batch, 1))
        for b, im in enumerate(images):
            data = im if isinstance(im, bytes) else im.read()
            arr = np.array(Image.open(io.BytesIO(data)).convert("RGB"), dtype=np.uint8)
            shapes[b] = arr.shape[:2]
            letterbox_image(torch.from_numpy(arr).to(e.device, non_blocking=True), self.input_shape, out=dst[b])
        if len(images) < e.batch:
            dst[len(images):].zero_()
        e.pp.set_image_shapes(shapes)
        if getattr(self, "_batch_graphs", None) is None:
            self._batch_graphs = {}
        if False not in self._batch_graphs:
            self._batch_graphs[False] = e.capture(self.score, self.nms, 0, False)
        self._batch_graphs[False].replay()
        return e.results()[:len(images)]

    def detect_stream(self, batches, image_shapes=None, unpack: bool = True, gather=None, gather_read: bool = True):
        """Pipelined ``detect_batch`` over an iterable of host batches (pinned memory recommended; uint8 or float32):
        yields one result per batch, in order.  The host->device upload of batch i+1 runs on a copy stream into a
        second input slot while batch i computes, and each batch's detections come back with ONE
        device->host copy on a side stream (the compute stream only snapshots the wire buffer), so in steady state
        a step costs max(compute, PCIe) instead of their sum.
        Every batch is still uploaded, computed and read back; nothing is cached between batches.

        Multi-GPU: pass ``gather`` (a ``parallel.DetectionGather`` over this engine's wire).  Every step then also
        all-gathers the ranks' detection wires on a side stream, overlapped with the next step's compute, and the
        generator yields ``(local_result, parts)`` with ``parts`` = one wire (numpy int32 view) per rank in rank =
        batch order (``None`` when ``gather_read`` is False: the gathered buffer stays on the device)."""
        e = self.engine
        dev = e.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
            self._graphs = {}
        e.pp.set_image_shapes(image_shapes if image_shapes is not None else self.input_shape)
        cs = self._copy_stream
        uploaded = [torch.cuda.Event(), torch.cuda.Event()]   # input slot filled
        consumed = [torch.cuda.Event(), torch.cuda.Event()]   # input slot free again (its graph finished)
        landed = [None, None]                                 # wire copy of that step on the host (side stream)
        kinds = [None, None]  # dtype flavour (uint8 / float32) of the batch sitting in each slot

        def upload(images, slot, first_use):
            dst, u8 = e.slot_for(images, slot)
            if (slot, u8) not in self._graphs:  # first batch of this flavour in this slot: capture its graph
                cs.synchronize()
                main.synchronize()
                self._graphs[(slot, u8)] = e.capture(self.score, self.nms, slot, u8)
                main.synchronize()
            kinds[slot] = u8
            with torch.cuda.stream(cs):
                if not first_use:
                    cs.wait_event(consumed[slot])
                dst.copy_(images, non_blocking=True)
                uploaded[slot].record(cs)

        def finish(p):
            ps, ph, tk = p
            landed[ps].synchronize()
            w = ph.numpy()
            local = e.pp.unpack_wire(w) if unpack else e.pp.padded_views(w)
            if gather is None:
                return local
            parts = gather.wait(tk) if gather_read else None
            return local, parts

        it = iter(batches)
        try:
            cur = next(it)
        except StopIteration:
            return
        cs.wait_stream(main)
        upload(cur, 0, True)
        i = 0
        pending = None  # (slot, host wire tensor, gather ticket) of the previous step, not yet yielded
        while cur is not None:
            slot = i & 1
            main.wait_event(uploaded[slot])
            self._graphs[(slot, kinds[slot])].replay()
            consumed[slot].record(main)
            host, landed[slot] = e.pp.enqueue_read(slot)
            ticket = gather.gather_async(read=gather_read) if gather is not None else None
            try:
                nxt = next(it)
            except StopIteration:
                nxt = None
            if nxt is not None:
                upload(nxt, (i + 1) & 1, i == 0)
            if pending is not None:
                yield finish(pending)
            pending = (slot, host, ticket)
            cur = nxt
            i += 1
        yield finish(pending)
        if gather is not None:
            gather.join()
