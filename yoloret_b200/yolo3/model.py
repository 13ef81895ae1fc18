"""Host-side mirror of reference code/yolo3/model.py on the B200 engine.

Same names, argument meaning and error behaviour as the reference functions; tensors
are CUDA ``torch.Tensor``s instead of ``tf.Tensor``s, and every arithmetic step runs in
the CUDA library behind include/yoloret_b200.h (no CPU fallback).

  yolov3_body / yolo_body   reference code/yolo3/model.py:170-342
  yolo_eval / YoloEval      code/yolo3/model.py:431-526
  YoloLoss / yolo_loss      code/yolo3/model.py:585-671 (+ AdvLossModel._compute_total_loss,
                            code/yolo3/train.py:11-16)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import YrLossParams, YrLoss3Params
from ..engine import Engine
from ..postprocess import PostProcess, ANCHOR_MASK
from ..weights import load_checkpoint
from .enums import BOX_LOSS

_BACKBONES = ("mobilenetv2x75", "mobilenetv2x14", "efficientnetb3", "efficientnetlite0")
_GLOBAL_PARAM_FIELDS = ("batch_norm_momentum", "batch_norm_epsilon", "dropout_rate", "data_format", "num_classes",
                        "width_coefficient", "depth_coefficient", "depth_divisor", "min_depth", "drop_connect_rate")


class YoloBody:
    """What ``yolov3_body`` returns: callable ``model(x) -> [y1, y2, y3]`` with
    ``load_weights`` / ``set_weights`` (reference: an ``AdvLossModel``, model.py:342)."""

    def __init__(self, input_shape, model_name, num_anchors, num_classes, micro_batch=None, device=None,
                 pw_variant=_lib.PW_AUTO, input_u8=False, num_scales=3, fuse_se=True, lanes=1, autotune=True, fuse_up2=True, fuse_dwpw=True, fold_linear=True,
                 stack_pw=True):
        self.batch, self.input_hw = int(input_shape[0]), (int(input_shape[1]), int(input_shape[2]))
        self.model_name, self.num_anchors, self.num_classes = model_name, num_anchors, num_classes
        self._kw = dict(micro_batch=micro_batch, device=device, pw_variant=pw_variant, input_u8=input_u8,
                        num_scales=num_scales, fuse_se=fuse_se, lanes=lanes, autotune=autotune, fuse_up2=fuse_up2, fuse_dwpw=fuse_dwpw, fold_linear=fold_linear,
                        stack_pw=stack_pw)
        self.engine: Optional[Engine] = None
        from ..netdef import NetDef
        self.netdef = NetDef(model_name, num_classes, self.input_hw, num_anchors)
        self.anchors = np.zeros((9, 2), np.float32)

    @property
    def weight_shapes(self):
        return self.netdef.weight_shapes

    def set_weights(self, weights: Dict[str, np.ndarray], anchors=None):
        if anchors is not None:
            self.anchors = np.asarray(anchors, np.float32).reshape(-1, 2)
        self.engine = Engine(self.model_name, self.num_classes, self.input_hw, self.batch, weights, self.anchors,
                             num_anchors=self.num_anchors, **self._kw)
        return self

    def load_weights(self, path: str, by_name: bool = False, anchors=None):
        """reference: model.load_weights(path) (code/yolo.py:87, code/yolo3/utils.py:390)."""
        return self.set_weights(load_checkpoint(path, self.netdef.weight_shapes), anchors)

    def __call__(self, x: torch.Tensor, copy: bool = True) -> List[torch.Tensor]:
        """``[y1, y2, y3]`` raw head logits, each [B, H/s, W/s, 3, C+5].  Like the reference, the returned tensors are
        fresh (the caller may keep them across calls); ``copy=False`` returns strided views of the engine's persistent
        output buffers instead, which the next forward / ``detect_*`` call / graph replay overwrites."""
        if self.engine is None:
            raise RuntimeError("weights not loaded: call load_weights()/set_weights() first")
        e = self.engine
        dst, u8 = e.slot_for(x)
        dst.copy_(x, non_blocking=True)
        e.run_network(0, u8)
        outs = e.raw_outputs()
        return [y.clone() for y in outs] if copy else outs


def yolov3_body(inputs, model_name, num_anchors, **kwargs):
    """reference code/yolo3/model.py:170.  ``inputs``: a (B,H,W,3) shape or a tensor of that shape
    (the reference passes a Keras Input).  kwargs are the GlobalParams overrides the reference
    accepts (``num_classes``, ``data_format``, ... ; ``drop_rate`` is dropped as in
    efficientnet.py:260-261) plus engine options (``micro_batch``, ``device``, ``pw_variant``,
    ``input_u8``)."""
    shape = tuple(inputs.shape) if hasattr(inputs, "shape") else tuple(inputs)
    if len(shape) != 4 or shape[3] != 3:
        raise ValueError("inputs must be (B,H,W,3) channels_last, got %s" % (shape,))
    kwargs = dict(kwargs)
    kwargs.pop("drop_rate", None)
    eng = {k: kwargs.pop(k) for k in ("micro_batch", "device", "pw_variant", "input_u8", "num_scales", "fuse_se",
                                      "lanes", "autotune", "fuse_up2", "fuse_dwpw", "fold_linear", "stack_pw") if k in kwargs}
    for k in kwargs:
        if k not in _GLOBAL_PARAM_FIELDS:  # namedtuple._replace raises ValueError in the reference
            raise ValueError("Got unexpected field names: %r" % [k])
    if kwargs.get("data_format", "channels_last") != "channels_last":
        raise ValueError("only data_format='channels_last' is supported (reference code/yolo.py:208)")
    if model_name not in _BACKBONES:
        raise ValueError("unknown model_name %r" % (model_name,))
    if int(num_anchors) != 3:
        # reference: num_anchors // num_scales per scale (code/yolo.py:214-216); every shipped anchor file gives 3
        raise ValueError("yolov3_body: this engine supports 3 anchors per scale, got %r" % (num_anchors,))
    return YoloBody(shape, model_name, num_anchors, kwargs.get("num_classes", 1000), **eng)


yolo_body = yolov3_body  # spelling used by BASELINE.json's north_star


# --------------------------------------------------------------------------------------
_pp_cache: Dict[Tuple, PostProcess] = {}


def _as_cells(t: torch.Tensor, num_classes: int) -> Tuple[torch.Tensor, int]:
    """Accepts [B,gh,gw,A,5+C] (contiguous or a padded-row strided view) -> (tensor, cell stride)."""
    if not t.is_cuda or t.dtype != torch.float32 or t.dim() != 5:
        raise ValueError("yolo outputs must be float32 CUDA tensors [B,gh,gw,A,5+C]")
    E = num_classes + 5
    if t.shape[4] != E or t.shape[3] != 3:
        raise ValueError("expected [...,3,%d], got %s" % (E, tuple(t.shape)))
    st = t.stride()
    ok = st[4] == 1 and st[3] == E and st[1] == t.shape[2] * st[2] and st[0] == t.shape[1] * st[1] and st[2] >= 3 * E
    if not ok:
        t = t.contiguous()
        st = t.stride()
    return t, st[2]


def yolo_head(feats: torch.Tensor, anchors, input_shape, calc_loss: bool = False):
    """reference code/yolo3/model.py:344.  ``feats`` [B,gh,gw,A,5+C] float32 CUDA, ``anchors`` [A,2] (w,h px),
    ``input_shape`` (h,w).  Returns ``(box_xy, box_wh, box_confidence, box_class_probs)`` or, with
    ``calc_loss=True``, ``(grid, box_xy, box_wh, box_confidence)`` (grid [gh,gw,1,2])."""
    anc = np.asarray(anchors, np.float32).reshape(-1, 2)
    if not feats.is_cuda or feats.dtype != torch.float32 or feats.dim() != 5 or feats.shape[3] != len(anc):
        raise ValueError("yolo_head expects a float32 CUDA tensor [B,gh,gw,%d,5+C]" % len(anc))
    B, gh, gw, A, E = (int(v) for v in feats.shape)
    f, ld = _as_cells_generic(feats, A, E)
    dev = feats.device
    anc_d = torch.from_numpy(anc).to(dev)
    xy = torch.empty(B, gh, gw, A, 2, dtype=torch.float32, device=dev)
    wh = torch.empty_like(xy)
    conf = torch.empty(B, gh, gw, A, 1, dtype=torch.float32, device=dev)
    cls = None if calc_loss else torch.empty(B, gh, gw, A, E - 5, dtype=torch.float32, device=dev)
    grid = torch.empty(gh, gw, 1, 2, dtype=torch.float32, device=dev) if calc_loss else None
    _lib.check(_lib.lib().yr_yolo_head(f.data_ptr(), ld, B, gh, gw, A, E - 5, anc_d.data_ptr(), int(input_shape[0]),
                                       int(input_shape[1]), xy.data_ptr(), wh.data_ptr(), conf.data_ptr(),
                                       cls.data_ptr() if cls is not None else None,
                                       grid.data_ptr() if grid is not None else None,
                                       torch.cuda.current_stream(dev).cuda_stream), "yr_yolo_head")
    if calc_loss:
        return grid, xy, wh, conf
    return xy, wh, conf, cls


def yolo_eval(yolo_outputs, anchors, num_scales, num_classes, image_shape, max_boxes=20, score_threshold=.6,
              iou_threshold=.5, zoom_outputs=None, sync=True):
    """reference code/yolo3/model.py:431.  Returns (boxes_ int32 [N,4] (ymin,xmin,ymax,xmax),
    scores_ f32 [N], classes_ int32 [N]) as CUDA tensors for a batch of one (as the reference);
    for B > 1 returns per-image lists (batch = independent per-image application)."""
    if zoom_outputs is not None:
        raise NotImplementedError("zoom_outputs is an unused experiment in the reference (model.py:408-417)")
    feats, lds = [], []
    for t in yolo_outputs[:num_scales]:
        t, ld = _as_cells(t, num_classes)
        feats.append(t)
        lds.append(ld)
    B = feats[0].shape[0]
    grids = tuple((int(t.shape[1]), int(t.shape[2])) for t in feats)
    key = (B, grids, num_classes, num_scales, max_boxes, feats[0].device.index, np.asarray(anchors).tobytes())
    pp = _pp_cache.get(key)
    if pp is None:
        if len(_pp_cache) > 8:
            _pp_cache.clear()
        pp = _pp_cache[key] = PostProcess(B, grids, num_classes, anchors, num_scales, max_boxes,
                                          device=feats[0].device)
    pp.set_image_shapes(image_shape)
    pp.run([t.data_ptr() for t in feats], lds, score_threshold, iou_threshold)
    if not sync:  # engine extension: leave the packed result on the device (PostProcess buffers), no host read
        return pp
    cnt = pp.out_count.cpu().numpy()
    if int(pp.status.item()) != 0:
        raise _lib.YrError("candidate list overflow")
    res = [(pp.out_boxes_i[b, :n].clone(), pp.out_scores[b, :n].clone(), pp.out_classes[b, :n].clone())
           for b, n in enumerate(cnt)]
    return res[0] if B == 1 else res


class YoloEval:
    """reference code/yolo3/model.py:494-526 (a Keras layer wrapper of yolo_eval)."""

    def __init__(self, anchors, num_scales, num_classes, max_boxes=20, score_threshold=.6, iou_threshold=.5, **kwargs):
        self.anchors, self.num_scales, self.num_classes = anchors, num_scales, num_classes
        self.max_boxes, self.score_threshold, self.iou_threshold = max_boxes, score_threshold, iou_threshold

    def __call__(self, yolo_outputs, image_shape, zoom_outputs=None):
        return yolo_eval(yolo_outputs, self.anchors, self.num_scales, self.num_classes, image_shape, self.max_boxes,
                         self.score_threshold, self.iou_threshold, zoom_outputs=zoom_outputs)

    call = __call__

    def get_config(self):
        return dict(anchors=self.anchors, num_scales=self.num_scales, num_classes=self.num_classes,
                    max_boxes=self.max_boxes, score_threshold=self.score_threshold,
                    iou_threshold=self.iou_threshold)


# --------------------------------------------------------------------------------------
class _YoloLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, yolo_output, y_true, loss_obj):
        parts, dlogits = loss_obj._run(y_true, yolo_output, need_grad=yolo_output.requires_grad)
        ctx.dlogits = dlogits
        loss_obj.last_parts = parts
        return parts[0] + parts[1] + parts[2]  # giou + confidence + class, model.py:669

    @staticmethod
    def backward(ctx, g):
        return (ctx.dlogits * g if ctx.dlogits is not None else None), None, None


class YoloLoss:
    """reference code/yolo3/model.py:585.  One instance per scale ``idx`` (strides 32,16,8).
    ``loss(y_true, yolo_output)`` returns a scalar CUDA tensor wired into torch autograd:
    ``.backward()`` delivers the fused kernel's analytic d(loss)/d(yolo_output)."""

    def __init__(self, idx, anchors, num_scales, ignore_thresh=.5, box_loss=BOX_LOSS.GIOU, print_loss=True):
        grid_steps = [32, 16, 8]
        anchor_masks = ANCHOR_MASK[-1 * num_scales:]
        self.idx, self.ignore_thresh, self.box_loss, self.print_loss = idx, ignore_thresh, box_loss, print_loss
        self.grid_step = grid_steps[idx]
        self.anchor = np.asarray(anchors, np.float32).reshape(-1, 2)[anchor_masks[idx]]
        if box_loss != BOX_LOSS.GIOU:
            # the reference's MSE branch references undefined names and would raise (model.py:672-690)
            raise NameError("BOX_LOSS.MSE is dead code in the reference (undefined names); only GIOU is supported")
        self.last_parts = None

    def _run(self, y_true: torch.Tensor, yolo_output: torch.Tensor, need_grad: bool):
        lib = _lib.lib()
        if not (yolo_output.is_cuda and y_true.is_cuda):
            raise _lib.YrError("YoloLoss needs CUDA tensors (no CPU fallback exists)")
        if yolo_output.dim() != 5 or tuple(y_true.shape) != tuple(yolo_output.shape):
            raise ValueError("YoloLoss: y_true %s and yolo_output %s must both be [B,gh,gw,A,5+C] with equal shapes"
                             % (tuple(y_true.shape), tuple(yolo_output.shape)))
        if y_true.device != yolo_output.device:
            raise ValueError("YoloLoss: y_true is on %s, yolo_output on %s" % (y_true.device, yolo_output.device))
        B, gh, gw, A, E = (int(v) for v in yolo_output.shape)
        if A > 3 or A != len(self.anchor) or E < 6:
            raise ValueError("YoloLoss: expected %d anchors per cell (<= 3) and 5+C >= 6 channels, got A=%d, 5+C=%d"
                             % (len(self.anchor), A, E))
        with torch.cuda.device(yolo_output.device):
            return self._launch(y_true, yolo_output, need_grad, B, gh, gw, A, E)

    def _launch(self, y_true, yolo_output, need_grad, B, gh, gw, A, E):
        lib = _lib.lib()
        logits, ldl = yolo_output.detach().contiguous(), A * E
        ytrue, ldt = _as_cells_generic(y_true.detach().to(torch.float32), A, E)
        p = YrLossParams()
        p.B, p.gh, p.gw, p.A, p.C = B, gh, gw, A, E - 5
        p.ld_logits, p.ld_true = ldl, ldt
        for k in range(A):
            p.anchors[k][0], p.anchors[k][1] = float(self.anchor[k][0]), float(self.anchor[k][1])
        p.input_h, p.input_w = gh * self.grid_step, gw * self.grid_step  # model.py:628
        p.ignore_thresh = float(self.ignore_thresh)
        p.max_true = B * gh * gw * A
        dev = logits.device
        st = torch.cuda.current_stream(dev).cuda_stream
        true_boxes = torch.empty(max(1, p.max_true), 4, dtype=torch.float32, device=dev)
        n_true = torch.zeros(1, dtype=torch.int32, device=dev)
        ws_bytes = int(lib.yr_yolo_loss_workspace(C.byref(p)))
        ws = torch.empty(max(4, ws_bytes // 4), dtype=torch.float32, device=dev)
        parts = torch.empty(4, dtype=torch.float32, device=dev)
        dl = None
        if need_grad:
            dl = torch.empty_like(logits)
        _lib.check(lib.yr_yolo_loss_gather_true(ytrue.data_ptr(), C.byref(p), true_boxes.data_ptr(),
                                                n_true.data_ptr(), st), "yr_yolo_loss_gather_true")
        _lib.check(lib.yr_yolo_loss(logits.data_ptr(), ytrue.data_ptr(), true_boxes.data_ptr(), n_true.data_ptr(),
                                    C.byref(p), parts.data_ptr(), dl.data_ptr() if dl is not None else None,
                                    ws.data_ptr(), ws.numel() * 4, st), "yr_yolo_loss")
        if self.print_loss:  # tf.print(idx, giou, conf, class, sum(ignore)), model.py:670-671
            v = parts.cpu().numpy()
            print("%d: %s %s %s %s" % (self.idx, v[0], v[1], v[2], v[3]))
        return parts, dl

    def __call__(self, y_true, yolo_output):
        return _YoloLossFn.apply(yolo_output, y_true, self)

    call = __call__


def _as_cells_generic(t: torch.Tensor, A: int, E: int):
    st = t.stride()
    ok = (t.dim() == 5 and st[4] == 1 and st[3] == E and st[1] == t.shape[2] * st[2]
          and st[0] == t.shape[1] * st[1] and st[2] >= A * E)
    if not ok:
        t = t.contiguous()
        st = t.stride()
    return t, st[2]


def yolo_loss(y_trues: Sequence[torch.Tensor], yolo_outputs: Sequence[torch.Tensor], anchors, num_scales=3,
              ignore_thresh=.5, box_loss=BOX_LOSS.GIOU, print_loss=False) -> torch.Tensor:
    """Functional form named by BASELINE.json: the sum over scales of the per-scale YoloLoss,
    exactly as AdvLossModel._compute_total_loss does (reference code/yolo3/train.py:11-16)."""
    loss = 0
    for idx, (yt, yo) in enumerate(zip(y_trues, yolo_outputs)):
        loss = loss + YoloLoss(idx, anchors, num_scales, ignore_thresh, box_loss, print_loss)(yt, yo)
    return loss


# --------------------------------------------------------------------------------------
class _FusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loss_obj, y_true, *outs):
        need = any(o.requires_grad for o in outs)
        parts, dls = loss_obj._run(y_true, outs, need)
        ctx.dls, ctx.n = dls, len(outs)
        return parts[:, :3].sum()  # sum over scales of giou + confidence + class (train.py:11-16, model.py:669)

    @staticmethod
    def backward(ctx, g):
        if ctx.dls is None:
            return (None, None) + (None,) * ctx.n
        return (None, None) + tuple(d * g for d in ctx.dls)


class FusedYoloLoss:
    """The training loss of the reference - ``sum(YoloLoss(idx, ...)(y_true[idx], yolo_outputs[idx]))`` over the scales
    (code/yolo3/model.py:585-671, code/yolo3/train.py:11-16) - in ONE kernel launch on a SPARSE y_true
    (``utils.encode_true_boxes_sparse``): no dense y_true tensors, no per-scale launches, workspaces allocated once,
    no host synchronisation, CUDA-graph capturable.  ``loss(y_true_sparse, [y1, y2, y3])`` returns the scalar total
    wired into autograd; ``last_parts`` [3,4] holds (giou, confidence, class, sum(ignore_mask)) per scale."""

    def __init__(self, anchors, num_scales=3, ignore_thresh=.5, box_loss=BOX_LOSS.GIOU):
        if box_loss != BOX_LOSS.GIOU:
            raise NameError("BOX_LOSS.MSE is dead code in the reference (undefined names); only GIOU is supported")
        self.anchors = np.asarray(anchors, np.float32).reshape(-1, 2)
        self.num_scales, self.ignore_thresh = int(num_scales), float(ignore_thresh)
        self._key, self._ws, self._parts, self._dl, self._p = None, None, None, None, None
        self.last_parts = None

    def _prepare(self, y_true, outs):
        key = (y_true.B, tuple(y_true.grids), y_true.num_classes, y_true.cap, tuple(tuple(o.shape) for o in outs),
               outs[0].device)
        if key == self._key:
            return
        p = YrLoss3Params()
        E = 5 + y_true.num_classes
        p.B, p.C, p.num_scales = y_true.B, y_true.num_classes, self.num_scales
        mask = ANCHOR_MASK[-self.num_scales:]
        steps = [32, 16, 8]
        for l in range(self.num_scales):
            gh, gw = y_true.grids[l]
            if tuple(outs[l].shape) != (y_true.B, gh, gw, 3, E):
                raise ValueError("yolo_outputs[%d] has shape %s, the sparse y_true expects %s"
                                 % (l, tuple(outs[l].shape), (y_true.B, gh, gw, 3, E)))
            p.gh[l], p.gw[l], p.ld_logits[l] = gh, gw, 3 * E
            for k in range(3):
                p.anchors[l][k][0] = float(self.anchors[mask[l][k]][0])
                p.anchors[l][k][1] = float(self.anchors[mask[l][k]][1])
        p.input_h, p.input_w = y_true.grids[0][0] * steps[0], y_true.grids[0][1] * steps[0]  # model.py:628
        p.ignore_thresh, p.max_records = self.ignore_thresh, y_true.cap
        dev = outs[0].device
        nbytes = int(_lib.lib().yr_yolo_loss3_workspace(C.byref(p)))
        if nbytes <= 0:
            raise _lib.YrError("yr_yolo_loss3: unsupported configuration")
        self._ws = torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device=dev)   # zeroed once: holds the arrival counter
        self._parts = torch.zeros(3, 4, dtype=torch.float32, device=dev)
        self._dl = [torch.empty(tuple(o.shape), dtype=torch.float32, device=dev) for o in outs[:self.num_scales]]
        self._p, self._key = p, key

    def _run(self, y_true, outs, need_grad: bool):
        if not all(o.is_cuda and o.dtype == torch.float32 for o in outs):
            raise _lib.YrError("FusedYoloLoss needs float32 CUDA tensors (no CPU fallback exists)")
        outs = [o.detach().contiguous() for o in outs[:self.num_scales]]
        self._prepare(y_true, outs)
        lib = _lib.lib()
        dev = outs[0].device
        with torch.cuda.device(dev):
            lp = (C.c_void_p * 3)(*[o.data_ptr() for o in outs] + [None] * (3 - len(outs)))
            mp = (C.c_void_p * 3)(*[m.data_ptr() for m in y_true.maps] + [None] * (3 - len(y_true.maps)))
            dp = (C.c_void_p * 3)(*[d.data_ptr() for d in self._dl] + [None] * (3 - len(self._dl))) if need_grad else None
            _lib.check(lib.yr_yolo_loss3(lp, mp, y_true.records.data_ptr(), y_true.counts.data_ptr(), C.byref(self._p),
                                         self._parts.data_ptr(), dp, self._ws.data_ptr(), self._ws.numel() * 4,
                                         torch.cuda.current_stream(dev).cuda_stream), "yr_yolo_loss3")
        self.last_parts = self._parts
        return self._parts, (self._dl if need_grad else None)

    def __call__(self, y_true, yolo_outputs):
        return _FusedLossFn.apply(self, y_true, *yolo_outputs[:self.num_scales])

    call = __call__
