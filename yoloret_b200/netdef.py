"""Network definition of YOLO-ReT as a flat list of fused device layers.

This is the host-side mirror of the reference graph builders:
  * ``yolov3_body``                         reference code/yolo3/model.py:170-342
  * ``rfcr_module`` / ``WeightedSum``       code/yolo3/model.py:117-168
  * ``make_last_layers_efficientnet_lite``  code/yolo3/model.py:91-115
  * ``MBConvBlock`` / ``SEBlock``           code/yolo3/efficientnet.py:406-438,467-536
  * ``EfficientNet``                        code/yolo3/efficientnet.py:203-267,364-388,611-677
  * Keras ``MobileNetV2`` behind ``mobilenet_v2``  code/yolo3/override.py:290-341

Instead of Keras layer objects it emits ``Layer`` records over ``View``s of
per-image NHWC buffers.  Differences from the reference that are pure layout:
every channel count is padded to a multiple of 8 (pad channels are exact
zeros), BatchNorm is folded into the preceding conv, Concatenate is realised by
producers writing into channel slices of one buffer, Add / SE-Multiply /
activations ride in the conv epilogue / operand loader.  Weight names are the
Keras auto-numbered names, so reference checkpoints load by name.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

ALIGN_C = 8


def pad_c(c: int) -> int:
    return (c + ALIGN_C - 1) // ALIGN_C * ALIGN_C


@dataclass
class Buf:
    name: str
    H: int
    W: int
    ld: int                 # floats per pixel
    full_batch: bool = False  # network outputs live for the whole batch, not one micro-batch


@dataclass
class View:
    buf: Buf
    off: int                          # channel offset inside the buffer
    segs: List[Tuple[int, int]]       # (logical, padded) channel counts, in order

    @property
    def C(self) -> int:               # padded channels
        return sum(p for _, p in self.segs)

    @property
    def Clog(self) -> int:
        return sum(l for l, _ in self.segs)

    @property
    def H(self):
        return self.buf.H

    @property
    def W(self):
        return self.buf.W


@dataclass
class Layer:
    kind: str                 # stem | pw | dw | resample | rfcr | se
    name: str
    inp: List[View]
    out: View
    act: str = "none"         # none | relu6 | swish
    conv: Optional[str] = None      # Keras layer holding kernel (/bias)
    bn: Optional[str] = None        # Keras BatchNormalization layer folded in
    k: int = 1
    stride: int = 1
    res: Optional[View] = None      # residual added after BN
    gate: Optional["Layer"] = None  # SE layer whose gate scales the A operand
    mode: Optional[str] = None      # resample: up2 | pool2 | pool4
    extra: Dict = field(default_factory=dict)
    flops: int = 0
    bytes_alg: int = 0        # algorithmic HBM bytes per image (SURVEY.md §8d formulas)


class _Namer:
    def __init__(self):
        self.counts: Dict[str, int] = {}

    def __call__(self, base):
        n = self.counts.get(base, 0)
        self.counts[base] = n + 1
        return base if n == 0 else "%s_%d" % (base, n)


def _make_divisible(v, divisor, min_value=None):
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def same_pad(size: int, k: int, s: int) -> Tuple[int, int]:
    """TF 'SAME': returns (output size, leading pad)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2


class NetDef:
    """Builds the layer list for one (backbone, classes, input size)."""

    def __init__(self, model_name: str, num_classes: int, input_hw: Tuple[int, int], num_anchors: int = 3):
        if input_hw[0] % 32 or input_hw[1] % 32:
            raise ValueError("input size must be a multiple of 32 (reference code/train.py:42)")
        self.model_name = model_name
        self.num_classes = num_classes
        self.num_anchors = num_anchors
        self.input_hw = tuple(input_hw)
        self.layers: List[Layer] = []
        self.bufs: List[Buf] = []
        self.namer = _Namer()
        self.weight_shapes: Dict[str, Tuple[int, ...]] = {}
        self.outputs: List[View] = []
        self._build()

    # ---- buffers ---------------------------------------------------------
    def new_buf(self, name, H, W, ld, full_batch=False) -> Buf:
        b = Buf("%02d_%s" % (len(self.bufs), name), H, W, ld, full_batch)
        self.bufs.append(b)
        return b

    def new_view(self, name, H, W, c_log) -> View:
        cp = pad_c(c_log)
        return View(self.new_buf(name, H, W, cp), 0, [(c_log, cp)])

    @staticmethod
    def slice_of(buf: Buf, off: int, c_log: int) -> View:
        return View(buf, off, [(c_log, pad_c(c_log))])

    # ---- layers ------------------------------------------------------------
    def _want(self, name, shape):
        self.weight_shapes[name] = tuple(shape)

    def _want_bn(self, bn, c):
        for leaf in ("gamma", "beta", "moving_mean", "moving_variance"):
            self._want("%s/%s" % (bn, leaf), (c,))

    def stem(self, cout, conv, bn, act) -> View:
        H, W = self.input_hw
        Ho, pt = same_pad(H, 3, 2)
        Wo, pl = same_pad(W, 3, 2)
        out = self.new_view(conv, Ho, Wo, cout)
        inp = View(Buf("input", H, W, 3, True), 0, [(3, 3)])
        self._want(conv + "/kernel", (3, 3, 3, cout))
        self._want_bn(bn, cout)
        L = Layer("stem", conv, [inp], out, act, conv, bn, 3, 2, extra=dict(pad_t=pt, pad_l=pl))
        L.flops = 2 * Ho * Wo * 27 * cout
        L.bytes_alg = (H * W * 3 + Ho * Wo * cout) * 4 + (27 * cout + 2 * cout) * 4
        self.layers.append(L)
        return out

    def pw(self, x: View, cout, conv, bn=None, act="none", out: Optional[View] = None, res: Optional[View] = None,
           gate: Optional[Layer] = None, name=None) -> View:
        if out is None:
            out = self.new_view(conv, x.H, x.W, cout)
        assert out.Clog == cout and out.H == x.H and out.W == x.W
        self._want(conv + "/kernel", (1, 1, x.Clog, cout))
        if bn:
            self._want_bn(bn, cout)
        L = Layer("pw", name or conv, [x], out, act, conv, bn, 1, 1, res=res, gate=gate)
        L.flops = 2 * x.H * x.W * x.Clog * cout
        L.bytes_alg = x.H * x.W * (x.Clog + cout) * 4 + (x.Clog * cout + 2 * cout) * 4
        self.layers.append(L)
        return out

    def dw(self, x: View, k, stride, conv, bn, act) -> View:
        Ho, pt = same_pad(x.H, k, stride)
        Wo, pl = same_pad(x.W, k, stride)
        out = self.new_view(conv, Ho, Wo, x.Clog)
        self._want(conv + "/depthwise_kernel", (k, k, x.Clog, 1))
        self._want_bn(bn, x.Clog)
        L = Layer("dw", conv, [x], out, act, conv, bn, k, stride, extra=dict(pad_t=pt, pad_l=pl))
        L.flops = 2 * Ho * Wo * k * k * x.Clog
        L.bytes_alg = (x.H * x.W * x.Clog + Ho * Wo * x.Clog) * 4 + (k * k * x.Clog + 2 * x.Clog) * 4
        self.layers.append(L)
        return out

    def resample(self, x: View, mode: str, out: Optional[View] = None, name="resample") -> View:
        if mode == "up2":
            Ho, Wo = x.H * 2, x.W * 2
        elif mode == "pool2":
            Ho, Wo = x.H // 2, x.W // 2
        elif mode == "pool4":
            Ho, Wo = x.H // 4, x.W // 4
        else:
            raise ValueError(mode)
        if out is None:
            out = self.new_view(name, Ho, Wo, x.Clog)
        assert out.C == x.C and out.H == Ho and out.W == Wo
        L = Layer("resample", name, [x], out, mode=mode)
        L.bytes_alg = (x.H * x.W * x.Clog + Ho * Wo * x.Clog) * 4
        self.layers.append(L)
        return out

    def se(self, x: View, reduced: int, conv1: str, conv2: str) -> Layer:
        f = x.Clog
        self._want(conv1 + "/kernel", (1, 1, f, reduced))
        self._want(conv1 + "/bias", (reduced,))
        self._want(conv2 + "/kernel", (1, 1, reduced, f))
        self._want(conv2 + "/bias", (f,))
        gate = View(self.new_buf(conv1 + "_gate", 1, 1, x.C), 0, list(x.segs))
        L = Layer("se", conv1, [x], gate, extra=dict(conv1=conv1, conv2=conv2, reduced=reduced))
        L.flops = x.H * x.W * f + 4 * f * reduced
        L.bytes_alg = x.H * x.W * f * 4 + (2 * f * reduced + f + reduced) * 4
        self.layers.append(L)
        return L

    # ---- composite blocks ----------------------------------------------------
    def mbconv(self, x: View, k, stride, expand, in_f, out_f, se_ratio, act, out: Optional[View] = None) -> View:
        """MBConvBlock, efficientnet.py:467-536 (inference path)."""
        nm = self.namer
        inp = x
        if expand != 1:
            x = self.pw(x, in_f * expand, nm("conv2d"), nm("batch_normalization"), act)
        x = self.dw(x, k, stride, nm("depthwise_conv2d"), nm("batch_normalization"), act)
        gate = None
        if se_ratio is not None and 0 < se_ratio <= 1:
            gate = self.se(x, max(1, int(in_f * se_ratio)), nm("conv2d"), nm("conv2d"))
        res = inp if (stride == 1 and in_f == out_f) else None
        return self.pw(x, out_f, nm("conv2d"), nm("batch_normalization"), "none", out=out, res=res, gate=gate)

    def head_stage(self, x: View, filters, out_f, orphan_y: bool, x_out: Optional[View] = None,
                   y_out: Optional[View] = None):
        """make_last_layers_efficientnet_lite, model.py:91-115."""
        nm = self.namer
        x = self.pw(x, filters, nm("conv2d"), nm("batch_normalization"), "relu6")
        x = self.mbconv(x, 3, 1, 1, filters, out_f, 0.25, "swish", out=x_out)
        yname = nm("conv2d")  # created even when panet discards it (model.py:240-241)
        y = None if orphan_y else self.pw(x, out_f, yname, None, "none", out=y_out)
        return x, y

    # ---- backbones -------------------------------------------------------------
    def _mobilenet_v2(self, alpha, tap_out: Dict[str, View]):
        """Keras MobileNetV2(alpha) truncated at block_15_add; taps per model.py:186-189.
        ``tap_out``: views (slices of concat buffers) the tapped adds must write into."""
        first = _make_divisible(32 * alpha, 8)
        x = self.stem(first, "Conv1", "bn_Conv1", "relu6")
        cfg = [(16, 1, 1), (24, 2, 6), (24, 1, 6), (32, 2, 6), (32, 1, 6), (32, 1, 6), (64, 2, 6), (64, 1, 6),
               (64, 1, 6), (64, 1, 6), (96, 1, 6), (96, 1, 6), (96, 1, 6), (160, 2, 6), (160, 1, 6), (160, 1, 6)]
        taps = {}
        for bid, (f, s, e) in enumerate(cfg):
            in_c = x.Clog
            pw_f = _make_divisible(int(f * alpha), 8)
            prefix = "block_%d_" % bid if bid else "expanded_conv_"
            inp = x
            if bid:
                x = self.pw(x, e * in_c, prefix + "expand", prefix + "expand_BN", "relu6")
            x = self.dw(x, 3, s, prefix + "depthwise", prefix + "depthwise_BN", "relu6")
            has_add = in_c == pw_f and s == 1
            key = prefix + "add"
            out = tap_out[key](x.H, x.W, pw_f) if (has_add and key in tap_out) else None
            x = self.pw(x, pw_f, prefix + "project", prefix + "project_BN", "none", out=out,
                        res=inp if has_add else None, name=prefix + "project" + ("+add" if has_add else ""))
            if has_add:
                taps[key] = x
        return taps["block_15_add"], taps["block_12_add"], taps["block_5_add"], taps["block_2_add"]

    _EFF_BLOCKS = [(1, 3, 1, 1, 32, 16, 0.25), (2, 3, 2, 6, 16, 24, 0.25), (2, 5, 2, 6, 24, 40, 0.25),
                   (3, 3, 2, 6, 40, 80, 0.25), (3, 5, 1, 6, 80, 112, 0.25), (4, 5, 2, 6, 112, 192, 0.25),
                   (1, 3, 1, 6, 192, 320, 0.25)]

    def _efficientnet(self, width, depth, lite, tap_out):
        """EfficientNet body (efficientnet.py:611-677); taps = end of stages 6,5,3,2 (model.py:213-216).
        lite=True: the derived 'EfficientNet-lite0' of BASELINE.json config 3 (no SE, ReLU6)."""
        def rf(f):
            f *= width
            nf = max(8, int(f + 4) // 8 * 8)
            if nf < 0.9 * f:
                nf += 8
            return int(nf)
        nm = self.namer
        act = "relu6" if lite else "swish"
        x = self.stem(rf(32), nm("conv2d"), nm("batch_normalization"), act)
        stage_end = {5: "s6", 4: "s5", 2: "s3", 1: "s2"}
        taps = {}
        for si, (r, k, s, e, i, o, se) in enumerate(self._EFF_BLOCKS):
            i, o, r = rf(i), rf(o), int(math.ceil(depth * r))
            se = None if lite else se
            if si == 6:
                # stage 7 + the 1280-wide head conv exist in EfficientNet() but are unreachable from
                # the detector outputs: Keras neither saves nor executes them; only the auto-numbering
                # advances.
                for _ in range(r):
                    for _ in range((1 if e != 1 else 0) + (0 if se is None else 2) + 1):
                        nm("conv2d")
                    nm("depthwise_conv2d")
                    for _ in range(3 if e != 1 else 2):
                        nm("batch_normalization")
                nm("conv2d")
                nm("batch_normalization")
                break
            for rep in range(r):
                last = rep == r - 1
                key = stage_end.get(si) if last else None
                out = tap_out[key](x.H // (s if rep == 0 else 1), x.W // (s if rep == 0 else 1), o) \
                    if (key and key in tap_out) else None
                x = self.mbconv(x, k, s if rep == 0 else 1, e, i if rep == 0 else o, o, se, act, out=out)
                if key:
                    taps[key] = x
        return taps["s6"], taps["s5"], taps["s3"], taps["s2"]

    # ---- whole graph -------------------------------------------------------------
    def _build(self):
        H, W = self.input_hw
        out_f = self.num_anchors * (self.num_classes + 5)
        outp = pad_c(out_f)
        g32, g16, g8 = (H // 32, W // 32), (H // 16, W // 16), (H // 8, W // 8)
        nm = self.namer

        # concat buffers are allocated lazily once the tap widths are known
        cat = {}

        def tap_slot(which, grid, tail):
            def make(h, w, c):
                assert (h, w) == grid, (which, h, w, grid)
                buf = self.new_buf(which, h, w, pad_c(c) + tail)
                cat[which] = buf
                return self.slice_of(buf, 0, c)
            return make

        # CAT1 = [b1 | pool2(bc)], CAT2r = [b2 | bc], CAT3r = [b3 | up2(bc)]  (model.py:164-166)
        if self.model_name in ("mobilenetv2x75", "mobilenetv2x14"):
            alpha = 0.75 if self.model_name.endswith("x75") else 1.4
            slots = {"block_15_add": tap_slot("cat1", g32, 96)}
            b1, b2, b3, b4 = self._mobilenet_v2(alpha, slots)
        elif self.model_name in ("efficientnetb3", "efficientnetlite0"):
            lite = self.model_name == "efficientnetlite0"
            width, depth = (1.0, 1.0) if lite else (1.2, 1.4)
            slots = {"s6": tap_slot("cat1", g32, 96)}
            b1, b2, b3, b4 = self._efficientnet(width, depth, lite, slots)
        else:
            raise ValueError("unknown backbone %r" % self.model_name)
        self.taps = dict(b1=b1, b2=b2, b3=b3, b4=b4)

        # --- RFCR (model.py:146-168); b4's MaxPool4 (model.py:190) is fused in the kernel
        rf_names = [nm("conv2d") for _ in range(4)]
        for cn, b in zip(rf_names, (b1, b2, b3, b4)):
            self._want("%s/kernel" % cn, (1, 1, b.Clog, 48))
        self._want("weighted_sum/alpha", (4,))
        bc0 = self.new_view("rfcr_sum", g16[0], g16[1], 48)
        L = Layer("rfcr", "rfcr_fuse", [b1, b2, b3, b4], bc0, extra=dict(convs=rf_names))
        L.flops = 2 * 48 * (g32[0] * g32[1] * b1.Clog + g16[0] * g16[1] * b2.Clog + g8[0] * g8[1] * b3.Clog
                            + g16[0] * g16[1] * b4.Clog)
        L.bytes_alg = 4 * (g32[0] * g32[1] * b1.Clog + g16[0] * g16[1] * b2.Clog + g8[0] * g8[1] * b3.Clog
                           + b4.H * b4.W * b4.Clog + g16[0] * g16[1] * 48)
        self.layers.append(L)
        x = self.dw(bc0, 5, 1, nm("depthwise_conv2d"), nm("batch_normalization"), "relu6")

        # CAT2 = [up2(block_20) 256 | b2 | bc 96]   (model.py:254-255 with b2' = [b2 | bc])
        cat2 = self.new_buf("cat2", g16[0], g16[1], 256 + b2.C + 96)
        bc = self.pw(x, 96, nm("conv2d"), nm("batch_normalization"), "relu6",
                     out=self.slice_of(cat2, 256 + b2.C, 96))
        # b2 was produced into its own buffer; copy-free alternative needs the tap to know cat2 up
        # front, so the tap is re-homed here: its producer layer is redirected into the cat2 slice.
        b2_slot = self.slice_of(cat2, 256, b2.Clog)
        self._rehome(b2, b2_slot)
        b2 = b2_slot
        cat1 = cat["cat1"]
        self.resample(bc, "pool2", out=self.slice_of(cat1, b1.C, 96), name="rfcr_pool2")
        # CAT3 = [up2(block_24) 128 | b3 | up2(bc) 96]   (model.py:274-275 with b3' = [b3 | up2(bc)])
        cat3 = self.new_buf("cat3", g8[0], g8[1], 128 + b3.C + 96)
        b3_slot = self.slice_of(cat3, 128, b3.Clog)
        self._rehome(b3, b3_slot)
        b3 = b3_slot
        self.resample(bc, "up2", out=self.slice_of(cat3, 128 + b3.C, 96), name="rfcr_up2")

        # --- top-down (model.py:238-281)
        cat6 = self.new_buf("cat6", g32[0], g32[1], 256 + outp)     # [pool2(conv2d_31) | c1]
        cat5 = self.new_buf("cat5", g16[0], g16[1], 128 + outp)     # [pool2(conv2d_25) | c2]
        s1_in = View(cat1, 0, [(b1.Clog, b1.C), (96, 96)])
        c1, _ = self.head_stage(s1_in, 512, out_f, True, x_out=self.slice_of(cat6, 256, out_f))
        x = self.pw(c1, 256, "block_20_conv", "block_20_BN", "relu6")
        self.resample(x, "up2", out=self.slice_of(cat2, 0, 256), name="up2_block_20")
        s2_in = View(cat2, 0, [(256, 256), (b2.Clog, b2.C), (96, 96)])
        c2, _ = self.head_stage(s2_in, 256, out_f, True, x_out=self.slice_of(cat5, 128, out_f))
        x = self.pw(c2, 128, "block_24_conv", "block_24_BN", "relu6")
        self.resample(x, "up2", out=self.slice_of(cat3, 0, 128), name="up2_block_24")
        s3_in = View(cat3, 0, [(128, 128), (b3.Clog, b3.C), (96, 96)])
        c3, _ = self.head_stage(s3_in, 128, out_f, True)

        # --- bottom-up (model.py:283-323); outputs are full-batch buffers
        ybuf = [self.new_buf("y%d" % (i + 1), g[0], g[1], outp, True) for i, g in enumerate((g32, g16, g8))]
        x, y3 = self.head_stage(c3, 128, out_f, False, y_out=self.slice_of(ybuf[2], 0, out_f))
        x = self.pw(x, 128, nm("conv2d"), nm("batch_normalization"), "relu6")
        self.resample(x, "pool2", out=self.slice_of(cat5, 0, 128), name="pool2_bu8")
        s5_in = View(cat5, 0, [(128, 128), (out_f, outp)])
        x, y2 = self.head_stage(s5_in, 256, out_f, False, y_out=self.slice_of(ybuf[1], 0, out_f))
        x = self.pw(x, 256, nm("conv2d"), nm("batch_normalization"), "relu6")
        self.resample(x, "pool2", out=self.slice_of(cat6, 0, 256), name="pool2_bu16")
        s6_in = View(cat6, 0, [(256, 256), (out_f, outp)])
        x, y1 = self.head_stage(s6_in, 512, out_f, False, y_out=self.slice_of(ybuf[0], 0, out_f))
        self.outputs = [y1, y2, y3]

    def _rehome(self, old: View, new: View):
        """Redirects every layer that produced / consumed ``old`` to the concat slice ``new``."""
        assert old.C == new.C
        dead = old.buf
        for L in self.layers:
            if L.out.buf is dead:
                L.out = View(new.buf, new.off + L.out.off, L.out.segs)
            L.inp = [View(new.buf, new.off + v.off, v.segs) if v.buf is dead else v for v in L.inp]
            if L.res is not None and L.res.buf is dead:
                L.res = View(new.buf, new.off + L.res.off, L.res.segs)
        self.bufs.remove(dead)
        for k, v in list(getattr(self, "taps", {}).items()):
            if v.buf is dead:
                self.taps[k] = View(new.buf, new.off + v.off, v.segs)

    # ---- graph-level folding ---------------------------------------------------------
    def fold_linear_pairs(self) -> List[str]:
        """Folds a LINEAR 1x1 conv (+BN, no activation, no residual) into the 1x1 convs that are its only readers.

        Every head stage ends in such a conv - the MBConv project conv + BN, `x` of
        make_last_layers_efficientnet_lite (reference code/yolo3/model.py:91-115) - and in the bottom-up half of
        yolov3_body that `x` is read by 1x1 convs only: the `y` conv (model.py:110-114) and the next 1x1 + BN + ReLU6
        (model.py:296-305, 309-318), or the first conv of the next stage (c3, model.py:283-295).  With A: t = x W_A + b_A
        and B: act(t W_B + b_B), B(A(x)) = act(x (W_A W_B) + (b_A W_B + b_B)): the 255-channel tensor t - at 52x52 the
        widest tensor of the head - is never written or read, and the flops drop too (128 -> 255 -> {255, 128} becomes
        128 -> {255, 128}).  Same function in exact arithmetic; like the BatchNorm folding it is done once on the
        host in float64.  The folded conv inherits A's input view and SE gate.  Returns the names of the folded layers."""
        done: List[str] = []
        changed = True
        while changed:
            changed = False
            for A in list(self.layers):
                if A.kind != "pw" or A.act != "none" or A.res is not None or "fold_from" in A.extra:
                    continue
                buf = A.out.buf
                if buf.full_batch or A.out.off != 0 or buf.ld != A.out.C:
                    continue
                if sum(1 for L in self.layers if L.out.buf is buf) != 1:
                    continue
                cons = [L for L in self.layers if any(v.buf is buf for v in L.inp) or (L.res is not None and L.res.buf is buf)]
                if not cons or not all(L.kind == "pw" and L.gate is None and len(L.inp) == 1 and L.inp[0].buf is buf
                                       and L.inp[0].off == 0 and L.inp[0].C == A.out.C and L.res is None
                                       and "fold_from" not in L.extra for L in cons):
                    continue
                ka, na = A.inp[0].Clog, A.out.Clog
                before = (ka + na) + sum(na + L.out.Clog for L in cons)
                after = sum(ka + L.out.Clog for L in cons)
                if after >= before:
                    continue
                x = A.inp[0]
                for L in cons:
                    L.extra["fold_from"] = A
                    L.inp = [x]
                    L.gate = A.gate
                    L.name = "%s*%s" % (A.name, L.name)
                    cout = L.out.Clog
                    L.flops = 2 * x.H * x.W * x.Clog * cout
                    L.bytes_alg = x.H * x.W * (x.Clog + cout) * 4 + (x.Clog * cout + 2 * cout) * 4
                self.layers.remove(A)
                self.bufs.remove(buf)
                done.append(A.name)
                changed = True
                break
        return done

    # ---- accounting ------------------------------------------------
    def totals(self) -> Dict[str, float]:
        t: Dict[str, float] = {"flops": 0, "bytes": 0}
        for L in self.layers:
            t["flops"] += L.flops
            t["bytes"] += L.bytes_alg
            t["flops_" + L.kind] = t.get("flops_" + L.kind, 0) + L.flops
            t["bytes_" + L.kind] = t.get("bytes_" + L.kind, 0) + L.bytes_alg
        return t
