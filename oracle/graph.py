"""torch-CPU restatement of the reference network graph (TEST INFRASTRUCTURE).

Follows, function by function:
  * ``yolov3_body``                          reference code/yolo3/model.py:170-342
  * ``rfcr_module`` / ``WeightedSum``        code/yolo3/model.py:117-168
  * ``MobilenetSeparableConv2D``             code/yolo3/model.py:14-30
  * ``make_last_layers_efficientnet_lite``   code/yolo3/model.py:91-115
  * ``MBConvBlock`` / ``SEBlock`` / ``Swish`` code/yolo3/efficientnet.py:327-331,406-438,467-536
  * ``EfficientNet`` (B0..B7 scaling)        code/yolo3/efficientnet.py:203-267,364-388,611-710
  * ``mobilenet_v2``                         code/yolo3/override.py:290-341 -> tf.keras.applications.MobileNetV2
    (third-party, un-vendored, unpinned: its published graph is restated here;
    de-facto pin = the shipped checkpoint's layer names / shapes)

Layer *names* are generated exactly as Keras auto-numbers them in creation
order, so the shipped ``.h5`` loads by name.  Tensors are NHWC at the
interface; NCHW inside (torch conv).  fp32 by default, fp64 on request for
error-budget studies.  Parity is unpinned by the reference (see
oracle/__init__.py).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # every BN in the graph (Keras default; override.py:207 ignores the kwarg)


# --------------------------------------------------------------------------
# Keras-style auto naming
# --------------------------------------------------------------------------
class Namer:
    """Reproduces Keras' per-class auto numbering: conv2d, conv2d_1, ..."""

    def __init__(self):
        self.counts: Dict[str, int] = {}

    def __call__(self, base: str) -> str:
        n = self.counts.get(base, 0)
        self.counts[base] = n + 1
        return base if n == 0 else "%s_%d" % (base, n)


def _make_divisible(v, divisor, min_value=None):
    # keras.applications.mobilenet_v2._make_divisible (same as model.py:32-39)
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


# --------------------------------------------------------------------------
# Weight spec: (name, kind, shape) in creation order, for a given config
# --------------------------------------------------------------------------
class SpecRecorder:
    """Stand-in weight source that records the shapes the graph asks for."""

    def __init__(self):
        self.spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()


class Net:
    """Executes the graph either on real tensors or in shape-recording mode."""

    def __init__(self, weights: Optional[Dict[str, np.ndarray]], dtype=torch.float32, record=False):
        self.w = weights
        self.dtype = dtype
        self.record = record
        self.spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
        self.namer = Namer()
        self.taps: Dict[str, torch.Tensor] = {}

    # ---- weight access ---------------------------------------------------
    def get(self, name: str, shape: Tuple[int, ...]) -> torch.Tensor:
        self.spec[name] = tuple(shape)
        if self.record:
            return torch.zeros(shape, dtype=self.dtype)
        arr = self.w[name]
        if tuple(arr.shape) != tuple(shape):
            raise ValueError("weight %s: have %s want %s" % (name, arr.shape, shape))
        return torch.from_numpy(np.ascontiguousarray(arr)).to(self.dtype)

    # ---- primitive layers (NCHW tensors) -----------------------------------
    @staticmethod
    def _same_pad(x, k, s):
        h, w = x.shape[2], x.shape[3]
        oh, ow = -(-h // s), -(-w // s)
        ph = max((oh - 1) * s + k - h, 0)
        pw = max((ow - 1) * s + k - w, 0)
        if ph or pw:
            x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        return x

    def conv(self, x, name, cout, k=1, s=1, bias=False):
        cin = x.shape[1]
        w = self.get(name + "/kernel", (k, k, cin, cout)).permute(3, 2, 0, 1).contiguous()
        b = self.get(name + "/bias", (cout,)) if bias else None
        return F.conv2d(self._same_pad(x, k, s), w, b, stride=s)

    def dwconv(self, x, name, k=3, s=1):
        c = x.shape[1]
        w = self.get(name + "/depthwise_kernel", (k, k, c, 1)).permute(2, 3, 0, 1).contiguous()
        return F.conv2d(self._same_pad(x, k, s), w, None, stride=s, groups=c)

    def bn(self, x, name):
        c = x.shape[1]
        g = self.get(name + "/gamma", (c,))
        b = self.get(name + "/beta", (c,))
        m = self.get(name + "/moving_mean", (c,))
        v = self.get(name + "/moving_variance", (c,))
        if self.record:
            return x
        return F.batch_norm(x, m, v, g, b, training=False, eps=BN_EPS)

    @staticmethod
    def relu6(x):
        return torch.clamp(x, 0.0, 6.0)

    @staticmethod
    def swish(x):  # efficientnet.py:327-331
        return x * torch.sigmoid(x)

    @staticmethod
    def up2(x):  # tf.keras.layers.UpSampling2D() default: nearest x2
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)

    @staticmethod
    def maxpool(x, s):  # downsample_layer, model.py:139-144 (pool=stride=s, 'valid')
        return F.max_pool2d(x, s, s)


# --------------------------------------------------------------------------
# Backbones
# --------------------------------------------------------------------------
def mobilenet_v2(net: Net, x, alpha: float):
    """tf.keras.applications.MobileNetV2(alpha, include_top=False) up to block_15_add.

    Reference call: code/yolo3/model.py:180,193 via code/yolo3/override.py:339.
    Returns the four taps (block_15_add, block_12_add, block_5_add, block_2_add).
    """
    first = _make_divisible(32 * alpha, 8)
    x = net.conv(x, "Conv1", first, k=3, s=2)
    x = net.relu6(net.bn(x, "bn_Conv1"))

    def inverted_res_block(x, expansion, stride, filters, block_id):
        in_ch = x.shape[1]
        pw_filters = _make_divisible(int(filters * alpha), 8)
        prefix = "block_%d_" % block_id if block_id else "expanded_conv_"
        inp = x
        if block_id:
            x = net.conv(x, prefix + "expand", expansion * in_ch)
            x = net.relu6(net.bn(x, prefix + "expand_BN"))
        x = net.dwconv(x, prefix + "depthwise", k=3, s=stride)  # ZeroPadding(correct_pad)+valid == SAME
        x = net.relu6(net.bn(x, prefix + "depthwise_BN"))
        x = net.conv(x, prefix + "project", pw_filters)
        x = net.bn(x, prefix + "project_BN")
        if in_ch == pw_filters and stride == 1:
            x = inp + x
            net.taps[prefix + "add"] = x
        return x

    cfg = [  # (filters, stride, expansion)
        (16, 1, 1),
        (24, 2, 6), (24, 1, 6),
        (32, 2, 6), (32, 1, 6), (32, 1, 6),
        (64, 2, 6), (64, 1, 6), (64, 1, 6), (64, 1, 6),
        (96, 1, 6), (96, 1, 6), (96, 1, 6),
        (160, 2, 6), (160, 1, 6), (160, 1, 6),
    ]
    for bid, (f, s, e) in enumerate(cfg):
        x = inverted_res_block(x, e, s, f, bid)
    t = net.taps
    return t["block_15_add"], t["block_12_add"], t["block_5_add"], t["block_2_add"]


_EFFNET_BLOCKS = [  # efficientnet.py:208-216  (r, k, s, e, i, o, se)
    (1, 3, 1, 1, 32, 16, 0.25),
    (2, 3, 2, 6, 16, 24, 0.25),
    (2, 5, 2, 6, 24, 40, 0.25),
    (3, 3, 2, 6, 40, 80, 0.25),
    (3, 5, 1, 6, 80, 112, 0.25),
    (4, 5, 2, 6, 112, 192, 0.25),
    (1, 3, 1, 6, 192, 320, 0.25),
]
_EFFNET_SCALE = {"efficientnetb0": (1.0, 1.0), "efficientnetb3": (1.2, 1.4)}


def _round_filters(filters, width, divisor=8):  # efficientnet.py:364-380
    filters *= width
    new_filters = max(divisor, int(filters + divisor / 2) // divisor * divisor)
    if new_filters < 0.9 * filters:
        new_filters += divisor
    return int(new_filters)


def _round_repeats(repeats, depth):  # efficientnet.py:383-388
    return int(math.ceil(depth * repeats))


def se_block(net: Net, x, input_filters, se_ratio):
    """SEBlock, efficientnet.py:406-438 (squeeze width from the *block input* filters)."""
    reduced = max(1, int(input_filters * se_ratio))
    f = x.shape[1]
    s = x.mean(dim=(2, 3), keepdim=True)
    s = net.swish(net.conv(s, net.namer("conv2d"), reduced, bias=True))
    s = torch.sigmoid(net.conv(s, net.namer("conv2d"), f, bias=True))
    return s * x


def mbconv_block(net: Net, x, k, stride, expand, in_f, out_f, se_ratio, act="swish", id_skip=True):
    """MBConvBlock, efficientnet.py:467-536 (inference: DropConnect is identity)."""
    a = net.swish if act == "swish" else net.relu6
    inp = x
    filters = in_f * expand
    if expand != 1:
        x = net.conv(x, net.namer("conv2d"), filters)
        x = a(net.bn(x, net.namer("batch_normalization")))
    x = net.dwconv(x, net.namer("depthwise_conv2d"), k=k, s=stride)
    x = a(net.bn(x, net.namer("batch_normalization")))
    if se_ratio is not None and 0 < se_ratio <= 1:
        x = se_block(net, x, in_f, se_ratio)
    x = net.conv(x, net.namer("conv2d"), out_f)
    x = net.bn(x, net.namer("batch_normalization"))
    if id_skip and stride == 1 and in_f == out_f:
        x = x + inp
    return x


def efficientnet(net: Net, x, model_name: str, lite: bool = False):
    """EfficientNet(include_top=False) body, efficientnet.py:611-677, with the
    taps ``yolov3_body`` takes for B3 (model.py:213-216): end of stages 6,5,3,2.

    ``lite=True`` is the *derived* "EfficientNet-lite0" of BASELINE.json config 3
    (no such backbone in the reference): B0 widths, no SE, ReLU6.
    """
    width, depth = _EFFNET_SCALE["efficientnetb0" if lite else model_name]
    act = "relu6" if lite else "swish"
    a = net.relu6 if lite else net.swish
    x = net.conv(x, net.namer("conv2d"), _round_filters(32, width), k=3, s=2)
    x = a(net.bn(x, net.namer("batch_normalization")))
    stage_out = []
    for si, (r, k, s, e, i, o, se) in enumerate(_EFFNET_BLOCKS):
        i, o, r = _round_filters(i, width), _round_filters(o, width), _round_repeats(r, depth)
        se = None if lite else se
        if si == 6:
            # stage 7 and the 1280-wide head conv are created by EfficientNet() but are not
            # reachable from [y1,y2,y3] (taps end at stage 6), so Keras neither saves nor runs
            # them; they only advance the auto-numbering.
            for _ in range(r):
                for _ in range((1 if e != 1 else 0) + (0 if se is None else 2) + 1):
                    net.namer("conv2d")
                net.namer("depthwise_conv2d")
                for _ in range(3 if e != 1 else 2):
                    net.namer("batch_normalization")
            net.namer("conv2d")
            net.namer("batch_normalization")
            break
        x = mbconv_block(net, x, k, s, e, i, o, se, act)
        for _ in range(r - 1):
            x = mbconv_block(net, x, k, 1, e, o, o, se, act)
        stage_out.append(x)
    # stages are 1-indexed in the reference's add_N bookkeeping: taps = stages 6,5,3,2
    return stage_out[5], stage_out[4], stage_out[2], stage_out[1]


# --------------------------------------------------------------------------
# RFCR + heads
# --------------------------------------------------------------------------
def rfcr_module(net: Net, b1, b2, b3, b4):
    """model.py:146-168.  b4 arrives already max-pooled by 4 (model.py:190)."""
    b1c = net.conv(b1, net.namer("conv2d"), 48)
    b2c = net.conv(b2, net.namer("conv2d"), 48)
    b3c = net.conv(b3, net.namer("conv2d"), 48)
    b4c = net.conv(b4, net.namer("conv2d"), 48)
    a = net.get("weighted_sum/alpha", (4,))
    # WeightedSum.call, model.py:133-134 (left-to-right adds)
    bc = a[0] * net.up2(b1c) + a[1] * b2c + a[2] * net.maxpool(b3c, 2) + a[3] * b4c
    # MobilenetSeparableConv2D(96, 5x5, use_bias=False, 'same'), model.py:14-30
    bc = net.dwconv(bc, net.namer("depthwise_conv2d"), k=5, s=1)
    bc = net.relu6(net.bn(bc, net.namer("batch_normalization")))
    bc = net.conv(bc, net.namer("conv2d"), 96)
    bc = net.relu6(net.bn(bc, net.namer("batch_normalization")))
    net.taps["rfcr_bc"] = bc
    o1 = torch.cat([b1, net.maxpool(bc, 2)], dim=1)
    o2 = torch.cat([b2, bc], dim=1)
    o3 = torch.cat([b3, net.up2(bc)], dim=1)
    return o1, o2, o3


def head_stage(net: Net, x, filters, out_filters, orphan_y: bool):
    """make_last_layers_efficientnet_lite, model.py:91-115, with
    BlockArgs(k=3, e=1, se=.25, input_filters=filters, output_filters=A*(C+5)).

    ``orphan_y``: the top-down stages create a y-conv that ``panet=True`` then
    discards (model.py:240-241); it still consumes a Keras auto-name."""
    x = net.conv(x, net.namer("conv2d"), filters)
    x = net.relu6(net.bn(x, net.namer("batch_normalization")))
    x = mbconv_block(net, x, 3, 1, 1, filters, out_filters, 0.25, "swish")
    yname = net.namer("conv2d")
    y = None if orphan_y else net.conv(x, yname, out_filters)
    return x, y


def conv_bn_relu6(net: Net, x, cout, conv_name=None, bn_name=None):
    x = net.conv(x, conv_name or net.namer("conv2d"), cout)
    return net.relu6(net.bn(x, bn_name or net.namer("batch_normalization")))


BACKBONES = ("mobilenetv2x75", "mobilenetv2x14", "efficientnetb3", "efficientnetlite0")


def yolov3_body(net: Net, inputs_nhwc: torch.Tensor, model_name: str, num_anchors: int, num_classes: int):
    """Reference ``yolov3_body`` (model.py:170-342).  Returns [y1, y2, y3],
    each [B, H/s, W/s, A, C+5] raw logits, s = 32, 16, 8."""
    x = inputs_nhwc.to(net.dtype).permute(0, 3, 1, 2).contiguous()
    if model_name == "mobilenetv2x75":
        b1, b2, b3, b4 = mobilenet_v2(net, x, 0.75)
    elif model_name == "mobilenetv2x14":
        b1, b2, b3, b4 = mobilenet_v2(net, x, 1.4)
    elif model_name == "efficientnetb3":
        b1, b2, b3, b4 = efficientnet(net, x, model_name)
    elif model_name == "efficientnetlite0":
        b1, b2, b3, b4 = efficientnet(net, x, model_name, lite=True)
    else:
        raise ValueError("unknown backbone %r" % model_name)
    for n, t in zip(("b1", "b2", "b3", "b4"), (b1, b2, b3, b4)):
        net.taps[n] = t
    b4 = net.maxpool(b4, 4)  # downsample_layer(b4, stride=4), model.py:190
    b1, b2, b3 = rfcr_module(net, b1, b2, b3, b4)

    out = num_anchors * (num_classes + 5)
    # top-down (FPN), model.py:238-281
    x, _ = head_stage(net, b1, 512, out, orphan_y=True)
    c1 = x
    x = conv_bn_relu6(net, x, 256, "block_20_conv", "block_20_BN")
    x = torch.cat([net.up2(x), b2], dim=1)
    x, _ = head_stage(net, x, 256, out, orphan_y=True)
    c2 = x
    x = conv_bn_relu6(net, x, 128, "block_24_conv", "block_24_BN")
    x = torch.cat([net.up2(x), b3], dim=1)
    x, _ = head_stage(net, x, 128, out, orphan_y=True)
    c3 = x
    # bottom-up (PAN), model.py:283-323
    x, y3 = head_stage(net, c3, 128, out, orphan_y=False)
    x = conv_bn_relu6(net, x, 128)
    x = torch.cat([net.maxpool(x, 2), c2], dim=1)
    x, y2 = head_stage(net, x, 256, out, orphan_y=False)
    x = conv_bn_relu6(net, x, 256)
    x = torch.cat([net.maxpool(x, 2), c1], dim=1)
    x, y1 = head_stage(net, x, 512, out, orphan_y=False)
    net.taps.update(c1=c1, c2=c2, c3=c3)

    def to_out(y):  # Lambda reshape, model.py:325-340
        y = y.permute(0, 2, 3, 1).contiguous()
        return y.reshape(y.shape[0], y.shape[1], y.shape[2], num_anchors, num_classes + 5)

    return [to_out(y1), to_out(y2), to_out(y3)]


# --------------------------------------------------------------------------
# Weight spec
# --------------------------------------------------------------------------
def weight_spec(model_name: str, num_classes: int, num_anchors: int = 3) -> "OrderedDict[str, Tuple[int, ...]]":
    """(name -> shape) of every weight, in creation order, for a config."""
    net = Net(None, record=True)
    with torch.no_grad():
        yolov3_body(net, torch.zeros(1, 64, 64, 3), model_name, num_anchors, num_classes)
    return net.spec


def forward(weights: Dict[str, np.ndarray], x_nhwc, model_name: str, num_classes: int,
            num_anchors: int = 3, dtype=torch.float32, return_taps: bool = False):
    """Run the oracle network.  x_nhwc: [B,H,W,3] float in [0,1]."""
    net = Net(weights, dtype=dtype)
    x = torch.as_tensor(x_nhwc)
    with torch.no_grad():
        ys = yolov3_body(net, x, model_name, num_anchors, num_classes)
    if return_taps:
        return ys, {k: v.permute(0, 2, 3, 1).contiguous() for k, v in net.taps.items()}
    return ys
