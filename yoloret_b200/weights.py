"""Weights for the engine: reference Keras checkpoints and seeded synthetic init.

``load_checkpoint`` replaces ``model.load_weights(path)`` of the reference
(code/yolo.py:87, code/yolo3/utils.py:389-391) without h5py/Keras.
``synthetic_weights`` produces random-init weights of the named architecture
(there is no network access for datasets/checkpoints other than the three the
reference ships); initialisers follow code/yolo3/efficientnet.py:285-291 and
code/yolo3/model.py:127.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Tuple

import numpy as np

from .h5lite import H5File


def load_checkpoint(path: str, weight_shapes: Mapping[str, Tuple[int, ...]]) -> Dict[str, np.ndarray]:
    """Reads a Keras weights-only .h5 and returns ``{name: array}`` for ``weight_shapes``.

    MobileNetV2 checkpoints match by layer name.  If names do not line up (the
    reference builds EfficientNet backbones twice, code/yolo3/model.py:206-207, which
    shifts Keras' auto-numbering), weights are matched per kind in creation order, like
    Keras' topological ``load_weights`` does.
    """
    return align_weights(H5File(path).weights(), weight_shapes, path)


def align_weights(have: Mapping[str, np.ndarray], weight_shapes: Mapping[str, Tuple[int, ...]],
                  path: str = "<arrays>") -> Dict[str, np.ndarray]:
    """``have``: the arrays of a checkpoint by their stored Keras names -> ``{name: array}`` for ``weight_shapes``
    (by name when every name and shape lines up, else per layer kind in creation order; see ``load_checkpoint``)."""
    if all(k in have and tuple(have[k].shape) == tuple(s) for k, s in weight_shapes.items()):
        return {k: have[k] for k in weight_shapes}

    def key(n):  # 'conv2d_12/kernel' -> ('conv2d', 12)
        layer = n.split("/")[0]
        base, _, num = layer.rpartition("_")
        return (base, int(num)) if num.isdigit() and base else (layer, 0)

    out: Dict[str, np.ndarray] = {}
    consumed = set()
    by_base_have: Dict[str, list] = {}
    for n in have:
        b, i = key(n)
        by_base_have.setdefault(b, set()).add(i)
    by_base_want: Dict[str, list] = {}
    for n in weight_shapes:
        b, i = key(n)
        by_base_want.setdefault(b, set()).add(i)
    remap: Dict[Tuple[str, int], int] = {}
    for b, want_ids in by_base_want.items():
        have_ids = sorted(by_base_have.get(b, []))
        want_sorted = sorted(want_ids)
        if len(have_ids) < len(want_sorted):
            raise KeyError("checkpoint %s lacks %s layers (%d < %d)" % (path, b, len(have_ids), len(want_sorted)))
        # candidate alignments: same order, possibly with extra (orphan) layers in the file
        hi = 0
        for wi in want_sorted:
            probe = next(n for n in weight_shapes if key(n) == (b, wi))
            leaf = probe.split("/")[1]
            while hi < len(have_ids):
                hn = "%s/%s" % (b if have_ids[hi] == 0 else "%s_%d" % (b, have_ids[hi]), leaf)
                if hn in have and tuple(have[hn].shape) == tuple(weight_shapes[probe]):
                    break
                hi += 1
            if hi >= len(have_ids):
                raise KeyError("cannot align %s with checkpoint %s" % (probe, path))
            remap[(b, wi)] = have_ids[hi]
            hi += 1
    for n, shape in weight_shapes.items():
        b, i = key(n)
        j = remap[(b, i)]
        hn = "%s/%s" % (b if j == 0 else "%s_%d" % (b, j), n.split("/")[1])
        if hn not in have:
            raise KeyError("cannot align %s: checkpoint %s has no %s" % (n, path, hn))
        arr = have[hn]
        if tuple(arr.shape) != tuple(shape):
            raise ValueError("weight %s <- %s: shape %s, expected %s" % (n, hn, arr.shape, shape))
        out[n] = arr
        consumed.add(hn)
    # order-based matching is only trustworthy as a bijection: a checkpoint array left over (or used twice) means a
    # layer was skipped and everything behind it may be shifted onto same-shaped neighbours
    left = sorted(set(have) - consumed)
    if left or len(consumed) != len(out):
        raise KeyError("cannot align checkpoint %s with the graph: %d arrays unmatched (e.g. %s)"
                       % (path, len(left), ", ".join(left[:4])))
    return out


HEAD_GAIN = 3.0  # y-conv gain: logit std ~2 like a trained head


def synthetic_weights(weight_shapes: Mapping[str, Tuple[int, ...]], num_classes: int, seed: int = 1234,
                      num_anchors: int = 3, calibrate_head: bool = True) -> Dict[str, np.ndarray]:
    """Seeded random weights for every entry of ``weight_shapes`` (creation order).

    conv kernels N(0, sqrt(2/fan_out)); BN statistics randomised (so folding is
    exercised); ``calibrate_head`` shifts the beta of the BNs feeding the y-convs so
    objectness / class logits average -4 / -3, i.e. a realistic ~1 % of
    boxes pass score 0.2 instead of the ~25 % of unshifted noise."""
    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    for name, shape in weight_shapes.items():
        leaf = name.split("/")[-1]
        if leaf == "kernel":
            kh, kw, _cin, cout = shape
            w[name] = (rng.standard_normal(shape) * math.sqrt(2.0 / (kh * kw * cout))).astype(np.float32)
        elif leaf == "depthwise_kernel":
            kh, kw, _c, _ = shape
            w[name] = (rng.standard_normal(shape) * (0.7 * math.sqrt(2.0 / (kh * kw)))).astype(np.float32)
        elif leaf == "bias":
            w[name] = (rng.standard_normal(shape) * 0.1).astype(np.float32)
        elif leaf == "gamma":
            w[name] = rng.uniform(0.7, 1.3, shape).astype(np.float32)
        elif leaf in ("beta", "moving_mean"):
            w[name] = rng.uniform(-0.2, 0.2, shape).astype(np.float32)
        elif leaf == "moving_variance":
            w[name] = rng.uniform(0.5, 1.5, shape).astype(np.float32)
        elif leaf == "alpha":
            w[name] = rng.uniform(0.6, 1.6, shape).astype(np.float32)
        else:
            raise ValueError("unknown weight kind: %s" % name)
    if calibrate_head:
        out = num_anchors * (num_classes + 5)
        target = np.tile(np.concatenate([np.zeros(4), [-4.0], np.full(num_classes, -3.0)]), num_anchors)
        names = list(weight_shapes.keys())
        for yk in [n for n in names if n.endswith("/kernel") and tuple(weight_shapes[n]) == (1, 1, out, out)]:
            idx = names.index(yk)
            beta = next(n for n in reversed(names[:idx]) if n.endswith("/beta") and tuple(weight_shapes[n]) == (out,))
            W = w[yk][0, 0].astype(np.float64)
            W = W / max(1e-6, float(np.linalg.norm(W, axis=0).mean())) * HEAD_GAIN
            w[yk] = W[None, None].astype(np.float32)
            b = np.linalg.lstsq(W.T, target, rcond=None)[0]
            base = beta[: -len("/beta")]
            w[beta] = b.astype(np.float32)
            w[base + "/gamma"] = (w[base + "/gamma"] * 0.5).astype(np.float32)
            w[base + "/moving_mean"] = np.zeros_like(w[base + "/moving_mean"])
    return w
