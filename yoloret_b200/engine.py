"""Device engine: lowers a ``NetDef`` to ``yr_op`` plans and runs them through the C-ABI.

PyTorch is used for device memory, streams and CUDA-graph capture only; every
arithmetic kernel on this path lives in ``csrc/`` behind ``include/yoloret_b200.h``.
There is no CPU fallback: constructing an ``Engine`` without CUDA raises.

Data layout in HBM: one activation arena sized for a
micro-batch (reused across micro-batches so producer->consumer tensors stay
L2-resident), full-batch head outputs y1..y3, and a post-process workspace
(boxes, per-(image,class) candidate lists, per-class detections, packed output).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import YrOp, YrDecodeParams
from .netdef import NetDef, Layer, View, pad_c
from .postprocess import PostProcess

BN_EPS = 1e-3  # every BatchNormalization in the reference graph
_ACT = {"none": _lib.ACT_NONE, "relu6": _lib.ACT_RELU6, "swish": _lib.ACT_SWISH}
_MODE = {"up2": _lib.UP2, "pool2": _lib.POOL2, "pool4": _lib.POOL4}


def _fold_bn(w: Dict[str, np.ndarray], bn: Optional[str], cout: int):
    """Returns per-output-channel (scale, bias) in float64."""
    if bn is None:
        return np.ones(cout), np.zeros(cout)
    g = w[bn + "/gamma"].astype(np.float64)
    b = w[bn + "/beta"].astype(np.float64)
    m = w[bn + "/moving_mean"].astype(np.float64)
    v = w[bn + "/moving_variance"].astype(np.float64)
    s = g / np.sqrt(v + BN_EPS)
    return s, b - m * s


def _expand_rows(mat: np.ndarray, segs) -> np.ndarray:
    """[sum(logical), N] -> [sum(padded), N] with zero rows at the pad positions."""
    out = np.zeros((sum(p for _, p in segs), mat.shape[1]), dtype=mat.dtype)
    ri = ro = 0
    for l, p in segs:
        out[ro:ro + l] = mat[ri:ri + l]
        ri += l
        ro += p
    return out


def _pad_cols(mat: np.ndarray, n_pad: int) -> np.ndarray:
    out = np.zeros(mat.shape[:-1] + (n_pad,), dtype=mat.dtype)
    out[..., :mat.shape[-1]] = mat
    return out


class Engine:
    def __init__(self, model_name: str, num_classes: int, input_hw: Tuple[int, int], batch: int,
                 weights: Dict[str, np.ndarray], anchors: np.ndarray, micro_batch: Optional[int] = None,
                 num_scales: int = 3, max_boxes: int = 20, cand_cap: Optional[int] = None,
                 device: Optional[torch.device] = None, pw_variant: int = _lib.PW_AUTO, input_u8: bool = False,
                 fuse_se: bool = True, lanes: int = 1, autotune: bool = True, fuse_up2: bool = True,
                 num_anchors: int = 3, fuse_dwpw: bool = True, fold_linear: bool = True, stack_pw: bool = True):
        if not torch.cuda.is_available():
            raise _lib.YrError("yoloret_b200.Engine needs a CUDA device (no CPU fallback exists)")
        self.lib = _lib.lib()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        if self.device.type != "cuda":
            raise _lib.YrError("yoloret_b200.Engine needs a CUDA device, got %s" % (self.device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):  # allocations, attribute caches and the autotuner run on that GPU
            self._init(model_name, num_classes, input_hw, batch, weights, anchors, micro_batch, num_scales, max_boxes,
                       cand_cap, pw_variant, input_u8, fuse_se, lanes, autotune, fuse_up2, num_anchors, fuse_dwpw,
                       fold_linear, stack_pw)

    def _init(self, model_name, num_classes, input_hw, batch, weights, anchors, micro_batch, num_scales, max_boxes,
              cand_cap, pw_variant, input_u8, fuse_se, lanes, autotune, fuse_up2, num_anchors, fuse_dwpw=True, fold_linear=True,
              stack_pw=True):
        # the reference derives anchors per scale as num_anchors // num_scales (code/yolo.py:214-216); decode, the head
        # width and the y_true layout of this engine are written for 3 per scale (every shipped anchor file: 9 / 3)
        if int(num_anchors) != 3:
            raise ValueError("this engine supports 3 anchors per scale (got %d): pass 9 anchors with num_scales=3"
                             % int(num_anchors))
        self.net = NetDef(model_name, num_classes, input_hw, int(num_anchors))
        # linear 1x1 convs whose only readers are 1x1 convs are folded into them (NetDef.fold_linear_pairs): the
        # 255-channel head tensors between a stage's project conv and its y / next conv never touch HBM
        self.folded = self.net.fold_linear_pairs() if fold_linear else []
        self.model_name, self.num_classes, self.input_hw = model_name, num_classes, tuple(input_hw)
        self.batch = int(batch)
        # lanes > 1: the micro-batches of a step run concurrently on that many CUDA streams (fork/join inside the
        # captured graph), each lane on its own activation arena: the small late layers leave most SMs idle when run
        # alone (a 13x13 layer at batch 64 is 85 tiles for 148 SMs) and overlap with the other lanes' layers instead
        self.lanes = max(1, int(lanes))
        if micro_batch:
            self.micro = int(micro_batch)
        else:
            self.micro = (self.batch + self.lanes - 1) // self.lanes
        self._lane_streams: List[torch.cuda.Stream] = []
        self.num_scales, self.max_boxes = num_scales, max_boxes
        self.anchors = np.asarray(anchors, dtype=np.float32).reshape(-1, 2)
        self.input_u8 = input_u8
        self.pw_variant = pw_variant
        self.fuse_se = fuse_se
        self.autotune = bool(autotune)
        self.fuse_up2 = bool(fuse_up2)
        self.fuse_dwpw = bool(fuse_dwpw)
        self.stack_pw = bool(stack_pw)
        self.cand_cap_arg = cand_cap
        self._check_weights(weights)
        self._alloc()
        self._prep_weights(weights)
        self._plans: Dict[Tuple[int, int], Tuple] = {}
        self._graph = None
        self.launches_per_forward = 0
        if self.autotune and self.pw_variant == _lib.PW_AUTO:
            self._autotune_pw()

    # ---- setup -----------------------------------------------------------------
    def _check_weights(self, w):
        for name, shape in self.net.weight_shapes.items():
            if name not in w:
                raise KeyError("missing weight %s" % name)
            if tuple(w[name].shape) != tuple(shape):
                raise ValueError("weight %s has shape %s, expected %s" % (name, w[name].shape, shape))

    def _alloc(self):
        net, dev = self.net, self.device
        self.buf_t: Dict[str, torch.Tensor] = {}
        offs, total = {}, 0
        for b in net.bufs:
            if b.full_batch:
                continue
            n = self.micro * b.H * b.W * b.ld
            offs[b.name] = total
            total += (n + 63) // 64 * 64
        self.arena = torch.zeros(self.lanes, total, dtype=torch.float32, device=dev)
        self.lane_buf_t: List[Dict[str, torch.Tensor]] = [dict() for _ in range(self.lanes)]
        for b in net.bufs:
            if b.full_batch:
                self.buf_t[b.name] = torch.zeros(self.batch, b.H, b.W, b.ld, dtype=torch.float32, device=dev)
                for lane in range(self.lanes):
                    self.lane_buf_t[lane][b.name] = self.buf_t[b.name]
            else:
                n = self.micro * b.H * b.W * b.ld
                for lane in range(self.lanes):
                    self.lane_buf_t[lane][b.name] = self.arena[lane, offs[b.name]:offs[b.name] + n].view(
                        self.micro, b.H, b.W, b.ld)
                self.buf_t[b.name] = self.lane_buf_t[0][b.name]
        self.arena_bytes = total * 4 * self.lanes
        # input slots, keyed (slot, is_u8): float32 images in [0,1] or raw uint8 images (the stem scales by 1/255 like
        # tf.io.decode_image(dtype=float32), reference code/yolo.py:106); a second slot lets a streaming caller upload
        # batch i+1 while batch i computes.  ``self.input`` is slot 0 of the engine's default dtype.
        self.inputs: Dict[Tuple[int, bool], torch.Tensor] = {}
        self.input = self.input_slot(0)
        grids = [(v.H, v.W) for v in net.outputs]
        self.pp = PostProcess(self.batch, grids, self.num_classes, self.anchors, self.num_scales, self.max_boxes,
                              device=dev, cand_cap=self.cand_cap_arg)
        self.total_boxes = self.pp.total_boxes

    def _dev(self, arr: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(arr.astype(np.float32))).to(self.device)

    def _prep_weights(self, w):
        """Folds BN, pads channels to the device layout and uploads."""
        self.wdev: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}
        self.wtc: Dict[int, torch.Tensor] = {}   # tensor-core weight images of the pointwise layers (SS kernel)
        self.wts: Dict[int, torch.Tensor] = {}   # ... for the A-in-TMEM kernel (variant 3)
        self.pw_choice: Dict[int, int] = {}      # per-layer kernel variant (set by the autotuner)
        self.se_fused: Dict[int, bool] = {}      # SE layers whose squeeze rides in the depthwise epilogue
        self.se_part: Dict[Tuple[int, int], torch.Tensor] = {}   # (lane, SE layer) -> per-CTA partial sums
        for i, L in enumerate(self.net.layers):
            if L.kind == "pw":
                k = w[L.conv + "/kernel"][0, 0].astype(np.float64)          # [Cin, Cout]
                s, b = _fold_bn(w, L.bn, k.shape[1])
                k = k * s[None, :]
                A = L.extra.get("fold_from")
                if A is not None:  # B(A(x)) = x (W_A W_B) + (b_A W_B + b_B), in float64 (NetDef.fold_linear_pairs)
                    ka = w[A.conv + "/kernel"][0, 0].astype(np.float64)
                    sa, ba = _fold_bn(w, A.bn, ka.shape[1])
                    b = ba @ k + b
                    k = (ka * sa[None, :]) @ k
                mat = _pad_cols(_expand_rows(k, L.inp[0].segs), L.out.C)
                self.wdev[i] = (self._dev(mat), self._dev(_pad_cols(b, L.out.C)))
                if self.pw_variant in (_lib.PW_AUTO, _lib.PW_TC):
                    self.wtc[i] = self._pack_tc(self.wdev[i][0], _lib.PW_TC)
                if self.pw_variant in (_lib.PW_AUTO, _lib.PW_TS, _lib.PW_TS2):
                    self.wts[i] = self._pack_tc(self.wdev[i][0], _lib.PW_TS)
            elif L.kind == "dw":
                k = w[L.conv + "/depthwise_kernel"][:, :, :, 0].astype(np.float64)  # [k,k,C]
                s, b = _fold_bn(w, L.bn, k.shape[2])
                mat = _pad_cols((k * s[None, None, :]).reshape(L.k * L.k, -1), L.out.C)
                self.wdev[i] = (self._dev(mat), self._dev(_pad_cols(b, L.out.C)))
            elif L.kind == "stem":
                k = w[L.conv + "/kernel"].astype(np.float64)                # [3,3,3,Cout]
                s, b = _fold_bn(w, L.bn, k.shape[3])
                mat = _pad_cols((k * s).reshape(27, -1), L.out.C)
                self.wdev[i] = (self._dev(mat), self._dev(_pad_cols(b, L.out.C)))
            elif L.kind == "se":
                c1, c2, R = L.extra["conv1"], L.extra["conv2"], L.extra["reduced"]
                w1 = _expand_rows(w[c1 + "/kernel"][0, 0].astype(np.float64), L.inp[0].segs)      # [Fp, R]
                w2 = _pad_cols(w[c2 + "/kernel"][0, 0].astype(np.float64), L.inp[0].C)            # [R, Fp]
                b2 = _pad_cols(w[c2 + "/bias"].astype(np.float64), L.inp[0].C)
                prev = self.net.layers[i - 1] if i > 0 else None
                fused = (self.fuse_se and prev is not None and prev.kind == "dw" and prev.out.buf is L.inp[0].buf
                         and prev.out.off == L.inp[0].off)
                if fused:  # squeeze fused into the producing depthwise op; first FC stored transposed
                    self.se_fused[i] = True
                    w1 = np.ascontiguousarray(w1.T)                                               # [R, Fp]
                wcat = np.concatenate([w1.ravel(), w2.ravel()])
                bcat = np.concatenate([w[c1 + "/bias"].astype(np.float64), b2])
                self.wdev[i] = (self._dev(wcat), self._dev(bcat))
            elif L.kind == "rfcr":
                mats = [_expand_rows(w[cn + "/kernel"][0, 0].astype(np.float64), v.segs)
                        for cn, v in zip(L.extra["convs"], L.inp)]
                mat = _pad_cols(np.concatenate(mats, 0), L.out.C)
                self.wdev[i] = (self._dev(mat), self._dev(w["weighted_sum/alpha"]))
        self._find_dwpw_pairs()
        self._find_stacked_pw()
        torch.cuda.synchronize(self.device)
        self.weight_bytes = sum(a.numel() * 4 + b.numel() * 4 for a, b in self.wdev.values())

    def _layer_consumers(self) -> Dict[str, int]:
        if getattr(self, "_consumers", None) is None:
            self._consumers = {}
            for L in self.net.layers:
                for v in L.inp + ([L.res] if L.res is not None else []):
                    self._consumers[v.buf.name] = self._consumers.get(v.buf.name, 0) + 1
        return self._consumers

    def _find_dwpw_pairs(self):
        """3x3 depthwise conv (+BN +act) directly followed by the 1x1 conv that is its only reader - block_N_depthwise ->
        block_N_project of MobileNetV2 (reference code/yolo3/override.py:339-341), the SE-less MBConv blocks of
        EfficientNet-lite (code/yolo3/efficientnet.py:501-533) - run as ONE yr_op (YR_OP_DWPW): the depthwise result goes
        from the converter warps' registers straight into tensor memory as the GEMM's A operand, so the widest tensor
        of the block is written to HBM once (by the expand conv) and read once, instead of twice each.  Bit-identical to
        the two separate ops (tests/test_gpu_ops.py).  Pairs with an SE gate in between, 5x5 depthwise kernels or more
        than 192 output channels keep the separate kernels."""
        self.dwpw_blob: Dict[int, torch.Tensor] = {}
        if not self.fuse_dwpw or self.pw_variant == _lib.PW_SIMT:
            return
        Ls = self.net.layers
        cons = self._layer_consumers()
        for i in range(len(Ls) - 1):
            d, b = Ls[i], Ls[i + 1]
            ok = (d.kind == "dw" and d.k == 3 and d.stride in (1, 2) and b.kind == "pw" and b.gate is None
                  and b.inp[0].buf is d.out.buf and b.inp[0].off == d.out.off and b.inp[0].C == d.out.C
                  and cons.get(d.out.buf.name, 0) == 1 and not d.out.buf.full_batch and not self.se_fused.get(i + 1)
                  and not self._up2_fused_into(i + 1))
            if not ok:
                continue
            C_, N_ = d.out.C, b.out.C
            if not self.lib.yr_dwpw_supported(C_, N_, d.stride, d.out.H, d.out.W):
                continue
            n = int(self.lib.yr_dwpw_packed_floats(C_, N_))
            if n <= 0:
                continue
            blob = torch.zeros(n, dtype=torch.float32, device=self.device)
            wd, bd = self.wdev[i]
            wp, _bp = self.wdev[i + 1]
            assert int(wd.shape[1]) == C_ and tuple(wp.shape) == (C_, N_), (wd.shape, wp.shape, C_, N_)
            _lib.check(self.lib.yr_dwpw_pack(wp.data_ptr(), C_, N_, wd.data_ptr(), bd.data_ptr(), blob.data_ptr(),
                                             self._stream()), "yr_dwpw_pack")
            self.dwpw_blob[i] = blob

    def _find_stacked_pw(self):
        """Two consecutive 1x1 convs that read the SAME tensor through the same SE gate - after the linear folding that is
        a head stage's y conv and the next bottom-up conv (reference code/yolo3/model.py:296-305, here
        conv2d_23*conv2d_24 | conv2d_23*conv2d_25 and conv2d_29*conv2d_30 | conv2d_29*conv2d_31) - run as ONE GEMM over
        [W1 | W2] with two destinations (yr_op.K2 / aux): the shared input is read, gated and split into TF32 once.
        Bit-identical to the two ops (tests/test_gpu_ops.py::test_pw_stacked_outputs).  Needs the tensor-memory-A
        kernel; with the autotuner on, a pair stays stacked only if that measures faster than its two ops."""
        self.pw_stack: Dict[int, Dict] = {}
        if not self.stack_pw or self.pw_variant not in (_lib.PW_AUTO, _lib.PW_TS, _lib.PW_TS2):
            return
        Ls = self.net.layers
        taken = set()
        for i in range(len(Ls) - 1):
            a, b = Ls[i], Ls[i + 1]
            if i in taken or a.kind != "pw" or b.kind != "pw" or (i - 1) in self.dwpw_blob:
                continue
            va, vb = a.inp[0], b.inp[0]
            if not (va.buf is vb.buf and va.off == vb.off and va.C == vb.C and a.gate is b.gate):
                continue
            if a.res is not None or b.res is not None or self._up2_fused_into(i) or self._up2_fused_into(i + 1):
                continue
            if a.act != "none" and a.act != b.act:
                continue
            if self.wts.get(i) is None or self.wts.get(i + 1) is None:
                continue
            w = torch.cat([self.wdev[i][0], self.wdev[i + 1][0]], 1).contiguous()
            bias = torch.cat([self.wdev[i][1], self.wdev[i + 1][1]]).contiguous()
            n1 = int(self.wdev[i][0].shape[1])
            if n1 % 4 or int(self.lib.yr_pw_ts_packed_floats(int(w.shape[0]), int(w.shape[1]))) <= 0:
                continue
            img = self._pack_tc(w, _lib.PW_TS)
            if img is None:
                continue
            pair_ok = bool(self.lib.yr_pw_ts2_supported(int(w.shape[0]), int(w.shape[1])))
            variant = _lib.PW_TS2 if (self.pw_variant == _lib.PW_TS2 and pair_ok) else _lib.PW_TS
            self.pw_stack[i] = {"w": w, "bias": bias, "img": img, "n1": n1, "variant": variant, "pair_ok": pair_ok,
                                "first_linear": 1 if (a.act == "none" and b.act != "none") else 0}
            taken.add(i + 1)

    def _pack_tc(self, w_kn: torch.Tensor, variant: int) -> Optional[torch.Tensor]:
        """yr_pw_tc_pack / yr_pw_ts_pack: [K,N] fp32 -> split/swizzled TF32 (hi, lo) image for a tcgen05 kernel."""
        K, N = int(w_kn.shape[0]), int(w_kn.shape[1])
        sizer, packer = ((self.lib.yr_pw_ts_packed_floats, self.lib.yr_pw_ts_pack) if variant == _lib.PW_TS
                         else (self.lib.yr_pw_tc_packed_floats, self.lib.yr_pw_tc_pack))
        n = int(sizer(K, N))
        if n <= 0:
            if self.pw_variant == variant:
                raise _lib.YrError("no tensor-core tiling for a %dx%d pointwise layer" % (K, N))
            return None  # auto: this layer stays on another kernel (the exact-fp32 SIMT one if neither fits)
        packed = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(packer(w_kn.data_ptr(), K, N, packed.data_ptr(), self._stream()), "yr_pw_pack")
        return packed

    def _up2_fused_into(self, i: int) -> bool:
        """True when pointwise layer ``i`` feeds only the nearest-x2 UpSampling2D that follows it (reference
        code/yolo3/model.py:254,274): the tensor-core kernels then store every output row straight to its four
        upsampled pixels in the concat slice and the resample op (one launch, one read and one write of the tensor) goes away."""
        Ls = self.net.layers
        if not self.fuse_up2 or i + 1 >= len(Ls):
            return False
        a, r = Ls[i], Ls[i + 1]
        if not (a.kind == "pw" and r.kind == "resample" and r.mode == "up2" and a.res is None):
            return False
        if not (r.inp[0].buf is a.out.buf and r.inp[0].off == a.out.off and r.inp[0].C == a.out.C):
            return False
        if self._layer_consumers().get(a.out.buf.name, 0) != 1 or a.out.buf.full_batch:
            return False
        return self._pw_variant_of(i) in (_lib.PW_TC, _lib.PW_TS, _lib.PW_TS2)

    def _pw_variant_of(self, i: int) -> int:
        """Kernel variant layer ``i`` runs with (explicit engine setting, else the autotuner's pick, else the
        first tensor-core kernel that has a tiling, else SIMT)."""
        if self.pw_variant == _lib.PW_TS2:  # the CTA-pair form where it has a tiling, else the single-CTA form
            L = self.net.layers[i]
            ok = self.wts.get(i) is not None and self.lib.yr_pw_ts2_supported(L.inp[0].C, L.out.C)
            return _lib.PW_TS2 if ok else (_lib.PW_TS if self.wts.get(i) is not None else _lib.PW_SIMT)
        if self.pw_variant != _lib.PW_AUTO:
            return self.pw_variant
        if i in self.pw_choice:
            return self.pw_choice[i]
        if self.wtc.get(i) is not None:
            return _lib.PW_TC
        if self.wts.get(i) is not None:
            return _lib.PW_TS
        return _lib.PW_SIMT

    def _autotune_pw(self, reps: int = 3):
        """Picks, per pointwise layer, the fastest of the tcgen05 kernels (shared-memory A, tensor-memory A, and the
        CTA-pair form of the latter) by timing them on the layer's real buffers with CUDA events.  They issue the
        same MMAs in the same order and give bit-identical outputs (tests/test_gpu_ops.py), so the pick changes speed
        only."""
        both = [i for i, L in enumerate(self.net.layers)
                if L.kind == "pw" and self.wtc.get(i) is not None and self.wts.get(i) is not None]
        if not both:
            return
        nb = min(self.micro, self.batch)
        stack_cands, self.pw_stack = self.pw_stack, {}   # time the layers one by one first, then the stacked pairs
        self._plans.clear()
        ops, cnt = self.build_plan(0, nb)
        op_of = {li: k for k, (_kind, _name, _b, _f, li) in enumerate(self._plan_meta)}
        st = self._stream()
        cache: Dict[Tuple, int] = {}
        best_ms: Dict[int, float] = {}
        tcache: Dict[Tuple, float] = {}
        self.autotune_log: List[Tuple] = []
        for i in both:
            if i not in op_of:
                continue
            L = self.net.layers[i]
            key = (L.inp[0].H, L.inp[0].W, L.inp[0].C, L.out.C, L.inp[0].buf.ld, L.out.buf.ld, L.res is not None,
                   L.gate is not None, L.act, self._up2_fused_into(i))
            if key not in cache:
                o = ops[op_of[i]]
                t = {}
                cands = [(_lib.PW_TC, self.wtc[i]), (_lib.PW_TS, self.wts[i])]
                if self.lib.yr_pw_ts2_supported(int(key[2]), int(key[3])):
                    cands.append((_lib.PW_TS2, self.wts[i]))   # the CTA-pair form reads the same weight image
                for rnd in range(2):  # two interleaved rounds, best time per kernel: less sensitive to what ran just before
                    for v, img in cands:
                        o.variant, o.w_tc = v, img.data_ptr()
                        _lib.check(self.lib.yr_run_ops(C.byref(o), 1, st), "yr_run_ops")
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(reps):
                            _lib.check(self.lib.yr_run_ops(C.byref(o), 1, st), "yr_run_ops")
                        e1.record()
                        torch.cuda.synchronize(self.device)
                        t[v] = min(t.get(v, 1e9), e0.elapsed_time(e1) / reps)
                # hysteresis: leave the default (tensor-memory A, one CTA) only for a kernel that is clearly faster, so
                # run-to-run noise does not flip the picks (outputs are bit-identical either way)
                best = min(t, key=lambda v: t[v])
                cache[key] = best if t[best] < 0.98 * t[_lib.PW_TS] else _lib.PW_TS
                tcache[key] = t[cache[key]]
                self.autotune_log.append((L.name, key[:4]) + tuple(round(t.get(v, 0.0) * 1e3, 1)
                                                                    for v in (_lib.PW_TC, _lib.PW_TS, _lib.PW_TS2)))
            self.pw_choice[i] = cache[key]
            best_ms[i] = tcache[key]
        # stacked pairs ([W1 | W2] as one GEMM): keep one only if it beats its two ops by more than the noise
        self.stack_log: List[Tuple] = []
        for i, cand in stack_cands.items():
            if i not in best_ms or (i + 1) not in best_ms:
                continue
            self.pw_stack = {i: cand}
            self._plans.clear()
            ops2, _cnt2 = self.build_plan(0, nb)
            k = {li: kk for kk, (_kind, _name, _b, _f, li) in enumerate(self._plan_meta)}[i]
            o = ops2[k]
            t = {}
            for rnd in range(2):
                for v in [_lib.PW_TS] + ([_lib.PW_TS2] if cand["pair_ok"] else []):
                    o.variant = v
                    _lib.check(self.lib.yr_run_ops(C.byref(o), 1, st), "yr_run_ops")
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        _lib.check(self.lib.yr_run_ops(C.byref(o), 1, st), "yr_run_ops")
                    e1.record()
                    torch.cuda.synchronize(self.device)
                    t[v] = min(t.get(v, 1e9), e0.elapsed_time(e1) / reps)
            best = min(t, key=lambda v: t[v])
            sep = best_ms[i] + best_ms[i + 1]
            keep = t[best] < 0.97 * sep
            self.stack_log.append((self.net.layers[i].name, self.net.layers[i + 1].name, round(t[best] * 1e3, 1),
                                   round(sep * 1e3, 1), keep))
            if keep:
                cand["variant"] = best
            else:
                stack_cands = {j: c for j, c in stack_cands.items() if j != i}
        self.pw_stack = {i: c for i, c in stack_cands.items() if i in best_ms and (i + 1) in best_ms}
        self._plans.clear()  # plans were built with the provisional variants

    # ---- plan ---------------------------------------------------------------------
    def input_slot(self, slot: int = 0, u8: Optional[bool] = None) -> torch.Tensor:
        """Device input buffer ``slot`` for uint8 (``u8=True``) or float32 batches (default: the engine's dtype)."""
        key = (int(slot), self.input_u8 if u8 is None else bool(u8))
        t = self.inputs.get(key)
        if t is None:
            t = self.inputs[key] = torch.zeros(self.batch, self.input_hw[0], self.input_hw[1], 3,
                                               dtype=torch.uint8 if key[1] else torch.float32, device=self.device)
        return t

    def slot_for(self, images: torch.Tensor, slot: int = 0) -> Tuple[torch.Tensor, bool]:
        """Validates a [B,H,W,3] uint8 / float32 batch and returns (its device input buffer, is_u8)."""
        want = (self.batch, self.input_hw[0], self.input_hw[1], 3)
        if tuple(images.shape) != want or images.dtype not in (torch.uint8, torch.float32):
            raise ValueError("expected %s uint8 or float32, got %s %s" % (want, tuple(images.shape), images.dtype))
        u8 = images.dtype == torch.uint8
        return self.input_slot(slot, u8), u8

    def _ptr(self, v: View, chunk0: int, slot: int = 0, lane: int = 0, u8: Optional[bool] = None) -> int:
        """Device address of a view for the micro-batch starting at image ``chunk0`` (arena of ``lane``)."""
        t = self.input_slot(slot, u8) if v.buf.name == "input" else self.lane_buf_t[lane][v.buf.name]
        base = t.data_ptr() + v.off * t.element_size()
        if v.buf.full_batch:
            base += chunk0 * v.buf.H * v.buf.W * v.buf.ld * t.element_size()
        return base

    @_lib.on_device
    def build_plan(self, chunk0: int, nb: int, slot: int = 0, lane: int = 0, u8: Optional[bool] = None):
        """yr_op array for images [chunk0, chunk0+nb) reading input slot ``slot`` (uint8 or float32 flavour),
        activations in arena ``lane``."""
        u8 = self.input_u8 if u8 is None else bool(u8)
        key = (chunk0, nb, slot, lane, u8)
        if key in self._plans:
            return self._plans[key]
        _ptr0 = self._ptr

        def _ptr(v, c0, sl=0):
            return _ptr0(v, c0, sl, lane, u8)
        ops = (YrOp * len(self.net.layers))()
        self._ref_bytes: Dict[str, int] = getattr(self, "_ref_bytes", {})
        gate_ptr: Dict[int, int] = {}
        meta = []   # per emitted op: (kind, name, algorithmic bytes / image, flops / image, layer index)
        n_ops = 0
        skip = 0
        for i, L in enumerate(self.net.layers):
            if skip:
                skip -= 1
                continue
            o = ops[n_ops]
            n_ops += 1
            x = L.inp[0]
            if i in self.dwpw_blob:
                b = self.net.layers[i + 1]
                o.kind = _lib.OP_DWPW
                o.B, o.H, o.W, o.C = nb, x.H, x.W, x.C
                o.Ho, o.Wo, o.N = b.out.H, b.out.W, b.out.C
                o.k, o.stride = 3, L.stride
                o.pad_t, o.pad_l = L.extra.get("pad_t", 0), L.extra.get("pad_l", 0)
                o.mode, o.act = _ACT[L.act], _ACT[b.act]
                o.ld_in, o.ld_out = x.buf.ld, b.out.buf.ld
                o.in_ = _ptr(x, chunk0, slot)
                o.out = _ptr(b.out, chunk0)
                o.w_tc = self.dwpw_blob[i].data_ptr()
                o.bias = self.wdev[i + 1][1].data_ptr()
                if b.res is not None:
                    o.res, o.ld_res = _ptr(b.res, chunk0), b.res.buf.ld
                # algorithmic bytes of the fused pair: depthwise input + 1x1 output (+ residual) + both layers' weights
                fused_bytes = 4 * (x.H * x.W * x.Clog + b.out.H * b.out.W * b.out.Clog * (2 if b.res is not None else 1)
                                   + (L.k * L.k + 2) * x.Clog + x.Clog * b.out.Clog + 2 * b.out.Clog)
                meta.append(("dwpw", L.name + "+" + b.name.split("_")[-1], fused_bytes, L.flops + b.flops, i))
                self._ref_bytes[meta[-1][1]] = L.bytes_alg + b.bytes_alg  # SURVEY 8d per-layer accounting of the two layers
                skip = 1
                continue
            if i in self.pw_stack:
                st_, b = self.pw_stack[i], self.net.layers[i + 1]
                o.kind, o.act, o.variant = _lib.OP_PW, _ACT[b.act], st_["variant"]
                o.B, o.H, o.W, o.C = nb, x.H, x.W, x.C
                o.Ho, o.Wo, o.N = L.out.H, L.out.W, int(st_["w"].shape[1])
                o.k, o.stride = 1, 1
                o.ld_in, o.ld_out, o.ld_in2 = x.buf.ld, L.out.buf.ld, b.out.buf.ld
                o.K2, o.K3 = st_["n1"], st_["first_linear"]
                o.in_, o.out, o.aux = _ptr(x, chunk0, slot), _ptr(L.out, chunk0), _ptr(b.out, chunk0)
                o.w, o.bias, o.w_tc = st_["w"].data_ptr(), st_["bias"].data_ptr(), st_["img"].data_ptr()
                if L.gate is not None:
                    o.scale = gate_ptr[id(L.gate)]
                shared = 4 * x.H * x.W * x.Clog   # the input both convs read: counted once
                name = L.name + " | " + b.name.split("*")[-1]
                meta.append(("pw", name, L.bytes_alg + b.bytes_alg - shared, L.flops + b.flops, i))
                self._ref_bytes[name] = L.bytes_alg + b.bytes_alg
                skip = 1
                continue
            meta.append((L.kind, L.name, L.bytes_alg, L.flops, i))
            o.act = _ACT[L.act]
            o.B, o.H, o.W, o.C = nb, x.H, x.W, x.C
            o.Ho, o.Wo, o.N = L.out.H, L.out.W, L.out.C
            o.k, o.stride = L.k, L.stride
            o.pad_t, o.pad_l = L.extra.get("pad_t", 0), L.extra.get("pad_l", 0)
            o.ld_in, o.ld_out = x.buf.ld, L.out.buf.ld
            o.in_ = _ptr(x, chunk0, slot)
            o.out = _ptr(L.out, chunk0)
            if i in self.wdev:
                o.w, o.bias = self.wdev[i][0].data_ptr(), self.wdev[i][1].data_ptr()
            if L.kind == "stem":
                o.kind = _lib.OP_STEM
                o.C = 3
                o.in_is_u8 = 1 if u8 else 0
            elif L.kind == "pw":
                o.kind = _lib.OP_PW
                o.variant = self._pw_variant_of(i)
                img = (self.wts if o.variant in (_lib.PW_TS, _lib.PW_TS2) else self.wtc).get(i)
                if o.variant != _lib.PW_SIMT and img is not None:
                    o.w_tc = img.data_ptr()
                if L.res is not None:
                    o.res, o.ld_res = _ptr(L.res, chunk0), L.res.buf.ld
                if L.gate is not None:
                    o.scale = gate_ptr[id(L.gate)]
                if self._up2_fused_into(i):
                    up = self.net.layers[i + 1]
                    o.Ho, o.Wo = up.out.H, up.out.W
                    o.out, o.ld_out = _ptr(up.out, chunk0), up.out.buf.ld
                    kind, name, byts, flops, li = meta[-1]
                    meta[-1] = (kind, name + "+up2", L.bytes_alg + up.out.H * up.out.W * up.out.Clog * 4 - L.out.H * L.out.W * L.out.Clog * 4,
                                flops, li)
                    skip = 1
            elif L.kind == "dw":
                o.kind = _lib.OP_DW
                if self.se_fused.get(i + 1):
                    if (lane, i + 1) not in self.se_part:
                        slots = int(self.lib.yr_dw_se_slots(C.byref(o)))
                        if slots <= 0:
                            _lib.check(slots, "yr_dw_se_slots")
                        self.se_part[(lane, i + 1)] = torch.zeros(self.micro, slots, o.C, dtype=torch.float32,
                                                                  device=self.device)
                    o.aux = self.se_part[(lane, i + 1)].data_ptr()
            elif L.kind == "resample":
                o.kind = _lib.OP_RESAMPLE
                o.mode = _MODE[L.mode]
            elif L.kind == "se":
                o.kind = _lib.OP_SE
                o.N = L.extra["reduced"]
                gate_ptr[id(L)] = o.out
                if self.se_fused.get(i):
                    o.kind = _lib.OP_SE_FC
                    o.in_ = self.se_part[(lane, i)].data_ptr()
                    o.K2 = int(self.se_part[(lane, i)].shape[1])
            elif L.kind == "rfcr":
                o.kind = _lib.OP_RFCR
                b1, b2, b3, b4 = L.inp
                o.C, o.K2, o.K3, o.K4 = b1.C, b2.C, b3.C, b4.C
                o.ld_in, o.ld_in2, o.ld_in3, o.ld_in4 = b1.buf.ld, b2.buf.ld, b3.buf.ld, b4.buf.ld
                o.in_, o.in2, o.in3, o.in4 = (_ptr(v, chunk0) for v in (b1, b2, b3, b4))
            else:
                raise ValueError(L.kind)
        self._plans[key] = (ops, n_ops)
        self._plan_meta = meta
        return self._plans[key]

    # ---- execution ------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @_lib.on_device
    def run_network(self, slot: int = 0, u8: Optional[bool] = None):
        """yolov3_body forward over input slot ``slot`` -> y buffers, micro-batch by micro-batch."""
        chunks = [(c0, min(self.micro, self.batch - c0)) for c0 in range(0, self.batch, self.micro)]
        n = 0
        if self.lanes == 1 or len(chunks) == 1:
            st = self._stream()
            for c0, nb in chunks:
                ops, cnt = self.build_plan(c0, nb, slot, 0, u8)
                _lib.check(self.lib.yr_run_ops(ops, cnt, st), "yr_run_ops")
                n += cnt
            return n
        # fork: every lane's stream waits for the caller's stream, runs its micro-batches in order on its own arena,
        # and the caller's stream joins them all (under graph capture this becomes a fork/join sub-graph)
        main = torch.cuda.current_stream(self.device)
        while len(self._lane_streams) < self.lanes:
            self._lane_streams.append(torch.cuda.Stream(self.device))
        fork = torch.cuda.Event()
        fork.record(main)
        for lane in range(self.lanes):
            ls = self._lane_streams[lane]
            ls.wait_event(fork)
            for j in range(lane, len(chunks), self.lanes):
                c0, nb = chunks[j]
                ops, cnt = self.build_plan(c0, nb, slot, lane, u8)
                _lib.check(self.lib.yr_run_ops(ops, cnt, ls.cuda_stream), "yr_run_ops")
                n += cnt
            done = torch.cuda.Event()
            done.record(ls)
            main.wait_event(done)
        return n

    def raw_outputs(self) -> List[torch.Tensor]:
        """[y1, y2, y3] as [B, gh, gw, A, C+5] strided views of the padded buffers."""
        E = self.num_classes + 5
        outs = []
        for v in self.net.outputs:
            t = self.buf_t[v.buf.name]
            ld = v.buf.ld
            outs.append(t.as_strided((self.batch, v.H, v.W, 3, E), (v.H * v.W * ld, v.W * ld, ld, E, 1)))
        return outs

    @_lib.on_device
    def run_postprocess(self, score_threshold: float, iou_threshold: float) -> int:
        """yolo_eval (model.py:431-491) on the y buffers."""
        ptrs = [self.buf_t[v.buf.name].data_ptr() for v in self.net.outputs[:self.num_scales]]
        ld = [v.buf.ld for v in self.net.outputs[:self.num_scales]]
        return self.pp.run(ptrs, ld, score_threshold, iou_threshold, self._stream())

    @_lib.on_device
    def step(self, score_threshold: float, iou_threshold: float, slot: int = 0, u8: Optional[bool] = None) -> int:
        """One full pass: network + post-process on whatever is in the input slot. Returns #kernel launches."""
        n = self.run_network(slot, u8)
        self.run_postprocess(score_threshold, iou_threshold)
        self.launches_per_forward = n + 3
        return n + 3

    @_lib.on_device
    def capture(self, score_threshold: float, iou_threshold: float, slot: int = 0, u8: Optional[bool] = None):
        """Captures step() into a CUDA graph (launch-bound at small batch: ~90 kernels / micro-batch)."""
        self.input_slot(slot, u8)
        self.step(score_threshold, iou_threshold, slot, u8)  # warm-up: sets func attributes outside capture
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.step(score_threshold, iou_threshold, slot, u8)
        self._graph = g
        return g

    @_lib.on_device
    def results(self, with_float_boxes: bool = False):
        return self.pp.results(with_float_boxes)

    # ---- per-layer timing (bench.py's roofline leg) ------------------------------------------
    @_lib.on_device
    def profile_layers(self, score_threshold: float, iou_threshold: float, reps: int = 3):
        """Times every launch of one full step with CUDA events on the launching stream, inside the
        real pipeline order (so each layer sees the cache state it sees in production).  A spin
        kernel ahead of each pass lets the host enqueue ahead of the device, so the event deltas are
        kernel durations, not launch gaps.  Returns a list of dicts, one per layer kind + layer name:
        ``{"kind", "name", "ms" (summed over micro-batches, averaged over reps), "bytes", "flops",
        "launches"}`` for one step over the whole batch."""
        st = self._stream()
        net = self.net
        chunks = [(c0, min(self.micro, self.batch - c0)) for c0 in range(0, self.batch, self.micro)]
        n_layers = self.build_plan(chunks[0][0], chunks[0][1])[1]
        meta = list(self._plan_meta)
        acc = np.zeros(n_layers + 3)
        for _ in range(reps):
            evs = []
            torch.cuda._sleep(20_000_000)  # ~10 ms head start for the host
            for c0, nb in chunks:
                ops, cnt = self.build_plan(c0, nb)
                row = [torch.cuda.Event(enable_timing=True) for _ in range(cnt + 1)]
                row[0].record()
                for i in range(cnt):
                    _lib.check(self.lib.yr_run_ops(C.byref(ops[i]), 1, st), "yr_run_ops")
                    row[i + 1].record()
                evs.append(row)
            pp_ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ptrs = [self.buf_t[v.buf.name].data_ptr() for v in net.outputs[:self.num_scales]]
            ld = [v.buf.ld for v in net.outputs[:self.num_scales]]
            self.pp.run(ptrs, ld, score_threshold, iou_threshold, st, events=pp_ev)
            torch.cuda.synchronize(self.device)
            for row in evs:
                for i in range(n_layers):
                    acc[i] += row[i].elapsed_time(row[i + 1])
            for j in range(3):
                acc[n_layers + j] += pp_ev[j].elapsed_time(pp_ev[j + 1])
        acc /= reps
        out = []
        for i, (kind, name, byts, flops, _li) in enumerate(meta):
            out.append(dict(kind=kind, name=name, ms=float(acc[i]), bytes=int(byts) * self.batch,
                            flops=int(flops) * self.batch, launches=len(chunks),
                            ref_bytes=int(self._ref_bytes.get(name, byts)) * self.batch))
        E = self.num_classes + 5
        dec_bytes = self.batch * self.total_boxes * (E + 4) * 4
        out.append(dict(kind="decode", name="decode_filter", ms=float(acc[n_layers]), bytes=dec_bytes, flops=0, launches=1))
        out.append(dict(kind="nms", name="nms_classwise", ms=float(acc[n_layers + 1]), bytes=0, flops=0, launches=1))
        out.append(dict(kind="pack", name="pack_detections", ms=float(acc[n_layers + 2]),
                        bytes=self.pp.wire_words * 4, flops=0, launches=1))
        return out
