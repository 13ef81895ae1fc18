"""Error-budget study (run on the GPU box): head logits of the CUDA engine (SIMT-fp32 and tcgen05-3xTF32
pointwise variants) and of the fp32 oracle, each against the fp64 oracle, on the golden demo images."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yoloret_b200.yolo3.model import yolov3_body  # noqa: E402
from oracle import graph as ograph, letterbox as olb  # noqa: E402

g = np.load(os.path.join(ROOT, "tests/golden/demo_golden.npz"))
z = np.load(os.path.join(ROOT, "tests/golden/voc_mbv2x75_weights.npz"))
w = {k.replace("__", "/"): z[k] for k in z.files}
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    img = olb.decode_image_u8(g["jpeg_%d" % i].tobytes())
    x = olb.letterbox_image(olb.u8_to_float(img), (320, 320))[None]
    ref64 = [y.numpy() for y in ograph.forward(w, torch.from_numpy(x).double(), "mobilenetv2x75", 20, dtype=torch.float64)]
    ref32 = [y.numpy() for y in ograph.forward(w, x, "mobilenetv2x75", 20)]
    out = {}
    for name, variant in (("simt", 1), ("tc3x", 2)):
        m = yolov3_body((1, 320, 320, 3), "mobilenetv2x75", 3, num_classes=20, pw_variant=variant).set_weights(w, g["anchors"])
        out[name] = [y.cpu().numpy() for y in m(torch.from_numpy(x).cuda())]
    for s in range(3):
        mag = np.abs(ref64[s]).max()
        e = {k: np.abs(v[s] - ref64[s]).max() for k, v in out.items()}
        e["oracle32"] = np.abs(ref32[s] - ref64[s]).max()
        e["simt_vs_o32"] = np.abs(out["simt"][s] - ref32[s]).max()
        e["tc_vs_o32"] = np.abs(out["tc3x"][s] - ref32[s]).max()
        print("img %d scale %d  max|logit| %.2f  " % (i, s, mag) + "  ".join("%s %.3e" % kv for kv in e.items()))

# ---- relative logit error and end-to-end detection deltas (tc3x vs oracle fp32) ----
from oracle import postprocess as opp  # noqa: E402
for i in range(len(g["names"])):
    img = olb.decode_image_u8(g["jpeg_%d" % i].tobytes())
    x = olb.letterbox_image(olb.u8_to_float(img), (320, 320))[None]
    ref32 = [y.numpy() for y in ograph.forward(w, x, "mobilenetv2x75", 20)]
    line = "img %d" % i
    for name, variant in (("simt", 1), ("tc3x", 2)):
        m = yolov3_body((1, 320, 320, 3), "mobilenetv2x75", 3, num_classes=20, pw_variant=variant).set_weights(w, g["anchors"])
        ys = [y.cpu().numpy() for y in m(torch.from_numpy(x).cuda())]
        rel = max(float((np.abs(a - r) / np.maximum(1.0, np.abs(r))).max()) for a, r in zip(ys, ref32))
        _, s1, c1, b1 = opp.yolo_eval(ys, g["anchors"], 3, 20, img.shape[:2], score_threshold=0.3, iou_threshold=0.5, return_float_boxes=True)
        _, s0, c0, b0 = opp.yolo_eval(ref32, g["anchors"], 3, 20, img.shape[:2], score_threshold=0.3, iou_threshold=0.5, return_float_boxes=True)
        same = len(c0) == len(c1) and (c0 == c1).all()
        ds = float(np.abs(s1 - s0).max()) if same and len(s0) else -1
        db = float(np.abs(b1 - b0).max()) if same and len(s0) else -1
        dbr = float((np.abs(b1 - b0) / max(img.shape[:2])).max()) if same and len(s0) else -1
        line += "  | %s rel-logit %.2e same-dets %s dscore %.2e dbox_px %.2e dbox/imgsize %.2e" % (name, rel, same, ds, db, dbr)
    print(line)
