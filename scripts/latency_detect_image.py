"""Single-image latency of the reference's own entry point, YOLO(FLAGS).detect_image(bytes, draw=False), on the
shipped VOC checkpoint (golden fixture) and the 7 demo JPEGs: JPEG decode (PIL, host) + upload + GPU letterbox +
network + yolo_eval + read-back, per call.  usage: latency_detect_image.py [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yoloret_b200.yolo import YOLO  # noqa: E402
from yoloret_b200.yolo3.enums import BACKBONE  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = np.load(os.path.join(ROOT, "tests", "golden", "demo_golden.npz"))
z = np.load(os.path.join(ROOT, "tests", "golden", "voc_mbv2x75_weights.npz"))
weights = {k.replace("__", "/"): z[k] for k in z.files}
tmp = "/tmp/yr_lat"
os.makedirs(tmp, exist_ok=True)
open(os.path.join(tmp, "anchors.txt"), "w").write(",  ".join("%g,%g" % (a, b) for a, b in g["anchors"]))
open(os.path.join(tmp, "classes.txt"), "w").write("\n".join("c%d" % i for i in range(20)) + "\n")
yolo = YOLO({"backbone": BACKBONE.MOBILENETV2x75, "classes_path": os.path.join(tmp, "classes.txt"),
             "anchors_path": os.path.join(tmp, "anchors.txt"), "input_size": (320, 320), "score": 0.3, "nms": 0.5,
             "weights": weights, "model": "golden", "quiet": True})
jpegs = [g["jpeg_%d" % i].tobytes() for i in range(len(g["names"]))]
for j in jpegs:  # warm-up: graph capture, PIL import
    yolo.detect_image(j, draw=False)
ts = []
for _ in range(reps):
    for j in jpegs:
        t0 = time.perf_counter()
        boxes, scores, classes = yolo.detect_image(j, draw=False)
        ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print("detect_image (320x320 VOC checkpoint, batch 1): median %.2f ms, p90 %.2f ms, min %.2f ms over %d calls "
      "(%.0f images/s single stream)" % (np.median(ts), np.percentile(ts, 90), ts.min(), len(ts), 1e3 / np.median(ts)))
