"""Generates the committed golden fixtures from the reference's shipped artifacts.

Run in the build container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py
Inputs : reference code/checkpoints/mobilenetv2x75_320_voc.h5 (MIT-licensed, (c) Prakhar Ganesh),
         code/data_paths/demo_images/*.jpg, code/model_data/{yolo_anchors,voc_classes}.txt
Outputs: tests/golden/voc_mbv2x75_weights.npz   the checkpoint's 392 arrays (by Keras name)
         tests/golden/demo_golden.npz           JPEG bytes, oracle letterbox, oracle head logits,
                                                oracle detections (score 0.3, IoU 0.5, 320x320)
         tests/golden/demo_detections.json      the same detections, human readable
         tests/golden/b3_coco_weights.npz       code/checkpoints/efficientnetb3_416_coco.h5 as stored: 603 arrays under
                                                the CHECKPOINT's Keras names (141 of them differ from the names a
                                                single graph build produces - the reference builds the backbone twice,
                                                code/yolo3/model.py:205-217 - so loading exercises the re-alignment)
         tests/golden/demo_golden_b3.npz        oracle head logits (image 0) and detections (score 0.3, IoU 0.5,
                                                416x416, COCO-80) of that checkpoint on the same 7 demo JPEGs
The oracle (oracle/) produced every number; SURVEY.md §8c lists the same detections from an
independent probe, which is the pin for the oracle itself.
"""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from yoloret_b200.h5lite import load_keras_weights, H5File  # noqa: E402
from yoloret_b200.netdef import NetDef  # noqa: E402
from yoloret_b200.weights import align_weights  # noqa: E402
from oracle import graph, postprocess as pp, letterbox as lb  # noqa: E402

R = "/root/reference/code/"


def main():
    w = load_keras_weights(R + "checkpoints/mobilenetv2x75_320_voc.h5")
    np.savez_compressed(os.path.join(HERE, "voc_mbv2x75_weights.npz"), **{k.replace("/", "__"): v for k, v in w.items()})
    anchors = np.array([float(x) for x in open(R + "model_data/yolo_anchors.txt").readline().split(",")],
                       np.float32).reshape(-1, 2)
    classes = [c.strip() for c in open(R + "model_data/voc_classes.txt")]
    out, readable = {"anchors": anchors}, {}
    files = sorted(glob.glob(R + "data_paths/demo_images/*"))
    for i, f in enumerate(files):
        data = open(f, "rb").read()
        img = lb.decode_image_u8(data)
        x = lb.letterbox_image(lb.u8_to_float(img), (320, 320))
        ys = [y.numpy() for y in graph.forward(w, x[None], "mobilenetv2x75", 20)]
        bi, sc, cl, bf = pp.yolo_eval(ys, anchors, 3, 20, img.shape[:2], score_threshold=0.3, iou_threshold=0.5,
                                      return_float_boxes=True)
        name = os.path.basename(f)
        readable[name] = [dict(cls=classes[c], score=round(float(s), 4), box=b.tolist()) for b, s, c in zip(bi, sc, cl)]
        out["det_boxes_i_%d" % i], out["det_boxes_f_%d" % i] = bi, bf
        out["det_scores_%d" % i], out["det_classes_%d" % i] = sc, cl
        out["shape_%d" % i] = np.array(img.shape[:2], np.int32)
        out["jpeg_%d" % i] = np.frombuffer(data, np.uint8)
        if i < 2:  # head logits (fp32) for the first two images only
            out["y1_%d" % i], out["y2_%d" % i], out["y3_%d" % i] = ys
        if i == 0:
            out["letterbox_0"] = x
    out["names"] = np.array([os.path.basename(f) for f in files])
    np.savez_compressed(os.path.join(HERE, "demo_golden.npz"), **out)
    json.dump(readable, open(os.path.join(HERE, "demo_detections.json"), "w"), indent=1)
    print(json.dumps(readable, indent=1))
    main_b3(anchors, files)


def main_b3(anchors, files):
    """EfficientNet-B3 (SE blocks, 5x5 depthwise, Swish): the only shipped real-weights pin for that path."""
    have = H5File(R + "checkpoints/efficientnetb3_416_coco.h5").weights()
    np.savez_compressed(os.path.join(HERE, "b3_coco_weights.npz"), **{k.replace("/", "__"): v for k, v in have.items()})
    w = align_weights(have, NetDef("efficientnetb3", 80, (416, 416)).weight_shapes)
    classes = [c.strip() for c in open(R + "model_data/coco_classes.txt")]
    out, readable = {"anchors": anchors}, {}
    for i, f in enumerate(files):
        img = lb.decode_image_u8(open(f, "rb").read())
        x = lb.letterbox_image(lb.u8_to_float(img), (416, 416))
        ys = [y.numpy() for y in graph.forward(w, x[None], "efficientnetb3", 80)]
        bi, sc, cl, bf = pp.yolo_eval(ys, anchors, 3, 80, img.shape[:2], score_threshold=0.3, iou_threshold=0.5,
                                      return_float_boxes=True)
        readable[os.path.basename(f)] = [dict(cls=classes[c], score=round(float(s), 4), box=b.tolist())
                                         for b, s, c in zip(bi, sc, cl)]
        out["det_boxes_i_%d" % i], out["det_boxes_f_%d" % i] = bi, bf
        out["det_scores_%d" % i], out["det_classes_%d" % i] = sc, cl
        out["shape_%d" % i] = np.array(img.shape[:2], np.int32)
        if i == 0:
            out["y1_0"], out["y2_0"], out["y3_0"] = ys
    out["names"] = np.array([os.path.basename(f) for f in files])
    out["classes"] = np.array(classes)
    np.savez_compressed(os.path.join(HERE, "demo_golden_b3.npz"), **out)
    json.dump(readable, open(os.path.join(HERE, "demo_detections_b3.json"), "w"), indent=1)
    print(json.dumps(readable, indent=1))


if __name__ == "__main__":
    main()
