"""CPU oracle for the YOLO-ReT hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
(torch-CPU / numpy / plain C), written function-by-function after the
reference sources it cites.  It exists to check the CUDA product path and to
provide the reported CPU baseline.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it;
nothing under ``yoloret_b200/`` does.

PARITY PINNING: the reference ships no tests, golden vectors or published
numbers for this path, and TensorFlow/Keras are not installable
here, so the reference itself cannot be executed.  **Parity is therefore
unpinned by the reference.**  The pins this oracle does have:
  * the shipped checkpoint ``code/checkpoints/mobilenetv2x75_320_voc.h5`` loads
    by layer name into this graph with every shape matching and the parameter
    count (1 887 687) reproduced exactly;
  * with those weights the 7 demo images give the confident, correct VOC
    detections recorded in SURVEY.md §8c (tests/golden/demo_detections.json);
  * NMS: the plain-C restatement, a pure-Python restatement and
    ``torchvision.ops.nms``-independent brute force agree on randomized cases.
"""
