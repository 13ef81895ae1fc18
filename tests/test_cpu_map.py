"""VOC mAP harness: host implementation vs the oracle restatement of reference
code/yolo3/map.py, plus hand-computed known answers (the reference ships no expected values)."""
import numpy as np
import pytest

from oracle import map_ref
from yoloret_b200.yolo3 import map as ymap


def test_voc_ap_known_answers():
    # one detection, one ground truth, matched: precision 1 at recall 1 -> AP 1
    assert ymap.voc_ap(np.array([1.0]), np.array([1.0])) == pytest.approx(1.0)
    # TP, FP, TP over 2 positives: rec .5,.5,1  prec 1,.5,.667 -> envelope: .5*1 + .5*.667
    rec, prec = np.array([0.5, 0.5, 1.0]), np.array([1.0, 0.5, 2.0 / 3.0])
    assert ymap.voc_ap(rec, prec) == pytest.approx(0.5 * 1.0 + 0.5 * 2.0 / 3.0)
    assert ymap.voc_ap(rec, prec) == pytest.approx(map_ref.voc_ap(rec, prec))
    # nothing found
    assert ymap.voc_ap(np.array([0.0]), np.array([0.0])) == 0.0


def test_parse_text_line_reference_format():
    line = "VOCdevkit/VOC2007/JPEGImages/000001.jpg 48 240 195 371 11 8 12 352 498 14\n"
    path, boxes = ymap.parse_text_line(line)
    rpath, rboxes = map_ref.parse_text_line(line)
    assert path == rpath == "VOCdevkit/VOC2007/JPEGImages/000001.jpg"
    assert boxes.dtype == np.float32 and boxes.shape == (2, 5)
    assert np.array_equal(boxes, rboxes) and boxes[1].tolist() == [8, 12, 352, 498, 14]
    with pytest.raises(ValueError):
        ymap.parse_text_line("img.jpg 1 2 3 4")


def test_class_aps_hand_case():
    # class 0: two ground truths in image 0; detections: exact hit (.9), duplicate of the same box (.8, a FP: the
    # ground truth is already taken), miss (.7), hit on the second box (.6).  class 1: no detections -> AP 0.
    true_res = {0: np.array([[10, 10, 50, 50, 0], [100, 100, 150, 160, 0], [0, 0, 5, 5, 1]], np.float32),
                1: np.zeros((0, 5), np.float32)}
    pred = [[0, 0, .9, 10, 10, 50, 50], [0, 0, .8, 11, 11, 50, 50], [1, 0, .7, 0, 0, 20, 20], [0, 0, .6, 100, 100, 150, 158]]
    aps = ymap.class_aps(np.array(pred), true_res, 2)
    # tp = 1,0,0,1  fp = 0,1,1,0 -> rec .5,.5,.5,1  prec 1,.5,.333,.5 -> AP = .5*1 + .5*.5
    assert aps[0] == pytest.approx(0.75) and aps[1] == 0
    ref = map_ref.class_aps(pred, true_res, 2)
    assert aps[0] == pytest.approx(ref[0]) and ref[1] == 0


@pytest.mark.parametrize("seed", range(8))
def test_class_aps_matches_oracle_on_random_detections(seed):
    rng = np.random.default_rng(seed)
    ncls, nimg = 4, 12
    true_res, pred = {}, []
    for i in range(nimg):
        n = rng.integers(0, 5)
        xy = rng.integers(0, 300, (n, 2)).astype(np.float32)
        wh = rng.integers(8, 120, (n, 2)).astype(np.float32)
        true_res[i] = np.concatenate([xy, xy + wh, rng.integers(0, ncls, (n, 1)).astype(np.float32)], 1)
        for t in true_res[i]:                      # jittered copies of the truth (some pass IoU .5, some do not)
            for _ in range(rng.integers(0, 3)):
                j = rng.normal(0, 12, 4)
                pred.append([i, int(t[4]) if rng.random() < .8 else int(rng.integers(0, ncls)),
                             float(np.round(rng.random(), 2)),  # rounded: score ties
                             t[0] + j[0], t[1] + j[1], t[2] + j[2], t[3] + j[3]])
        for _ in range(rng.integers(0, 3)):        # background detections
            x, y = rng.integers(0, 300, 2)
            pred.append([i, int(rng.integers(0, ncls)), float(np.round(rng.random(), 2)), x, y, x + 30, y + 40])
    if seed == 0:
        pred = [p for p in pred if p[1] != 3]      # a class with ground truth but no detections
    got = ymap.class_aps(np.array(pred, dtype=np.float64), true_res, ncls)
    ref = map_ref.class_aps(pred, true_res, ncls)
    assert set(got) == set(ref) == set(range(ncls))
    for c in range(ncls):
        assert got[c] == pytest.approx(ref[c], abs=1e-12), c


def test_callback_needs_a_model_and_a_list(tmp_path):
    m = ymap.MAPCallback(str(tmp_path / "none*.txt"), (416, 416), ["a", "b"])
    with pytest.raises(RuntimeError):
        m.calculate_aps()
    m.set_model(lambda x: (np.zeros((0, 4), np.int32), np.zeros(0, np.float32), np.zeros(0, np.int32)))
    with pytest.raises(FileNotFoundError):
        m.calculate_aps()


def test_callback_end_to_end_with_a_stub_model(tmp_path):
    """The harness plumbing (list -> bytes -> model -> (top,left,bottom,right) boxes -> AP) without a GPU."""
    imgs = []
    for i in range(3):
        p = tmp_path / ("img%d.bin" % i)
        p.write_bytes(bytes([i]))
        imgs.append(p)
    (tmp_path / "list.txt").write_text(
        "%s 10 20 110 220 1\n%s 5 5 50 50 0 60 60 90 90 1\n%s 0 0 10 10 0\n" % tuple(p.name for p in imgs))

    def model(batch):   # detections keyed by the image byte: boxes come back as (top, left, bottom, right)
        k = batch[0][0]
        if k == 0:
            return np.array([[20, 10, 220, 110]], np.int32), np.array([.9], np.float32), np.array([1], np.int32)
        if k == 1:
            return (np.array([[5, 5, 50, 50], [200, 200, 240, 240]], np.int32), np.array([.8, .7], np.float32),
                    np.array([0, 1], np.int32))
        return np.zeros((0, 4), np.int32), np.zeros(0, np.float32), np.zeros(0, np.int32)

    m = ymap.MAPCallback(str(tmp_path / "list.txt"), (416, 416), ["a", "b"], image_root=str(tmp_path))
    m.set_model(model)
    aps = m.calculate_aps()
    assert aps[0] == pytest.approx(0.5)    # 1 of 2 class-0 truths found, no false positives
    assert aps[1] == pytest.approx(0.5)    # TP (.9) then FP (.7) over 2 truths: rec .5,.5  prec 1,.5
    assert m.on_train_end({}) == pytest.approx(0.5)
