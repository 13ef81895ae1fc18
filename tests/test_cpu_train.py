"""CPU tests of the training-step host pieces: Keras-HDF5 weight writer <-> reader round trip, the oracle's restatement
of Adam / CosineDecay (reference code/train.py:92-100,158-160) against independent formulas."""
import math
import os

import numpy as np
import pytest
import torch

from yoloret_b200.h5lite import H5File, load_keras_weights
from yoloret_b200.h5write import save_keras_weights
from yoloret_b200.train import cosine_decay
from oracle import optim as ooptim

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("fixture", ["voc_mbv2x75_weights.npz", "b3_coco_weights.npz"])
def test_h5_writer_round_trip_is_bit_exact(tmp_path, fixture):
    """model.save_weights (reference code/train.py:74-79,182-186): every array of the shipped checkpoints written by
    h5write and read back by h5lite is bit-identical, names and layer order preserved; groups with more than 8 and more
    than 256 entries exercise multi-node and two-level B-trees."""
    z = np.load(os.path.join(GOLD, fixture))
    w = {k.replace("__", "/"): z[k] for k in z.files}
    order = []
    for k in w:
        if k.split("/")[0] not in order:
            order.append(k.split("/")[0])
    order.insert(3, "a_layer_without_weights")
    path = str(tmp_path / "out.h5")
    save_keras_weights(path, w, order)
    back = load_keras_weights(path)
    assert set(back) == set(w)
    for k, v in w.items():
        assert back[k].dtype == np.float32 and back[k].shape == v.shape
        assert np.array_equal(back[k].view(np.uint32), np.asarray(v, np.float32).view(np.uint32)), k
    f = H5File(path)
    assert f.layer_names() == order
    assert str(f.root.attrs["backend"].ravel()[0]) == "tensorflow"
    kids = f.children(f.root)
    assert set(kids) == set(order) and kids["a_layer_without_weights"].is_group
    some = order[0]
    names = [str(s) for s in kids[some].attrs["weight_names"].ravel()]
    assert names == ["%s/%s:0" % (some, k.split("/")[1]) for k in w if k.split("/")[0] == some]


def test_h5_writer_many_children_and_shapes(tmp_path):
    rng = np.random.default_rng(0)
    w = {"layer_%04d/kernel" % i: rng.standard_normal((1, 1, i % 5 + 1, 3)).astype(np.float32) for i in range(300)}
    w["layer_0007/bias"] = np.zeros((3,), np.float32)
    w["scalar_like/alpha"] = np.array([1.5, -2.0, 0.25, 1e-30], np.float32)
    w["empty/kernel"] = np.zeros((0, 4), np.float32)
    path = str(tmp_path / "many.h5")
    save_keras_weights(path, w)
    back = load_keras_weights(path)
    assert set(back) == set(w)
    for k in w:
        assert back[k].shape == w[k].shape and np.array_equal(back[k], w[k]), k
    with pytest.raises(ValueError):
        save_keras_weights(path, {"no_slash": np.zeros(3, np.float32)})


def test_adam_restatement_matches_torch_adam_up_to_epsilon_placement():
    """The oracle's ApplyAdam (epsilon added to sqrt(v) BEFORE the bias correction is folded into alpha, as TF does)
    against torch.optim.Adam (epsilon after the correction): identical when epsilon = 0-ish, and within 1e-6 relative for
    the reference's epsilon = 1e-8 on ordinary gradients."""
    rng = np.random.default_rng(1)
    p0 = rng.standard_normal(1000).astype(np.float32)
    grads = [rng.standard_normal(1000).astype(np.float32) * 0.1 for _ in range(5)]
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    tp = torch.tensor(p0.copy(), requires_grad=True)
    opt = torch.optim.Adam([tp], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for t, g in enumerate(grads, 1):
        ooptim.adam_step(p, g, m, v, 1e-3, t, eps=1e-8)
        tp.grad = torch.tensor(g)
        opt.step()
        big = np.abs(np.stack(grads[:t])).min(0) > 1e-3   # epsilon placement only matters where sqrt(v) ~ epsilon
        np.testing.assert_allclose(p[big], tp.detach().numpy()[big], rtol=2e-6, atol=2e-7)
        np.testing.assert_allclose(p, tp.detach().numpy(), atol=1e-4)
    # first step moves every weight by ~lr * sign(g)
    p1, m1, v1 = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    ooptim.adam_step(p1, grads[0], m1, v1, 1e-3, 1)
    big = np.abs(grads[0]) > 1e-3
    np.testing.assert_allclose((p1 - p0)[big], -1e-3 * np.sign(grads[0])[big], rtol=1e-3, atol=2e-7)


def test_cosine_decay_schedule():
    for lr0, epochs in ((1e-3, 50), (1e-4, 7)):
        for e in range(epochs + 3):
            want = lr0 * 0.5 * (1 + math.cos(math.pi * min(e, epochs) / epochs))
            assert cosine_decay(lr0, epochs, e) == pytest.approx(want, rel=1e-5, abs=lr0 * 2e-7)  # float32 like TF: 1 + cos cancels near the end
            assert float(ooptim.cosine_decay(lr0, epochs, e)) == pytest.approx(cosine_decay(lr0, epochs, e), rel=1e-6, abs=1e-12)
    assert cosine_decay(1e-3, 50, 0) == pytest.approx(1e-3) and cosine_decay(1e-3, 50, 50) == pytest.approx(0.0, abs=1e-10)
