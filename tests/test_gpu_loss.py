"""GPU parity of the fused YoloLoss forward/backward against the torch-autograd oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from yoloret_b200.yolo3.model import YoloLoss, yolo_loss  # noqa: E402
from oracle import loss as oloss  # noqa: E402


def _make(B, hw, ncls, anchors, n_boxes, seed):
    rng = np.random.default_rng(seed)
    yts = [[] for _ in range(3)]
    for _ in range(B):
        wh = rng.uniform(0.05, 0.6, (n_boxes, 2)) * np.array(hw[::-1])
        cxy = rng.uniform(0.15, 0.85, (n_boxes, 2)) * np.array(hw[::-1])
        lim = np.array([hw[1] - 1, hw[0] - 1])
        tb = np.concatenate([np.clip(cxy - wh / 2, 0, lim), np.clip(cxy + wh / 2, 0, lim),
                             rng.integers(0, ncls, (n_boxes, 1))], 1)
        yt = oloss.preprocess_true_boxes(tb, hw, anchors, ncls) if n_boxes else \
            [np.zeros((hw[0] // s, hw[1] // s, 3, 5 + ncls), np.float32) for s in (32, 16, 8)]
        for s in range(3):
            yts[s].append(yt[s])
    yts = [torch.tensor(np.stack(y)) for y in yts]
    g = torch.Generator().manual_seed(seed)
    yos = [torch.randn(y.shape, generator=g) * 1.5 for y in yts]
    return yts, yos


@pytest.mark.parametrize("B,hw,ncls,n_boxes", [(2, (128, 128), 20, 6), (3, (96, 160), 80, 8), (2, (64, 64), 4, 0)])
def test_loss_and_grad_match_oracle(built_lib, anchors, B, hw, ncls, n_boxes):
    yts, yos = _make(B, hw, ncls, anchors, n_boxes, seed=B)
    exercised = 0
    for idx in range(3):
        ref_in = yos[idx].double().requires_grad_(True)
        ref, parts = oloss.yolo_loss_scale(yts[idx].double(), ref_in, idx, anchors)
        ref.backward()
        out = yos[idx].cuda().requires_grad_(True)
        L = YoloLoss(idx, anchors, 3, print_loss=False)
        loss = L(yts[idx].cuda(), out)
        loss.backward()
        got_parts = L.last_parts.cpu().numpy()
        np.testing.assert_allclose(got_parts[:3], [float(p) for p in parts[:3]], rtol=2e-4, atol=1e-5)
        assert got_parts[3] == float(parts[3])  # sum(ignore_mask)
        np.testing.assert_allclose(float(loss.detach()), float(ref.detach()), rtol=2e-4)
        gref = ref_in.grad.float().numpy()
        ggot = out.grad.cpu().numpy()
        np.testing.assert_allclose(ggot, gref, rtol=2e-3, atol=2e-6)
        if float(yts[idx][..., 4].sum()) > 0:  # this scale holds objects: the GIoU branch is exercised
            assert np.abs(gref[..., :4]).max() > 0
            exercised += 1
    assert exercised > 0 or n_boxes == 0


def test_functional_yolo_loss_sums_scales(built_lib, anchors):
    yts, yos = _make(2, (96, 96), 20, anchors, 5, seed=9)
    ref = oloss.yolo_loss([y.double() for y in yts], [y.double() for y in yos], anchors)
    got = yolo_loss([y.cuda() for y in yts], [y.cuda() for y in yos], anchors)
    np.testing.assert_allclose(float(got), float(ref), rtol=2e-4)
