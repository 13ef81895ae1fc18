"""GPU tests of the training-step kernels that are not the loss: the Adam update on a parameter shard."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from yoloret_b200 import _lib  # noqa: E402
from yoloret_b200.parallel import GradBucket  # noqa: E402
from yoloret_b200.train import ShardedAdam, cosine_decay  # noqa: E402
from oracle import optim as ooptim  # noqa: E402


@pytest.mark.parametrize("n", [1, 3, 4, 1001, 2630000])
def test_adam_step_matches_oracle(built_lib, n):
    """yr_adam_step == the numpy restatement of TF's ApplyAdam (reference code/train.py:158-160: Adam(lr, epsilon=1e-8)),
    five consecutive steps, sizes with a ragged float4 tail and the 2.63 M-parameter bucket of MobileNetV2-0.75 COCO."""
    rng = np.random.default_rng(n)
    p = rng.standard_normal(n).astype(np.float32)
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    dp, dm, dv = (torch.from_numpy(a.copy()).cuda() for a in (p, m, v))
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    for t in range(1, 6):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 1)).astype(np.float32)
        ooptim.adam_step(p, g, m, v, 3e-4, t)
        dg = torch.from_numpy(g).cuda()
        _lib.check(lib.yr_adam_step(dp.data_ptr(), dg.data_ptr(), dm.data_ptr(), dv.data_ptr(), n, 3e-4, 0.9, 0.999, 1e-8,
                                    t, st), "yr_adam_step")
        torch.cuda.synchronize()
        np.testing.assert_allclose(dm.cpu().numpy(), m, rtol=1e-6, atol=1e-30)
        np.testing.assert_allclose(dv.cpu().numpy(), v, rtol=1e-6, atol=1e-38)
        np.testing.assert_allclose(dp.cpu().numpy(), p, rtol=1e-6, atol=1e-7)
    with pytest.raises(_lib.YrError):
        _lib.check(lib.yr_adam_step(dp.data_ptr(), dg.data_ptr(), dm.data_ptr(), dv.data_ptr(), n, 3e-4, 0.9, 0.999, 1e-8, 0, st))


def test_sharded_adam_single_rank_follows_cosine_schedule(built_lib):
    n, epochs = 1001, 4
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(n).astype(np.float32)
    params = torch.from_numpy(p0.copy()).cuda()
    bucket = GradBucket(n, 1, 0, device="cuda")
    opt = ShardedAdam(params, bucket, lr=1e-3, epochs=epochs)
    p, m, v = p0.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    t = 0
    for epoch in range(epochs):
        opt.on_epoch_begin(epoch)
        assert opt.lr == pytest.approx(cosine_decay(1e-3, epochs, epoch))
        for _ in range(2):
            g = rng.standard_normal(n).astype(np.float32)
            bucket.view().copy_(torch.from_numpy(g))
            opt.step()
            t += 1
            ooptim.adam_step(p, g, m, v, cosine_decay(1e-3, epochs, epoch), t)
    np.testing.assert_allclose(params.cpu().numpy(), p, rtol=2e-6, atol=1e-7)
    assert not np.allclose(p, p0)
