"""CPU tests of the host logic: network definition / accounting, weight handling, file-format
contract, batch sharding and the world_size-2 detection all-gather (gloo)."""
import os
import socket

import numpy as np
import pytest
import torch

from yoloret_b200.netdef import NetDef, same_pad, pad_c
from yoloret_b200.weights import synthetic_weights
from yoloret_b200.yolo3.utils import get_anchors, get_classes, letterbox_geometry
from yoloret_b200.yolo3.enums import BACKBONE, BOX_LOSS
from yoloret_b200 import parallel

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_netdef_matches_survey_accounting():
    """cfg2 = 3.448 GFLOP, 259.3 MB/img; 23 dw + 55 pw (+4 RFCR 1x1 fused)."""
    nd = NetDef("mobilenetv2x75", 80, (416, 416))
    kinds = [L.kind for L in nd.layers]
    assert kinds.count("dw") == 23 and kinds.count("pw") == 55 - 4 and kinds.count("stem") == 1
    assert kinds.count("se") == 6 and kinds.count("rfcr") == 1
    t = nd.totals()
    assert abs(t["flops"] / 1e9 - 3.448) < 0.02
    assert abs(t["bytes_dw"] / 1e6 - 81.8) < 0.5
    assert [(v.H, v.W) for v in nd.outputs] == [(13, 13), (26, 26), (52, 52)]
    assert all(v.Clog == 255 and v.C == 256 for v in nd.outputs)
    # cfg4 and the derived lite0 build too
    assert abs(NetDef("mobilenetv2x14", 80, (608, 608)).totals()["flops"] / 1e9 - 12.18) < 0.1
    # SURVEY's 8.70 GFLOP for B3 counts stage 7 (0.88 GFLOP), which is unreachable from the detector outputs
    assert abs(NetDef("efficientnetb3", 80, (416, 416)).totals()["flops"] / 1e9 - (8.70 - 0.88)) < 0.1
    NetDef("efficientnetlite0", 80, (320, 320))
    with pytest.raises(ValueError):
        NetDef("mobilenetv2x75", 80, (400, 416))
    with pytest.raises(ValueError):
        NetDef("vgg", 80, (416, 416))


def test_netdef_weight_names_equal_shipped_checkpoint():
    z = np.load(os.path.join(GOLD, "voc_mbv2x75_weights.npz"))
    have = {k.replace("__", "/"): z[k].shape for k in z.files}
    want = NetDef("mobilenetv2x75", 20, (320, 320)).weight_shapes
    assert set(want) == set(have)
    assert all(tuple(have[k]) == tuple(v) for k, v in want.items())
    assert sum(int(np.prod(s)) for s in want.values()) == 1887687


def test_netdef_buffers_are_consistent():
    """Every layer reads channels some earlier layer (or the input) wrote; concat slices tile their buffer."""
    for name, ncls, hw in (("mobilenetv2x75", 80, (416, 416)), ("efficientnetb3", 20, (96, 96)),
                           ("mobilenetv2x14", 80, (64, 64)), ("efficientnetlite0", 80, (64, 64))):
        nd = NetDef(name, ncls, hw)
        written = {}
        for L in nd.layers:
            for v in L.inp + ([L.res] if L.res is not None else []):
                if v.buf.name == "input":
                    continue
                w = written.get(v.buf.name, set())
                need = set(range(v.off, v.off + v.C))
                assert need <= w, "%s reads unwritten channels of %s" % (L.name, v.buf.name)
                assert v.off + v.C <= v.buf.ld and v.off % 4 == 0
            assert (L.out.H, L.out.W) == (L.out.buf.H, L.out.buf.W)
            written.setdefault(L.out.buf.name, set()).update(range(L.out.off, L.out.off + L.out.C))
            assert L.out.off + L.out.C <= L.out.buf.ld


def test_same_pad_rule():
    # TF 'SAME': stride-2 on an even input pads (0,1); on an odd input (1,1); stride 1 k=5 pads (2,2)
    assert same_pad(416, 3, 2) == (208, 0) and same_pad(13, 3, 2) == (7, 1)
    assert same_pad(26, 5, 1) == (26, 2) and same_pad(26, 3, 1) == (26, 1)
    assert pad_c(255) == 256 and pad_c(24) == 24 and pad_c(75) == 80


def test_synthetic_weights_are_seeded_and_calibrated():
    nd = NetDef("mobilenetv2x75", 80, (96, 96))
    a = synthetic_weights(nd.weight_shapes, 80, seed=5)
    b = synthetic_weights(nd.weight_shapes, 80, seed=5)
    c = synthetic_weights(nd.weight_shapes, 80, seed=6)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert any(not np.array_equal(a[k], c[k]) for k in a)
    assert set(a) == set(nd.weight_shapes) and all(a[k].dtype == np.float32 for k in a)
    assert all((a[k] > 0).all() for k in a if k.endswith("moving_variance"))


def test_model_data_file_formats(tmp_path):
    (tmp_path / "a.txt").write_text("10,13,  16,30,  33,23,  30,61,  62,45,  59,119,  116,90,  156,198,  373,326\nignored\n")
    (tmp_path / "c.txt").write_text("person \nbicycle\n car\n")
    a = get_anchors(str(tmp_path / "a.txt"))
    assert a.shape == (9, 2) and a.dtype == np.float32 and a[8].tolist() == [373.0, 326.0]
    assert get_classes(str(tmp_path / "c.txt")) == ["person", "bicycle", "car"]
    # letterbox geometry, reference utils.py:75-79 (float64 scale, int() truncation)
    assert letterbox_geometry(375, 500, (320, 320)) == (240, 320, 40, 0)
    assert letterbox_geometry(500, 375, (416, 416)) == (416, 312, 0, 52)
    assert BACKBONE.MOBILENETV2x75 != BACKBONE.EFFICIENTNETB3 and BOX_LOSS.GIOU != BOX_LOSS.MSE


def test_shard_range_partitions_the_batch():
    for batch in (0, 1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, words, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wire = (torch.arange(words, dtype=torch.int32) + 1000 * rank)
        g = parallel.DetectionGather(wire, world, rank)
        g.all_gather()
        parts = g.read()
        ok = all(np.array_equal(parts[r], np.arange(words, dtype=np.int32) + 1000 * r) for r in range(world))
        wire += 7  # a second step re-uses the buffers
        g.all_gather()
        ok = ok and all(np.array_equal(p, np.arange(words, dtype=np.int32) + 1000 * r + 7)
                        for r, p in enumerate(g.read()))
        # pipelined form: results are picked up one step late; the double buffers keep step i intact while i+1 runs
        t_prev, seen = None, []
        for step in range(4):
            wire.copy_(torch.arange(words, dtype=torch.int32) + 1000 * rank + 31 * step)
            t = g.gather_async()
            if t_prev is not None:
                seen.append([p.copy() for p in g.wait(t_prev)])
            t_prev = t
        seen.append([p.copy() for p in g.wait(t_prev)])
        g.join()
        for step, parts in enumerate(seen):
            ok = ok and all(np.array_equal(p, np.arange(words, dtype=np.int32) + 1000 * r + 31 * step)
                            for r, p in enumerate(parts))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_detection_all_gather_world2_gloo():
    """The N>1 path on CPU: two ranks all-gather their detection wires; every rank sees both, in rank order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, 1234, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def _bucket_worker(rank, world, port, numel, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = parallel.GradBucket(numel, world, rank)
        g = torch.arange(numel, dtype=torch.float32) * (rank + 1)   # rank r holds (r+1) * [0..numel)
        b.view().copy_(g)
        shard = b.reduce_scatter().clone()
        lo = rank * b.shard_numel
        expect = torch.arange(lo, lo + b.shard_numel, dtype=torch.float32) * 3.0   # 1x + 2x
        expect[max(0, numel - lo):] = 0.0                                          # padding stays zero
        ok = torch.equal(shard, expect)
        b.shard.mul_(0.5)                                                          # "optimizer step" on the shard
        full = b.all_gather()
        ok = ok and torch.equal(full, torch.arange(numel, dtype=torch.float32) * 1.5)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_grad_bucket_reduce_scatter_world2_gloo():
    """Training config: flat gradient bucket, SUM reduce-scatter + all-gather, world size 2 on CPU (gloo)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, 1001, q)) for r in range(2)]   # odd size: padding
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
    b = parallel.GradBucket(10, 1, 0)
    b.view().fill_(2.0)
    assert b.shard_numel % 32 == 0   # shards start on 128-byte boundaries; the padding stays zero
    assert torch.equal(b.reduce_scatter()[:10], torch.full((10,), 2.0)) and float(b.shard[10:].abs().sum()) == 0.0
    assert torch.equal(b.all_gather(), torch.full((10,), 2.0))
    with pytest.raises(ValueError):
        parallel.GradBucket(0, 1, 0)


def test_preprocess_true_boxes_matches_oracle():
    """Host y_true encoder (reference utils.py:298-376) against the oracle's restatement, incl. padding rows."""
    from yoloret_b200.yolo3.utils import preprocess_true_boxes, encode_true_boxes_batch
    from oracle import loss as oloss
    anchors = np.array([10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326], np.float32).reshape(-1, 2)
    rng = np.random.default_rng(0)
    for hw, ncls, n in (((416, 416), 80, 8), ((320, 224), 20, 5), ((96, 96), 4, 0)):
        wh = rng.uniform(0.02, 0.7, (n, 2)) * np.array(hw[::-1])
        c = rng.uniform(0.1, 0.9, (n, 2)) * np.array(hw[::-1])
        lim = np.array([hw[1] - 1, hw[0] - 1])
        tb = np.concatenate([np.clip(c - wh / 2, 0, lim), np.clip(c + wh / 2, 0, lim), rng.integers(0, ncls, (n, 1))], 1)
        tb = np.concatenate([tb, np.zeros((3, 5))], 0)  # zero-width padding rows
        got = preprocess_true_boxes(tb, hw, anchors, ncls)
        ref = oloss.preprocess_true_boxes(tb, hw, anchors, ncls)
        assert len(got) == 3
        for g, r in zip(got, ref):
            assert g.shape == r.shape and np.array_equal(g, r)
        assert sum(int(g[..., 4].sum()) for g in got) <= n
    b = encode_true_boxes_batch([tb, tb], hw, anchors, ncls)
    assert b[0].shape == (2, 3, 3, 3, 9)
