#!/usr/bin/env python
"""Headline benchmark: images/sec of the YOLO-ReT detection hot path at 416x416.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (yolov3_body forward -> yolo_eval: decode, class-wise NMS,
packing) over one batch of synthetic input.  Workload = BASELINE.json configs[1]:
MobileNetV2-0.75x, 416x416, COCO-80 classes, batch 64 per GPU, seeded random-init weights,
``torch.rand`` images.  Multi-GPU shards the batch (64 images per rank, weak scaling) and ends
every step with an NCCL all-gather of the packed detections.

Keys of the JSON line (see the task contract):
  value     images/s, inputs resident in HBM, one CUDA-graph replay (+ all-gather) per step
  e2e       images/s through ``YOLO.detect_stream`` with pinned HOST fp32 batches: every step uploads its
            batch (H2D), replays the same graph and reads its packed detections back (one D2H); the
            upload of step i+1 overlaps the compute of step i.  ``e2e.blocking_call`` is the same through
            the blocking ``YOLO.detect_batch`` (upload, compute, read back in series)
  roofline  the dominant kernel (largest share of the step): algorithmic bytes / CUDA-event time
  cpu_baseline  the torch-CPU oracle (restatement of the reference TF graph; TF is not
            installable here) on the host cores, bounded sample
``--impl reference`` times that same oracle as the reference arm (the reference is pure Python on
TensorFlow, which is not installed in this image - see DESIGN.md).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ANCHORS = [10, 13, 16, 30, 33, 23, 30, 61, 62, 45, 59, 119, 116, 90, 156, 198, 373, 326]
METRIC = "images/sec at 416x416"  # the headline workload; other --workload values override the size in `config`
UNIT = "images/s"
SCORE, IOU = 0.2, 0.5  # YOLO defaults, reference code/yolo.py:176-177


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="mobilenetv2x75")
    ap.add_argument("--size", type=int, default=416)
    ap.add_argument("--classes", type=int, default=80)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--micro", type=int, default=0, help="micro-batch (0 = engine default)")
    ap.add_argument("--pw-variant", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=1, help="concurrent micro-batch lanes (streams) inside a step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of the workload's batch")
    ap.add_argument("--ref-images", type=int, default=4, help="images per step of the reference arm")
    ap.add_argument("--ncu-step", action="store_true",
                    help="profiling aid: run ONE eager step between cudaProfilerStart/Stop and exit "
                         "(use with ncu --profile-from-start off); prints no bench line")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5", "post", "post0"],
                    help="cfg2 = the headline (BASELINE.json configs[1]); cfg3/cfg4 = the other inference configs "
                         "(per-GPU shard); cfg5 = training-loss step (yolo_loss fwd+bwd + gradient reduce-scatter); "
                         "post / post0 = yolo_eval alone on synthetic head logits at score 0.2 / 0.0")
    a = ap.parse_args()
    if a.workload == "cfg3":    # derived EfficientNet-lite0 320x320, batch 256 over 8 GPUs -> 32 per GPU
        a.model, a.size, a.batch = "efficientnetlite0", 320, 32
    elif a.workload == "cfg4":  # MobileNetV2-1.4 608x608, batch 128 over 8 GPUs -> 16 per GPU
        a.model, a.size, a.batch = "mobilenetv2x14", 608, 16
    elif a.workload == "cfg5":  # training step, batch 256 over 8 GPUs -> 32 per GPU
        a.batch = 32
    return a


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0))), "measured"
    return 6650.0, 1400.0, "fallback"


def measured_traffic(kind):
    """DRAM bytes per step of one kernel kind, from the committed ncu capture of this workload
    (profiles/r2_step_ncu.json: dram__bytes_read.sum + dram__bytes_write.sum summed over the kind's launches)."""
    p = os.path.join(ROOT, "profiles", "r2_step_ncu.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    k = d.get("kinds", {}).get(kind)
    return None if k is None else {"bytes_per_step": k["dram_bytes"], "launches": k["launches"],
                                   "source": "profiles/r2_step_ncu.json (ncu, one eager step, batch %d)" % d.get("batch", 0)}


def workload_name(a):
    return "%s %dx%d %d-class inference, batch=%d per GPU" % (a.model, a.size, a.size, a.classes, a.batch)


def make_inputs(a, batch, seed):
    import torch
    return torch.rand(batch, a.size, a.size, 3, generator=torch.Generator().manual_seed(seed))


def make_weights(a):
    from yoloret_b200.netdef import NetDef
    from yoloret_b200.weights import synthetic_weights
    nd = NetDef(a.model, a.classes, (a.size, a.size))
    return nd, synthetic_weights(nd.weight_shapes, a.classes, seed=1234)


# ---------------------------------------------------------------------------------------------
def oracle_step(weights, x, a, anchors):
    """The CPU restatement of the hot path on a batch: network + per-image yolo_eval."""
    from oracle import graph as ograph, postprocess as opp
    ys = [y.numpy() for y in ograph.forward(weights, x, a.model, a.classes)]
    n = 0
    for b in range(x.shape[0]):
        _, s, _ = opp.yolo_eval([y[b:b + 1] for y in ys], anchors, 3, a.classes, (a.size, a.size),
                                score_threshold=SCORE, iou_threshold=IOU)
        n += len(s)
    return n


def cpu_baseline(a, weights, anchors, budget_s=12.0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    nimg = 4
    x = make_inputs(a, nimg, 99)
    oracle_step(weights, x, a, anchors)  # warm-up
    t0 = time.perf_counter()
    done = 0
    while True:  # a bounded sample: ~12 s of CPU work on all host cores
        oracle_step(weights, x, a, anchors)
        done += nimg
        el = time.perf_counter() - t0
        if el > budget_s:
            break
    return {"value": done / el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d images (%d passes of batch %d) of the same workload through oracle/ "
                      "(torch-CPU restatement of the reference TF graph + numpy/C yolo_eval), %.1f s"
                      % (done, done // nimg, nimg, el)}


def run_reference(a):
    """Reference arm: the reference's CPU implementation of the path.  The reference is pure Python
    on TensorFlow/Keras, not installable here, so this is the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    anchors = np.array(ANCHORS, np.float32).reshape(-1, 2)
    _, weights = make_weights(a)
    nimg = max(1, a.ref_images)
    x = make_inputs(a, nimg, 1234)
    for _ in range(max(1, min(a.warmup, 2))):
        oracle_step(weights, x, a, anchors)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        oracle_step(weights, x, a, anchors)
    el = time.perf_counter() - t0
    v = a.steps * nimg / el
    sample = "%d images per step (bounded sample of the batch-%d workload), oracle/ torch-CPU port" % (nimg, a.batch)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": el / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
def quantise_u8(x):
    """float32 images in [0,1] -> the uint8 images a decoder would deliver (round to nearest)."""
    import torch
    return torch.clamp(torch.round(x * 255.0), 0, 255).to(torch.uint8)


ARITH = ("pointwise 1x1: 3xTF32 products (tcgen05 kind::tf32 on a hi/lo split of both operands, ~21 mantissa bits), "
         "fp32 accumulation in TMEM; everything else fp32 SIMT; fp32 storage")


def verify_against_oracle(a, yolo, weights, anchors, x_host, xu_host):
    """Outside every timed region: the engine's results for THIS workload's batch against the CPU oracle on four
    images spread over the batch (oracle/verify.py: head logits, post-process on identical inputs, end to end), for
    the float32 batch and for the uint8 batch the e2e leg uploads."""
    from oracle import verify as overify
    B = a.batch
    sample = sorted({0, B // 3, (2 * B) // 3, B - 1})
    rep = {}
    try:
        for tag, host, ref_x in (("f32", x_host, x_host), ("u8", xu_host, xu_host.float() * (1.0 / 255.0))):
            dets = yolo.detect_batch(host)
            logits = [y.cpu().numpy() for y in yolo.engine.raw_outputs()]
            rep[tag] = overify.verify_batch(weights, ref_x, a.model, a.classes, anchors, logits, dets, sample, SCORE, IOU)
        return True, rep
    except AssertionError as e:  # report, do not hide: the line then says "verified": false
        rep["error"] = str(e)[:400]
        return False, rep


def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from yoloret_b200.yolo import YOLO
    from yoloret_b200 import parallel, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    anchors = np.array(ANCHORS, np.float32).reshape(-1, 2)
    nd, weights = make_weights(a)

    import tempfile
    tmp = tempfile.mkdtemp(prefix="yr_bench_")
    with open(os.path.join(tmp, "anchors.txt"), "w") as f:
        f.write(", ".join("%d" % v for v in ANCHORS))
    with open(os.path.join(tmp, "classes.txt"), "w") as f:
        f.write("\n".join("class%d" % i for i in range(a.classes)) + "\n")
    flags = {"backbone": a.model, "classes_path": os.path.join(tmp, "classes.txt"),
             "anchors_path": os.path.join(tmp, "anchors.txt"), "input_size": (a.size, a.size), "score": SCORE,
             "nms": IOU, "weights": weights, "batch": a.batch, "pw_variant": a.pw_variant, "quiet": True}
    if a.micro:
        flags["micro_batch"] = a.micro
    flags["lanes"] = a.lanes
    yolo = YOLO(flags)
    eng = yolo.engine
    gather = parallel.DetectionGather(eng.pp, world, rank) if world > 1 else None

    # rank-distinct synthetic images, pinned on the host for the e2e legs: float32 in [0,1] (the workload as
    # SURVEY.md section 8d states it) and the same images as uint8 (what an image decoder hands to the reference,
    # code/yolo.py:106 - the stem kernel applies the 1/255)
    x_host = make_inputs(a, a.batch, 1234 + rank).pin_memory()
    xu_host = quantise_u8(x_host).pin_memory()
    eng.input_slot(0, False).copy_(x_host, non_blocking=True)
    eng.pp.set_image_shapes((a.size, a.size))
    if a.ncu_step:
        eng.step(SCORE, IOU, 0, False)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        eng.step(SCORE, IOU, 0, False)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        return
    graph = eng.capture(SCORE, IOU, 0, False)
    launches_per_step = eng.launches_per_forward

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step():
        graph.replay()
        if gather is not None:  # side stream: overlaps the next replay; joined before the closing event
            gather.gather_async(read=False)

    # ---- value: inputs resident in HBM -----------------------------------------------------
    for _ in range(max(a.warmup, 3)):
        device_step()
    if gather is not None:
        gather.join()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        device_step()
    if gather is not None:
        gather.join()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    n_det = int(eng.pp.out_count.sum().item())

    # ---- e2e: public API, host buffers, H2D + D2H inside -------------------------------------
    # YOLO.detect_stream: every step uploads its batch from pinned host memory, computes, and reads its
    # detections back; the upload of step i+1 overlaps the compute of step i (double-buffered input).  Multi-GPU:
    # every step also all-gathers the ranks' detections (side stream) and rank 0 reads ALL ranks' detections back.
    def e2e_run(n, host):
        for _res in yolo.detect_stream((host for _ in range(n)), unpack=False, gather=gather, gather_read=(rank == 0)):
            pass

    def e2e_sync_step():  # the same through the blocking call, for reference (upload, compute, read back in series)
        yolo.detect_batch(xu_host, use_graph=True, unpack=False)

    def timed_e2e(host):
        e2e_run(max(a.warmup, 3), host)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        t0 = time.perf_counter()
        e2e_run(a.steps, host)
        f1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return max(f0.elapsed_time(f1), wall)  # host-side waits count: take the longer clock

    ms_e2e = timed_e2e(xu_host)
    ms_e2e_f32 = timed_e2e(x_host)
    for _ in range(3):
        e2e_sync_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_sync_step()
    barrier()
    ms_e2e_sync = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_e2e_f32], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e_f32 = float(t[0]), float(t[1]), float(t[2])
    total_images = a.batch * world * a.steps
    value = total_images / (ms * 1e-3)
    e2e_value = total_images / (ms_e2e * 1e-3)
    h2d = xu_host.numel() * xu_host.element_size()
    h2d_f32 = x_host.numel() * x_host.element_size()
    # device->host per step: every rank reads its own wire; rank 0 additionally reads the gathered wires of all ranks
    d2h = eng.pp.d2h_bytes() * ((2 * world) if (gather is not None) else 1)

    # ---- multi-GPU equivalence (outside the timed regions): rank r's slice of the gathered wire == what ONE GPU
    # returns for rank r's images -------------------------------------------------------
    equivalence = None
    if gather is not None:
        yolo.detect_batch(x_host, unpack=False)
        gather.all_gather()
        parts = gather.read()
        if rank == 0:
            same, checked = True, 0
            for r in range(world):
                xr = x_host if r == 0 else make_inputs(a, a.batch, 1234 + r).pin_memory()
                mine = yolo.detect_batch(xr)
                theirs = eng.pp.unpack_wire(parts[r])
                for p, q in zip(mine, theirs):
                    same = same and all(np.array_equal(u, v) for u, v in zip(p, q))
                checked += 1
            equivalence = {"ranks_checked": checked, "identical_to_single_gpu": bool(same)}
        barrier()

    verified, verify_report = (None, None)
    if rank == 0 and not a.no_verify:
        verified, verify_report = verify_against_oracle(a, yolo, weights, anchors, x_host, xu_host)
        if equivalence is not None:
            verified = bool(verified and equivalence["identical_to_single_gpu"])

    picks = {}
    for i, v in sorted(eng.pw_choice.items()):
        picks[eng.net.layers[i].name] = {_lib.PW_TC: "tc", _lib.PW_TS: "ts", _lib.PW_TS2: "ts2"}.get(v, str(v))
    out = {
        "metric": METRIC if a.size == 416 else "images/sec at %dx%d" % (a.size, a.size), "value": value, "unit": UNIT,
        "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "arith": ARITH, "data": "synthetic",
        "config": {"workload": workload_name(a), "global_batch": a.batch * world, "micro_batch": eng.micro, "lanes": eng.lanes,
                   "parallelism": "batch-sharded x%d, NCCL all-gather of packed detections (side stream)" % world if world > 1
                   else "single GPU", "weights": "seeded random init (head biases calibrated to ~1% boxes > 0.2)",
                   "score_threshold": SCORE, "iou_threshold": IOU, "detections_per_step_rank0": n_det,
                   "l2": "inputs larger than L2: fp32 batch = %.0f MB and one step moves ~%.1f GB of activations "
                         "through a 126 MB L2, so nothing survives between timed iterations"
                         % (h2d_f32 / 1e6, nd.totals()["bytes"] * a.batch / 1e9)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / a.steps,
                "api": "YOLO.detect_stream(pinned host uint8 [B,%d,%d,3] batches, as an image decoder delivers them; the "
                       "stem kernel applies 1/255): per step H2D of the batch, graph replay, one D2H of the packed "
                       "detections (+ all-gather and rank 0's read of every rank's detections when N > 1); upload of "
                       "step i+1 overlaps compute of step i" % (a.size, a.size),
                "fp32_upload": {"value": total_images / (ms_e2e_f32 * 1e-3), "ms_per_step": ms_e2e_f32 / a.steps,
                                "h2d_bytes_per_step": h2d_f32,
                                "api": "the same with float32 [0,1] host batches (4x the upload bytes)"},
                "blocking_call": {"value": total_images / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync / a.steps,
                                  "api": "YOLO.detect_batch on the uint8 batch (upload, compute, read back in series)"}},
        "gpu_launches": launches_per_step * a.steps,
        "clocks": clocks,
        "verified": verified, "verify": verify_report, "multi_gpu_equivalence": equivalence,
        "pw_kernel_picks": {"tc": sum(1 for v in picks.values() if v == "tc"), "ts": sum(1 for v in picks.values() if v == "ts"),
                            "ts2_cta_pair": sum(1 for v in picks.values() if v == "ts2"),
                            "per_layer": picks,
                            # pairs of 1x1 convs reading the same tensor, run as one GEMM with two destinations:
                            # (first, second, stacked us, separate us, kept) as the autotuner measured them
                            "stacked_pairs": [list(t) for t in getattr(eng, "stack_log", [])]},
    }

    if rank == 0 and not a.no_roofline:
        hbm, tf, which = measured_peaks()
        prof = eng.profile_layers(SCORE, IOU, reps=3)
        kinds = {}
        for r in prof:
            k = kinds.setdefault(r["kind"], {"ms": 0.0, "bytes": 0, "flops": 0, "launches": 0, "ref_bytes": 0})
            k["ms"] += r["ms"]
            k["bytes"] += r["bytes"]
            k["ref_bytes"] += r.get("ref_bytes", r["bytes"])
            k["flops"] += r["flops"]
            k["launches"] += r["launches"]
        tot = sum(k["ms"] for k in kinds.values())
        top = max(kinds, key=lambda k: kinds[k]["ms"])
        K = kinds[top]
        achieved = K["bytes"] / (K["ms"] * 1e-3) / 1e9
        names = {"pw": "pw_tc_kernel / pw_ts_kernel (tcgen05 3xTF32 pointwise 1x1 conv, per-layer autotuned)",
                 "dw": "dw_tma_kernel (TMA-staged depthwise conv)",
                 "dwpw": "pw_ts_kernel<FRONT> (fused 3x3 depthwise -> tcgen05 pointwise, YR_OP_DWPW)"}
        traffic = measured_traffic(top) if (a.workload == "cfg2" and a.batch == 64) else None
        out["roofline"] = {"bound": "hbm", "kernel": names.get(top, top), "achieved": achieved, "peak": hbm,
                           "unit": "GB/s", "frac": achieved / hbm,
                           "traffic": None if traffic is None else traffic["bytes_per_step"] / max(1, traffic["launches"]),
                           "traffic_per_step": None if traffic is None else traffic["bytes_per_step"],
                           "traffic_source": None if traffic is None else traffic["source"],
                           "algorithmic_bytes_per_launch": K["bytes"] / max(1, K["launches"]), "peak_source": which,
                           "share_of_step": K["ms"] / tot, "launches_per_step": K["launches"],
                           "algorithmic_bytes_per_step": K["bytes"], "tflops": K["flops"] / (K["ms"] * 1e-3) / 1e12,
                           "accounting": "bytes = compulsory traffic of each launch AS EXECUTED (a fused depthwise->pointwise "
                                         "launch counts its input, output, residual and weights only - the fused minimum)",
                           "per_layer_accounting": {
                               "bytes_per_step": K["ref_bytes"], "achieved": K["ref_bytes"] / (K["ms"] * 1e-3) / 1e9,
                               "frac": K["ref_bytes"] / (K["ms"] * 1e-3) / 1e9 / hbm,
                               "note": "the same launches against SURVEY.md 8d's per-layer formulas (depthwise layer bytes + "
                                       "pointwise layer bytes, as if the intermediate tensor existed); > what the kernel moves"}}
        out["kernels"] = {k: {"ms_per_step": round(v["ms"], 4), "share": round(v["ms"] / tot, 4),
                              "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None,
                              "launches": v["launches"]} for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]["ms"])}
        out["whole_net"] = {"algorithmic_GB_per_step": nd.totals()["bytes"] * a.batch / 1e9,
                            "GBps_at_value": nd.totals()["bytes"] * a.batch / 1e9 / (ms / a.steps * 1e-3),
                            "frac_of_hbm_peak": nd.totals()["bytes"] * a.batch / 1e9 / (ms / a.steps * 1e-3) / hbm}
        if os.environ.get("YR_BENCH_LAYERS"):
            with open(os.environ["YR_BENCH_LAYERS"], "w") as f:
                json.dump(prof, f, indent=1)

    if rank == 0 and not a.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(a, weights, anchors)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# ---------------------------------------------------------------------------------------------
def synthetic_head_logits(a, batch, seed):
    """Head tensors as SURVEY.md section 8d prescribes: t_xy~N(0,1), t_wh~N(0,.5), t_obj~N(-4,2), t_cls~N(-3,2)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    ys = []
    for s in (32, 16, 8):
        gh = a.size // s
        t = torch.randn(batch, gh, gh, 3, a.classes + 5, generator=g)
        t[..., 2:4] *= 0.5
        t[..., 4] = t[..., 4] * 2 - 4
        t[..., 5:] = t[..., 5:] * 2 - 3
        ys.append(t)
    return ys


def run_aux(a):
    """Secondary workloads (not the headline line): the training-loss step and the standalone post-process."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from yoloret_b200 import parallel
    from yoloret_b200.yolo3.model import YoloLoss, yolo_eval
    from yoloret_b200.yolo3.utils import encode_true_boxes_batch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    anchors = np.array(ANCHORS, np.float32).reshape(-1, 2)
    hbm, _, which = measured_peaks()
    ys = [y.to(dev) for y in synthetic_head_logits(a, a.batch, 1234 + rank)]
    if a.workload == "cfg5":
        # BASELINE.json configs[4]: the training step's loss part on the per-GPU shard (batch 256 / 8 = 32): y_true
        # encoding (reference preprocess_true_boxes, code/yolo3/utils.py:298-376) + yolo_loss forward and backward of
        # all three scales (code/yolo3/model.py:585-671, train.py:11-16) captured in ONE CUDA graph (sparse y_true, one
        # loss launch), then the data-parallel optimizer plumbing at full size: SUM reduce-scatter of the 2.63 M-float
        # gradient bucket, Adam(eps=1e-8) on this rank's shard, all-gather of the parameters (train.py:55-56,158-160).
        # The layer backward (K10) is not built, so the bucket's CONTENT is a placeholder; sizes and kernels are real.
        from yoloret_b200.yolo3.model import FusedYoloLoss
        from yoloret_b200.yolo3.utils import encode_true_boxes_sparse
        from yoloret_b200.train import ShardedAdam
        rng = np.random.default_rng(1234 + rank)
        boxes = np.zeros((a.batch, 8, 5), np.float32)
        for b in range(a.batch):  # 8 boxes per image, wh ~ U(0.05, 0.6), SURVEY.md section 8d
            wh = rng.uniform(0.05, 0.6, (8, 2)) * a.size
            c = rng.uniform(0.0, 1.0, (8, 2)) * a.size
            lo, hi = np.clip(c - wh / 2, 0, a.size - 1), np.clip(c + wh / 2, 0, a.size - 1)
            boxes[b] = np.concatenate([lo, hi, rng.integers(0, a.classes, (8, 1))], 1)
        boxes_host = torch.from_numpy(boxes).pin_memory()
        tb = boxes_host.to(dev)
        sp = encode_true_boxes_sparse(tb, (a.size, a.size), anchors, a.classes)
        fused = FusedYoloLoss(anchors, 3)
        fused._run(sp, ys, True)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            encode_true_boxes_sparse(tb, (a.size, a.size), anchors, a.classes, out=sp)
            fused._run(sp, ys, True)
        nparam = 2630000
        bucket = parallel.GradBucket(nparam, world, rank, device=dev)  # the 2.63 M-parameter gradient buffer
        params = torch.zeros(nparam, dtype=torch.float32, device=dev)
        opt = ShardedAdam(params, bucket, 1e-3, epochs=50)
        loss_host = torch.zeros(3, 4).pin_memory()

        def step():
            g.replay()
            opt.step()

        def e2e_step():  # host boxes in, loss value out
            tb.copy_(boxes_host, non_blocking=True)
            step()
            loss_host.copy_(fused.last_parts, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return float(loss_host[:, :3].sum())
        what = ("yolo_loss fwd+bwd, 3 scales in one launch on a sparse y_true, encoder included (batch %d per GPU, CUDA "
                "graph) + SUM reduce-scatter of a 2.63M-float gradient bucket + Adam on the shard + parameter all-gather"
                % a.batch)
        alg_bytes = sum(3 * y.numel() * 4 for y in ys)  # read logits + y_true, write dlogits
        own_bytes = sum(y.numel() * 4 + y.numel() // (a.classes + 5) * 5 * 4 for y in ys)  # dlogits + 5 logits per slot
    else:
        thr = 0.2 if a.workload == "post" else 0.0

        def step():
            return yolo_eval(ys, anchors, 3, a.classes, (a.size, a.size), score_threshold=thr, iou_threshold=IOU,
                             sync=False)
        e2e_step = None
        what = "yolo_eval (decode + class-wise NMS + pack) on synthetic head logits, batch %d, score_threshold %.1f" % (
            a.batch, thr)
        alg_bytes = sum(y.numel() * 4 for y in ys)
        own_bytes = alg_bytes
    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    ms_e2e = None
    if e2e_step is not None:
        for _ in range(3):
            e2e_step()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            last_loss = e2e_step()
        ms_e2e = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms, ms_e2e or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), (float(t[1]) if ms_e2e is not None else None)
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        per = ms / a.steps
        print(json.dumps({"metric": "images/sec (%s)" % a.workload, "value": a.batch * world * a.steps / (ms * 1e-3),
                          "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                          "ms_per_step": per, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": {"workload": what},
                          "e2e": None if ms_e2e is None else {
                              "value": a.batch * world * a.steps / (ms_e2e * 1e-3), "unit": UNIT,
                              "ms_per_step": ms_e2e / a.steps, "h2d_bytes_per_step": a.batch * 8 * 5 * 4,
                              "d2h_bytes_per_step": 48, "loss": last_loss,
                              "api": "pinned host true boxes -> encode_true_boxes_sparse + FusedYoloLoss (graph) + "
                                     "ShardedAdam.step -> loss value on the host"},
                          "roofline": {"bound": "hbm", "achieved": alg_bytes / (per * 1e-3) / 1e9, "peak": hbm,
                                       "unit": "GB/s", "frac": alg_bytes / (per * 1e-3) / 1e9 / hbm, "traffic": None,
                                       "peak_source": which,
                                       "algorithmic_bytes_per_step": alg_bytes, "bytes_this_implementation_moves": own_bytes,
                                       "note": "whole step (loss graph + collectives + optimizer) against SURVEY.md 8d's "
                                               "logits + y_true + dlogits bytes; the sparse y_true path does not read "
                                               "y_true or the class logits of empty slots, so > 1.0 is possible"}}))


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload in ("cfg5", "post", "post0"):
        run_aux(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
