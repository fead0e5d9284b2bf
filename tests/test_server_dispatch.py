"""The multi-model, multi-GPU dispatcher (csrc/server.cc, fastdet_b200/server.py; BASELINE config 5).

CPU: routing, stream pinning, micro-batching and the closed-loop generator on the library's host-only stand-in backend
(every "detection" records where and how its frame was served).  GPU (-m gpu): two real models co-resident on one device,
results equal to the single-model detector's."""
import threading

import numpy as np
import pytest

from fastdet_b200 import _native, modelgen
from fastdet_b200.server import DetectServer


def _frame(tag, size=8):
    f = np.zeros((size, size, 3), np.uint8)
    f[0, 0, 0] = tag
    return f


def test_routing_pinning_and_batching_on_the_stand_in_backend():
    srv = DetectServer.fake(n_models=2, n_devices=3, max_batch=4, max_delay_ms=100.0, latency_us=2000)
    out = {}

    def call(stream, model, tag):
        out[(stream, model, tag)] = srv.perform_records(model, stream, _frame(tag), 0.1)

    # 4 streams pinned to device slot 1 (ids 1, 4, 7, 10), model 1, at once: one batch of 4 (max_batch) on device 1
    threads = [threading.Thread(target=call, args=(s, 1, 10 + i)) for i, s in enumerate((1, 4, 7, 10))]
    # and two lone requests elsewhere
    threads += [threading.Thread(target=call, args=(0, 0, 99)), threading.Thread(target=call, args=(5, 1, 98))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for i, s in enumerate((1, 4, 7, 10)):
        r = out[(s, 1, 10 + i)]
        assert len(r) == 1 and r[0]["klass"] == 10 + i           # each caller got ITS frame's answer
        assert r[0]["conf"] == 1.0 and r[0]["x"] == 1.0          # served on device slot s % 3 = 1 by model 1
        assert r[0]["box"] == 4                                   # the four concurrent callers shared one batch
    r = out[(0, 0, 99)]
    assert r[0]["klass"] == 99 and r[0]["conf"] == 0.0 and r[0]["x"] == 0.0 and r[0]["box"] == 1
    r = out[(5, 1, 98)]
    assert r[0]["klass"] == 98 and r[0]["conf"] == 2.0 and r[0]["x"] == 1.0
    stats = srv.lane_stats()
    assert stats[(1, "m1")] == (1, 4) and stats[(0, "m0")] == (1, 1) and stats[(2, "m1")] == (1, 1) and stats[(0, "m1")] == (0, 0)
    with pytest.raises(ValueError, match="invalid image size"):  # reference detector.py:131-132
        srv.perform_records(0, 0, np.zeros((4, 8, 3), np.uint8))
    with pytest.raises(_native.NativeError):
        srv.perform_records(7, 0, _frame(1))
    srv.close()


def test_thresholds_do_not_share_a_batch_and_slots_alternate():
    srv = DetectServer.fake(n_models=1, n_devices=1, max_batch=8, max_delay_ms=50.0, latency_us=1000)
    out = {}

    def call(i):
        out[i] = srv.perform_records(0, 0, _frame(i), 0.1 if i % 2 else 0.3)

    threads = [threading.Thread(target=call, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert all(out[i][0]["klass"] == i for i in range(6))
    assert all(out[i][0]["box"] <= 3 for i in range(6))  # a batch holds one threshold: at most the 3 callers that share it
    b, f = srv.lane_stats()[(0, "m0")]
    assert f == 6 and b >= 2
    srv.close()


def test_closed_loop_generator_statistics():
    srv = DetectServer.fake(n_models=2, n_devices=4, max_batch=16, latency_us=500)
    srv.warm(16)
    frames = np.stack([_frame(i) for i in range(5)])
    st = srv.closed_loop([i % 2 for i in range(16)], frames, seconds=0.5, warmup_seconds=0.1)
    assert st["streams"] == 16 and st["frames"] > 100 and st["detections"] == st["frames"]
    assert sum(st["frames_per_device"]) == st["frames"] and min(st["frames_per_device"]) > 0
    assert sum(st["frames_per_model"].values()) == st["frames"]
    assert 0.4 < st["latency_ms"]["p50"] < 50 and st["latency_ms"]["p99"] >= st["latency_ms"]["p50"]
    assert st["mean_batch"] >= 1.0
    # 16 streams in closed loop, each call >= 0.5 ms: at most 32k frames/s however the lanes batch
    assert st["frames_per_second"] < 16 / 0.0005 * 1.05
    srv.close()


@pytest.mark.gpu
def test_two_models_co_resident_on_one_gpu_match_the_detector():
    full = modelgen.build_onnx("tiny", 80, 416, 1)
    rsu = modelgen.build_onnx("tiny", 9, 416, 4)
    srv = DetectServer({"a": (full, 80), "b": (rsu, 9)}, devices=[0], max_batch=8, max_det=256, max_delay_ms=20.0)
    srv.warm(8)
    frames = [modelgen.synthetic_frame(600 + i, 416) for i in range(6)]
    out = {}

    def call(i):
        out[i] = srv.perform("a" if i % 2 == 0 else "b", i, frames[i], 0.1)

    threads = [threading.Thread(target=call, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    da = _native.Model(full, 80, (416, 416), device=0)
    db = _native.Model(rsu, 9, (416, 416), device=0)
    for i in range(6):
        m = da if i % 2 == 0 else db
        dets, counts = m.detect(np.stack([frames[j] for j in range(i % 2, 6, 2)]), 0.1, max_det=256)  # the same batch of 3
        want = dets[i // 2, :counts[i // 2]][["klass", "conf", "x", "y", "w", "h"]].tolist()
        from tests.compare import same_detections
        assert len(want) > 0 and same_detections(out[i], want, 0.1), (out[i], want)
    st = srv.closed_loop(["a", "b"] * 4, np.stack(frames), seconds=1.0, warmup_seconds=0.3)
    assert st["frames"] > 50 and st["frames_per_model"]["a"] > 0 and st["frames_per_model"]["b"] > 0
    srv.close()


@pytest.mark.gpu
def test_lanes_on_every_visible_gpu():
    """One process driving several GPUs (config 5): every device needs its own kernel attributes and execution state — a
    process-wide 'already initialised' guard once made every launch on the second GPU fail.  Skipped on single-GPU boxes."""
    n_dev = _native.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    data = modelgen.build_onnx("tiny", 80, 416, 1)
    devices = list(range(min(n_dev, 8)))
    srv = DetectServer({"a": (data, 80)}, devices=devices, max_batch=8, max_det=256)
    frame = modelgen.synthetic_frame(610, 416)
    want = srv.perform("a", 0, frame, 0.1)
    assert want
    for s_id in range(1, len(devices)):
        assert srv.perform("a", s_id, frame, 0.1) == want  # same frame, same model, another GPU: identical records
    st = srv.closed_loop(["a"] * (4 * len(devices)), np.stack([frame] * 2), seconds=1.0, warmup_seconds=0.5)
    assert min(st["frames_per_device"]) > 0
    srv.close()
