"""Many-session serving through the reference's call signature: N client threads each call
BatchingService.perform(jpeg_bytes, threshold) in a loop (what DetectService.process_data does per UDP payload,
server/server.py:225-241, if sessions ran concurrently).  Prints aggregate frames/s and per-call latency.

    python tools/serve_bench.py [--arch full] [--clients 64] [--seconds 5] [--max-batch 64] [--max-delay 0.002]
"""
import argparse
import io
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image  # noqa: E402

from fastdet_b200 import detector as fdet, modelgen  # noqa: E402
from fastdet_b200.service import BatchingService  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="full")
    ap.add_argument("--clients", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--max-batch", type=int, default=64)
    ap.add_argument("--max-delay", type=float, default=0.002)
    a = ap.parse_args()
    det = fdet.ONNXDetector(modelgen.build_onnx(a.arch, 80, 416, 2), num_classes=80, max_det=256)
    payloads = []
    for i in range(8):
        b = io.BytesIO()
        Image.fromarray(modelgen.synthetic_frame(100 + i, 416)).save(b, "JPEG", quality=75)
        payloads.append(b.getvalue())
    # the reference's way, one session: perform() per payload, one after another
    det.perform(payloads[0])
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < 1.5:
        det.perform(payloads[k % 8])
        k += 1
    single = k / (time.perf_counter() - t0)
    svc = BatchingService(det, max_batch=a.max_batch, max_delay=a.max_delay)
    for p in payloads:
        svc.perform(p)
    stop = time.perf_counter() + a.seconds
    lat = [[] for _ in range(a.clients)]

    def client(i):
        j = i
        while time.perf_counter() < stop:
            t = time.perf_counter()
            svc.perform(payloads[j % 8], 0.1)
            lat[i].append(time.perf_counter() - t)
            j += 1

    th = [threading.Thread(target=client, args=(i,)) for i in range(a.clients)]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    secs = time.perf_counter() - t0
    svc.close()
    all_lat = np.concatenate([np.array(x) for x in lat]) * 1e3
    print({"arch": a.arch, "clients": a.clients, "frames_per_s": round(len(all_lat) / secs, 1),
           "single_session_perform_per_s": round(single, 1), "latency_ms_p50": round(float(np.percentile(all_lat, 50)), 2),
           "latency_ms_p99": round(float(np.percentile(all_lat, 99)), 2), "mean_batch": round(svc.frames_run / max(svc.batches_run, 1), 1),
           "jpeg_device_frames": det.jpeg_device_frames, "jpeg_host_frames": det.jpeg_host_frames})


if __name__ == "__main__":
    main()
