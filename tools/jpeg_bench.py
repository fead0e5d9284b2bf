"""JPEG-in throughput on the GPU box: fd_submit_jpeg / fd_collect (Huffman on the host pool, the rest on the device)
against the reference's decode step (PIL, one frame per call) and against the raw-frame pipeline.

    python tools/jpeg_bench.py [--arch full] [--batch 64] [--batches 12] [--quality 75] [--subsampling 2]
"""
import argparse
import io
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from PIL import Image  # noqa: E402

from fastdet_b200 import _native, modelgen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="full")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--batches", type=int, default=12)
    ap.add_argument("--quality", type=int, default=75)
    ap.add_argument("--subsampling", type=int, default=2)
    ap.add_argument("--smooth", type=int, default=1, help="1: camera-like frames (blurred synthetic), 0: raw synthetic noise")
    a = ap.parse_args()
    m = _native.Model(modelgen.build_onnx(a.arch, 80, 416, 2), 80, (416, 416), device=0)
    n = a.batch
    frames = []
    for i in range(8):
        f = modelgen.synthetic_frame(100 + i, 416)
        if a.smooth:
            from PIL import ImageFilter
            f = np.array(Image.fromarray(f).filter(ImageFilter.GaussianBlur(1.5)))
        frames.append(f)
    datas = []
    for f in frames:
        b = io.BytesIO()
        Image.fromarray(f).save(b, "JPEG", quality=a.quality, subsampling=a.subsampling)
        datas.append(b.getvalue())
    batch = [datas[i % len(datas)] for i in range(n)]
    kb = sum(len(d) for d in batch) / n / 1024
    # reference decode step, single core
    t = time.perf_counter()
    for d in batch[:16]:
        np.array(Image.open(io.BytesIO(d)))
    pil_ms = (time.perf_counter() - t) / 16 * 1e3
    decoded = np.stack([np.array(Image.open(io.BytesIO(d))) for d in batch])
    # decode only (synchronous): host entropy decode + H2D + kernels
    m.decode_jpeg(batch, want_rgb=False)
    t = time.perf_counter()
    for _ in range(5):
        m.decode_jpeg(batch, want_rgb=False)
    dec_ms = (time.perf_counter() - t) / 5 * 1e3
    got = m.decode_jpeg(batch)
    exact = bool(np.array_equal(got, decoded))

    def pipelined(submit):
        for s in range(2):
            submit(s)
        for s in range(2):
            m.collect(s)
        t0 = time.perf_counter()
        pend = []
        for b in range(a.batches):
            s = b % 2
            if len(pend) == 2:
                m.collect(pend.pop(0))
            submit(s)
            pend.append(s)
        for s in pend:
            m.collect(s)
        return n * a.batches / (time.perf_counter() - t0)

    fps_jpeg = pipelined(lambda s: m.submit_jpeg(s, batch, 0.1, max_det=256))
    fps_raw = pipelined(lambda s: m.submit(s, decoded, 0.1, max_det=256))
    print({"arch": a.arch, "batch": n, "kb_per_frame": round(kb, 1), "host_threads": os.cpu_count(),
           "pil_decode_ms_per_frame_1core": round(pil_ms, 3), "decode_jpeg_ms_per_batch": round(dec_ms, 3),
           "decode_jpeg_fps": round(n / dec_ms * 1e3), "bit_exact_vs_pil": exact,
           "pipelined_fps_from_jpeg": round(fps_jpeg), "pipelined_fps_from_raw_frames": round(fps_raw),
           "pil_fps_all_cores_upper_bound": round(os.cpu_count() / pil_ms * 1e3)})


if __name__ == "__main__":
    main()
