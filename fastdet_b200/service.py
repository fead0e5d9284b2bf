"""service.py — micro-batching front end behind the reference's ``perform`` signature.

The reference serves every UDP payload with one blocking ``detector.perform(data, threshold)`` call
(server/server.py:225-241), so all sessions of a model share batch-1 latency.  ``BatchingService`` keeps that call
signature — ``perform(data, threshold) -> [(klass, conf, x, y, w, h), ...]``, same exceptions — but lets many threads
(one per session / stream) call it at once: requests that arrive within ``max_delay`` seconds are decoded in their
caller's thread, stacked, and run as ONE batch through the detector (pipelined ``perform_stream`` when several batches
are waiting), and every caller gets exactly the list ``detector.perform`` would have returned for its frame — or the
exception it would have raised: a payload the batch path cannot take (damaged entropy data behind a valid header, say) is
taken out of its batch, answered on its own, and the rest of the batch runs as a batch.
"""
import inspect
import io
import threading
import time
from collections import deque

import numpy as np


class _Request:
    __slots__ = ("frame", "threshold", "event", "result", "error")

    def __init__(self, frame, threshold):
        self.frame = frame  # decoded RGB array, or the JPEG payload itself (bytes) when the library will decode it
        self.threshold = threshold
        self.event = threading.Event()
        self.result = None
        self.error = None


class BatchingService:
    """detector: an object with ``image_size`` and ``perform_frames(frames_u8[n,h,w,3], threshold) -> [list per frame]``
    (fastdet_b200.detector.ONNXDetector).  If it also has ``jpeg_probe(data)`` and ``perform_jpegs(datas, threshold)``,
    baseline-JPEG payloads are queued as bytes and decoded by the library (fd_detect_jpeg) as part of the batch.
    Thread-safe; ``close()`` stops the worker."""

    def __init__(self, detector, max_batch=64, max_delay=0.002):
        self.detector = detector
        self.image_size = tuple(detector.image_size)
        self.max_batch = int(max_batch)
        self.max_delay = float(max_delay)
        self.batches_run = 0
        self.frames_run = 0
        self._queue = deque()
        self._cond = threading.Condition()
        self._closed = False
        self.isolated = 0  # requests answered on their own after their batch was refused or failed
        sj = getattr(detector, "submit_jpegs", None)
        self._strict = sj is not None and "strict" in inspect.signature(sj).parameters
        pj = getattr(detector, "perform_jpegs", None)
        self._ret_exc = pj is not None and "return_exceptions" in inspect.signature(pj).parameters
        self._worker = threading.Thread(target=self._run, name="fastdet-batcher", daemon=True)
        self._worker.start()

    # -- the reference entry point (server/detector.py:126-146), callable from many threads
    def perform(self, data, threshold=0.1):
        probe = getattr(self.detector, 'jpeg_probe', None)
        if probe is not None:
            verdict = probe(data)  # header parse only (microseconds), in the caller's thread
            if verdict == 'size':
                raise ValueError('invalid image size')
            if verdict == 'device':  # the batch worker hands the bytes to the library's JPEG path
                return self._wait(_Request(bytes(data), float(threshold)))
        from PIL import Image
        (width, height) = self.image_size
        img = Image.open(io.BytesIO(data))  # decode in the caller's thread: it parallelises across sessions
        if img.size != self.image_size:
            raise ValueError('invalid image size')
        frame = np.array(img)
        if frame.ndim != 3 or frame.shape[2] != 3:
            raise ValueError(f'cannot reshape array of size {frame.size} into shape (1,{height},{width},3)')
        return self.perform_frame(frame, threshold)

    def perform_frame(self, frame, threshold=0.1):
        """One decoded RGB frame [h, w, 3] u8."""
        return self._wait(_Request(np.ascontiguousarray(frame, np.uint8), float(threshold)))

    def _wait(self, req):
        with self._cond:
            if self._closed:
                raise RuntimeError("BatchingService is closed")
            self._queue.append(req)
            self._cond.notify()
        req.event.wait()
        if req.error is not None:
            raise req.error
        return req.result

    def close(self):
        with self._cond:
            self._closed = True
            self._cond.notify()
        self._worker.join()

    # -- worker: one batch per distinct threshold among the waiting requests
    def _take(self, block=True):
        with self._cond:
            if not block and not self._queue:
                return None
            while not self._queue and not self._closed:
                self._cond.wait()
            if not self._queue:
                return None
            deadline = time.monotonic() + self.max_delay
            while len(self._queue) < self.max_batch and not self._closed:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._cond.wait(left)
            first = self._queue[0].threshold
            batch, rest = [], deque()
            while self._queue and len(batch) < self.max_batch:
                r = self._queue.popleft()
                (batch if r.threshold == first else rest).append(r)
            rest.extend(self._queue)
            self._queue = rest
            return batch

    @staticmethod
    def _finish(part, results=None, error=None):
        for i, r in enumerate(part):
            r.result, r.error = (results[i] if error is None else None), error
            r.event.set()

    def _run(self):
        self._pipelined = all(hasattr(self.detector, a) for a in ("submit_jpegs", "submit_frames", "collect"))
        self._pending = deque()  # (slot, requests) in submission order; at most two: the library's two ring slots
        self._free = [0, 1]
        while True:
            # with work in flight do not wait for new requests: whatever has arrived is taken, else the oldest batch is
            # collected and delivered
            batch = self._take(block=not self._pending)
            if batch is None and not self._pending:
                if self._closed:
                    return
                continue
            if batch:
                self.batches_run += 1
                self.frames_run += len(batch)
                for part in ([r for r in batch if isinstance(r.frame, bytes)], [r for r in batch if not isinstance(r.frame, bytes)]):
                    if part:
                        self._run_part(part)
            else:
                slot, old = self._pending.popleft()
                self._free.append(slot)
                self._deliver(slot, old)

    def _run_part(self, part):
        """One homogeneous group of requests (all encoded payloads, or all decoded frames) as one batch."""
        encoded = isinstance(part[0].frame, bytes)
        thr = part[0].threshold
        try:
            if self._pipelined:
                if not self._free:
                    slot, old = self._pending.popleft()
                    self._free.append(slot)
                    self._deliver(slot, old)
                slot = self._free.pop(0)
                try:
                    if encoded and self._strict:
                        self.detector.submit_jpegs(slot, [r.frame for r in part], thr, strict=True)
                    elif encoded:
                        self.detector.submit_jpegs(slot, [r.frame for r in part], thr)
                    else:
                        self.detector.submit_frames(slot, np.stack([r.frame for r in part]), thr)
                except Exception:
                    self._free.insert(0, slot)
                    raise
                self._pending.append((slot, part))
            elif encoded and self._ret_exc:
                res = self.detector.perform_jpegs([r.frame for r in part], threshold=thr, return_exceptions=True)
                for r, out in zip(part, res):
                    if isinstance(out, BaseException):
                        self.isolated += 1
                        self._finish([r], error=out)
                    else:
                        self._finish([r], [out])
            elif encoded:
                self._finish(part, self.detector.perform_jpegs([r.frame for r in part], threshold=thr))
            else:
                self._finish(part, self.detector.perform_frames(np.stack([r.frame for r in part]), threshold=thr))
        except Exception as e:  # noqa: BLE001
            self._isolate(part, e)

    def _isolate(self, part, err):
        """A batch was refused or failed before anything ran: answer the offenders on their own — each gets its own result
        or its own exception — and run the others as a batch again.  The refusal's per-frame status (JpegRefused.status)
        names the offenders; without one every request of the part is answered on its own."""
        if len(part) == 1:
            self._finish(part, error=err)
            return
        status = getattr(err, "status", None)
        good, bad = [], list(part)
        if status is not None and len(status) == len(part) and any(int(st) != 0 for st in status):
            good = [r for r, st in zip(part, status) if int(st) == 0]
            bad = [r for r, st in zip(part, status) if int(st) != 0]
        for r in bad:
            self.isolated += 1
            try:
                if isinstance(r.frame, bytes):
                    out = self.detector.perform_jpegs([r.frame], threshold=r.threshold)[0]
                else:
                    out = self.detector.perform_frames(r.frame[None], threshold=r.threshold)[0]
                self._finish([r], [out])
            except Exception as e:  # noqa: BLE001 — this caller's own failure
                self._finish([r], error=e)
        if good:
            self._run_part(good)

    def _deliver(self, slot, part):
        try:
            self._finish(part, self.detector.collect(slot))
        except Exception as e:  # noqa: BLE001 — a device-side failure of a batch that did run: nobody to single out
            self._finish(part, error=e)
