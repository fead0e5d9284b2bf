#!/usr/bin/env python
"""detector.py — drop-in for the reference's ``server/detector.py`` with the model executed by
libfastdet_b200.so (hand-written sm_100a kernels) instead of ONNX Runtime.

Same surface as the reference module (server/detector.py): ``Detector`` (:64-76), ``DummyDetector``
(:78-92), ``ONNXDetector(path, mode=None, num_classes=80, dbgout=None)`` (:94-146) with
``perform(data, threshold=0.1) -> [(klass, conf, x, y, w, h), ...]`` in network-input pixels, klass 1-based,
Soft-NMS order; same exception types (``ValueError('invalid image size')`` :132, PIL's
``UnidentifiedImageError`` for undecodable bytes, ``KeyError`` when the graph has neither 2 nor 3 outputs
:136); the same CLI (:169-195).  ``server/server.py`` imports ``DummyDetector, ONNXDetector`` from a module
named ``detector`` (:17) and needs nothing else.

Additive extras (not in the reference): ``image_size=`` and ``device=`` keyword arguments,
``perform_batch`` / ``perform_frames`` for decoded RGB frames, ``perform_stream`` (pipelined batches),
``perform_wire`` (results already in the server's wire format),
``forward_raw`` for parity tests.
``max_det`` bounds the records copied back per frame (2048 by default); a frame with more detections than that is run
again uncapped, so results never differ from the reference's (which has no cap) — with a logged warning, because it costs
a second pass.  ``mode`` is accepted and stored like the reference does; every mode runs on the B200 — there is no
CPU execution provider here and no fallback.
"""
import io
import logging
import sys
import time

import numpy as np

from . import _native


class Detector:

    def __init__(self, image_size=(416, 416), num_classes=80, dbgout=None):
        self.image_size = image_size
        self.num_classes = num_classes
        self.dbgout = dbgout
        return

    def perform(self, data, threshold=0.1):
        if self.dbgout is not None:
            with open(self.dbgout, 'wb') as fp:
                fp.write(data)
        return


class DummyDetector(Detector):
    """Fixed answer regardless of input — lets the transport be exercised without a model."""

    def __repr__(self):
        return '<DummyDetector>'

    def perform(self, data, threshold=0.1):
        super().perform(data)
        (width, height) = self.image_size
        return [(16, 1.0, 0.5 * width, 0.5 * height, 0.4 * width, 0.4 * height)]


class ONNXDetector(Detector):

    ANCHORS = {
        3: (((116, 90), (156, 198), (373, 326)),
            ((30, 61), (62, 45), (59, 119)),
            ((10, 13), (16, 30), (33, 23))),
        2: (((81, 82), (135, 169), (344, 319)),
            ((10, 14), (23, 27), (37, 58))),
    }

    def __init__(self, path, mode=None, num_classes=80, dbgout=None, image_size=(416, 416), device=0,
                 max_det=2048):
        super().__init__(image_size=tuple(image_size), num_classes=num_classes, dbgout=dbgout)
        self.mode = mode
        self.path = path
        self.max_det = max_det
        self.jpeg_device_frames = 0  # payloads decoded by the library / by PIL (refused by the library)
        self.jpeg_host_frames = 0
        if isinstance(path, (bytes, bytearray)):
            data = bytes(path)
            self.path = '<bytes>'
        else:
            with open(path, 'rb') as fp:
                data = fp.read()
        self.model = _native.Model(data, num_classes, self.image_size, device=device)
        self.logger = logging.getLogger()
        self.logger.info(f'load: path={self.path}, providers=[\'fastdet_b200:sm_100a\'], '
                         f'layers={self.model.info.n_layers}, heads={self.model.head_shapes}')
        return

    def __repr__(self):
        return (f'<ONNXDetector mode={self.mode}, path={self.path}, num_classes={self.num_classes}>')

    # -- the reference entry point -------------------------------------------------------------
    def perform(self, data, threshold=0.1):
        super().perform(data)
        results = self.perform_jpegs([data], threshold=threshold)[0]
        self.logger.info(f'perform: results={results}')
        return results

    def _decode_host(self, data, check_size=True):
        """The reference's own decode lines (server/detector.py:128-133), with their exceptions: used for payloads the
        device JPEG decoder refuses (PNG, progressive or grey JPEG, damaged streams ...)."""
        from PIL import Image
        img = Image.open(io.BytesIO(data))
        if check_size and img.size != self.image_size:
            raise ValueError('invalid image size')
        (width, height) = img.size
        frame = np.array(img)
        if frame.ndim != 3 or frame.shape[2] != 3:
            # the reference's reshape(1,height,width,3) raises ValueError for non-RGB modes (:133)
            raise ValueError(f'cannot reshape array of size {frame.size} into shape (1,{height},{width},3)')
        return frame

    def jpeg_probe(self, data):
        """'device' if the library's JPEG path takes this payload, 'size' if it is such a JPEG of the wrong size
        (the reference raises ValueError('invalid image size')), else 'host' (decode it the reference's way)."""
        info = _native.jpeg_probe(bytes(data))
        if info.status != _native.FD_JPEG_OK:
            return 'host'
        return 'device' if (info.width, info.height) == tuple(self.image_size) else 'size'

    def perform_jpegs(self, datas, threshold=0.1, allow_resize=False, source_coords=False, return_exceptions=False):
        """perform() for a batch of encoded payloads: one result list per payload.  Baseline JPEGs are decoded by the
        library (Huffman on its host thread pool, IDCT / upsampling / colour on the device: fd_detect_jpeg), bit-identical
        to PIL.  Payloads the library refuses (its per-frame status says which) go, one by one, through the reference's own
        decode lines and raise what the reference raises; the others still run as one batch, so one bad payload neither
        slows nor fails its neighbours.  An exception is raised for the first failing payload, or — with
        return_exceptions=True — stands in that payload's place in the returned list.
        allow_resize / source_coords (extension): as in perform_frames, for payloads of one common size other than the
        network's."""
        self.ANCHORS[self.model.n_heads]  # KeyError exactly where the reference raises it (:136)
        datas = list(datas)
        results = [None] * len(datas)
        todo = list(range(len(datas)))
        for _ in range(3):  # a refusal names its offenders; parse problems surface before entropy-decode problems
            if not todo:
                break
            batch = [datas[i] for i in todo]
            try:
                dets, counts = self.model.detect_jpeg(batch, threshold, max_det=self.max_det, allow_resize=allow_resize)
            except _native.JpegRefused as e:
                bad = [i for i, st in zip(todo, e.status) if int(st) != _native.FD_JPEG_OK] or list(todo)
                for i in bad:
                    results[i] = self._perform_host_one(datas[i], threshold, allow_resize, source_coords)
                todo = [i for i in todo if i not in set(bad)]
                continue
            if (counts >= self.max_det).any():
                # a frame filled its record buffer: it may hold more detections than max_det, and the reference has no cap
                self.logger.warning(f'perform: a frame reached max_det={self.max_det}; re-running uncapped')
                dets, counts = self.model.detect_jpeg(batch, threshold, max_det=int(self.model.info.boxes_per_frame), allow_resize=allow_resize)
            self.jpeg_device_frames += len(batch)
            if allow_resize and source_coords and batch:
                info = _native.jpeg_probe(bytes(batch[0]))
                if (info.width, info.height) != tuple(self.image_size):
                    for f in range(dets.shape[0]):
                        dets[f, :counts[f]] = _native.unmap_letterbox(dets[f, :counts[f]], (info.width, info.height), self.image_size)
            for i, r in zip(todo, self._tuples(dets, counts)):
                results[i] = r
            todo = []
        for i in todo:  # (not reached in practice: three refusals in a row)
            results[i] = self._perform_host_one(datas[i], threshold, allow_resize, source_coords)
        if not return_exceptions:
            for r in results:
                if isinstance(r, BaseException):
                    raise r
        return results

    def _perform_host_one(self, data, threshold, allow_resize=False, source_coords=False):
        """One payload through the reference's decode lines; returns its result list, or the exception the reference's
        perform() would have raised for it."""
        try:
            frame = self._decode_host(data, check_size=not allow_resize)
            self.jpeg_host_frames += 1
            return self.perform_frames(frame[None], threshold=threshold, allow_resize=allow_resize, source_coords=source_coords)[0]
        except Exception as e:  # noqa: BLE001 — delivered to the owner of this payload only
            return e

    # -- extras --------------------------------------------------------------------------------
    def perform_frames(self, frames, threshold=0.1, allow_resize=False, source_coords=False):
        """frames: [n, h, w, 3] u8 decoded RGB.  Returns one reference-style result list per frame.  allow_resize
        (extension): frames of another size are letterboxed on the device; boxes then come back in network pixels, or in
        pixels of the frames that were passed in with source_coords=True."""
        self.ANCHORS[self.model.n_heads]  # KeyError exactly where the reference raises it (:136)
        frames = np.asarray(frames)
        dets, counts, total = self.model.detect(frames, threshold, allow_resize=allow_resize, max_det=self.max_det, with_total=True)
        if (total > counts).any():
            # the reference has no cap (at threshold 0 it returns one object per anchor box): run again with room for all
            self.logger.warning(f'perform: {int(total.max())} detections exceed max_det={self.max_det}; re-running uncapped')
            dets, counts = self.model.detect(frames, threshold, allow_resize=allow_resize, max_det=int(self.model.info.boxes_per_frame))
        if source_coords and allow_resize and frames.shape[1:3] != (self.image_size[1], self.image_size[0]):
            src = (frames.shape[2], frames.shape[1])
            for f in range(dets.shape[0]):
                dets[f, :counts[f]] = _native.unmap_letterbox(dets[f, :counts[f]], src, self.image_size)
        return self._tuples(dets, counts)

    perform_batch = perform_frames

    _RESULT_FIELDS = ['klass', 'conf', 'x', 'y', 'w', 'h']

    @classmethod
    def _tuples(cls, dets, counts):
        """fd_det records -> the reference's result lists: [(klass, conf, x, y, w, h), ...] per frame (:142-144)."""
        return [dets[f, :counts[f]][cls._RESULT_FIELDS].tolist() for f in range(dets.shape[0])]

    # -- two batches in flight (fd_submit / fd_submit_jpeg + fd_collect): what BatchingService drives
    def submit_jpegs(self, slot, datas, threshold=0.1, strict=False):
        """Starts a batch of encoded payloads in ring slot `slot` and returns; collect(slot) yields the result lists.
        The entropy decode happens in this call (library's host threads) while the device works on the other slot.
        strict=True: a refusal (JpegRefused, with the per-frame status) is raised instead of decoding the whole batch the
        reference's way — nothing was launched and the slot is free; BatchingService uses it to isolate the offenders."""
        self.ANCHORS[self.model.n_heads]
        datas = list(datas)
        try:
            self.model.submit_jpeg(slot, datas, threshold, max_det=self.max_det)
            self.jpeg_device_frames += len(datas)
        except _native.JpegRefused:
            if strict:
                raise
            frames = np.stack([self._decode_host(d) for d in datas])
            self.jpeg_host_frames += len(datas)
            self.model.submit(slot, frames, threshold, max_det=self.max_det)

    def submit_frames(self, slot, frames, threshold=0.1):
        self.ANCHORS[self.model.n_heads]
        self.model.submit(slot, np.ascontiguousarray(frames, np.uint8), threshold, max_det=self.max_det)

    def collect(self, slot):
        dets, counts, _ = self.model.collect(slot)
        return self._tuples(dets, counts)

    def perform_wire(self, data, threshold=0.1, reqid=0, saturate=False):
        """perform() + the reference server's response packing (server/server.py:231-239) in one call: returns the
        bytes DetectService.send() would be given (16-byte 'YOLO' header + 10 bytes per detection), built by the
        native library straight from the detection records (no per-detection Python objects)."""
        super().perform(data)
        self.ANCHORS[self.model.n_heads]
        t0 = time.time()
        try:
            dets, counts = self.model.detect_jpeg([data], threshold, max_det=self.max_det)
            self.jpeg_device_frames += 1
        except _native.JpegRefused:
            frame = self._decode_host(data)
            self.jpeg_host_frames += 1
            (width, height) = self.image_size
            dets, counts = self.model.detect(frame.reshape(1, height, width, 3), threshold, max_det=self.max_det)
        msec = int((time.time() - t0) * 1000)
        return _native.pack_wire(dets[0, :counts[0]], reqid, msec, saturate)

    def perform_stream(self, batches, threshold=0.1, allow_resize=False):
        """Generator over an iterable of batches, each an [n, h, w, 3] u8 array or a list of JPEG payloads: yields one list of per-frame result lists per
        batch, in order, with two batches in flight (fd_submit / fd_collect) so the host->device copy of batch
        i+1 overlaps the compute of batch i.  Same results as perform_frames on each batch."""
        self.ANCHORS[self.model.n_heads]
        pending = []  # slots in submission order
        slot = 0
        for frames in batches:
            encoded = not isinstance(frames, np.ndarray) and len(frames) and isinstance(frames[0], (bytes, bytearray, memoryview))
            if not encoded:
                frames = np.ascontiguousarray(frames, np.uint8)
            if len(pending) == _native.FD_MAX_SLOTS:
                dets, counts, _ = self.model.collect(pending.pop(0))
                yield self._tuples(dets, counts)
            if encoded:  # a list of JPEG payloads: entropy decode here (host pool) while the device runs the other slot
                try:
                    self.model.submit_jpeg(slot, frames, threshold, max_det=self.max_det)
                    self.jpeg_device_frames += len(frames)
                except _native.JpegRefused:
                    decoded = np.stack([self._decode_host(d) for d in frames])
                    self.jpeg_host_frames += len(frames)
                    self.model.submit(slot, decoded, threshold, max_det=self.max_det)
            else:
                self.model.submit(slot, frames, threshold, allow_resize=allow_resize, max_det=self.max_det)
            pending.append(slot)
            slot = (slot + 1) % _native.FD_MAX_SLOTS
        for s in pending:
            dets, counts, _ = self.model.collect(s)
            yield self._tuples(dets, counts)

    def forward_raw(self, frames):
        """Raw head tensors (what ``model.run`` returns in the reference): list of f32 [n, C, H, W]."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, h, w, _ = frames.shape
        if (w, h) != self.image_size:
            raise ValueError('invalid image size')
        self.model.preprocess(frames, n, (w, h))
        self.model.forward(n)
        return self.model.heads(n)


USAGE = 'usage: {prog} [-m mode] [-c num_classes] [-t threshold] onnx images ...'


def main(argv):
    """Command line of the reference (server/detector.py:169-195): one line ``<seconds> <results>`` per image."""
    import getopt
    settings = {'-m': None, '-c': '80', '-t': '0.1'}
    try:
        flags, rest = getopt.getopt(argv[1:], 'm:c:t:')
    except getopt.GetoptError:
        rest = []
    else:
        settings.update(flags)
    if not rest:
        print(USAGE.format(prog=argv[0]))
        return 100
    onnx_path, images = rest[0], rest[1:]
    det = ONNXDetector(onnx_path, mode=settings['-m'], num_classes=int(settings['-c']))
    thr = float(settings['-t'])
    for image in images:
        with open(image, 'rb') as fp:
            payload = fp.read()
        started = time.time()
        found = det.perform(payload, threshold=thr)
        print(time.time() - started, found)
    return None


if __name__ == '__main__':
    sys.exit(main(sys.argv))
