"""Oracle (test infrastructure only): the response payload of the reference server, restated from
server/server.py:234-239 (DetectService.process_data) — per detection ``struct.pack('>BBhhhh', klass, int(conf*255),
int(x), int(y), int(w), int(h))`` and the header ``struct.pack('>4sLLL', b'YOLO', reqid, msec, len(buf))``.
Pinned: tests/test_wire_service.py replays the reference's own DummyDetector answer (server/detector.py:83-92) through both, and
tests/test_reference_fixtures.py compares it with what the reference's unchanged server.py sends over loopback."""
import struct


def pack_results(results, reqid: int, msec: int) -> bytes:
    buf = b''
    for (klass, conf, x, y, w, h) in results:
        buf += struct.pack('>BBhhhh', klass, int(conf * 255), int(x), int(y), int(w), int(h))
    return struct.pack('>4sLLL', b'YOLO', reqid, msec, len(buf)) + buf
