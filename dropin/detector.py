"""Drop-in `detector` module: put this directory ahead of the reference's `server/` on sys.path (or copy this
file over `server/detector.py`) and `server/server.py` runs unchanged on the B200 kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200.detector import *  # noqa: F401,F403,E402
from fastdet_b200.detector import Detector, DummyDetector, ONNXDetector, main  # noqa: F401,E402

if __name__ == '__main__':
    sys.exit(main(sys.argv))
