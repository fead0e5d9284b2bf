"""Drop-in `detector` module: put this directory ahead of the reference's `server/` on sys.path (or copy this
file over `server/detector.py`) and `server/server.py` runs unchanged on the B200 kernels."""
import os
import sys

try:
    import fastdet_b200  # noqa: F401  (installed, or on PYTHONPATH)
except ImportError:  # run from a checkout: FASTDET_B200_HOME, or this file still sits in <checkout>/dropin/
    sys.path.insert(0, os.environ.get("FASTDET_B200_HOME") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastdet_b200.detector import *  # noqa: F401,F403,E402
from fastdet_b200.detector import Detector, DummyDetector, ONNXDetector, main  # noqa: F401,E402

if __name__ == '__main__':
    sys.exit(main(sys.argv))
