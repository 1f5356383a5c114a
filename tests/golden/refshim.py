"""The reference import shim lives in oracle/refshim.py (bench.py's reference arm uses it too)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.refshim import *  # noqa: F401,F403,E402
from oracle.refshim import REF, _Registry, _build_from_cfg, available, install_stubs, load  # noqa: F401,E402
