"""Moved to oracle/ref_harness.py (bench.py's reference arm uses it too); kept importable under the old name."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle.ref_harness import *          # noqa: F401,F403,E402
from oracle.ref_harness import REFERENCE_ROOT, reference_available, load_reference   # noqa: F401,E402
