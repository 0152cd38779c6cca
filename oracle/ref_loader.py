"""Import the UNMODIFIED reference from /root/reference with its I/O + plotting deps stubbed.

TEST INFRASTRUCTURE ONLY (see oracle/restatement.py).  The reference imports soundfile,
sounddevice, matplotlib and seaborn at module scope (classes/signal.py:9-11,
audio_io/audio_io.py:18, beamforming/_beamforming.py:6-7, plots/plots.py:5-9); none of the
hot path's arithmetic lives there, so MagicMock stand-ins are enough.
/root/reference only exists in the build container.  `oracle/install_reference.py` installs the same
unmodified package into baseline/_ref/ (git-ignored, but it travels with gpurun), so the GPU box can
run the real reference too: the roots are tried in the order $B2DSP_REFERENCE_ROOT, /root/reference,
<repo>/baseline/_ref.  The committed golden vectors under tests/golden/ carry its outputs as well.
"""

from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ROOTS = [r for r in (os.environ.get("B2DSP_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")) if r]
REFERENCE_ROOT = next((r for r in _ROOTS if os.path.isdir(os.path.join(r, "dsptoolbox"))), _ROOTS[0])
_STUBS = (
    "soundfile", "sounddevice", "seaborn", "matplotlib", "matplotlib.pyplot",
    "matplotlib.figure", "matplotlib.axes", "matplotlib.ticker", "matplotlib.colormaps",
    "matplotlib.colors", "matplotlib.lines", "matplotlib.patches", "matplotlib.gridspec",
)
_cached = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dsptoolbox"))


def load():
    """Return the reference `dsptoolbox` module (v0.8)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import dsptoolbox  # noqa: E402

    _cached = dsptoolbox
    return dsptoolbox


def read_array_xml(path: str | None = None):
    """Mic positions of example_data/array.xml as a dict(x=, y=, z=) of float64 arrays."""
    import numpy as np
    import xml.etree.ElementTree as ET

    path = path or os.path.join(REFERENCE_ROOT, "example_data", "array.xml")
    pos = [(float(p.get("x")), float(p.get("y")), float(p.get("z"))) for p in ET.parse(path).getroot().iter("pos")]
    a = np.asarray(pos, dtype=np.float64)
    return dict(x=a[:, 0], y=a[:, 1], z=a[:, 2])
