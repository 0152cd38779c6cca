"""dsptoolbox_b200: B200-native (sm_100a) drop-in for dsptoolbox's Welch-CSM + DAS-map path.

    import dsptoolbox_b200 as dsp
    sig = dsp.Signal(None, time_data, 48_000)
    f, csm = sig.get_csm()
    bf = dsp.beamforming.BeamformerDASFrequency(sig, dsp.beamforming.MicArray(pos), grid,
                                                dsp.beamforming.SteeringVector())
    m = bf.get_beamformer_map(4000, 3)

The names above behave like their dsptoolbox v0.8 counterparts; all arithmetic runs in
hand-written CUDA behind the C ABI of include/b2dsp.h.  There is no CPU fallback.
"""

from . import beamforming  # noqa: F401
from . import transfer_functions  # noqa: F401
from .classes import Signal  # noqa: F401
from .classes.spectrum import Spectrum  # noqa: F401
from .engine import config  # noqa: F401
from .standard.enums import SpectrumMethod, SpectrumScaling, Window  # noqa: F401

__version__ = "0.1.0"
__all__ = ["Signal", "Spectrum", "SpectrumMethod", "SpectrumScaling", "Window", "beamforming", "config",
           "transfer_functions"]
