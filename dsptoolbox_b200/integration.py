"""The binding a dsptoolbox maintainer adds to run the reference's OWN classes on the B200 path
(INTEGRATION.md section 2 shows this file).

`install(dsptoolbox)` patches the two places where the reference enters the hot path:

* `dsptoolbox.classes.signal._csm_welch` -- `classes/signal.py:39` binds the function by name at import,
  so THAT binding is replaced (and the defining module's, for direct callers) by a function with the
  signature of `standard/_spectral_methods.py:285-294` that runs K0/K1/K2/K2e and returns host arrays;
* `BeamformerDASFrequency.get_beamformer_map` (`beamforming/beamforming.py:805-880`) -- replaced as a
  method; it keeps the reference's side effects (`center_frequency_hz`, `octave_fraction`, `f_range_hz`,
  `map`, returned copy) and its own helpers (`_get_fractional_octave_bandwidth`,
  `find_nearest_points_index_in_vector`, `grid.reconstruct_map_shape`), takes the CSM parameters from the
  reference `Signal`, and keeps the matrix on the device between the two stages.

`uninstall()` restores the originals.  The reference's enums are translated by NAME (the reference's
`Window` / `SpectrumScaling` members and ours are spelled alike); its steering functions are identified by
identity, like `beamforming.py:634-642` does.
"""

from __future__ import annotations

import numpy as np

from . import _native as N
from . import engine
from .beamforming.beamforming import simpson_weights
from .standard import _spectral_methods as b2sm
from .standard.enums import SpectrumScaling as B2Scaling

_saved: dict = {}


def _window_spec(window_type):
    """scipy window spec of a reference `Window` member (enums.py:395-437) or a plain string."""
    return window_type.to_scipy_format() if hasattr(window_type, "to_scipy_format") else window_type


def _scaling(scaling):
    return B2Scaling[scaling.name] if hasattr(scaling, "name") else scaling


def csm_welch_b200(time_data, sampling_rate_hz, window_length_samples, window_type, overlap_percent, detrend,
                   average, scaling):
    """Drop-in for the reference's `_csm_welch` (same positional signature, host arrays out)."""
    f, csm = b2sm.csm_welch_device(time_data, sampling_rate_hz, window_length_samples, _window_spec(window_type),
                                   overlap_percent, detrend, average, _scaling(scaling))
    return f, engine.to_host(csm)


def _make_get_beamformer_map(ref):
    bfm = ref.beamforming.beamforming
    steer = {bfm.classic_steering: N.STEER_CLASSIC, bfm.inverse_steering: N.STEER_INVERSE,
             bfm.true_power_steering: N.STEER_TRUE_POWER, bfm.true_location_steering: N.STEER_TRUE_LOCATION}

    def get_beamformer_map(self, center_frequency_hz, octave_fraction=3, remove_csm_diagonal=True):
        self.center_frequency_hz, self.octave_fraction = center_frequency_hz, octave_fraction
        self.f_range_hz = bfm._get_fractional_octave_bandwidth(center_frequency_hz, octave_fraction)
        txt = "Beamformer computation has started successfully:"
        print("\n" + txt)
        print("-" * len(txt))
        print("...csm...")
        assert self.signal.number_of_channels > 1, (
            "Cross spectral matrix can only be computed when at least two channels are available")
        p = self.signal._spectrum_parameters
        if p["method"].name != "WelchPeriodogram":
            raise NotImplementedError("the patched map runs on the Welch CSM")
        f, csm = b2sm.csm_welch_device(self.signal.time_data, self.signal.sampling_rate_hz,
                                       p["window_length_samples"], _window_spec(p["window_type"]),
                                       p["overlap_percent"], p["detrend"], p["average"], _scaling(p["scaling"]))
        print("...Steering vector...")
        ids = bfm.find_nearest_points_index_in_vector(self.f_range_hz, f)
        id1, id2 = int(ids[0]), int(ids[1])
        if id1 == id2:
            id2 += 1
        fsel = f[id1:id2]
        self.f_range_hz = np.array([fsel[0], fsel[-1]])
        w = simpson_weights(len(fsel), fsel[1] - fsel[0]) if len(fsel) > 1 else np.ones(1)
        mode = engine.resolve_map_mode()
        op = engine.das_prepare(csm, id1, len(fsel), remove_csm_diagonal, mode)
        print("...Apply...")
        out, _ = engine.das_map(op, self.mics.number_of_points, fsel, w, self.c, self.mics.coordinates,
                                self.mics.array_center_coordinates, self.grid.coordinates,
                                steer[self.st_vec.get_vector], remove_csm_diagonal, mode)
        self.map = self.grid.reconstruct_map_shape(out.cpu().numpy())
        return self.map.copy()

    return get_beamformer_map


def install(ref=None):
    """Patch the reference package `ref` (default: `import dsptoolbox`).  Idempotent."""
    if ref is None:
        import dsptoolbox as ref
    if _saved.get("ref") is ref:
        return ref
    sig_mod = ref.classes.signal
    sm_mod = ref.standard._spectral_methods
    cls = ref.beamforming.beamforming.BeamformerDASFrequency
    _saved.update(ref=ref, sig=sig_mod._csm_welch, sm=sm_mod._csm_welch, gbm=cls.get_beamformer_map)
    sig_mod._csm_welch = csm_welch_b200                              # the signal.py:39 binding
    sm_mod._csm_welch = csm_welch_b200
    cls.get_beamformer_map = _make_get_beamformer_map(ref)
    return ref


def uninstall():
    ref = _saved.pop("ref", None)
    if ref is None:
        return
    ref.classes.signal._csm_welch = _saved.pop("sig")
    ref.standard._spectral_methods._csm_welch = _saved.pop("sm")
    ref.beamforming.beamforming.BeamformerDASFrequency.get_beamformer_map = _saved.pop("gbm")
