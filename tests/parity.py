"""Comparison metrics used by the parity tests.

CSM tolerance (BASELINE.json north_star): rtol 1e-5 element-wise.  Bin 0 is excluded when the
frames are de-meaned: it is pure rounding noise (~1e-14) in the reference itself.
Map tolerance: within 1e-4 of the map maximum everywhere, and within 0.01 dB on every point
that lies inside the top 20 dB of the map (below that a dB comparison is ill-conditioned: the
reference clamps negative powers to exactly 0).
"""

import numpy as np

CSM_RTOL = 1e-5
MAP_REL_TO_MAX = 1e-4
MAP_DB = 0.01
MAP_DB_RANGE = 20.0


def csm_errors(got, ref, skip_dc=True):
    got, ref = np.asarray(got), np.asarray(ref)
    sl = slice(1, None) if skip_dc else slice(None)
    g, r = got[sl], ref[sl]
    diff = np.abs(g - r)
    mag = np.abs(r)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(mag > 0, diff / mag, np.where(diff == 0, 0.0, np.inf))
    d = np.sqrt(np.abs(np.einsum("fii->fi", r)))
    coh_norm = diff / (d[:, :, None] * d[:, None, :] + 1e-300)
    return dict(max_rel=float(rel.max()), pass_fraction=float(np.mean(rel <= CSM_RTOL)),
                max_coherence_normalised=float(coh_norm.max()),
                max_rel_to_bin_max=float((diff.reshape(len(diff), -1).max(1) / mag.reshape(len(mag), -1).max(1)).max()))


def assert_csm_close(got, ref, skip_dc=True, rtol=CSM_RTOL, dc_atol_rel=1e-9):
    e = csm_errors(got, ref, skip_dc)
    assert e["max_rel"] <= rtol, f"CSM element-wise rtol {rtol:g} violated: {e}"
    if skip_dc:
        scale = np.abs(np.asarray(ref)[1:]).max()
        assert np.abs(np.asarray(got)[0] - np.asarray(ref)[0]).max() <= dc_atol_rel * scale, "DC bin is not negligible"
    return e


def map_errors(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    mx = ref.max()
    rel = float(np.max(np.abs(got - ref)) / mx)
    sel = ref >= mx * 10 ** (-MAP_DB_RANGE / 10)
    db = float(np.max(np.abs(10 * np.log10(got[sel] / ref[sel])))) if sel.any() else 0.0
    return dict(rel_to_max=rel, max_db_top20=db)


def assert_map_close(got, ref, rel=MAP_REL_TO_MAX, db=MAP_DB):
    assert np.asarray(got).shape == np.asarray(ref).shape
    e = map_errors(got, ref)
    assert e["rel_to_max"] <= rel, f"map differs by more than {rel:g} of its maximum: {e}"
    assert e["max_db_top20"] <= db, f"map differs by more than {db} dB inside its top {MAP_DB_RANGE} dB: {e}"
    return e
