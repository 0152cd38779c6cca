"""Batched Hermitian Jacobi eigensolver (csrc/heig_jacobi.cu) behind engine.hermitian_eig: the contract of
np.linalg.eigh, which the reference calls per bin (beamforming/beamforming.py:1107; the SVD of :1201 is the same
decomposition for a Hermitian matrix)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from dsptoolbox_b200 import engine, _native as N   # noqa: E402


def _matrices(rng, M, kind, nb=7):
    out = []
    for _ in range(nb):
        r = {"psd": 3 * M, "rank_deficient": max(1, M // 4), "indefinite": 3 * M}[kind]
        x = rng.standard_normal((M, r)) + 1j * rng.standard_normal((M, r))
        a = x @ x.conj().T
        if kind == "indefinite":
            a -= np.eye(M) * 0.6 * np.trace(a).real / M
        out.append(a)
    return np.stack(out)


@pytest.mark.parametrize("kind", ["psd", "rank_deficient", "indefinite"])
@pytest.mark.parametrize("M", [1, 2, 5, 24, 64, 80])
def test_jacobi_against_lapack(M, kind):
    assert M <= N.lib().b2_heig_max_order()
    rng = np.random.default_rng(100 * M + len(kind))
    a = _matrices(rng, M, kind)
    w, v = engine.hermitian_eig(torch.from_numpy(a).cuda())
    w, v = w.cpu().numpy(), v.cpu().numpy()
    ref = np.linalg.eigvalsh(a)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.all(np.diff(w, axis=1) >= 0)                                     # ascending, like eigh
    assert np.max(np.abs(w - ref) / scale) <= 1e-13
    assert np.max(np.abs(a @ v - v * w[:, None, :]) / scale[:, :, None]) <= 1e-13      # A v = lambda v
    assert np.max(np.abs(np.swapaxes(v.conj(), 1, 2) @ v - np.eye(M))) <= 1e-13        # unitary


def test_only_the_lower_triangle_is_read_and_large_orders_fall_through():
    rng = np.random.default_rng(3)
    a = _matrices(rng, 16, "psd", nb=3)
    junk = a.copy()
    iu = np.triu_indices(16, 1)
    junk[:, iu[0], iu[1]] = 123.0 + 7j
    w0, _ = engine.hermitian_eig(torch.from_numpy(a).cuda())
    w1, _ = engine.hermitian_eig(torch.from_numpy(junk).cuda())
    assert torch.equal(w0, w1)
    M = N.lib().b2_heig_max_order() + 1                                         # torch.linalg.eigh beyond the limit
    b = _matrices(rng, M, "psd", nb=2)
    w, v = engine.hermitian_eig(torch.from_numpy(b).cuda())
    assert np.max(np.abs(w.cpu().numpy() - np.linalg.eigvalsh(b))) <= 1e-10 * np.abs(b).max() * M
    assert N.lib().b2_heig_jacobi(None, 4, 1, None, None, None, None) == N.E_INVAL
