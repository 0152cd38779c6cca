"""CPU models of the index arithmetic two kernels rely on (no GPU): the fused last pass + untangling step of the Welch
FFT kernel with its slot-interleaved spectra layout (csrc/welch_fft.cu: last_pass_untangle, common.cuh: b2_bin_slot), and
the round-robin pair schedule + rotation of the Jacobi eigensolver (csrc/heig_jacobi.cu).  The GPU tests check the
kernels themselves; these pin the algebra they implement."""
import numpy as np
import pytest


def bin_slot(f, nc):                                   # common.cuh: b2_bin_slot
    if 2 * f < nc:
        return f
    if 2 * f == nc:
        return nc
    return nc // 2 + (nc - f)


@pytest.mark.parametrize("nc,radix", [(8, 2), (64, 4), (512, 2), (2048, 8)])
def test_fused_last_pass_untangling_model(nc, radix):
    """Two real frames of 2*nc samples -> their rfft bins, the way a thread of the kernel sees it: after the last
    radix-R butterfly thread j holds Z[j + J r] (J = nc / R); it keeps r < R/2, reads the partners Z[nc - p] from the
    exchange buffer E[q][J - j] written by thread J - j (q = R/2 - 1 - r; thread 0 pairs with itself, q = R/2 - r) and
    writes bin p to slot p and bin nc - p to slot nc/2 + p; thread 0 also owns the self-paired bin nc/2 (slot nc)."""
    rng = np.random.default_rng(nc)
    wl, J, H = 2 * nc, nc // radix, radix // 2
    y = rng.standard_normal(wl)
    Z = np.fft.fft(y[0::2] + 1j * y[1::2])             # the packed complex FFT the passes before produce
    held = {j: [Z[j + J * r] for r in range(radix)] for j in range(J)}
    E = np.zeros((H, J), dtype=complex)
    for j in range(J):                                 # publish the upper half
        for q in range(H):
            E[q, j] = held[j][H + q]
    tw = np.exp(-2j * np.pi * np.arange(wl) / wl)
    slots = np.full(4 * (nc // 4 + 1), np.nan, dtype=complex)

    def untangle(a, bc, w):
        s, d = a + bc, w * (a - bc)
        return 0.5 * (s - 1j * d), 0.5 * np.conj(s + 1j * d)

    for j in range(J):
        own = j == 0
        for r in range(H):
            p = j + J * r
            a = held[j][r]
            if own and r == 0:
                bc = np.conj(a)
            else:
                bc = np.conj(E[H - r if own else H - 1 - r, 0 if own else J - j])
            lo, hi = untangle(a, bc, tw[p])
            slots[p], slots[nc // 2 + p] = lo, hi
        if own:
            a = held[0][H]
            slots[nc] = untangle(a, np.conj(a), -1j)[0]
    ref = np.fft.rfft(y)
    got = np.array([slots[bin_slot(f, nc)] for f in range(nc + 1)])
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.abs(ref).max()
    used = sorted(bin_slot(f, nc) for f in range(nc + 1))
    assert used == list(range(nc + 1))                 # a permutation: every slot below nc + 1 exactly once


@pytest.mark.parametrize("M", [2, 5, 8, 17, 64])
def test_jacobi_schedule_and_rotation_model(M):
    """Every sweep of the tournament schedule meets every pair (p, q) exactly once with disjoint pairs per step (what
    lets the kernel rotate M/2 pairs between two barriers), and the rotation it applies annihilates a_pq."""
    Me = M + (M & 1)
    seen = set()
    for s in range(Me - 1):
        step = []
        for t in range(Me // 2):
            p, q = (Me - 1, s) if t == 0 else ((s + t) % (Me - 1), (s - t + Me - 1) % (Me - 1))
            p, q = min(p, q), max(p, q)
            assert p != q
            step += [p, q]
            if q < M:
                seen.add((p, q))
        assert len(set(step)) == Me                    # disjoint: each index in exactly one pair
    assert seen == {(p, q) for p in range(M) for q in range(p + 1, M)}

    rng = np.random.default_rng(M)
    x = rng.standard_normal((M, M)) + 1j * rng.standard_normal((M, M))
    A = x @ x.conj().T
    p, q = 0, M - 1
    a, b, c = A[p, p].real, A[q, q].real, A[p, q]
    ac = abs(c)
    tau = (b - a) / (2 * ac)
    t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + np.sqrt(1 + tau * tau))
    cs = 1 / np.sqrt(1 + t * t)
    sn = t * cs
    e = np.conj(c) / ac
    Jm = np.eye(M, dtype=complex)
    Jm[p, p], Jm[p, q], Jm[q, p], Jm[q, q] = cs, sn, -sn * e, cs * e      # columns: x_p' = cs x_p - sn e x_q, x_q' = sn x_p + cs e x_q
    B = Jm.conj().T @ A @ Jm
    assert abs(B[p, q]) <= 1e-13 * np.abs(A).max()
    assert abs(B[p, p] - (a - t * ac)) <= 1e-12 * np.abs(A).max() and abs(B[q, q] - (b + t * ac)) <= 1e-12 * np.abs(A).max()
    assert np.allclose(np.linalg.eigvalsh(B), np.linalg.eigvalsh(A))
