"""Host-side model of the map kernel's mbarrier protocol (scripts/sim_protocol.py): the ring / accumulator
hand-shakes of the synthesis+reduction warps, the loader and the two MMA-issue warps must be free of deadlocks
and of phase aliasing for every block / ring shape the kernel instantiates, under randomly stalled roles; a
chunk's columns must be final when they are read and an accumulator fully read before it is overwritten."""

import importlib.util
import os

import pytest

_SPEC = importlib.util.spec_from_file_location(
    "sim_protocol", os.path.join(os.path.dirname(__file__), "..", "scripts", "sim_protocol.py"))
sim = importlib.util.module_from_spec(_SPEC)
_SPEC.loader.exec_module(sim)

# (M, NB, NSA, NSB, LAG, NG, bins, shared A tiles): the shapes of TcCfg in das_map_tc.cu
SHAPES = [
    (256, 128, 6, 4, 4, 3, 4, True),      # cfg5: one pair of column blocks (FP16x3 / TF32 single pass)
    (256, 128, 8, 3, 6, 4, 3, True),      # other ring shapes
    (300, 128, 6, 4, 4, 3, 3, True),      # a pair plus a lone block, chunk count off the block size
    (520, 128, 6, 4, 4, 3, 2, True),      # two pairs plus a lone block
    (144, 128, 6, 4, 4, 3, 4, True),      # second block of one chunk
    (64, 64, 10, 4, 8, 3, 5, True),       # single block
    (16, 32, 10, 4, 8, 3, 9, True),       # one chunk per bin
    (256, 128, 3, 2, 1, 1, 3, False),     # TF32 modes: shallow rings, no sharing, one group
    (100, 128, 4, 3, 2, 2, 4, False),
    (32, 32, 10, 4, 8, 3, 7, True),       # two chunks per bin
]


@pytest.mark.parametrize("issue3", [False, True])
@pytest.mark.parametrize("shape", SHAPES)
def test_protocol_has_no_deadlock_or_aliasing(shape, issue3):
    M, NB, NSA, NSB, LAG, NG, nbins, share = shape
    sim.ISSUE3 = issue3
    try:
        for seed in range(15):
            assert sim.main(M, NB, NSA, NSB, LAG, NG, nbins, share, seed), (shape, seed)
    finally:
        sim.ISSUE3 = False


def test_too_many_groups_for_the_ring_are_rejected_by_the_model():
    """2 NG > NSA lets a ring slot complete a further use while a group still has to wait for it (parity aliasing):
    the kernel's configuration clamps NG to NSA / 2 (TcCfg::NG); the model refuses the shape."""
    with pytest.raises(AssertionError):
        sim.main(256, 128, 3, 2, 1, 2, 2, False, 0)


def test_walk_pair_covers_every_step_once():
    """Every (chunk, block) product of the lower block triangle appears exactly once; the first step of an
    accumulator is its widest chunk; an A tile is first used before it is last used."""
    for NKC, kcA, kcB, LAG in [(16, 0, 8, 6), (19, 0, 8, 6), (16, 0, 16, 6), (9, 0, 8, 6), (33, 16, 24, 6), (4, 0, 4, 8)]:
        steps = list(sim.walk_pair(NKC, kcA, kcB, LAG))
        want = {(kc, 0) for kc in range(kcA, NKC)} | {(kc, 1) for kc in range(kcB, NKC)}
        assert sorted((kc, t) for kc, t, _, _ in steps) == sorted(want)
        for t in (0, 1):
            mine = [kc for kc, tt, _, _ in steps if tt == t]
            if mine:
                assert mine[0] == NKC - 1 and mine == sorted(mine, reverse=True)
        for kc in range(kcA, NKC):
            uses = [(i, fu, lu) for i, (k, _, fu, lu) in enumerate(steps) if k == kc]
            assert uses[0][1] and uses[-1][2] and sum(u[1] for u in uses) == 1 and sum(u[2] for u in uses) == 1
