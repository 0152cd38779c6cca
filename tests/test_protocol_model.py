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

# (M, NB, NSA, NSB0, NSB1, NG, DLAG, bins, shared A tiles): the shapes of TcCfg in das_map_tc.cu
SHAPES = [
    (256, 128, 7, 4, 3, 3, 2, 4, True),      # cfg5: one pair of column blocks (FP16x3)
    (256, 128, 6, 2, 2, 3, 2, 3, True),      # single-pass TF32
    (300, 128, 7, 4, 3, 3, 2, 3, True),      # a pair plus a lone block, chunk count off the block size
    (520, 128, 7, 4, 3, 3, 2, 2, True),      # two pairs plus a lone block
    (1024, 128, 7, 4, 3, 3, 2, 2, True),     # the largest array: tdone ring of 2 NKC barriers
    (144, 128, 7, 4, 3, 3, 2, 4, True),      # second block of one chunk
    (64, 64, 10, 4, 3, 3, 2, 5, True),       # single block
    (16, 32, 10, 4, 3, 1, 2, 9, True),       # one chunk per bin: one group (NB = 32)
    (10, 32, 10, 4, 3, 1, 1, 9, True),
    (32, 32, 10, 4, 3, 1, 2, 7, True),       # two chunks per bin
    (40, 64, 10, 4, 3, 3, 1, 6, True),       # as many chunks per bin as groups
    (256, 128, 7, 4, 3, 3, 1, 3, True),      # one own chunk in flight
    (256, 128, 7, 4, 3, 2, 3, 3, True),      # two groups, three own chunks in flight
    (256, 128, 3, 2, 0, 1, 2, 3, False),     # split-TF32 / mixed: shallow rings, no sharing, one group
    (64, 64, 5, 2, 0, 2, 2, 4, False),
    (100, 128, 4, 2, 0, 2, 1, 4, False),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_protocol_has_no_deadlock_or_aliasing(shape):
    M, NB, NSA, NSB0, NSB1, NG, DLAG, nbins, share = shape
    for seed in range(20):
        assert sim.main(M, NB, NSA, NSB0, NSB1, NG, nbins, share, seed, DLAG=DLAG), (shape, seed)


def test_shapes_the_kernel_configuration_must_not_produce_are_rejected_by_the_model():
    """An A ring shallower than the chunks the groups keep in flight (NSA < DLAG NG) starves the issue warps, and a
    tdone ring shorter than 2 NKC barriers lets a barrier complete a further use while a group still waits on it
    (parity aliasing): TcCfg clamps NG / DLAG and sizes the ring for the largest array; the model refuses such shapes."""
    with pytest.raises(AssertionError):
        sim.main(256, 128, 3, 2, 0, 2, 2, False, 0, DLAG=2)
    with pytest.raises(AssertionError):
        sim.main(256, 128, 7, 4, 3, 3, 2, True, 0, DLAG=2, ND=16)
