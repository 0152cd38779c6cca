"""CPU model of the map kernel's mbarrier protocol (roles as coroutines, randomly stalled) to find logical
deadlocks and phase-aliasing bugs before they cost GPU time.

Barriers only know the PARITY of a phase, exactly like mbarrier.try_wait.parity, so a role that gets a whole
ring round away from its partner is modelled faithfully (it passes or blocks on the wrong phase).  Besides
deadlock the model checks that every consumer sees the tile it expects and that an accumulator is complete
when the epilogue reads it.

usage: python scripts/sim_protocol.py M NB NSA NSB LAG nbins [share=1] [runs=300] [alternate]
"""
import random
import sys


class Bar:
    def __init__(self, count):
        self.count = count; self.pending = count; self.phase = 0

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals"
        if self.pending == 0:
            self.phase += 1; self.pending = self.count

    def done(self, parity):
        return (self.phase & 1) != parity


def walk_pair(NKC, kcA, kcB, LAG):
    """(kc, target, first_use_of_A, last_use_of_A) -- mirrors walk_pair in das_map_tc.cu"""
    n1, n0 = NKC - kcB, NKC - kcA
    lag = min(n1, LAG)
    for i in range(n0 + lag):
        if i < n1:
            yield (NKC - 1 - i, 1, True, False)
        j = i - lag
        if j >= 0:
            yield (NKC - 1 - j, 0, j >= n1, True)


ALTERNATE = False      # False: warp 0 multiplies the b0 stream, warp 1 the b1 stream (what das_map_tc.cu does);
                       # True: the steps of every stream alternate between the warps (tried: not faster)


def owner(tgt, i):
    """issue warp that multiplies step i of stream tgt"""
    if not ALTERNATE:
        return tgt
    return ((i + 1) & 1) if tgt else (i & 1)


def main(M, NB, NSA, NSB, LAG, nbins, share=True, seed=0, stall=0.9, quiet=True):
    KCH = 16
    NKC = (M + KCH - 1) // KCH
    NBLK = (M + NB - 1) // NB
    NPAIR = (NBLK + 1) // 2 if share else NBLK
    b0 = lambda p: 2 * p if share else p
    b1 = lambda p: (2 * p + 1) if (share and 2 * p + 1 < NBLK) else -1
    fc = lambda b: b * NB // KCH
    fullA = [Bar(1) for _ in range(NSA)]; emptyA = [Bar(2) for _ in range(NSA)]
    fullB = [Bar(1) for _ in range(NSB)]; emptyB = [Bar(2) for _ in range(NSB)]
    tfull = [Bar(2 if ALTERNATE else 1) for _ in range(2)]; tempty = [Bar(1), Bar(1)]; tinit = [Bar(1), Bar(1)]
    inA = [None] * NSA; inB = [None] * NSB            # what the producers last put into a stage
    acc_steps = {}                                       # accumulator -> MMA steps issued
    acc_need = {}

    def synth():
        sa = pha = 0; seq = 0
        total = nbins * sum(NKC - fc(b0(p)) for p in range(NPAIR))
        for n in range(total):
            while not emptyA[sa].done(pha ^ 1): yield ("synth wait emptyA", sa, pha ^ 1)
            inA[sa] = seq; seq += 1
            fullA[sa].arrive()
            sa += 1
            if sa == NSA: sa = 0; pha ^= 1
            yield None

    def loader():
        sb = phb = 0; q = 0
        for fi in range(nbins):
            for p in range(NPAIR):
                kcA = fc(b0(p)); kcB = fc(b1(p)) if b1(p) >= 0 else NKC
                for (kc, tgt, fu, lu) in walk_pair(NKC, kcA, kcB, LAG):
                    while not emptyB[sb].done(phb ^ 1): yield ("loader wait emptyB", sb)
                    inB[sb] = q; q += 1
                    fullB[sb].arrive()
                    sb += 1
                    if sb == NSB: sb = 0; phb ^= 1
                    yield None

    def issuer(my):
        sb = phb = 0; seqA = 0; acc_it = 0; q = 0
        for fi in range(nbins):
            for p in range(NPAIR):
                kcA = fc(b0(p)); kcB = fc(b1(p)) if b1(p) >= 0 else NKC
                acc1 = acc_it; acc0 = acc_it + (1 if b1(p) >= 0 else 0); acc_it = acc0 + 1
                seq0 = seqA; seqA += NKC - kcA
                for (kc, tgt, fu, lu) in walk_pair(NKC, kcA, kcB, LAG):
                    i = NKC - 1 - kc
                    seq = seq0 + i; sa = seq % NSA; pha = (seq // NSA) & 1
                    acc = acc1 if tgt else acc0; buf = acc & 1; use_par = (acc >> 1) & 1
                    n = NKC - (kcB if tgt else kcA)                     # steps of this stream
                    acc_need[acc] = n
                    mine = owner(tgt, i) == my
                    if i == 0 and (mine or ALTERNATE):                  # buffer drained?
                        while not tempty[buf].done(use_par ^ 1): yield (f"W{my} wait tempty", buf)
                        if ALTERNATE and not mine and n == 1: tfull[buf].arrive()     # no step of this stream is mine
                    while not fullB[sb].done(phb): yield (f"W{my} wait fullB", sb, phb, kc)
                    assert inB[sb] == q, f"W{my} sees B tile {inB[sb]} instead of {q}"
                    if mine or (tgt == 0 and fu):                       # tiles I multiply or must release
                        while not fullA[sa].done(pha): yield (f"W{my} wait fullA", sa, pha, kc)
                        assert inA[sa] == seq, f"W{my} sees A tile {inA[sa]} instead of {seq}"
                    if mine:
                        if ALTERNATE and i == 1:                        # first MMA into an accumulator the other warp started
                            while not tinit[buf].done(use_par): yield (f"W{my} wait tinit", buf)
                        acc_steps[acc] = acc_steps.get(acc, 0) + 1
                        emptyA[sa].arrive(); emptyB[sb].arrive()
                        if i == 0: tinit[buf].arrive()
                        if i == n - 1 or (ALTERNATE and i == n - 2): tfull[buf].arrive()   # my last step of this stream
                    else:
                        if tgt == 0 and fu: emptyA[sa].arrive()
                        emptyB[sb].arrive()
                    q += 1; sb += 1
                    if sb == NSB: sb = 0; phb ^= 1
                    yield None

    def epilogue():
        acc_it = 0
        for fi in range(nbins):
            for t in range(NBLK):
                buf = acc_it & 1
                while not tfull[buf].done((acc_it >> 1) & 1): yield ("epi wait tfull", buf, acc_it)
                assert acc_steps.get(acc_it, 0) == acc_need[acc_it], f"accumulator {acc_it} read after {acc_steps.get(acc_it, 0)} of {acc_need[acc_it]} steps"
                tempty[buf].arrive()
                acc_it += 1
                yield None

    roles = {"synth": synth(), "loader": loader(), "W0": issuer(0), "W1": issuer(1), "epi": epilogue()}
    rng = random.Random(seed)
    state = {}
    slow = {n: (rng.random() < 0.5) for n in roles}          # some roles are slow in this run
    while roles:
        progressed = stalled = False
        names = list(roles); rng.shuffle(names)
        for n in names:
            if slow[n] and rng.random() < stall:
                stalled = True
                continue
            try:
                r = next(roles[n])
            except StopIteration:
                del roles[n]; progressed = True; continue
            if r is None: progressed = True
            state[n] = r
        if not progressed and not stalled:
            print("DEADLOCK:", {n: state.get(n) for n in roles}); return False
    if not quiet: print("ok")
    return True


if __name__ == "__main__":
    a = sys.argv
    M, NB, NSA, NSB, LAG, nbins = (int(x) for x in a[1:7])
    share = len(a) < 8 or a[7] != "0"
    if len(a) > 9 and a[9] == "alternate":
        ALTERNATE = True
    bad = 0
    for seed in range(int(a[8]) if len(a) > 8 else 300):
        try:
            ok = main(M, NB, NSA, NSB, LAG, nbins, share, seed)
        except AssertionError as e:
            print("BUG (seed %d):" % seed, e); ok = False
        if not ok:
            bad += 1
            if bad >= 3: break
    print(f"M={M} NB={NB} NSA={NSA} NSB={NSB} LAG={LAG} share={share}: failures {bad}")
