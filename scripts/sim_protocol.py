"""CPU model of the map kernel's mbarrier protocol (roles as coroutines) to find logical deadlocks.

usage: python scripts/sim_protocol.py M NB NSA NSB LAG nbins [share]
"""
import sys, random

class Bar:
    def __init__(self, count): self.count = count; self.pending = count; self.phase = 0
    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals"
        if self.pending == 0: self.phase += 1; self.pending = self.count
    def done(self, parity): return (self.phase & 1) != parity

def walk_pair(NKC, kcA, kcB, LAG):
    n1, n0 = NKC - kcB, NKC - kcA
    lag = min(n1, LAG)
    for i in range(n0 + lag):
        if i < n1: yield (NKC - 1 - i, 1, True, False)
        j = i - lag
        if j >= 0: yield (NKC - 1 - j, 0, j >= n1, True)

def main(M, NB, NSA, NSB, LAG, nbins, share=True, seed=0, emptyb_count=2, stall=0.5, quiet=False):
    KCH = 16
    NKC = (M + KCH - 1) // KCH
    NBLK = (M + NB - 1) // NB
    NPAIR = (NBLK + 1) // 2 if share else NBLK
    b0 = lambda p: 2 * p if share else p
    b1 = lambda p: (2 * p + 1) if (share and 2 * p + 1 < NBLK) else -1
    fc = lambda b: b * NB // KCH
    fullA = [Bar(1) for _ in range(NSA)]; emptyA = [Bar(2) for _ in range(NSA)]
    fullB = [Bar(1) for _ in range(NSB)]; emptyB = [Bar(emptyb_count) for _ in range(NSB)]
    tfull = [Bar(1), Bar(1)]; tempty = [Bar(1), Bar(1)]
    log = []

    def synth():
        sa = pha = 0
        total = nbins * sum(NKC - fc(b0(p)) for p in range(NPAIR))
        for n in range(total):
            while not emptyA[sa].done(pha ^ 1): yield ("synth wait emptyA", sa, pha ^ 1)
            fullA[sa].arrive()
            sa += 1
            if sa == NSA: sa = 0; pha ^= 1
            yield None

    def loader():
        sb = phb = 0
        for fi in range(nbins):
            for p in range(NPAIR):
                kcA = fc(b0(p)); kcB = fc(b1(p)) if b1(p) >= 0 else NKC
                for (kc, tgt, fu, lu) in walk_pair(NKC, kcA, kcB, LAG):
                    while not emptyB[sb].done(phb ^ 1): yield ("loader wait emptyB", sb)
                    fullB[sb].arrive()
                    sb += 1
                    if sb == NSB: sb = 0; phb ^= 1
                    yield None

    def issuer(my):
        sb = phb = 0; seqA = 0; acc_it = 0
        for fi in range(nbins):
            for p in range(NPAIR):
                kcA = fc(b0(p)); kcB = fc(b1(p)) if b1(p) >= 0 else NKC
                acc1 = acc_it; acc0 = acc_it + (1 if b1(p) >= 0 else 0); acc_it = acc0 + 1
                seq0 = seqA; seqA += NKC - kcA
                for (kc, tgt, fu, lu) in walk_pair(NKC, kcA, kcB, LAG):
                    seq = seq0 + (NKC - 1 - kc); sa = seq % NSA; pha = (seq // NSA) & 1
                    if tgt != my:
                        if my == 1 and fu:
                            while not fullA[sa].done(pha): yield (f"D{my} skip wait fullA", sa, pha, kc)
                            emptyA[sa].arrive()
                        # skipped B tile: wait for it (keeps this warp inside the ring's current round) ...
                        while not fullB[sb].done(phb): yield (f"D{my} skip wait fullB", sb, phb, kc)
                        if emptyb_count == 2: emptyB[sb].arrive()      # ... and let the loader recycle it
                        sb += 1
                        if sb == NSB: sb = 0; phb ^= 1
                        yield None
                        continue
                    acc = acc1 if tgt else acc0; buf = acc & 1
                    if kc == NKC - 1:
                        while not tempty[buf].done(((acc >> 1) & 1) ^ 1): yield (f"D{my} wait tempty", buf)
                    while not fullB[sb].done(phb): yield (f"D{my} wait fullB", sb, phb, kc)
                    while not fullA[sa].done(pha): yield (f"D{my} wait fullA", sa, pha, kc)
                    emptyA[sa].arrive(); emptyB[sb].arrive()
                    kbase = kcB if tgt else kcA
                    if kc == kbase: tfull[buf].arrive()
                    sb += 1
                    if sb == NSB: sb = 0; phb ^= 1
                    yield None

    def epilogue():
        acc_it = 0
        for fi in range(nbins):
            for t in range(NBLK):
                buf = acc_it & 1
                while not tfull[buf].done((acc_it >> 1) & 1): yield ("epi wait tfull", buf, acc_it)
                tempty[buf].arrive()
                acc_it += 1
                yield None

    roles = {"synth": synth(), "loader": loader(), "D0": issuer(0), "D1": issuer(1), "epi": epilogue()}
    rng = random.Random(seed)
    state = {}
    slow = {n: (rng.random() < 0.5) for n in roles}          # some roles are slow in this run
    while roles:
        progressed = False
        names = list(roles); rng.shuffle(names)
        stalled = False
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
    ebc = int(a[8]) if len(a) > 8 else 2
    bad = 0
    for seed in range(int(a[9]) if len(a) > 9 else 300):
        if not main(M, NB, NSA, NSB, LAG, nbins, share, seed, ebc, 0.9, True):
            bad += 1
            if bad >= 3: break
    print("deadlocks:", bad)
