"""CPU model of the map kernel's mbarrier protocol (roles as coroutines, randomly stalled) to find logical
deadlocks and phase-aliasing bugs before they cost GPU time.

Roles of das_map_tc_kernel: NG synthesis/reduction warp groups that take the chunks round-robin (group j
synthesises chunks j, j + NG, ...; before it synthesises its next chunk it passes all chunks up to its previous
one -- reducing that one after waiting for its `tdone`, signalling `tempty` when a block's lowest chunk goes
by), the B loader, and the three MMA issue warps (W1: stream of block b1; W2: the steps of block b0 on the chunks it
shares with b1; W0: the rest of b0's stream, accumulating behind W2's first step -- `tinit`).  A B tile is waited
for and released by its step's warp only.
Barriers only know the PARITY of a phase, exactly like mbarrier.try_wait.parity, so a role that gets a whole
ring round away from its partner is modelled faithfully (it passes or blocks on the wrong phase).  Besides
deadlock the model checks that every consumer sees the tile it expects, that the columns of a chunk are final
when a synthesis warp reads them (every step that writes them has been issued) and that an accumulator is not
overwritten before every synthesis warp has read all of its columns.

usage: python scripts/sim_protocol.py M NB NSA NSB LAG NG nbins [share=1] [runs=300]
"""
import random
import sys


class Bar:
    arrivals = 0          # any arrival counts as progress of the whole system (a role may arrive and block in one step)

    def __init__(self, count):
        self.count = count; self.pending = count; self.phase = 0

    def arrive(self):
        Bar.arrivals += 1
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


ISSUE3 = False     # False: warp 0 issues the whole stream of block b0; True: warp 2 takes the chunks it shares with b1


def step_owner(tgt, kc, kcB):
    return 1 if tgt else (2 if (ISSUE3 and kc >= kcB) else 0)


def main(M, NB, NSA, NSB, LAG, NG, nbins, share=True, seed=0, stall=0.9, quiet=True, units=2):
    KCH = 16
    NKC = (M + KCH - 1) // KCH
    NBLK = (M + NB - 1) // NB
    NPAIR = (NBLK + 1) // 2 if share else NBLK
    b0 = lambda p: 2 * p if share else p
    b1 = lambda p: (2 * p + 1) if (share and 2 * p + 1 < NBLK) else -1
    fc = lambda b: b * NB // KCH
    assert NG == 1 or 2 * NG <= NSA, "a ring slot could complete a further use while a group still waits for it"
    fullA = [Bar(1) for _ in range(NSA)]; emptyA = [Bar(3) for _ in range(NSA)]      # one arrival per owning group / every issue warp releases
    fullB = [[Bar(1) for _ in range(NSB)] for _ in range(3)]                          # one set of full barriers per issue warp
    emptyB = [Bar(1) for _ in range(NSB)]                                              # released by the step's issue warp
    tdone = [Bar(1) for _ in range(NSA)]; tempty = [Bar(NG), Bar(NG)]; tinit = [Bar(1), Bar(1)]
    inA = [None] * NSA; inB = [None] * NSB              # what the producers last put into a stage
    issued = {}                   # accumulator sequence number -> set of chunks multiplied into it
    readers = {}                  # accumulator sequence number -> groups that have finished reading it
    buf_owner = [None, None]      # accumulator currently held by a TMEM buffer

    def chunks_of_unit():
        for fi in range(nbins):
            for p in range(NPAIR):
                for kc in range(NKC - 1, fc(b0(p)) - 1, -1):
                    yield (fi, p, kc)

    def group(j):
        seq_base = 0; r_acc = 0
        for _u in range(units):
            todo = list(chunks_of_unit())
            total = len(todo)
            e = 0

            def pass_chunk():
                nonlocal e, r_acc
                fi, p, kc = todo[e]
                seq = seq_base + e; slot = seq % NSA; par = (seq // NSA) & 1
                blk = kc * KCH // NB
                if blk in (b0(p), b1(p)):
                    acc = r_acc + (0 if (blk == b1(p) or b1(p) < 0) else 1)
                    if e % NG == j:
                        while not tdone[slot].done(par): yield (f"G{j} wait tdone", slot, par, kc)
                        assert buf_owner[acc & 1] == acc, f"G{j} reads buffer of accumulator {buf_owner[acc & 1]}, wants {acc}"
                        need = set(range(kc, NKC))
                        assert need <= issued.get(acc, set()), f"G{j} reads columns of chunk {kc} of accumulator {acc} before {sorted(need - issued.get(acc, set()))}"
                    if kc == fc(blk):
                        if e % NG != j:       # not ours: still wait until the accumulator is complete before calling it read
                            while not tdone[slot].done(par): yield (f"G{j} wait tdone (lowest chunk)", slot, par, kc)
                        readers.setdefault(acc, set()).add(j)
                        tempty[acc & 1].arrive()
                if kc == fc(b0(p)):
                    r_acc += 2 if b1(p) >= 0 else 1
                e += 1

            i = j
            while i < total + NG:
                upto = min(i - NG, total - 1)
                while e <= upto:
                    yield from pass_chunk()
                if i < total:
                    seq = seq_base + i; slot = seq % NSA; par = (seq // NSA) & 1
                    while not emptyA[slot].done(par ^ 1): yield (f"G{j} wait emptyA", slot, par ^ 1)
                    inA[slot] = seq
                    fullA[slot].arrive()
                yield None
                i += NG
            while e < total:
                yield from pass_chunk()
            seq_base += total

    def loader():
        sb = phb = 0; q = 0
        mbw = [0, 0, 0]
        for _u in range(units):
            for fi in range(nbins):
                for p in range(NPAIR):
                    kcA = fc(b0(p)); kcB = fc(b1(p)) if b1(p) >= 0 else NKC
                    for (kc, tgt, fu, lu) in walk_pair(NKC, kcA, kcB, LAG):
                        owner = step_owner(tgt, kc, kcB)
                        while not emptyB[sb].done(phb ^ 1): yield ("loader wait emptyB", sb)
                        inB[sb] = q; q += 1
                        fullB[owner][mbw[owner]].arrive()
                        mbw[owner] = (mbw[owner] + 1) % NSB
                        sb += 1
                        if sb == NSB: sb = 0; phb ^= 1
                        yield None

    def issuer(my):
        sb = 0; mb = phm = 0; seqA = 0; acc_it = 0; q = 0
        tin = [0, 0]
        for _u in range(units):
            for fi in range(nbins):
                for p in range(NPAIR):
                    blk0, blk1 = b0(p), b1(p)
                    kcA = fc(blk0); kcB = fc(blk1) if blk1 >= 0 else NKC
                    acc1 = acc_it; acc0 = acc_it + (1 if blk1 >= 0 else 0); acc_it = acc0 + 1
                    seq0 = seqA; seqA += NKC - kcA
                    tin_par = 0
                    if ISSUE3 and blk1 >= 0:
                        tin_par = tin[acc0 & 1] & 1; tin[acc0 & 1] += 1
                    for (kc, tgt, fu, lu) in walk_pair(NKC, kcA, kcB, LAG):
                        i = NKC - 1 - kc
                        seq = seq0 + i; sa = seq % NSA; pha = (seq // NSA) & 1
                        acc = acc1 if tgt else acc0; buf = acc & 1; use_par = (acc >> 1) & 1
                        owner = step_owner(tgt, kc, kcB)
                        if owner != my:
                            uses_tile = (my == 1 or my == (2 if ISSUE3 else 0)) if kc >= kcB else (my == 0)
                            if fu and not uses_tile:                        # release of a tile this warp never multiplies
                                while not fullA[sa].done(pha): yield (f"W{my} wait fullA (release)", sa, pha, kc)
                                assert inA[sa] == seq, f"W{my} releases A tile {inA[sa]} instead of {seq}"
                                emptyA[sa].arrive()
                        else:
                            if i == 0:                                      # previous contents read by everybody?
                                while not tempty[buf].done(use_par ^ 1): yield (f"W{my} wait tempty", buf)
                                prev = buf_owner[buf]
                                assert prev is None or len(readers.get(prev, ())) == NG, f"accumulator {prev} overwritten before it was read"
                                buf_owner[buf] = acc
                            if ISSUE3 and my == 0 and kc == kcB - 1 and kcB < NKC:     # continue behind W2's last step
                                while not tinit[buf].done(tin_par): yield (f"W{my} wait tinit", buf)
                                assert buf_owner[buf] == acc and set(range(kcB, NKC)) <= issued.get(acc, set()), "W0 continues an unfinished accumulator"
                            while not fullB[my][mb].done(phm): yield (f"W{my} wait fullB", mb, phm, kc)
                            assert inB[sb] == q, f"W{my} sees B tile {inB[sb]} instead of {q}"
                            while not fullA[sa].done(pha): yield (f"W{my} wait fullA", sa, pha, kc)
                            assert inA[sa] == seq, f"W{my} sees A tile {inA[sa]} instead of {seq}"
                            issued.setdefault(acc, set()).add(kc)
                            cblk = kc * KCH // NB
                            own = cblk == (blk1 if tgt else blk0)
                            if own or (tgt == 0 and cblk != blk0 and cblk != blk1):
                                tdone[sa].arrive()
                            emptyA[sa].arrive(); emptyB[sb].arrive()
                            if ISSUE3 and my == 2 and kc == kcB: tinit[buf].arrive()
                            mb += 1
                            if mb == NSB: mb = 0; phm ^= 1
                        q += 1; sb += 1
                        if sb == NSB: sb = 0
                        yield None

    roles = {"loader": loader(), "W0": issuer(0), "W1": issuer(1), "W2": issuer(2)}
    for j in range(NG):
        roles[f"G{j}"] = group(j)
    rng = random.Random(seed)
    state = {}
    slow = {n: (rng.random() < 0.5) for n in roles}          # some roles are slow in this run
    while roles:
        progressed = stalled = False
        before = Bar.arrivals
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
        if not progressed and not stalled and Bar.arrivals == before:
            print("DEADLOCK:", {n: state.get(n) for n in roles}); return False
    if not quiet: print("ok")
    return True


if __name__ == "__main__":
    a = sys.argv
    M, NB, NSA, NSB, LAG, NG, nbins = (int(x) for x in a[1:8])
    share = len(a) < 9 or a[8] != "0"
    bad = 0
    for seed in range(int(a[9]) if len(a) > 9 else 300):
        try:
            ok = main(M, NB, NSA, NSB, LAG, NG, nbins, share, seed)
        except AssertionError as e:
            print("BUG (seed %d):" % seed, e); ok = False
        if not ok:
            bad += 1
            if bad >= 3: break
    print(f"M={M} NB={NB} NSA={NSA} NSB={NSB} LAG={LAG} NG={NG} share={share}: failures {bad}")
