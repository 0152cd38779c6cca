"""CPU model of the map kernel's mbarrier protocol (roles as coroutines, randomly stalled) to find logical
deadlocks and phase-aliasing bugs before they cost GPU time.

Roles of das_map_tc_kernel: NG synthesis/reduction warp groups that take the chunks round-robin (group j
synthesises chunks j, j + NG, ...; before it synthesises chunk i it passes all chunks up to its own chunk
i - DLAG NG -- reducing that one after waiting for its `tdone`, signalling `tempty` when a block's lowest chunk
goes by), two MMA issue warps (W0: stream of block b0 = every chunk of the pair; W1: stream of block b1 = the chunks
from kcB up, plus the release of the tiles only b0 uses) and one B loader per stream with its own ring.
Barriers only know the PARITY of a phase, exactly like mbarrier.try_wait.parity, so a role that gets a whole
ring round away from its partner is modelled faithfully (it passes or blocks on the wrong phase).  Besides
deadlock the model checks that every consumer sees the tile it expects, that the columns of a chunk are final
when a synthesis warp reads them (every step that writes them has been issued) and that an accumulator is not
overwritten before every synthesis warp has read all of its columns.

usage: python scripts/sim_protocol.py M NB NSA NSB0 NSB1 NG nbins [share=1] [runs=300] [DLAG=1] [ND=128]
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


def main(M, NB, NSA, NSB0, NSB1, NG, nbins, share=True, seed=0, stall=0.9, quiet=True, units=2, DLAG=1, ND=128):
    KCH = 16
    NKC = (M + KCH - 1) // KCH
    NBLK = (M + NB - 1) // NB
    NPAIR = (NBLK + 1) // 2 if share else NBLK
    b0 = lambda p: 2 * p if share else p
    b1 = lambda p: (2 * p + 1) if (share and 2 * p + 1 < NBLK) else -1
    fc = lambda b: b * NB // KCH
    assert NSA >= NG and NSA >= DLAG * NG, "the A ring is too shallow for the chunks the groups keep in flight"
    assert ND >= 2 * NKC, "a tdone barrier could complete a further use while a group still waits for it"
    NSBs = (NSB0, NSB1)
    NBUF = 4 if NB <= 64 else 2                # TMEM accumulator buffers (512 columns / 2 NB, at most 4)
    fullA = [Bar(1) for _ in range(NSA)]; emptyA = [Bar(2) for _ in range(NSA)]      # one arrival per owning group / both issue warps release
    fullB = [[Bar(1) for _ in range(n)] for n in NSBs]; emptyB = [[Bar(1) for _ in range(n)] for n in NSBs]   # one ring per stream
    tdone = [Bar(1) for _ in range(ND)]; tempty = [Bar(NG) for _ in range(NBUF)]      # tdone: its own ring, one barrier per chunk in flight
    pfull = Bar(max(1, NG - 1)); pfree = Bar(1)               # hand-off of the bin sums (one quadrant's pair of barriers)
    handoff = {'hpar': [0] * NG}
    inA = [None] * NSA; inB = [[None] * n for n in NSBs]   # what the producers last put into a stage
    issued = {}                   # accumulator sequence number -> set of chunks multiplied into it
    readers = {}                  # accumulator sequence number -> groups that have finished reading it
    buf_owner = [None] * NBUF     # accumulator currently held by a TMEM buffer

    def chunks_of_unit():
        for fi in range(nbins):
            for p in range(NPAIR):
                for kc in range(NKC - 1, fc(b0(p)) - 1, -1):
                    yield (fi, p, kc)

    def group(j):
        # no walk over the other groups' chunks: per own chunk the group waits for its tdone, first releases the
        # accumulator uses it has no chunk in (after waiting for their lowest chunk), reads, and releases the
        # accumulator right away if this was its last chunk of the block
        seq_base = 0; accs = 0; arrived_acc = 0
        accoff = {}; lastpos = {}; n = 0; accbase = 0
        for p in range(NPAIR):
            for kc in range(NKC - 1, fc(b0(p)) - 1, -1):
                blk = kc * KCH // NB
                if blk in (b0(p), b1(p)):
                    off = accbase + (0 if (blk == b1(p) or b1(p) < 0) else 1)
                    accoff[n] = off
                    if kc == fc(blk): lastpos[off] = n
                else:
                    accoff[n] = accbase + 8      # above both blocks: the accumulators of the earlier pairs are complete (8 = not a read)
                n += 1
            accbase += 2 if b1(p) >= 0 else 1
        cpb = n
        pend = {'on': False, 'par': 0}
        for _u in range(units):
            todo = list(chunks_of_unit())
            total = len(todo)
            cur = {'bin': 0}

            def finish_bin(block):
                if NG > 1:
                    if block:
                        while not pfull.done(pend['par']): yield (f"G{j} wait pfull", pend['par'])
                    elif not pfull.done(pend['par']):
                        return
                    pfree.arrive()
                pend['on'] = False

            def flush_bins(upto):
                while cur['bin'] < upto:
                    hp = handoff['hpar'][j]
                    if NG > 1 and j > 0:
                        while not pfree.done(hp ^ 1): yield (f"G{j} wait pfree", hp ^ 1)
                        pfull.arrive()
                    else:
                        if pend['on']: yield from finish_bin(True)
                        pend['on'] = True; pend['par'] = hp
                        yield from finish_bin(NG == 1)
                    handoff['hpar'][j] ^= 1
                    cur['bin'] += 1

            def release_skipped(upto):
                nonlocal arrived_acc
                while arrived_acc < upto:
                    u = arrived_acc - accs
                    ls = seq_base + (u // NBLK) * cpb + lastpos[u % NBLK]
                    while not tdone[ls % ND].done((ls // ND) & 1): yield (f"G{j} wait tdone (lowest chunk of a skipped use)", ls)
                    readers.setdefault(arrived_acc, set()).add(j)
                    tempty[arrived_acc % NBUF].arrive()
                    arrived_acc += 1

            def reduce_own(idx):
                nonlocal arrived_acc
                fi, p, kc = todo[idx]
                cib = idx % cpb
                if NG > 1 and pend['on']: yield from finish_bin(False)
                yield from release_skipped(accs + fi * NBLK + accoff[cib] % 8)
                if accoff[cib] < 8:
                    blk = kc * KCH // NB
                    acc = accs + fi * NBLK + accoff[cib]
                    seq = seq_base + idx
                    while not tdone[seq % ND].done((seq // ND) & 1): yield (f"G{j} wait tdone", seq, kc)
                    yield from flush_bins(fi)
                    assert buf_owner[acc % NBUF] == acc, f"G{j} reads buffer of accumulator {buf_owner[acc % NBUF]}, wants {acc}"
                    need = set(range(kc, NKC))
                    assert need <= issued.get(acc, set()), f"G{j} reads columns of chunk {kc} of accumulator {acc} before {sorted(need - issued.get(acc, set()))}"
                    if kc - fc(blk) < NG:         # this group's last own chunk of the block
                        readers.setdefault(acc, set()).add(j)
                        tempty[acc % NBUF].arrive()
                        arrived_acc = acc + 1

            i = j
            while i < total + DLAG * NG:
                if i - DLAG * NG >= 0:
                    yield from reduce_own(i - DLAG * NG)
                if i < total:
                    seq = seq_base + i; slot = seq % NSA; par = (seq // NSA) & 1
                    while not emptyA[slot].done(par ^ 1): yield (f"G{j} wait emptyA", slot, par ^ 1)
                    inA[slot] = seq
                    fullA[slot].arrive()
                yield None
                i += NG
            yield from release_skipped(accs + nbins * NBLK)
            accs += nbins * NBLK
            yield from flush_bins(nbins)
            if pend['on']: yield from finish_bin(True)
            seq_base += total

    def loader(st):
        sb = phb = 0; q = 0
        for _u in range(units):
            for fi in range(nbins):
                for p in range(NPAIR):
                    blk = b1(p) if st else b0(p)
                    if blk < 0: continue
                    for kc in range(NKC - 1, fc(blk) - 1, -1):
                        while not emptyB[st][sb].done(phb ^ 1): yield (f"loader{st} wait emptyB", sb)
                        inB[st][sb] = q; q += 1
                        fullB[st][sb].arrive()
                        sb += 1
                        if sb == NSBs[st]: sb = 0; phb ^= 1
                        yield None

    def issuer(my):
        # up to TWO chunks per step: the warp waits for the B and A tiles of both before it issues either
        sb = phb = 0; seqA = 0; acc_it = 0; q = 0
        nsb_my = NSBs[my]
        for _u in range(units):
            for fi in range(nbins):
                for p in range(NPAIR):
                    blk0, blk1 = b0(p), b1(p)
                    kcA = fc(blk0); kcB = fc(blk1) if blk1 >= 0 else NKC
                    acc = acc_it if my else acc_it + (1 if blk1 >= 0 else 0)
                    acc_it += 2 if blk1 >= 0 else 1
                    buf = acc % NBUF; use_par = (acc // NBUF) & 1
                    myblk = blk1 if my else blk0
                    kbase = kcB if my else kcA
                    kc = NKC - 1
                    while kc >= kcA:
                        sa = seqA % NSA; pha = (seqA // NSA) & 1
                        if my == 1 and kc < kcB:                            # release of a tile only b0 uses
                            while not fullA[sa].done(pha): yield (f"W{my} wait fullA (release)", sa, pha, kc)
                            assert inA[sa] == seqA, f"W{my} releases A tile {inA[sa]} instead of {seqA}"
                            emptyA[sa].arrive()
                            kc -= 1; seqA += 1
                            yield None
                            continue
                        if kc == NKC - 1:                                   # previous contents read by everybody?
                            while not tempty[buf].done(use_par ^ 1): yield (f"W{my} wait tempty", buf)
                            prev = buf_owner[buf]
                            assert prev is None or len(readers.get(prev, ())) == NG, f"accumulator {prev} overwritten before it was read"
                            buf_owner[buf] = acc
                        two = kc - 1 >= kbase and nsb_my >= 3
                        steps = []
                        s_b, p_b = sb, phb
                        for d in range(2 if two else 1):
                            sq = seqA + d
                            steps.append((kc - d, sq, sq % NSA, (sq // NSA) & 1, s_b, p_b, q + d))
                            s_b += 1
                            if s_b == nsb_my: s_b = 0; p_b ^= 1
                        for (k_, sq, sa_, pha_, sb_, phb_, q_) in steps:   # all waits first
                            while not fullB[my][sb_].done(phb_): yield (f"W{my} wait fullB", sb_, phb_, k_)
                            assert inB[my][sb_] == q_, f"W{my} sees B tile {inB[my][sb_]} instead of {q_}"
                            while not fullA[sa_].done(pha_): yield (f"W{my} wait fullA", sa_, pha_, k_)
                            assert inA[sa_] == sq, f"W{my} sees A tile {inA[sa_]} instead of {sq}"
                        for (k_, sq, sa_, pha_, sb_, phb_, q_) in steps:   # then issue + commits, chunk by chunk
                            issued.setdefault(acc, set()).add(k_)
                            cblk = k_ * KCH // NB
                            if cblk == myblk or (my == 0 and cblk != blk0 and cblk != blk1):
                                tdone[sq % ND].arrive()
                            emptyA[sa_].arrive(); emptyB[my][sb_].arrive()
                        n = len(steps)
                        kc -= n; seqA += n; q += n; sb, phb = s_b, p_b
                        yield None

    roles = {"loader0": loader(0), "W0": issuer(0), "W1": issuer(1)}
    if NSB1 > 0:
        roles["loader1"] = loader(1)
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
    M, NB, NSA, NSB0, NSB1, NG, nbins = (int(x) for x in a[1:8])
    share = len(a) < 9 or a[8] != "0"
    dlag = int(a[10]) if len(a) > 10 else 1
    nd = int(a[11]) if len(a) > 11 else 128
    bad = 0
    for seed in range(int(a[9]) if len(a) > 9 else 300):
        try:
            ok = main(M, NB, NSA, NSB0, NSB1, NG, nbins, share, seed, DLAG=dlag, ND=nd)
        except AssertionError as e:
            print("BUG (seed %d):" % seed, e); ok = False
        if not ok:
            bad += 1
            if bad >= 3: break
    print(f"M={M} NB={NB} NSA={NSA} NSB={NSB0}+{NSB1} NG={NG} share={share}: failures {bad}")
