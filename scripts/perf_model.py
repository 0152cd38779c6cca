"""Discrete-event timing model of one persistent CTA of the map kernel (das_map_tc.cu), to rank design options
without GPU time.  Roles and hand-shakes are those of scripts/sim_protocol.py; durations are calibrated against
the -DB2_MAP_TRACE timeline of the final round-1 kernel (FP16x3, M=256: 25.6 K cycles per bin):

  synthesis   one A tile per chunk: T_SYNTH_COMPUTE (can start before the ring slot is free) + T_SYNTH_STORE
  loader      B tile = latency + bytes / bandwidth, NSB loads in flight
  issue warp  per own step: T_STEP_PRE, barrier waits, N_MMA x (T_ISSUE each, blocks when more than QUEUE MMAs are
              pending), commits, T_STEP_POST;  per skipped step T_SKIP + waits
  tensor pipe executes MMAs in issue order, 128 * N / 256 cycles each (N = 2w columns)
  epilogue    T_EPI per 128-column accumulator

How far to trust it (model vs measurement on the final r1 kernel): cycles per bin 29.6 K vs 25.6 K; one issue
warp +3 % vs ~0; half the MMAs -19 % vs -21 %; all MMAs at full width +6 % vs +14 %; 10 A + 2 B stages +15 % vs
+17 %; lag 4 instead of 6 +9 % vs ~0.  It reproduces that the b0-stream issue warp is the pacing role and the
sensitivity to the number of tcgen05 instructions, but it ignores the competition of the roles for issue slots
(synthesis and epilogue durations are constants), so it is a first-order ranking aid, not a predictor.

usage: python scripts/perf_model.py            (prints the calibration run and a table of what-if variants)
"""
import heapq
from dataclasses import dataclass, replace


@dataclass
class Cfg:
    M: int = 256
    NB: int = 128
    NSA: int = 8
    NSB: int = 3
    LAG: int = 6
    share: bool = True
    issuers: int = 2            # 1: one warp issues everything; 2: warp 0 = b0 stream, warp 1 = b1 stream
    n_mma: int = 6              # tcgen05.mma per step (FP16x3: 3 products x 2 K-steps)
    mma_cycles_full: float = 128.0   # one MMA at N = 256
    t_synth_compute: float = 1050.0
    t_synth_store: float = 300.0
    t_issue: float = 80.0       # issue cost of one tcgen05 instruction
    queue: int = 2              # MMAs that may be pending before the issue thread blocks
    t_step_pre: float = 300.0
    t_step_post: float = 300.0
    t_skip: float = 150.0
    n_commit: int = 2
    t_epi: float = 5500.0       # per 128 columns
    b_latency: float = 1500.0
    b_bytes_per_cycle: float = 50.0
    full_width: bool = False    # what-if: no trimming along the diagonal
    kch: int = 16


def walk_pair(NKC, kcA, kcB, LAG):
    n1, n0 = NKC - kcB, NKC - kcA
    lag = min(n1, LAG)
    for i in range(n0 + lag):
        if i < n1:
            yield (NKC - 1 - i, 1, True, False)
        j = i - lag
        if j >= 0:
            yield (NKC - 1 - j, 0, j >= n1, True)


def simulate(c: Cfg, nbins=12):
    """Returns cycles per bin in steady state (second half of the run)."""
    KCH = c.kch
    NKC = (c.M + KCH - 1) // KCH
    NBLK = (c.M + c.NB - 1) // c.NB
    NPAIR = (NBLK + 1) // 2 if c.share else NBLK
    b0 = lambda p: 2 * p if c.share else p
    b1 = lambda p: (2 * p + 1) if (c.share and 2 * p + 1 < NBLK) else -1
    fc = lambda b: b * c.NB // KCH
    # the full step list (bin, pair, kc, tgt, first_use, w, acc index)
    steps = []
    acc_it = 0
    acc_blocks = []
    for fi in range(nbins):
        for p in range(NPAIR):
            kcA = fc(b0(p)); kcB = fc(b1(p)) if b1(p) >= 0 else NKC
            acc1 = acc_it; acc0 = acc_it + (1 if b1(p) >= 0 else 0); acc_it = acc0 + 1
            if b1(p) >= 0: acc_blocks.append(acc1)
            acc_blocks.append(acc0)
            for (kc, tgt, fu, lu) in walk_pair(NKC, kcA, kcB, c.LAG):
                kbase = kcB if tgt else kcA
                w = c.NB if c.full_width else min(c.NB, KCH * (kc - kbase + 1))
                steps.append(dict(bin=fi, kc=kc, tgt=tgt, fu=fu, w=w, acc=acc1 if tgt else acc0,
                                  first=(kc == NKC - 1), last=(kc == kbase), seq=None, pair=(fi, p)))
    # A tile sequence numbers: per (bin, pair) chunks NKC-1 .. kcA
    seq_of = {}
    seq = 0
    for fi in range(nbins):
        for p in range(NPAIR):
            for kc in range(NKC - 1, fc(b0(p)) - 1, -1):
                seq_of[(fi, p, kc)] = seq; seq += 1
    n_tiles = seq
    for s in steps:
        s["seq"] = seq_of[(s["pair"][0], s["pair"][1], s["kc"])]
    # uses per A tile (1 or 2) and release times
    uses = [0] * n_tiles
    for s in steps: uses[s["seq"]] += 1
    INF = float("inf")
    a_ready = [INF] * n_tiles          # fullA time
    a_released = [INF] * n_tiles       # emptyA time (all users done)
    a_done_cnt = [0] * n_tiles
    a_done_t = [0.0] * n_tiles
    b_ready = [INF] * len(steps)       # fullB time of step q's tile
    b_released = [INF] * len(steps)
    acc_full = {}                      # acc -> time tfull
    acc_empty = {}                     # acc -> time tempty (epilogue done)
    # ---- iterate to a fixed point in time order: simple sequential relaxation works because every role is
    # in-order; we advance the roles round-robin until all are done
    t_synth = 0.0; i_synth = 0
    t_load_issue = 0.0; i_load = 0
    pipe_free = 0.0
    mma_start_times = []               # start times of issued MMAs (for the queue limit)
    t_warp = [0.0, 0.0]; i_warp = [0, 0]
    t_epi = 0.0; i_epi = 0
    bin_done = {}
    progress = True
    while progress:
        progress = False
        # synthesis
        while i_synth < n_tiles:
            slot_prev = i_synth - c.NSA
            free_t = 0.0 if slot_prev < 0 else a_released[slot_prev]
            if free_t == INF: break
            start = t_synth
            done = max(start + c.t_synth_compute, free_t) + c.t_synth_store
            a_ready[i_synth] = done; t_synth = done; i_synth += 1; progress = True
        # loader
        while i_load < len(steps):
            prev = i_load - c.NSB
            free_t = 0.0 if prev < 0 else b_released[prev]
            if free_t == INF: break
            issue = max(t_load_issue, free_t)
            nbytes = 2 * steps[i_load]["w"] * 128
            b_ready[i_load] = issue + c.b_latency + nbytes / c.b_bytes_per_cycle
            t_load_issue = issue + 50.0; i_load += 1; progress = True
        # issue warps
        for wid in range(c.issuers):
            while i_warp[wid] < len(steps):
                q = i_warp[wid]; s = steps[q]
                mine = (c.issuers == 1) or (s["tgt"] == wid)
                t = t_warp[wid]
                if b_ready[q] == INF: break
                if mine:
                    if a_ready[s["seq"]] == INF: break
                    if s["first"]:
                        prev_use = s["acc"] - 2
                        if prev_use >= 0 and prev_use not in acc_empty: break
                        t = max(t, acc_empty.get(prev_use, 0.0))
                    t += c.t_step_pre
                    t = max(t, b_ready[q], a_ready[s["seq"]])
                    dur = c.mma_cycles_full * (2 * s["w"]) / 256.0
                    for k in range(c.n_mma):
                        if len(mma_start_times) >= c.queue:
                            t = max(t, mma_start_times[-c.queue])
                        t += c.t_issue
                        st = max(pipe_free, t)
                        mma_start_times.append(st)
                        pipe_free = st + dur
                    t += c.n_commit * c.t_issue
                    fin = pipe_free + 100.0
                    b_released[q] = fin if c.issuers == 1 else max(fin, b_released[q] if b_released[q] != INF else 0.0)
                    a_done_cnt[s["seq"]] += 1; a_done_t[s["seq"]] = max(a_done_t[s["seq"]], fin)
                    if s["last"]:
                        acc_full[s["acc"]] = fin
                    t += c.t_step_post
                else:
                    need_a = (wid == 1 and s["fu"] and s["tgt"] == 0)
                    if need_a and a_ready[s["seq"]] == INF: break
                    t += c.t_skip
                    t = max(t, b_ready[q])
                    if need_a:
                        t = max(t, a_ready[s["seq"]])
                        a_done_cnt[s["seq"]] += 1; a_done_t[s["seq"]] = max(a_done_t[s["seq"]], t)
                # releases need both warps when there are two
                need = uses[s["seq"]] if c.issuers == 1 else 2
                if a_done_cnt[s["seq"]] >= need and a_released[s["seq"]] == INF:
                    a_released[s["seq"]] = a_done_t[s["seq"]]
                if c.issuers == 2:
                    # the B tile is released when both warps have passed it
                    key = ("bpass", q)
                    cnt = acc_full.setdefault(key, 0) + 1
                    acc_full[key] = cnt
                    tt = max(acc_full.get(("bt", q), 0.0), (pipe_free + 100.0) if mine else t)
                    acc_full[("bt", q)] = tt
                    if cnt == 2: b_released[q] = tt
                t_warp[wid] = t; i_warp[wid] += 1; progress = True
        # epilogue
        while i_epi < len(acc_blocks):
            a = acc_blocks[i_epi]
            if a not in acc_full: break
            t_epi = max(t_epi, acc_full[a]) + c.t_epi
            acc_empty[a] = t_epi
            i_epi += 1; progress = True
            bin_done[i_epi // NBLK] = t_epi
    per_bin = NBLK
    n_done = len(acc_blocks) // per_bin
    t_half = acc_empty[acc_blocks[(n_done // 2) * per_bin - 1]]
    t_end = acc_empty[acc_blocks[n_done * per_bin - 1]]
    return (t_end - t_half) / (n_done - n_done // 2)


if __name__ == "__main__":
    base = Cfg()
    ref = simulate(base)
    print(f"calibration (final r1 kernel, FP16x3, M=256): {ref:8.0f} cycles per bin   (trace: 25 600)")
    variants = [
        ("one issue warp (measured: same time)", replace(base, issuers=1)),
        ("half the MMAs per step (measured -21 %)", replace(base, n_mma=3)),
        ("all MMAs at full width (measured +14 %)", replace(base, full_width=True)),
        ("lag 4 (measured: same)", replace(base, LAG=4)),
        ("10 A + 2 B stages, lag 8 (measured +17 %)", replace(base, NSA=10, NSB=2, LAG=8)),
        ("synthesis -24 % instructions (f32x2)", replace(base, t_synth_compute=1050 * 0.76, t_epi=5500 * 0.76)),
        ("32-microphone chunks", replace(base, kch=32, t_synth_compute=2100, t_synth_store=450, n_mma=12, NSA=4, NSB=2,
                                          LAG=3, b_latency=1500)),
        ("f32x2 + 32-microphone chunks", replace(base, kch=32, t_synth_compute=2100 * 0.76, t_synth_store=450, n_mma=12,
                                                   NSA=4, NSB=2, LAG=3, t_epi=5500 * 0.76)),
        ("tcgen05 issue at 40 cycles", replace(base, t_issue=40.0)),
        ("synthesis free (upper bound of synthesis work)", replace(base, t_synth_compute=1.0, t_synth_store=1.0)),
        ("epilogue free", replace(base, t_epi=1.0)),
    ]
    for name, cfg in variants:
        v = simulate(cfg)
        print(f"{name:55s} {v:8.0f} cycles per bin  {100.0 * (v / ref - 1):+6.1f} %")
