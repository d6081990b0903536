"""Capacity-aware re-levelling of a compiled Program.

`program.lower` places every lookup at the earliest level its inputs allow (as soon as possible).  The critical path of
the reference's inversions is long and thin, while a few levels hold hundreds of lookups that are
not needed for many levels to come.  A bootstrap launch costs the same from one ciphertext up to the number of
ciphertexts the low-latency kernel keeps resident (33 on one B200 at N = 2048), so lookups with slack are moved into
later, emptier levels: same lookups, same number of levels, the same results bit for bit (every lookup still reads the
same values), fewer launches beyond the resident capacity.

The pass works on the Program itself (levels, CSR rows over value slots), so it applies to programs loaded from .npz
and can be redone for another capacity (another GPU, or `world` GPUs sharing each level).
"""
from __future__ import annotations

import heapq

import numpy as np

from .program import Level, Program


def _dag(prog: Program):
    """reconstruct values and lookups from the levelled form.
    value ids: 0..n_inputs-1 are the inputs, then one per lookup in level order.
    Returns (groups, out_rows): groups = list of dicts {key=(terms, konst, full), deps, jobs=[(value id, lut)]}, one per
    keyswitch row (lookups that share their input ride together), out_rows = [(terms over value ids, konst)]"""
    cur = {int(s): i for i, s in enumerate(prog.input_slots)}
    groups, next_id = [], prog.n_inputs
    for lv in prog.levels:
        rows = []
        for r in range(len(lv.konst)):
            a, b = int(lv.row_ptr[r]), int(lv.row_ptr[r + 1])
            terms = tuple((cur[int(s)], int(c)) for s, c in zip(lv.idx[a:b], lv.coef[a:b]))
            rows.append({"terms": terms, "konst": int(lv.konst[r]), "full": bool(lv.full[r]),
                         "deps": sorted({v for v, _ in terms}), "jobs": []})
        writes = []
        for j in range(len(lv.job_ks)):
            rows[int(lv.job_ks[j])]["jobs"].append((next_id, int(lv.job_lut[j])))
            writes.append((int(lv.job_out[j]), next_id))
            next_id += 1
        for s, v in writes:
            cur[s] = v
        groups.extend(g for g in rows if g["jobs"])
    out_rows = []
    for r in range(len(prog.out_konst)):
        a, b = int(prog.out_row_ptr[r]), int(prog.out_row_ptr[r + 1])
        out_rows.append((tuple((cur[int(s)], int(c)) for s, c in zip(prog.out_idx[a:b], prog.out_coef[a:b])), int(prog.out_konst[r])))
    return groups, out_rows, next_id


def default_tiers(capacity: int):
    """launch sizes up to which a level costs about the same.  The steps of the measured cost curve (B200, N = 2048:
    33 ciphertexts on the 8-CTA kernel, 74 / 148 / 222 in one / two / three waves of the latency build, 296 resident on
    the throughput build, profiles/r2_pbs_sweep_pairs_w4.jsonl), expressed through the engine's launch capacity"""
    wave = capacity * 296 // 33
    return sorted({capacity, capacity * 74 // 33, capacity * 148 // 33, capacity * 222 // 33, wave}
                  | {wave * m for m in (2, 3, 4, 6, 8, 16, 32, 64, 1 << 12)})


def modelled_ms(prog: Program, capacity: int, world: int = 1) -> float:
    """rough wall time of a run, for choosing between level layouts: per level, the bootstrap launch of the busiest rank
    plus the fixed keyswitch / exchange cost.  Constants measured on B200 at N = 2048 with the pair rotation
    (profiles/r2_pbs_sweep_pairs_w4.jsonl), expressed through the engine's launch capacity so that they carry over:
    up to a quarter of the capacity 2.08 ms, up to about half 2.38, up to the capacity 2.48, then the wider builds."""
    cap = max(capacity // max(world, 1), 1)
    total = 0.0
    for lv in prog.levels:
        c = -(-len(lv.job_ks) // world)
        if c <= max(cap // 4, 1):
            t = 2.08
        elif c <= max((cap * 18) // 33, 1):
            t = 2.38
        elif c <= cap:
            t = 2.48
        elif c <= (cap * 74) // 33:
            t = 4.32
        elif c <= 3 * cap:
            t = 7.3
        elif c <= (cap * 148) // 33:
            t = 8.6
        else:
            t = max(9.4, 0.0496 * c)
        total += t + 0.12
    return total


def schedule_for(prog: Program, capacity: int, world: int = 1) -> Program:
    """the cheaper (modelled) of two layouts: levels filled up to the launch capacity, or first up to a quarter of it --
    the size a launch still runs at its very lowest latency -- which pays once several GPUs share each level"""
    if capacity <= 0:
        return rebalance(prog, 1 << 30)
    full = rebalance(prog, capacity)
    quarter = (capacity // max(world, 1)) // 4 * world          # every rank still within a quarter of its own capacity
    if world < 2 or quarter < 1:
        return full
    fine = rebalance(prog, capacity, [quarter] + default_tiers(capacity))
    return fine if modelled_ms(fine, capacity, world) < modelled_ms(full, capacity, world) else full


def rebalance(prog: Program, capacity: int = 36, tiers=None) -> Program:
    """move lookups with slack out of levels that exceed a launch-size tier into later levels with room.
    capacity: ciphertexts a bootstrap launch handles at its minimum latency (per GPU times the GPUs sharing a level)."""
    tiers = sorted(tiers or default_tiers(capacity))
    groups, out_rows, n_values = _dag(prog)
    n_in = prog.n_inputs
    producer = {}                                   # value id -> group index
    for gi, g in enumerate(groups):
        for v, _ in g["jobs"]:
            producer[v] = gi
    # as-soon-as-possible level of each group (inputs are level 0) and as-late-as-possible level for the same depth
    asap = np.zeros(len(groups), np.int64)
    for gi, g in enumerate(groups):                 # groups come in level order: producers precede consumers
        asap[gi] = 1 + max((asap[producer[v]] for v in g["deps"] if v >= n_in), default=0)
    depth = int(asap.max(initial=0))
    alap = np.full(len(groups), depth, np.int64)
    consumers = [[] for _ in groups]
    for gi, g in enumerate(groups):
        for v in g["deps"]:
            if v >= n_in:
                consumers[producer[v]].append(gi)
    for gi in range(len(groups) - 1, -1, -1):
        for c in consumers[gi]:
            alap[gi] = min(alap[gi], alap[c] - 1)
    # list scheduling, least slack first; a level takes everything that can wait no longer, then fills up to the
    # smallest tier that holds those
    size = np.array([len(g["jobs"]) for g in groups], np.int64)
    # a group becomes ready once every GROUP that produces one of its inputs has been scheduled
    pending = np.array([len({producer[v] for v in g["deps"] if v >= n_in}) for g in groups], np.int64)
    cons_groups = [sorted(set(c)) for c in consumers]
    ready = [(int(alap[gi]), gi) for gi in range(len(groups)) if pending[gi] == 0]
    heapq.heapify(ready)
    level_of = np.zeros(len(groups), np.int64)
    t = 0
    scheduled = 0
    while scheduled < len(groups):
        t += 1
        forced, rest = [], []
        while ready and ready[0][0] <= t:
            forced.append(heapq.heappop(ready)[1])
        count = int(size[forced].sum()) if forced else 0
        room = next(tr for tr in tiers if tr >= max(count, 1)) - count
        deferred = []
        while ready and room > 0:
            a, gi = heapq.heappop(ready)
            if size[gi] <= room:
                rest.append(gi)
                room -= int(size[gi])
            else:
                deferred.append((a, gi))
        for item in deferred:
            heapq.heappush(ready, item)
        newly = []
        for gi in forced + rest:
            level_of[gi] = t
            scheduled += 1
            for c in cons_groups[gi]:
                pending[c] -= 1
                if pending[c] == 0:
                    newly.append(c)
        for c in newly:                              # consumers become ready for the NEXT level
            heapq.heappush(ready, (int(alap[c]), c))
    n_levels = t
    return _emit(prog, groups, out_rows, level_of, n_levels)


def _emit(prog: Program, groups, out_rows, level_of, n_levels) -> Program:
    """levels + liveness-based slot allocation (as program.lower does) for a given level assignment"""
    n_in = prog.n_inputs
    by_level = [[] for _ in range(n_levels)]
    for gi, lv in enumerate(level_of):
        by_level[int(lv) - 1].append(gi)
    last_use = {v: 0 for v in range(n_in)}
    for gi, g in enumerate(groups):
        for v in g["deps"]:
            last_use[v] = max(last_use.get(v, 0), int(level_of[gi]))
    for terms, _k in out_rows:
        for v, _c in terms:
            last_use[v] = n_levels + 1
    slot_of, free, n_slots = {}, [], 0
    for v in range(n_in):
        slot_of[v] = int(prog.input_slots[v])
        n_slots = max(n_slots, slot_of[v] + 1)
    expiring = [[] for _ in range(n_levels + 2)]
    for v in range(n_in):
        expiring[last_use[v]].append(v)
    levels = []
    for li, gis in enumerate(by_level, start=1):
        row_ptr = np.zeros(len(gis) + 1, np.int32)
        idx, coef, konst, full, job_ks, job_lut, job_out = [], [], [], [], [], [], []
        for r, gi in enumerate(gis):
            g = groups[gi]
            for v, c in g["terms"]:
                idx.append(slot_of[v])
                coef.append(c)
            row_ptr[r + 1] = len(idx)
            konst.append(g["konst"])
            full.append(g["full"])
        for r, gi in enumerate(gis):                 # outputs are allocated after every read of the level was resolved
            for v, lut in groups[gi]["jobs"]:
                if free:
                    slot_of[v] = free.pop()
                else:
                    slot_of[v] = n_slots
                    n_slots += 1
                expiring[max(last_use.get(v, li), li)].append(v)
                job_ks.append(r)
                job_lut.append(lut)
                job_out.append(slot_of[v])
        levels.append(Level(row_ptr, np.asarray(idx, np.int32), np.asarray(coef, np.int64), np.asarray(konst, np.int64),
                            np.asarray(job_ks, np.int32), np.asarray(job_lut, np.int32), np.asarray(job_out, np.int32),
                            np.asarray(full, bool)))
        for v in expiring[li]:
            free.append(slot_of[v])
    out_ptr = np.zeros(len(out_rows) + 1, np.int32)
    oidx, ocoef, okonst = [], [], []
    for r, (terms, k) in enumerate(out_rows):
        for v, c in terms:
            oidx.append(slot_of[v])
            ocoef.append(c)
        out_ptr[r + 1] = len(oidx)
        okonst.append(k)
    new = Program(prog.width, n_in, n_slots, prog.input_slots.copy(), levels, out_ptr, np.asarray(oidx, np.int32),
                  np.asarray(ocoef, np.int64), np.asarray(okonst, np.int64), prog.out_shape, prog.tables, prog.nu2,
                  table_half=prog.table_half)
    new.stats = dict(prog.stats, levels=n_levels, slots=n_slots, pbs=new.n_pbs, keyswitches=new.n_ks,
                     max_level_pbs=max((len(l.job_ks) for l in levels), default=0), rebalanced=True)
    return new
