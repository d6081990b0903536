"""Lowering of a trace into a levelled TFHE program, and its clear-text evaluator.

A Program is what the engine executes (one batched launch group per level):

  level L:  lincomb rows (CSR over value slots)  ->  keyswitch  ->  PBS jobs (ks row, LUT, out slot)
  outputs:  lincomb rows over value slots

All ciphertexts of a program share one message width W (plaintext scale 2^(63-W)); a lookup
whose input ranges over [lo, hi] on the inputset reads the table at (value + offset) with the
range centred in [0, 2^W).  W is the smallest width that fits every lookup input range, the
quantity Concrete's bit-width assignment derives from the same inputset
(/root/reference/matrix_inversion/qfloat_matrix_inversion.py:981-1004).

Wide-lookup splitting.  The polynomial size (hence the cost of EVERY bootstrap) follows W, yet in the
reference's circuits a handful of lookups are one bit wider than all the others (22 of 29 k in the 3x3
inversion).  Such a lookup F on y in [0, 2^(W+1)) is lowered to three W-bit bootstraps that use the
padding bit and the negacyclic wrap T(y + 2^W) = -T(y) of the accumulator:
    s  = PBS_const(y)            = +2^(W-1) for y < 2^W, -2^(W-1) above           (the sign, for free)
    y' = y - 2^(W-1) + s         = y mod 2^W                                       (leveled)
    F(y) = PBS_A(y) + PBS_G(y'),   A(t) = (F(t) - F(t + 2^W)) / 2,  G(t) = (F(t) + F(t + 2^W)) / 2
A and G may be half-integers; accumulator entries are arbitrary torus elements, so those tables are stored
doubled and encoded at half the plaintext scale (`table_half`); their sum is always an integer message.
Keyswitch rows that legitimately use the padding bit are flagged `Level.full`.
"""
from __future__ import annotations

import hashlib
import os
from dataclasses import dataclass, field

import numpy as np

from .. import params as PR


@dataclass
class Level:
    row_ptr: np.ndarray      # [n_ks + 1] int32
    idx: np.ndarray          # [nnz] int32 value slots
    coef: np.ndarray         # [nnz] int64 signed coefficients
    konst: np.ndarray        # [n_ks] int64 signed constants (message units, offset included)
    job_ks: np.ndarray       # [n_pbs] int32 row of this level's keyswitch batch
    job_lut: np.ndarray      # [n_pbs] int32
    job_out: np.ndarray      # [n_pbs] int32 value slot
    full: np.ndarray = None  # [n_ks] bool: row ranges over the whole torus (padding bit used on purpose)

    def __post_init__(self):
        if self.full is None:
            self.full = np.zeros(len(self.konst), bool)


@dataclass
class Program:
    width: int
    n_inputs: int
    n_slots: int
    input_slots: np.ndarray
    levels: list
    out_row_ptr: np.ndarray
    out_idx: np.ndarray
    out_coef: np.ndarray
    out_konst: np.ndarray
    out_shape: tuple
    tables: np.ndarray       # [n_luts][2^W] int64 table outputs (message units)
    nu2: int                 # largest squared 2-norm of a lookup input's linear combination
    stats: dict = field(default_factory=dict)
    table_half: np.ndarray = None   # [n_luts] bool: table holds 2x its values (half-integer outputs)
    debug: dict = None

    def __post_init__(self):
        if self.table_half is None:
            self.table_half = np.zeros(len(self.tables), bool)

    @property
    def n_pbs(self):
        return int(sum(len(l.job_ks) for l in self.levels))

    @property
    def n_ks(self):
        return int(sum(len(l.konst) for l in self.levels))

    def lut_polynomials(self, N: int) -> np.ndarray:
        W = self.width
        return np.stack([PR.lut_polynomial([PR.encode(int(t), W + int(h)) for t in tab], W, N)
                         for tab, h in zip(self.tables, self.table_half)])

    def tables_half_units(self) -> np.ndarray:
        """every table in units of half a message (what evaluate_clear and clear-text stand-in engines compute in)"""
        return self.tables * np.where(self.table_half, 1, 2)[:, None]

    # ------------------------------------------------------------ (de)serialisation
    def save(self, path, **extra):
        """compiled programs travel as .npz (the compile step needs the circuit's Python source; running does not)"""
        lv = self.levels

        def cat(xs, dt):
            """concatenate and narrow to the on-disk type; a value that does not fit is an error, never a wrap-around"""
            if not xs:
                return np.zeros(0, dt)
            wide = np.concatenate([np.asarray(x, np.int64) for x in xs])
            if dt is not bool and wide.size:
                info = np.iinfo(dt)
                if wide.min() < info.min or wide.max() > info.max:
                    raise OverflowError(f"program field out of range for {np.dtype(dt).name}: [{wide.min()}, {wide.max()}]")
            return wide.astype(dt)

        np.savez_compressed(
            path, width=self.width, n_inputs=self.n_inputs, n_slots=self.n_slots, input_slots=self.input_slots,
            ks_counts=np.array([len(l.konst) for l in lv], np.int64), nz_counts=np.array([len(l.idx) for l in lv], np.int64),
            pbs_counts=np.array([len(l.job_ks) for l in lv], np.int64),
            row_ptr=cat([l.row_ptr for l in lv], np.int32), idx=cat([l.idx for l in lv], np.int32),
            coef=cat([l.coef for l in lv], np.int32), konst=cat([l.konst for l in lv], np.int32),
            job_ks=cat([l.job_ks for l in lv], np.int32), job_lut=cat([l.job_lut for l in lv], np.int16),
            job_out=cat([l.job_out for l in lv], np.int32), out_row_ptr=self.out_row_ptr, out_idx=self.out_idx,
            out_coef=self.out_coef, out_konst=self.out_konst, out_shape=np.array(self.out_shape, np.int64),
            tables=self.tables.astype(np.int16 if self.width < 14 else np.int64), nu2=self.nu2,
            table_half=self.table_half, ks_full=cat([l.full for l in lv], bool),
            stats=np.array(repr(self.stats)), **extra)

    @classmethod
    def load(cls, path):
        with np.load(path, allow_pickle=False) as npz:
            z = {k: npz[k] for k in npz.files}            # NpzFile decompresses on every access: read each array once
        ks, nz, pb = z["ks_counts"], z["nz_counts"], z["pbs_counts"]
        ko, no, po = np.cumsum(np.r_[0, ks]), np.cumsum(np.r_[0, nz]), np.cumsum(np.r_[0, pb])
        ro = np.cumsum(np.r_[0, ks + 1])
        row_ptr, idx, coef, konst = (z["row_ptr"].astype(np.int32), z["idx"].astype(np.int32), z["coef"].astype(np.int64),
                                     z["konst"].astype(np.int64))
        job_ks, job_lut, job_out = z["job_ks"].astype(np.int32), z["job_lut"].astype(np.int32), z["job_out"].astype(np.int32)
        full = z["ks_full"].astype(bool) if "ks_full" in z else np.zeros(len(konst), bool)
        levels = []
        for i in range(len(ks)):
            levels.append(Level(row_ptr[ro[i]:ro[i + 1]], idx[no[i]:no[i + 1]], coef[no[i]:no[i + 1]], konst[ko[i]:ko[i + 1]],
                                job_ks[po[i]:po[i + 1]], job_lut[po[i]:po[i + 1]], job_out[po[i]:po[i + 1]],
                                full[ko[i]:ko[i + 1]]))
        import ast
        prog = cls(int(z["width"]), int(z["n_inputs"]), int(z["n_slots"]), z["input_slots"].astype(np.int32), levels,
                   z["out_row_ptr"].astype(np.int32), z["out_idx"].astype(np.int32), z["out_coef"].astype(np.int64),
                   z["out_konst"].astype(np.int64), tuple(int(v) for v in z["out_shape"]), z["tables"].astype(np.int64),
                   int(z["nu2"]), table_half=z["table_half"].astype(bool) if "table_half" in z else None)
        prog.stats = ast.literal_eval(str(z["stats"]))
        return prog

    # ------------------------------------------------------------ clear evaluation
    def evaluate_clear(self, inputs: np.ndarray, strict: bool = True):
        """run the levelled program on clear integers (what the ciphertexts hold); inputs [n_inputs] or
        [batch][n_inputs] -> outputs shaped out_shape (with a leading batch axis if given).
        A lookup input outside the W-bit message space means the encrypted run would read a wrong table entry
        (the value left the ranges seen on the compile-time inputset): strict raises OverflowError, otherwise the
        result is (outputs, per-lane bool mask of such lanes) with the offending inputs wrapped like the ciphertext would."""
        x = np.asarray(inputs, dtype=np.int64)
        single = x.ndim == 1
        x = np.atleast_2d(x)
        B, W = x.shape[0], self.width
        size = 1 << W
        tab2 = self.tables_half_units()
        vals = np.zeros((self.n_slots, B), np.int64)          # half-message units throughout
        vals[self.input_slots] = 2 * x.T
        bad = np.zeros(B, bool)
        for lv in self.levels:
            ks2 = _csr_apply(lv.row_ptr, lv.idx, lv.coef, 2 * lv.konst, vals)
            assert not (ks2 & 1).any(), "half-integer lookup input: split tables combined with unequal coefficients"
            ks = ks2 >> 1
            # what the bootstrap sees: the message mod 2^(W+1) (table outputs are stored reduced, so only the residue
            # is meaningful); the upper half is the padding bit -- out of range unless the row uses it on purpose --
            # and reads the negated table (negacyclic)
            m = np.mod(ks, 2 * size)
            neg = m >= size
            oob = neg & ~lv.full[:, None]
            if oob.any():
                if strict:
                    raise OverflowError(f"lookup input outside the {W}-bit message space: {np.unique(m[oob] - 2 * size)[:8]}")
                bad |= oob.any(axis=0)
            t = tab2[lv.job_lut[:, None], np.where(neg, m - size, m)[lv.job_ks]]
            vals[lv.job_out] = np.where(neg[lv.job_ks], -t, t)
        out2 = _csr_apply(self.out_row_ptr, self.out_idx, self.out_coef, 2 * self.out_konst, vals)
        assert not (out2 & 1).any()
        out = (np.mod((out2 >> 1) + size, 2 * size) - size).T.reshape((B,) + tuple(self.out_shape))
        if strict:
            return out[0] if single else out
        return (out[0], bad[0]) if single else (out, bad)


class _Pseudo:
    """stand-in for a traced lookup [const + terms < 0] derived from another one"""
    __slots__ = ("base", "const")

    def __init__(self, base, const):
        self.base, self.const = base, const


_PROBE = np.arange(-70, 71, dtype=np.int64)


def _lt0_scale(fn) -> int:
    """c if the lookup function is x -> c [x < 0] with c != 0, else 0 (decided on values: the tracer keeps functions
    opaque, and it folds a constant factor such as the `p *` of `temp + p * borrow` into the table)"""
    try:
        v = np.asarray(fn(_PROBE))
    except Exception:
        return 0
    if v.shape != _PROBE.shape:
        return 0
    v = v.astype(np.int64)
    c = int(v[0])
    return c if c and np.array_equal(v, c * (_PROBE < 0)) else 0


def _narrowing_pays(wide: "Program", narrow: "Program") -> bool:
    """one bit of width doubles the polynomial size: bootstrap work grows 2 (W + 7) / (W + 6)-fold (transform
    butterflies), the latency of one level about 1.6-fold (measured)"""
    W = wide.width
    work = lambda p, w: p.n_pbs * (1 << w) * (w + 7)
    path = lambda p, w: len(p.levels) * 1.6 ** w
    return work(narrow, W - 1) < work(wide, W) and path(narrow, W - 1) < 1.1 * path(wide, W)


def _csr_apply(row_ptr, idx, coef, konst, vals):
    n = len(konst)
    out = np.repeat(konst[:, None], vals.shape[1], axis=1).astype(np.int64)
    if len(idx):
        rows = np.repeat(np.arange(n), np.diff(row_ptr))
        np.add.at(out, rows, coef[:, None] * vals[idx])
    return out


# ------------------------------------------------------------------ lowering
def lower(trace, outputs, out_shape, slack_bits: int = 0, min_width: int = 1, split_wide="auto", split_guard=None,
          collapse_borrows: bool = True) -> Program:
    """trace: fhe.tracing.Trace after the circuit function ran; outputs: flat list of scalars
    (ints / Aff) the function returned.
    split_wide: True lowers the lookups that alone need the top bit of width into three narrower bootstraps (module
    docstring); False never does; "auto" does when the narrower width halves the polynomial size (W >= 5), the wide
    lookups are a minority, and the narrowed program is cheaper both in total bootstrap work and along its critical
    path (`_narrowing_pays`).  split_guard: spare table entries required either side of a lookup's observed range in a narrowed
    program, lookups with less are split as well (default 2^W / 4).
    collapse_borrows: a borrow chain  b' = [A - b < 0],  b = [S < 0]  (the reference's digit-by-digit subtraction,
    base_p_arrays.py:119-121, 70 % of an inversion's critical path) is rewritten  b' = [M A + S < 0]  with
    M = max(-min S, max S + 1): the same bit for every integer A as long as S stays inside the bounds M was derived
    from, and no longer dependent on b -- so several digits resolve per level (True: as many as fit the message space,
    three base-2 digits in 4 bits; 2: two, with one step of slack on S where it fits; False: none).  The bounds are what
    borrow sources and digit differences span ANYWHERE in the program on the inputset (the recurrence is the same at
    every digit position; a single position may have shown less), and the rewrite is only done where M A + S fits."""
    if split_wide == "auto":
        wide = lower(trace, outputs, out_shape, slack_bits, min_width, False, None, collapse_borrows)
        if wide.width - slack_bits < 5 or wide.stats["top_width_lookups"] * 10 > wide.stats["live_lookups"]:
            return wide
        narrow = lower(trace, outputs, out_shape, slack_bits, min_width, True, split_guard, collapse_borrows)
        return narrow if narrow.width < wide.width and _narrowing_pays(wide, narrow) else wide
    n_in = trace.n_inputs
    jobs = trace.jobs

    # --- dead code elimination: only lookups the outputs depend on
    live = np.zeros(len(jobs), bool)
    stack = []
    for o in outputs:
        if not isinstance(o, int):
            stack.extend(b for b in o.terms if b >= n_in)
    while stack:
        b = stack.pop()
        j = b - n_in
        if live[j]:
            continue
        live[j] = True
        stack.extend(t for t in jobs[j].terms if t >= n_in)

    # --- message width: every live lookup's input range must fit (or be split), every output value too
    bits = {int(j): int(jobs[j].group.hi - jobs[j].group.lo).bit_length() for j in np.flatnonzero(live)}
    W_out = min_width
    for o in outputs:
        if not isinstance(o, int):
            W_out = max(W_out, int(max(abs(int(o.vals.min())), abs(int(o.vals.max())))).bit_length())
    W = max([W_out] + list(bits.values()))
    n_top = sum(1 for v in bits.values() if v == W)
    narrowed = False
    if split_wide and W > max(W_out, 1) and n_top:
        W -= 1
        narrowed = True
    W += slack_bits
    size = 1 << W
    dom = np.arange(size, dtype=np.int64)
    # a narrowed program keeps spare table entries either side of every observed range (else the lookup is split too):
    # unseen inputs stray a little outside the inputset's ranges, which the wider program absorbed for free
    guard = (size // 4 if narrowed else 0) if split_guard is None else int(split_guard)

    # --- tables, common-subexpression elimination of lookups and of keyswitches
    table_ids, tables, table_half = {}, [], []
    lookup_ids = {}            # (src key, table id) -> representative base
    subst = {}                 # traced base -> {representative base: coefficient} where they differ
    job_info = {}              # representative base -> (src key, table id)
    full_keys = set()          # keyswitch rows that use the padding bit on purpose
    state = {"nu2": 1, "next": n_in + len(jobs), "split": 0, "collapsed": 0, "bitwise_folded": 0}
    bit_inputs = {b for b, (lo, hi) in enumerate(getattr(trace, "input_ranges", [])[:n_in]) if lo >= 0 and hi <= 1}
    split_of = {}              # helper lookup of a split -> the traced lookup it serves (debug only)
    lt0_src = {}               # representative base of a lookup [S < 0] -> (terms of S, constant of S, observed min, max)
    # A borrow chain is the same digit recurrence at every position, but a given position may have shown only part of
    # its range on the inputset (a leading digit that happened to be 0 in all samples).  Every borrow source is
    # therefore assumed to span what borrow sources span ANYWHERE in the program (for base-2 digits: [-2, 1]).
    scales = {int(j): _lt0_scale(jobs[j].fn) for j in np.flatnonzero(live)} if collapse_borrows else {}
    lt0_jobs = [jobs[j] for j, c in scales.items() if c]
    span_lo = min((jb.group.lo for jb in lt0_jobs), default=0)
    span_hi = max((jb.group.hi for jb in lt0_jobs), default=0)
    # ... and every digit difference A what digit differences span anywhere (A = source + borrow, per sample)
    a_span = [0, 0]
    for jb in lt0_jobs:
        for b, c in jb.terms.items():
            if c == -1 and b >= n_in and scales.get(b - n_in) == 1:
                a = jb.src_vals.astype(np.int64) + jobs[b - n_in].out_vals
                a_span = [min(a_span[0], int(a.min())), max(a_span[1], int(a.max()))]
    # True: as many digits per level as the message space holds with the spans taken exactly (three base-2 digits in
    # 4 bits); 2: two digits, and the previous source may leave its span by one where that still fits
    margins = (1, 0) if collapse_borrows == 2 and collapse_borrows is not True else (0,)
    # "prefix": all borrows of a chain from block signs combined lexicographically, depth ~ 2 log3(digits)
    prefix_ok = (collapse_borrows == "prefix" and size >= 16 and a_span[0] >= -1 and a_span[1] <= 1
                 and span_lo >= -2 and span_hi <= 1)
    chain = {}                 # representative base of a chain borrow -> its leaves [(terms, const)], lowest digit first
    off16 = (size - 16) // 2 + 8 if size >= 16 else 0

    def resolve(src):
        terms = {}
        for b, c in src.items():
            for rb, rc in (subst[b].items() if b in subst else ((b, 1),)):
                v = terms.get(rb, 0) + c * rc
                if v:
                    terms[rb] = v
                else:
                    terms.pop(rb, None)
        return terms

    def add_lookup(base, key, tab, half=False):
        """registers lookup `tab` of keyswitch row `key`; returns the base that holds its result"""
        tab = np.asarray(tab, dtype=np.int64)
        if tab.shape != (size,):
            tab = np.broadcast_to(tab, (size,)).copy()
        # entries for inputs never seen on the inputset may be anything; arithmetic is mod 2^(W+1) anyway
        span = 4 * size if half else 2 * size
        tab = ((tab + span // 2) % span) - span // 2
        h = hashlib.blake2b(tab.tobytes() + bytes([half]), digest_size=16).digest()
        tid = table_ids.get(h)
        if tid is None:
            tid = table_ids[h] = len(tables)
            tables.append(tab)
            table_half.append(half)
        rep = lookup_ids.get((key, tid))
        if rep is not None:
            return rep
        if base is None:
            base = state["next"]
            state["next"] += 1
        lookup_ids[(key, tid)] = base
        job_info[base] = (key, tid)
        state["nu2"] = max(state["nu2"], sum(c * c for _b, c in key[0]))
        return base

    def _weighted(parts):
        """sum_j 2^j parts[j] as (terms, const)"""
        terms, const = {}, 0
        for j, (t, c) in enumerate(parts):
            for b, cf in t.items():
                v = terms.get(b, 0) + (cf << j)
                if v:
                    terms[b] = v
                else:
                    terms.pop(b, None)
            const += c << j
        return terms, const

    def _sign_digit(parts):
        """a value in {-1, 0, 1} with the lexicographic sign of up to three such digits (least significant first)"""
        if len(parts) == 1:
            return parts[0]
        terms, const = _weighted(parts)
        rep = add_lookup(None, (tuple(sorted(terms.items())), const + off16), np.sign(dom - off16))
        return {rep: 1}, 0

    def _segment(leaves, level, idx):
        """sign digit of the leaves [idx 3^level, (idx + 1) 3^level)"""
        if level == 0:
            return leaves[idx]
        return _sign_digit([_segment(leaves, level - 1, 3 * idx + j) for j in range(3)])

    def _prefix(leaves, level, count):
        """sign digit of the leaves [0, count 3^level)"""
        if count <= 3:
            return _sign_digit([_segment(leaves, level, j) for j in range(count)])
        q, r = divmod(count, 3)
        return _sign_digit([_prefix(leaves, level + 1, q)] + [_segment(leaves, level, 3 * q + j) for j in range(r)])

    def _prefix_borrow(jb, terms, uv):
        for b1, c1 in terms.items():
            if c1 != -1 or b1 not in chain:
                continue
            leaves = chain[b1] + [({b: c for b, c in terms.items() if b != b1}, jb.const)]
            k, r = divmod(len(leaves), 3)
            parts = ([_prefix(leaves, 1, k)] if k else []) + [leaves[3 * k + j] for j in range(r)]
            t_terms, t_const = _weighted(parts)
            rep = add_lookup(jb.base, (tuple(sorted(t_terms.items())), t_const + off16), np.where(dom - off16 < 0, uv[0], uv[1]))
            if rep != jb.base:
                subst[jb.base] = {rep: 1}
            if uv == (1, 0):
                chain.setdefault(rep, leaves)
            state["collapsed"] += 1
            return True
        return False

    def _is_bit(b):
        if b < n_in:
            return b in bit_inputs
        if b - n_in < len(jobs):
            v = jobs[b - n_in].out_vals
            return v.size > 0 and v.min() >= 0 and v.max() <= 1
        return b in chain                                      # derived lookups: only chain borrows are known bits

    def _bitwise_on_borrow(jb, terms):
        if len(terms) != 2:
            return None
        (b0, c0), (b1, c1) = terms.items()
        for (bb, cb), (ob, co) in (((b0, c0), (b1, c1)), ((b1, c1), (b0, c0))):
            if bb in chain and cb > 0 and co > 0 and _is_bit(ob):
                pts = np.array([jb.const, jb.const + co, jb.const + cb, jb.const + cb + co], dtype=np.int64)   # (B, o) = 00 01 10 11
                try:
                    tt = tuple(int(v) for v in np.broadcast_to(np.asarray(jb.fn(pts)), (4,)))
                except Exception:
                    return None
                if tt[0] == tt[1] == tt[2] != tt[3]:            # a function of B & o  (digit 1 - o)
                    return {bb: -1, ob: -1}, 1, (tt[3], tt[0])
                if tt[1] == tt[2] == tt[3] != tt[0]:            # a function of B | o  (digit -o)
                    return {bb: -1, ob: -1}, 0, (tt[3], tt[0])
        return None

    def _collapse_borrow(jb, terms, uv):
        """uv: the lookup is u where its source is negative, v elsewhere ((1, 0): a borrow bit)"""
        if prefix_ok:
            return _prefix_borrow(jb, terms, uv)
        for b1, c1 in terms.items():
            if c1 != -1 or b1 not in lt0_src:
                continue
            s_terms, s_const, s_lo, s_hi = lt0_src[b1]
            a_terms = {b: c for b, c in terms.items() if b != b1}
            a_lo, a_hi = a_span
            for margin in margins:                          # tolerate S one step outside its span, if it fits
                M = max(-(s_lo - margin), s_hi + margin + 1, 1)
                t_lo, t_hi = M * a_lo + s_lo - margin, M * a_hi + s_hi + margin
                if t_hi - t_lo + 1 <= size:
                    break
            else:
                continue
            t_terms = dict(s_terms)
            for b, c in a_terms.items():
                v = t_terms.get(b, 0) + M * c
                if v:
                    t_terms[b] = v
                else:
                    t_terms.pop(b, None)
            t_const = M * jb.const + s_const
            offset = (size - (t_hi - t_lo + 1)) // 2 - t_lo
            key = (tuple(sorted(t_terms.items())), t_const + offset)
            rep = add_lookup(jb.base, key, np.where(dom - offset < 0, uv[0], uv[1]))
            if rep != jb.base:
                subst[jb.base] = {rep: 1}
            if uv == (1, 0):
                lt0_src.setdefault(rep, (t_terms, t_const, t_lo, t_hi))
                chain.setdefault(rep, [(t_terms, t_const)])
            state["collapsed"] += 1
            return True
        return False

    for j in np.flatnonzero(live):        # jobs are in creation (topological) order
        jb = jobs[j]
        g = jb.group
        terms = resolve(jb.terms)
        lt0 = scales.get(int(j), 0)
        if lt0 and _collapse_borrow(jb, terms, (lt0, 0)):
            continue
        if collapse_borrows and not lt0 and not os.environ.get("BMI_NO_FOLD"):
            # borrow & bit, borrow | bit (the overflow test after a subtraction, base_p_arrays.py:134/137) are one more,
            # most significant, digit of the same chain: [1 - o - B < 0] = B & o, [-o - B < 0] = B | o
            as_digit = _bitwise_on_borrow(jb, terms)
            if as_digit is not None and _collapse_borrow(_Pseudo(jb.base, as_digit[1]), as_digit[0], as_digit[2]):
                state["bitwise_folded"] += 1
                continue
        if (g.hi - g.lo + 1) + 2 * guard <= size or not narrowed:
            offset = (size - (g.hi - g.lo + 1)) // 2 - g.lo           # centre [lo, hi] in [0, 2^W)
            key = (tuple(sorted(terms.items())), jb.const + offset)
            raw = np.broadcast_to(np.asarray(jb.fn(dom - offset), dtype=np.int64), (size,))
            rep = add_lookup(jb.base, key, raw)
            if rep != jb.base:
                subst[jb.base] = {rep: 1}
            if lt0 == 1:
                lt0_src.setdefault(rep, (dict(terms), jb.const, min(g.lo, span_lo), max(g.hi, span_hi)))
                chain.setdefault(rep, [(dict(terms), jb.const)])
            continue
        # one bit too wide: sign through the padding bit, then negacyclic + cyclic halves (module docstring)
        assert bits[int(j)] <= W + 1, "lookup more than one bit wider than the program width"
        offset = (2 * size - (g.hi - g.lo + 1)) // 2 - g.lo           # centre [lo, hi] in [0, 2^(W+1))
        f = np.asarray(jb.fn(np.arange(2 * size, dtype=np.int64) - offset), dtype=np.int64)
        f = np.broadcast_to(f, (2 * size,))
        key_y = (tuple(sorted(terms.items())), jb.const + offset)
        full_keys.add(key_y)
        sgn = add_lookup(None, key_y, np.full(size, size // 2))
        neg = add_lookup(None, key_y, f[:size] - f[size:], half=True)
        low = dict(terms)
        low[sgn] = low.get(sgn, 0) + 1
        key_low = (tuple(sorted(low.items())), jb.const + offset - size // 2)
        cyc = add_lookup(None, key_low, f[:size] + f[size:], half=True)
        subst[jb.base] = {neg: 1, cyc: 1}
        for helper in (sgn, neg, cyc):
            split_of.setdefault(helper, jb.base)
        state["split"] += 1
    nu2 = state["nu2"]

    # --- levels (as soon as possible)
    level_of = {}
    for base, (key, _tid) in job_info.items():
        lv = 1
        for b, _c in key[0]:
            if b >= n_in:
                lv = max(lv, level_of[b] + 1)
        level_of[base] = lv
    n_levels = max(level_of.values(), default=0)
    by_level = [[] for _ in range(n_levels)]
    for base, lv in level_of.items():
        by_level[lv - 1].append(base)

    # --- outputs as linear combinations of representatives
    out_rows = []
    for o in outputs:
        out_rows.append(({}, o) if isinstance(o, int) else (resolve(o.terms), o.const))

    # --- liveness -> slot reuse
    last_use = {b: 0 for b in range(n_in)}
    for base, (key, _tid) in job_info.items():
        for b, _c in key[0]:
            last_use[b] = max(last_use.get(b, 0), level_of[base])
    for terms, _k in out_rows:
        for b in terms:
            last_use[b] = n_levels + 1
    slot_of, free, n_slots = {}, [], 0
    for b in range(n_in):
        slot_of[b] = n_slots
        n_slots += 1
    expiring = [[] for _ in range(n_levels + 2)]
    for b in range(n_in):
        expiring[last_use[b]].append(b)
    levels = []
    for li, bases in enumerate(by_level, start=1):
        for b in bases:
            if free:
                slot_of[b] = free.pop()
            else:
                slot_of[b] = n_slots
                n_slots += 1
            expiring[max(last_use.get(b, li), li)].append(b)
        # keyswitch rows shared by lookups with the same input
        ks_rows, ks_index = [], {}
        job_ks, job_lut, job_out = [], [], []
        for b in bases:
            key, tid = job_info[b]
            r = ks_index.get(key)
            if r is None:
                r = ks_index[key] = len(ks_rows)
                ks_rows.append(key)
            job_ks.append(r)
            job_lut.append(tid)
            job_out.append(slot_of[b])
        row_ptr = np.zeros(len(ks_rows) + 1, np.int32)
        idx, coef, konst = [], [], []
        for r, (terms, k) in enumerate(ks_rows):
            for b, c in terms:
                idx.append(slot_of[b])
                coef.append(c)
            row_ptr[r + 1] = len(idx)
            konst.append(k)
        levels.append(Level(row_ptr, np.asarray(idx, np.int32), np.asarray(coef, np.int64), np.asarray(konst, np.int64),
                            np.asarray(job_ks, np.int32), np.asarray(job_lut, np.int32), np.asarray(job_out, np.int32),
                            np.array([key in full_keys for key in ks_rows], bool)))
        for b in expiring[li]:              # values last read at this level free their slot for the next one
            free.append(slot_of[b])

    out_ptr = np.zeros(len(out_rows) + 1, np.int32)
    oidx, ocoef, okonst = [], [], []
    for r, (terms, k) in enumerate(out_rows):
        for b, c in sorted(terms.items()):
            oidx.append(slot_of[b])
            ocoef.append(c)
        out_ptr[r + 1] = len(oidx)
        okonst.append(k)

    prog = Program(W, n_in, n_slots, np.arange(n_in, dtype=np.int32), levels, out_ptr, np.asarray(oidx, np.int32),
                   np.asarray(ocoef, np.int64), np.asarray(okonst, np.int64), tuple(out_shape),
                   np.stack(tables) if tables else np.zeros((1, size), np.int64), int(nu2),
                   table_half=np.asarray(table_half, bool) if tables else None)
    prog.stats = {"traced_lookups": len(jobs), "live_lookups": int(live.sum()), "pbs": prog.n_pbs, "keyswitches": prog.n_ks,
                  "levels": n_levels, "tables": len(tables), "slots": n_slots, "width": W, "nu2": int(nu2),
                  "max_level_pbs": max((len(l.job_ks) for l in levels), default=0), "split_lookups": state["split"], "top_width_lookups": n_top,
                  "collapsed_borrows": state["collapsed"], "bitwise_folded": state["bitwise_folded"]}
    prog.debug = {"level_of": level_of, "subst": subst, "job_info": job_info, "split_of": split_of}      # for scripts/critical_path.py; not saved
    return prog
