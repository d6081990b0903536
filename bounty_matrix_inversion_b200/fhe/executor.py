"""GPU execution of a levelled Program through the C ABI (one launch group per level).

PyTorch only owns device memory, streams and the process group here; every arithmetic step is a kernel of
libbmi_tfhe.so (bmi_lincomb -> bmi_keyswitch -> bmi_pbs, plus bmi_scatter_rows after an all-gather).

* The program is re-levelled for the engine's launch capacity first (fhe/schedule.py): lookups with slack leave
  levels that would exceed what one bootstrap launch handles at minimum latency.
* The whole static program is captured once per batch size as a CUDA graph and replayed (thousands of launches
  become one submission); `graphs=False` keeps the eager per-level loop.
* world > 1 (one process per GPU, keys replicated): rank r bootstraps a contiguous share of every level's lookups
  -- and keyswitches only the rows those lookups read -- straight into its slice of a gather buffer; one in-place
  NCCL all-gather per level, then a row scatter into the value slots.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import params as PR
from ..native import Engine
from .program import Program
from .schedule import schedule_for

P = PR.P


def _field(vals, scale=1):
    """signed python/numpy ints -> uint64 field elements (times scale), returned as an int64-typed view"""
    out = np.array([(int(v) * scale) % P for v in vals], dtype=np.uint64)
    return out.view(np.int64)


def shard_bounds(n_jobs: int, rank: int, world: int):
    """contiguous share of a level's lookups owned by `rank`: (rows per rank in the gather buffer, first, last+1)"""
    per = (n_jobs + world - 1) // world
    return per, min(rank * per, n_jobs), min((rank + 1) * per, n_jobs)


class Executor:
    def __init__(self, program: Program, params: PR.TfheParams, engine: Engine, device: int = 0,
                 rank: int = 0, world: int = 1, group=None, torch_device=None, level_capacity="auto", graphs=True):
        """engine: anything with load_luts / lincomb / keyswitch / pbs / scatter_rows taking torch tensors
        (native.Engine on a GPU; the CPU tests of the multi-rank plumbing inject a clear-text stand-in).
        rank/world/group: level sharding.  level_capacity: lookups per level the scheduler aims for ("auto": what the
        engine reports per GPU, times world; 0 / None: keep the program's own levels)."""
        self.params, self.eng = params, engine
        self.dev = torch.device("cuda", device) if torch_device is None else torch.device(torch_device)
        self.rank, self.world, self.group = rank, world, group
        if level_capacity == "auto":
            level_capacity = int(getattr(engine, "pbs_capacity", 0)) * world
        # every program goes through the scheduler: with no capacity it only normalises the level layout (lookups of
        # a level ordered by keyswitch row, which the sharded path relies on)
        program = schedule_for(program, int(level_capacity or 0), world)
        self.prog = program
        self.level_capacity = int(level_capacity or 0)
        W1 = getattr(engine, "words", params.big_dim + 1)
        self.W1 = W1
        # a clear-text stand-in engine computes in half-message units (Program.tables_half_units)
        clear = getattr(engine, "clear", False)
        engine.load_luts(program.tables_half_units() if clear else program.lut_polynomials(params.N))
        scale = 2 if clear else PR.delta(program.width)
        lv = program.levels

        def cat(arrs, dtype):
            return np.concatenate([np.asarray(a, dtype) for a in arrs]) if arrs else np.zeros(0, dtype)

        # per-level arrays concatenated once; each level addresses its slice
        self.ks_off = np.cumsum([0] + [len(l.konst) for l in lv])
        self.nz_off = np.cumsum([0] + [len(l.idx) for l in lv])
        self.pbs_off = np.cumsum([0] + [len(l.job_ks) for l in lv])
        row_ptr = cat([l.row_ptr.astype(np.int64) for l in lv], np.int64)      # each level keeps its own n_ks+1 entries
        self.rp_off = np.cumsum([0] + [len(l.row_ptr) for l in lv])
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        self.d_row_ptr = t(row_ptr.astype(np.int32))
        self.d_idx = t(cat([l.idx for l in lv], np.int32))
        self.d_coef = t(_field(cat([l.coef for l in lv], np.int64)))
        self.d_konst = t(_field(cat([l.konst for l in lv], np.int64), scale))
        self.d_job_lut = t(cat([l.job_lut for l in lv], np.int32))
        self.d_job_out = t(cat([l.job_out for l in lv], np.int32))
        # per level: this rank's lookups [lo, hi), the keyswitch rows [r0, r1) they read, and the lookups' rows
        # relative to r0
        self.shard = []
        local_ks = []
        for l in lv:
            per, lo, hi = shard_bounds(len(l.job_ks), rank, world)
            if hi > lo:
                assert np.all(np.diff(l.job_ks) >= 0), "lookups of a level must be ordered by keyswitch row"
                r0, r1 = int(l.job_ks[lo]), int(l.job_ks[hi - 1]) + 1
            else:
                r0 = r1 = 0
            self.shard.append((per, lo, hi, r0, r1))
            local_ks.append(l.job_ks - (r0 if world > 1 else 0))
        self.d_job_ks = t(cat(local_ks, np.int32))
        self.d_out_ptr = t(program.out_row_ptr.astype(np.int32))
        self.d_out_idx = t(program.out_idx.astype(np.int32))
        self.d_out_coef = t(_field(program.out_coef))
        self.d_out_konst = t(_field(program.out_konst, scale))
        self.max_ks = max((len(l.konst) for l in lv), default=1)
        self.max_pbs = max((len(l.job_ks) for l in lv), default=1)
        self._batch = 0
        self.stream = None           # torch.cuda.Stream for every launch of this executor (None = current stream)
        self.profile = None          # list of (start event, end event, jobs) around every PBS launch when profiling
        self.graphs = bool(graphs) and self.dev.type == "cuda"
        self._graph = None           # (batch, torch.cuda.CUDAGraph)
        self.level_events = None     # with collect_level_times(): CUDA events around lincomb+keyswitch / PBS / exchange

    def collect_profile(self, origin=None):
        """PBS-kernel time measured with CUDA events on the launching stream, and the algorithmic work it covers;
        with `origin` (an event recorded before the region) also the [start, end] offsets of every launch in ms"""
        if self.dev.type == "cuda":
            torch.cuda.synchronize(self.dev)
        p = self.params
        prof = self.profile or []
        ms = sum(a.elapsed_time(b) for a, b, _ in prof)
        jobs = sum(j for _, _, j in prof)
        per_pbs_bytes = p.bsk_bytes() + (p.n + 1) * 8 + p.N * 8 + (p.big_dim + 1) * 8
        out = {"pbs_ms": ms, "pbs_launches": len(prof), "pbs_jobs": jobs, "alg_bytes": jobs * per_pbs_bytes}
        if origin is not None:
            out["intervals"] = [(origin.elapsed_time(a), origin.elapsed_time(b)) for a, b, _ in prof]
        return out

    def _ensure(self, batch):
        if batch == self._batch:
            return
        p = self.params
        self._graph = None
        self.vals = torch.zeros((self.prog.n_slots, batch, self.W1), dtype=torch.int64, device=self.dev)
        self.ks_in = torch.empty((self.max_ks * batch, self.W1), dtype=torch.int64, device=self.dev)
        self.small = torch.empty((self.max_ks * batch, getattr(self.eng, "small_words", p.n + 1)), dtype=torch.int64, device=self.dev)
        n_out = len(self.prog.out_konst)
        self.outs = torch.empty((n_out, batch, self.W1), dtype=torch.int64, device=self.dev)
        if self.world > 1:
            rows = shard_bounds(self.max_pbs, 0, self.world)[0] * self.world       # padded so every rank owns `per` rows
            self.stage = torch.zeros((rows, batch, self.W1), dtype=torch.int64, device=self.dev)
            self.iota = torch.arange(rows, dtype=torch.int32, device=self.dev)
        self._batch = batch

    def run(self, input_cts: np.ndarray) -> np.ndarray:
        """input_cts: [n_inputs][kN+1] or [batch][n_inputs][kN+1] uint64 (host) -> output ciphertexts (host)"""
        x = np.ascontiguousarray(input_cts, dtype=np.uint64)
        single = x.ndim == 2
        if single:
            x = x[None]
        batch = x.shape[0]
        self._ensure(batch)
        host = torch.from_numpy(np.ascontiguousarray(x.transpose(1, 0, 2)).view(np.int64))
        with self._on_stream():
            self.vals[: self.prog.n_inputs].copy_(host, non_blocking=True)
            self.run_device(batch)
            out = self.outs.cpu()
        out = out.numpy().view(np.uint64).transpose(1, 0, 2)
        return out[0] if single else np.ascontiguousarray(out)

    # ---- streams: every torch op of this executor runs on the same stream as its kernels
    def _s(self):
        if self.dev.type != "cuda":
            return None
        return self.stream if self.stream is not None else torch.cuda.current_stream(self.dev)

    def _on_stream(self):
        import contextlib
        if self.dev.type != "cuda" or self.stream is None:
            return contextlib.nullcontext()
        return torch.cuda.stream(self.stream)

    def run_device(self, batch):
        """inputs already in self.vals[:n_inputs]; leaves output ciphertexts in self.outs"""
        if not self.graphs or self.profile is not None or self.level_events is not None:
            return self._issue(batch)
        if self._graph is None or self._graph[0] != batch:
            self._capture(batch)
        if self._graph[1] is None:       # capture was not possible on this setup: eager loop
            return self._issue(batch)
        with self._on_stream():
            self._graph[1].replay()

    def _capture(self, batch):
        # value slots are reused by liveness, so a run overwrites its own inputs: keep them across the warm-up run
        inputs = self.vals[: self.prog.n_inputs].clone()
        self._issue(batch)               # warm-up: lazy allocations and occupancy queries happen outside the capture
        torch.cuda.synchronize(self.dev)
        self.vals[: self.prog.n_inputs].copy_(inputs)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.dev)
        keep, self.stream = self.stream, side
        try:
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.graph(g, stream=side):
                self._issue(batch)
            self._graph = (batch, g)
        except Exception:                # noqa: BLE001 -- e.g. a collective that cannot be captured: fall back to eager
            torch.cuda.synchronize(self.dev)
            self._graph = (batch, None)
        finally:
            self.stream = keep
        torch.cuda.synchronize(self.dev)

    def _issue(self, batch):
        eng, prog = self.eng, self.prog
        for li in range(len(prog.levels)):
            self._level(li, batch)
        eng.lincomb(self.vals, self.d_out_ptr, self.d_out_idx, self.d_out_coef, self.d_out_konst, self.outs,
                    len(prog.out_konst), batch, stream=self._s())

    def _mark(self, st):
        e = torch.cuda.Event(enable_timing=True)
        e.record(st)
        return e

    def _level(self, li, batch):
        eng = self.eng
        st = self._s()
        k0 = self.ks_off[li]
        p0, p1 = self.pbs_off[li], self.pbs_off[li + 1]
        n_pbs = int(p1 - p0)
        if n_pbs == 0:
            return
        per, lo, hi, r0, r1 = self.shard[li]
        n_rows = r1 - r0
        ev = [self._mark(st)] if self.level_events is not None else None
        if hi > lo:
            # this level's CSR rows index the concatenated idx/coef arrays through their own row_ptr (level-local
            # offsets); a rank forms and keyswitches only the rows [r0, r1) its lookups read
            rp = self.d_row_ptr[self.rp_off[li] + r0: self.rp_off[li] + r1 + 1]
            idx = self.d_idx[self.nz_off[li]: self.nz_off[li + 1]]
            coef = self.d_coef[self.nz_off[li]: self.nz_off[li + 1]]
            konst = self.d_konst[k0 + r0: k0 + r1]
            eng.lincomb(self.vals, rp, idx, coef, konst, self.ks_in, n_rows, batch, stream=st)
            eng.keyswitch(self.ks_in, self.small, n_rows * batch, stream=st)
        if ev is not None:
            ev.append(self._mark(st))
        job_ks, job_lut = self.d_job_ks[p0 + lo: p0 + hi], self.d_job_lut[p0 + lo: p0 + hi]
        if self.world == 1:
            job_out = self.d_job_out[p0:p1]
            if self.profile is not None:
                a = self._mark(st)
                eng.pbs(self.small, job_ks, job_lut, job_out, self.vals, n_pbs, batch, stream=st)
                self.profile.append((a, self._mark(st), n_pbs * batch))
            else:
                eng.pbs(self.small, job_ks, job_lut, job_out, self.vals, n_pbs, batch, stream=st)
            if ev is not None:
                ev.append(self._mark(st))
                self.level_events.append((n_pbs, ev))
            return
        # ---- one box, several GPUs: bootstrap my share straight into my slice of the gather buffer, all-gather in
        # place over NVLink, scatter the level's outputs into their value slots
        import torch.distributed as dist
        stage = self.stage[: per * self.world]
        if hi > lo:
            eng.pbs(self.small, job_ks, job_lut, self.iota[lo:hi], stage, hi - lo, batch, stream=st)
        if ev is not None:
            ev.append(self._mark(st))
        with self._on_stream():
            dist.all_gather_into_tensor(stage, stage[self.rank * per: (self.rank + 1) * per], group=self.group)
        eng.scatter_rows(stage, self.d_job_out[p0:p1], self.vals, n_pbs, batch, stream=st)
        if ev is not None:
            ev.append(self._mark(st))
            self.level_events.append((n_pbs, ev))

    def collect_level_times(self):
        """(lookups, ms lincomb+keyswitch, ms bootstrap, ms exchange) per level of the last eager run"""
        torch.cuda.synchronize(self.dev)
        out = []
        for n_pbs, ev in self.level_events or []:
            out.append((n_pbs, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]) if len(ev) > 3 else 0.0))
        return out
