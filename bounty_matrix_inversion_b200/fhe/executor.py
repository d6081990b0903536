"""GPU execution of a levelled Program through the C ABI (one launch group per level).

PyTorch only owns device memory and streams here; every arithmetic step is a kernel of
libbmi_tfhe.so (bmi_lincomb -> bmi_keyswitch -> bmi_pbs).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import params as PR
from ..native import Engine
from .program import Program

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
                 rank: int = 0, world: int = 1, group=None, torch_device=None):
        """engine: anything with load_luts / lincomb / keyswitch / pbs taking torch tensors (native.Engine on a GPU;
        the CPU tests of the multi-rank plumbing inject a clear-text stand-in).  rank/world/group: level sharding."""
        self.prog, self.params, self.eng = program, params, engine
        self.dev = torch.device("cuda", device) if torch_device is None else torch.device(torch_device)
        self.rank, self.world, self.group = rank, world, group
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
        self.d_job_ks = t(cat([l.job_ks for l in lv], np.int32))
        self.d_job_lut = t(cat([l.job_lut for l in lv], np.int32))
        self.d_job_out = t(cat([l.job_out for l in lv], np.int32))
        self.d_out_ptr = t(program.out_row_ptr.astype(np.int32))
        self.d_out_idx = t(program.out_idx.astype(np.int32))
        self.d_out_coef = t(_field(program.out_coef))
        self.d_out_konst = t(_field(program.out_konst, scale))
        self.max_ks = max((len(l.konst) for l in lv), default=1)
        self.max_pbs = max((len(l.job_ks) for l in lv), default=1)
        self._batch = 0
        self.stream = None           # torch.cuda.Stream for every launch of this executor (None = current stream)
        self.profile = None          # list of (start event, end event, jobs) around every PBS launch when profiling

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
        butterflies = (p.k + 1) * (p.bsk_l + 1) * (p.N // 2) * p.logN
        per_step = butterflies * 26 + (p.k + 1) ** 2 * p.bsk_l * p.N * 22 + (p.k + 1) * p.bsk_l * p.N * 12
        if p.bsk_group == 2:
            # pair key: per slot and CTA one product + three subtractions for the monomials, six products + four
            # additions to combine the three keys for both output polynomials, two products with the digits
            per_step = butterflies * 26 + (p.k + 1) * p.N * (9 * 22 + 7 * 3) + (p.k + 1) * p.bsk_l * p.N * 12
        per_pbs_int = (p.n // p.bsk_group) * per_step
        out = {"pbs_ms": ms, "pbs_launches": len(prof), "pbs_jobs": jobs, "alg_bytes": jobs * per_pbs_bytes,
               "int_ops": jobs * per_pbs_int}
        if origin is not None:
            out["intervals"] = [(origin.elapsed_time(a), origin.elapsed_time(b)) for a, b, _ in prof]
        return out

    def _ensure(self, batch):
        if batch == self._batch:
            return
        p = self.params
        self.vals = torch.zeros((self.prog.n_slots, batch, self.W1), dtype=torch.int64, device=self.dev)
        self.ks_in = torch.empty((self.max_ks * batch, self.W1), dtype=torch.int64, device=self.dev)
        self.small = torch.empty((self.max_ks * batch, getattr(self.eng, "small_words", p.n + 1)), dtype=torch.int64, device=self.dev)
        n_out = len(self.prog.out_konst)
        self.outs = torch.empty((n_out, batch, self.W1), dtype=torch.int64, device=self.dev)
        if self.world > 1:
            rows = shard_bounds(self.max_pbs, 0, self.world)[0] * self.world       # padded so every rank owns `per` rows
            self.stage = torch.empty((rows, batch, self.W1), dtype=torch.int64, device=self.dev)
            self.mine = torch.empty((rows // self.world, batch, self.W1), dtype=torch.int64, device=self.dev)
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
        self.vals[: self.prog.n_inputs].copy_(host, non_blocking=True)
        self.run_device(batch)
        out = self.outs.cpu().numpy().view(np.uint64).transpose(1, 0, 2)
        return out[0] if single else np.ascontiguousarray(out)

    def run_device(self, batch):
        """inputs already in self.vals[:n_inputs]; leaves output ciphertexts in self.outs"""
        eng, prog = self.eng, self.prog
        for li in range(len(prog.levels)):
            self._level(li, batch)
        eng.lincomb(self.vals, self.d_out_ptr, self.d_out_idx, self.d_out_coef, self.d_out_konst, self.outs,
                    len(prog.out_konst), batch, stream=self._s())

    def _s(self):
        if self.dev.type != "cuda":
            return None
        return self.stream if self.stream is not None else torch.cuda.current_stream(self.dev)

    def _level(self, li, batch):
        eng = self.eng
        st = self._s()
        k0, k1 = self.ks_off[li], self.ks_off[li + 1]
        p0, p1 = self.pbs_off[li], self.pbs_off[li + 1]
        n_ks, n_pbs = int(k1 - k0), int(p1 - p0)
        if n_pbs == 0:
            return
        # this level's CSR rows index the concatenated idx/coef arrays through their own row_ptr (level-local offsets)
        rp = self.d_row_ptr[self.rp_off[li]: self.rp_off[li + 1]]
        idx = self.d_idx[self.nz_off[li]: self.nz_off[li + 1]]
        coef = self.d_coef[self.nz_off[li]: self.nz_off[li + 1]]
        konst = self.d_konst[k0:k1]
        job_ks, job_lut, job_out = self.d_job_ks[p0:p1], self.d_job_lut[p0:p1], self.d_job_out[p0:p1]
        if self.world == 1:
            eng.lincomb(self.vals, rp, idx, coef, konst, self.ks_in, n_ks, batch, stream=st)
            eng.keyswitch(self.ks_in, self.small, n_ks * batch, stream=st)
            if self.profile is not None:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st)
                eng.pbs(self.small, job_ks, job_lut, job_out, self.vals, n_pbs, batch, stream=st)
                b.record(st)
                self.profile.append((a, b, n_pbs * batch))
            else:
                eng.pbs(self.small, job_ks, job_lut, job_out, self.vals, n_pbs, batch, stream=st)
            return
        self._level_sharded(li, batch, rp, idx, coef, konst, job_ks, job_lut, job_out, n_ks, n_pbs)

    # ---- one box, several GPUs: each rank bootstraps a contiguous share of the level's lookups, then the
    # output ciphertexts are all-gathered over NVLink (keys are replicated on every GPU)
    def _level_sharded(self, li, batch, rp, idx, coef, konst, job_ks, job_lut, job_out, n_ks, n_pbs):
        import torch.distributed as dist
        eng = self.eng
        per, lo, hi = shard_bounds(n_pbs, self.rank, self.world)
        # every rank forms all keyswitch inputs it needs; for simplicity all rows (cheap next to the bootstraps)
        st = self._s()
        eng.lincomb(self.vals, rp, idx, coef, konst, self.ks_in, n_ks, batch, stream=st)
        eng.keyswitch(self.ks_in, self.small, n_ks * batch, stream=st)
        stage = self.stage[: per * self.world]
        if hi > lo:
            eng.pbs(self.small, job_ks[lo:hi], job_lut[lo:hi], self.iota[lo:hi], stage, hi - lo, batch, stream=st)
        mine = stage[self.rank * per: (self.rank + 1) * per]
        self.mine[:per].copy_(mine)                       # all_gather needs an input that does not alias the output
        dist.all_gather_into_tensor(stage, self.mine[:per], group=self.group)
        self.vals.index_copy_(0, job_out.long(), stage[:n_pbs])
