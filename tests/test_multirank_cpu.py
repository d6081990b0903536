"""world_size-2 (and 3) CPU test of the level-sharded execution path over gloo: every rank bootstraps its share
of each level, output "ciphertexts" are all-gathered and scattered into the value slots.  The engine is replaced
by a clear-text stand-in (one word per ciphertext) so the multi-rank plumbing runs without a GPU; the result must
equal the single-process clear evaluation and the reference's golden digits."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bounty_matrix_inversion_b200 import params as PR
from bounty_matrix_inversion_b200.fhe.executor import Executor, shard_bounds
from bounty_matrix_inversion_b200.fhe.program import Program

HERE = os.path.dirname(os.path.abspath(__file__))


class ClearEngine:
    """same call surface as native.Engine, operating on clear messages in half-message units (1 word per ciphertext)"""
    clear, words, small_words = True, 1, 1

    def __init__(self, width):
        self.size = 1 << width

    def load_luts(self, tables):
        self.tables = torch.from_numpy(np.asarray(tables, dtype=np.int64))

    def lincomb(self, vals, row_ptr, idx, coef, konst, out, njobs, batch=1, stream=None):
        # field elements arrive as the int64 view of a uint64 in [0, p): x >= 2^63 means the negative number x - p
        unfield = lambda v: torch.where(v < 0, v + (2 ** 32 - 1), v)
        signed, konst = unfield(coef), unfield(konst)
        rp = row_ptr.tolist()
        for j in range(njobs):
            acc = konst[j].expand(batch).clone()
            for t in range(rp[j], rp[j + 1]):
                acc = acc + signed[t] * vals[idx[t].item(), :, 0]
            out.view(-1, 1)[j * batch:(j + 1) * batch, 0] = acc

    def keyswitch(self, big, small, count, stream=None):
        small.view(-1, 1)[:count] = big.view(-1, 1)[:count]

    def pbs(self, small, job_in, job_lut, job_out, out, njobs, batch=1, stream=None):
        s, o = small.view(-1, batch, 1), out.view(-1, batch, 1)
        for q in range(njobs):
            m = torch.remainder(s[job_in[q].item(), :, 0] >> 1, 2 * self.size)      # negacyclic over the padding bit
            t = self.tables[job_lut[q].item(), torch.where(m >= self.size, m - self.size, m)]
            o[job_out[q].item(), :, 0] = torch.where(m >= self.size, -t, t)

    def scatter_rows(self, src, dst_row, dst, count, batch=1, stream=None):
        dst.view(-1, batch, 1)[dst_row[:count].long()] = src.view(-1, batch, 1)[:count]


def _messages(half_units, width):
    """what decryption returns: the signed message mod 2^(width+1)"""
    size = 1 << width
    return np.mod(half_units // 2 + size, 2 * size) - size


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, path, capacity, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z, prog = np.load(path), Program.load(path)
    ex = Executor(prog, PR.TOY_1024, ClearEngine(prog.width), rank=rank, world=world, torch_device="cpu", level_capacity=capacity)
    if capacity:
        over = lambda p: sum(len(l.job_ks) > capacity for l in p.levels)
        assert over(ex.prog) < over(prog) and len(ex.prog.levels) == len(prog.levels) and ex.prog.n_pbs == prog.n_pbs
    x = z["golden_inputs"].astype(np.int64)[:3]
    batch = x.shape[0]
    ex._ensure(batch)
    ex.vals[: prog.n_inputs, :, 0] = torch.from_numpy(2 * x.T)
    ex.run_device(batch)
    got = _messages(ex.outs[:, :, 0].numpy().T, prog.width)
    ok = bool(np.array_equal(got, z["golden_outputs"].astype(np.int64)[:3])) and bool(np.array_equal(got, prog.evaluate_clear(x)))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,capacity", [(2, 0), (3, 0), (2, 6)])
def test_level_sharded_execution_over_gloo(world, capacity):
    """capacity > 0: the levels are re-balanced for that many lookups per level first (fhe/schedule.py)"""
    path = os.path.join(HERE, "golden", "qf_add_medium.npz")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), path, capacity, ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}


def test_shard_bounds_cover_every_job_once():
    for n in (0, 1, 2, 7, 8, 9, 124, 1444):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                per, lo, hi = shard_bounds(n, r, world)
                assert hi - lo <= per and per * world >= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def test_single_rank_clear_engine_matches_program():
    """the stand-in engine itself is faithful (world = 1 uses the unsharded path)"""
    path = os.path.join(HERE, "golden", "qf_mul_medium.npz")
    z, prog = np.load(path), Program.load(path)
    ex = Executor(prog, PR.TOY_4096, ClearEngine(prog.width), torch_device="cpu")
    x = z["golden_inputs"].astype(np.int64)[:2]
    ex._ensure(2)
    ex.vals[: prog.n_inputs, :, 0] = torch.from_numpy(2 * x.T)
    ex.run_device(2)
    assert np.array_equal(_messages(ex.outs[:, :, 0].numpy().T, prog.width), z["golden_outputs"].astype(np.int64)[:2])
