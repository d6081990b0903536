"""CPU tests of the capacity-aware re-levelling (fhe/schedule.py): same lookups, same depth, same clear results on the
reference's golden inputs, fewer levels above the launch capacity; and the round trip through the level layout is the
identity when there is no capacity."""
import os

import numpy as np
import pytest

from bounty_matrix_inversion_b200.fhe.program import Program
from bounty_matrix_inversion_b200.fhe.schedule import _dag, rebalance

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    path = os.path.join(GOLDEN, name + ".npz")
    z = np.load(path)
    return Program.load(path), z["golden_inputs"].astype(np.int64), z["golden_outputs"].astype(np.int64)


@pytest.mark.parametrize("name,capacity", [("inv2_low", 36), ("inv2_low_prefix", 8), ("qf_div_medium", 36), ("qf_mul_medium", 36),
                                           ("inv3_low_prefix", 36)])
def test_rebalanced_program_is_the_same_circuit(name, capacity):
    prog, x, want = _load(name)
    new = rebalance(prog, capacity)
    assert len(new.levels) == len(prog.levels), "re-levelling must not lengthen the critical path"
    assert new.n_pbs == prog.n_pbs and new.n_ks == prog.n_ks
    over = lambda p: sum(len(l.job_ks) > capacity for l in p.levels)
    assert over(new) < over(prog)
    assert np.array_equal(new.evaluate_clear(x[:2]), want[:2])
    # the same multiset of (lookup table, keyswitch row content) groups, only placed at other levels
    sig = lambda p: sorted((g["konst"], g["full"], tuple(c for _, c in g["terms"]), tuple(l for _, l in g["jobs"])) for g in _dag(p)[0])
    assert sig(new) == sig(prog)
    for l in new.levels:
        assert np.all(np.diff(l.job_ks) >= 0), "lookups of a level are ordered by keyswitch row (the sharded executor relies on it)"


def test_no_capacity_keeps_every_lookup_at_its_level():
    prog, x, want = _load("inv2_low")
    new = rebalance(prog, 1 << 30)
    assert [len(l.job_ks) for l in new.levels] == [len(l.job_ks) for l in prog.levels]
    assert np.array_equal(new.evaluate_clear(x[:1]), want[:1])


def test_forced_lookups_are_never_delayed():
    """a level may exceed the capacity only with lookups that cannot wait: a capacity of one still yields the depth"""
    prog, x, want = _load("qf_add_medium")
    new = rebalance(prog, 1)
    assert len(new.levels) == len(prog.levels)
    assert np.array_equal(new.evaluate_clear(x[:2]), want[:2])


def test_program_save_refuses_values_that_do_not_fit(tmp_path):
    prog, _x, _w = _load("qf_add_medium")
    prog.save(str(tmp_path / "ok.npz"))
    back = Program.load(str(tmp_path / "ok.npz"))
    assert all(np.array_equal(a.coef, b.coef) and np.array_equal(a.job_lut, b.job_lut) for a, b in zip(prog.levels, back.levels))
    prog.levels[0].coef[0] = 2 ** 40
    with pytest.raises(OverflowError):
        prog.save(str(tmp_path / "bad.npz"))


def test_layout_choice_for_several_gpus():
    """with several GPUs sharing each level the scheduler may stop filling at a quarter of the per-GPU capacity (the size
    a launch runs at its lowest latency); whichever layout it picks, it is the cheaper one by its own model and the
    same circuit"""
    from bounty_matrix_inversion_b200.fhe.schedule import modelled_ms, schedule_for
    prog, x, want = _load("inv3_low_prefix")
    for world in (1, 2, 8):
        cap = 33 * world
        new = schedule_for(prog, cap, world)
        assert len(new.levels) == len(prog.levels) and new.n_pbs == prog.n_pbs
        assert modelled_ms(new, cap, world) <= modelled_ms(rebalance(prog, cap), cap, world) + 1e-9
        assert modelled_ms(new, cap, world) < modelled_ms(prog, cap, world)
    assert max(len(l.job_ks) for l in new.levels) <= 8 * 8          # 8 GPUs: every rank at most 8 lookups per level
    assert np.array_equal(new.evaluate_clear(x[:1]), want[:1])
    assert len(schedule_for(prog, 0, 1).levels) == len(prog.levels)
