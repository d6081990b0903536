#!/usr/bin/env python
"""bench.py -- PBS/s per GPU on the QFloat microbench (BASELINE.json configs[1]).

Workload: 4,096 independent encrypted QFloat pairs, base 2, medium precision (length 31, 16 integer
digits), each pair going through the reference's QFloat add, mul and div circuits (compiled programs
traced from the unmodified reference: tests/golden/qf_{add,mul,div}_medium.npz).  One STEP pushes
`--pairs` of those pairs (one batch lane each) through all three programs: every circuit level is one
batched lincomb -> keyswitch -> PBS launch group.  PBS/s = table lookups executed / time.

  value : inputs (ciphertexts) already resident in HBM, CUDA-event timed
  e2e   : same step through Circuit.run(), i.e. HOST ciphertext buffers in, HOST ciphertext buffers out
  --impl reference : the CPU restatement (oracle/tfhe_oracle_fast.c, all host threads) on a bounded sample
  --inversion X    : additionally time one encrypted inversion (default on one GPU: inv3_low_prefix; X = 2|3|4 -> tests/golden/inv{X}_low.npz, or a program
                     name such as inv3_medium / inv4_high)

Multi-GPU: launched by torchrun; the pairs are independent, so every rank takes its own `--pairs`
lanes with replicated keys and no data-path collective ("weak" scaling).  The level-sharded
all-gather mode used for a single inversion is exercised by --inversion with world size > 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
OPS = ("add", "mul", "div")
METRIC = "PBS/s per GPU; encrypted 3x3 LU-inversion wall time at 1/2/4/8 B200"


def load_programs():
    from bounty_matrix_inversion_b200.fhe.program import Program
    progs = {}
    for op in OPS:
        path = os.path.join(GOLDEN, f"qf_{op}_medium.npz")
        z = np.load(path)
        progs[op] = (Program.load(path), z["golden_inputs"].astype(np.int64), z["golden_outputs"].astype(np.int64))
    return progs


def synthetic_pairs(count, seed):
    """`count` QFloat pairs uniform in +-[0,100) as arrays+signs rows (the reference's test sampler,
    tests/test_qfloat_fhe.py:124), quantised to length 31 / 16 integer digits, base 2"""
    rs = np.random.RandomState(seed)
    qlen, ints = 31, 16
    f = rs.uniform(0, 100, (count, 2)) * rs.choice([-1, 1], (count, 2))
    mag = np.floor(np.abs(f) * 2.0 ** (qlen - ints)).astype(np.int64)
    digits = (mag[..., None] >> np.arange(qlen - 1, -1, -1)) & 1            # most significant digit first
    signs = np.where(f < 0, -1, 1).astype(np.int64)
    return np.concatenate([digits.reshape(count, 2 * qlen), signs], axis=1)


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


_CPU_CTX = {}


def cpu_context(progs):
    """keys and transform-domain key of the CPU restatement, built once per process"""
    if not _CPU_CTX:
        from bounty_matrix_inversion_b200 import native, params as PR
        from oracle import oracle as orc
        orc.build()
        for op in OPS:
            prog = progs[op][0]
            prm = PR.for_width(prog.width, prog.nu2)
            keys = native.ClientKeys(prm, seed=11)
            _CPU_CTX[op] = (prm, keys, orc.Fast(prm, keys.bsk, keys.ksk), prog.lut_polynomials(prm.N)[:1])
    return _CPU_CTX


def cpu_sample(progs, threads, seconds_hint=15.0):
    """time the CPU restatement (keyswitch + PBS, all host threads) on a bounded sample of the workload's lookups"""
    from bounty_matrix_inversion_b200 import params as PR
    ctx = cpu_context(progs)
    total_pbs = sum(p.n_pbs for p, _, _ in progs.values())
    done, spent, parts = 0, 0.0, []
    for op in OPS:
        prog = progs[op][0]
        prm, keys, fast, luts = ctx[op]
        share = prog.n_pbs / total_pbs
        count = threads
        cts = keys.encrypt([PR.encode(i % 2, prog.width) for i in range(count)])
        t0 = time.time()
        fast.batch(luts, np.zeros(count, np.int32), cts, with_ks=True, threads=threads)     # calibrate
        dt = time.time() - t0
        reps = max(1, int(seconds_hint * share / max(dt, 1e-3)))
        count = threads * reps
        cts = np.tile(cts, (reps, 1))
        t0 = time.time()
        fast.batch(luts, np.zeros(count, np.int32), cts, with_ks=True, threads=threads)
        dt = time.time() - t0
        parts.append((prog.n_pbs, count / dt))
        done += count
        spent += dt
    # PBS/s on the workload's own mix of parameter sets (harmonic mean weighted by lookups per pair)
    rate = total_pbs / sum(n / r for n, r in parts)
    return rate, done, spent


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    progs = load_programs()
    rates = []
    for _ in range(args.warmup):
        cpu_sample(progs, threads, seconds_hint=2.0)
    t0 = time.time()
    done = 0
    for _ in range(args.steps):
        r, d, _s = cpu_sample(progs, threads, seconds_hint=args.cpu_seconds)
        rates.append(r)
        done += d
    wall = time.time() - t0
    value = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "PBS/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 (mod 2^64-2^32+1)", "data": "synthetic",
            "config": {"workload": "qfloat_microbench_4096pairs_base2_medium", "ops": "add,mul,div",
                       "pbs_per_pair": sum(p.n_pbs for p, _, _ in progs.values()),
                       "params": {op: __import__("bounty_matrix_inversion_b200.params", fromlist=["x"]).for_width(
                           progs[op][0].width, progs[op][0].nu2).name for op in OPS},
                       "note": "concrete-python (the reference's FHE runtime) is not installable here; this arm times the "
                               "CPU restatement of the same keyswitch+PBS path on the workload's parameter sets"},
            "cpu_baseline": {"value": value, "unit": "PBS/s", "cores": threads, "kind": "port",
                             "sample": f"{done} keyswitch+PBS over {args.steps} steps, mixed as in the workload"},
            "e2e": {"value": value, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--pairs", type=int, default=48, help="QFloat pairs (batch lanes) per step and per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--inversion", default="auto", help="also time one encrypted inversion: 2 | 3 | 4 (low precision), a "
                                                       "compiled program in tests/golden (e.g. inv3_medium, inv4_high_prefix), "
                                                       "none, or auto = the 3x3 low-precision one on a single GPU")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from bounty_matrix_inversion_b200 import fhe, params as PR

    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    progs = load_programs()
    P = args.pairs
    circuits, inputs, outputs = {}, {}, {}
    pairs = synthetic_pairs(4096, seed=2026)
    my = pairs[(rank * P) % 4096: (rank * P) % 4096 + P]
    if len(my) < P:
        my = np.concatenate([my, pairs[: P - len(my)]])
    for op in OPS:
        prog = progs[op][0]
        c = fhe.Circuit.from_program(prog, configuration=fhe.Configuration(device=local, seed=100 + len(circuits)))
        c.keygen()
        ex = c.executor(device=local)
        enc = c.encrypt_batch([(row,) for row in my])
        circuits[op], inputs[op] = c, enc
        ex._ensure(P)
        ex.profile = []
    pinned = {op: torch.from_numpy(np.ascontiguousarray(inputs[op].cts.transpose(1, 0, 2)).view(np.int64)).pin_memory() for op in OPS}
    pbs_per_step = P * sum(progs[op][0].n_pbs for op in OPS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the three circuits are independent: each runs its levels on its own stream so small levels overlap
    streams = {op: torch.cuda.Stream(device=local) for op in OPS}
    for op in OPS:
        circuits[op]._executor.stream = streams[op]

    def fork():
        cur = torch.cuda.current_stream()
        for op in OPS:
            streams[op].wait_stream(cur)

    def join():
        cur = torch.cuda.current_stream()
        for op in OPS:
            cur.wait_stream(streams[op])

    def step_device():
        fork()
        for op in OPS:
            circuits[op]._executor.run_device(P)
        join()

    host_out = {op: None for op in OPS}

    def step_e2e():
        fork()
        for op in OPS:
            ex = circuits[op]._executor
            with torch.cuda.stream(streams[op]):
                ex.vals[: ex.prog.n_inputs].copy_(pinned[op], non_blocking=True)
                ex.run_device(P)
                if host_out[op] is None:
                    host_out[op] = torch.empty(ex.outs.shape, dtype=ex.outs.dtype, pin_memory=True)
                host_out[op].copy_(ex.outs, non_blocking=True)
        join()
        torch.cuda.current_stream().synchronize()         # the caller holds the result ciphertexts in host memory
        return host_out

    # inputs resident before the device-timed region
    for op in OPS:
        ex = circuits[op]._executor
        ex.vals[: ex.prog.n_inputs].copy_(pinned[op])
    for _ in range(args.warmup):
        step_device()
    launches0 = sum(circuits[op]._executor.eng.launch_count for op in OPS)
    for op in OPS:
        circuits[op]._executor.profile = []
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    launches = sum(circuits[op]._executor.eng.launch_count for op in OPS) - launches0
    prof = {op: circuits[op]._executor.collect_profile(ev0) for op in OPS}
    for op in OPS:
        circuits[op]._executor.profile = None

    barrier()
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        outs = step_e2e()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.time() - t0) * 1e3)

    # correctness of what was just timed: decrypt lanes 0..3 of the last end-to-end step and compare with the
    # clear evaluation of the same compiled program (itself pinned to the reference's clear path by tests/golden)
    check = {}
    for op in OPS:
        c = circuits[op]
        got = c.decrypt(fhe.EncryptedData(outs[op].numpy().view(np.uint64).transpose(1, 0, 2)[:4], batch=True))
        try:
            want = c.program.evaluate_clear(my[:4])
            check[op] = bool(np.array_equal(np.stack(got), want))
        except OverflowError as e:                 # a fresh input left the ranges seen on the compile-time inputset
            check[op] = f"range: {e}"

    times = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = times.tolist()

    inv = None
    which = args.inversion if args.inversion != "auto" else ("inv3_low_prefix" if world == 1 else "none")
    if which and which != "none":
        if world > 1:
            inv = time_inversion(which, fhe, PR, local, rank, world, dist)
        else:
            try:                                   # the second half of the headline metric; never at the price of the line
                inv = time_inversion(which, fhe, PR, local, rank, world, None)
            except Exception as e:                 # noqa: BLE001
                inv = {"program": which, "error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        total = pbs_per_step * world * args.steps
        value = total / dev_ms * 1e3
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        # dominant kernel: pbs_kernel.  Algorithmic bytes per lookup = the whole bootstrapping key once + the
        # switched input + the LUT + the output ciphertext
        alg_bytes = sum(prof[op]["alg_bytes"] for op in OPS)
        # the three circuits run on three streams: PBS time = length of the union of all launch intervals
        iv = sorted(t for op in OPS for t in prof[op]["intervals"])
        pbs_ms, cur_a, cur_b = 0.0, None, None
        for a_, b_ in iv:
            if cur_b is None or a_ > cur_b:
                if cur_b is not None:
                    pbs_ms += cur_b - cur_a
                cur_a, cur_b = a_, b_
            else:
                cur_b = max(cur_b, b_)
        if cur_b is not None:
            pbs_ms += cur_b - cur_a
        n_launch = sum(prof[op]["pbs_launches"] for op in OPS)
        achieved = alg_bytes / (pbs_ms * 1e-3) / 1e9 if pbs_ms else None
        int_ops = sum(prof[op]["int_ops"] for op in OPS)
        sm_mhz = clocks["sm_mhz"] or 1965.0
        int_peak = 148 * 128 * sm_mhz * 1e6                       # 4 SMSPs x 32 lanes, one integer instruction / clk / lane
        line = {
            "metric": METRIC, "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 (mod 2^64-2^32+1)", "data": "synthetic",
            "config": {"workload": "qfloat_microbench_4096pairs_base2_medium", "ops": "add,mul,div", "pairs_per_step_per_gpu": P,
                       "pbs_per_pair": pbs_per_step // P, "levels_per_pair": sum(len(progs[op][0].levels) for op in OPS),
                       "params": {op: circuits[op].params.name for op in OPS}, "parallelism": f"dp{world} (pairs sharded, keys replicated)",
                       "l2": "bootstrapping keys (46-110 MB per op) plus value store exceed what stays resident across the three programs; "
                             "each step re-streams all three key sets"},
            "e2e": {"value": total / e2e_ms * 1e3, "unit": "PBS/s",
                    "h2d_bytes_per_step": int(sum(pinned[op].numel() * 8 for op in OPS)) * world,
                    "d2h_bytes_per_step": int(sum(circuits[op]._executor.outs.numel() * 8 for op in OPS)) * world},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak if achieved else None, "traffic": None, "peak_source": peak_src,
                         "kernel": "pbs_cluster_kernel", "launches": n_launch, "avg_launch_ms": sum(prof[op]["pbs_ms"] for op in OPS) / max(n_launch, 1),
                         "pbs_busy_ms": pbs_ms,
                         "pbs_share_of_step": pbs_ms / dev_ms,
                         "int_pipe": {"achieved_Tops": int_ops / (pbs_ms * 1e-3) / 1e12 if pbs_ms else None,
                                      "peak_Tops": int_peak / 1e12, "frac": int_ops / (pbs_ms * 1e-3) / int_peak if pbs_ms else None,
                                      "note": "model count of 32-bit integer instructions the NTT external product needs (DESIGN.md section 5)"}},
            "parity_check": check,
            "blind_rotation": {op: "pairs" if circuits[op].params.bsk_group == 2 else "single" for op in OPS},
        }
        if inv:
            line["inversion"] = inv
        if not args.no_cpu:
            rate, done, spent = cpu_sample(progs, os.cpu_count() or 1, args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": "PBS/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"{done} keyswitch+PBS in {spent:.1f}s, mixed as in the workload"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def time_inversion(n, fhe, PR, local, rank, world, dist):
    """one encrypted n x n QFloat inversion (low precision), levels sharded over the ranks when world > 1"""
    import torch
    from bounty_matrix_inversion_b200.fhe.program import Program
    name = f"inv{n}_low" if str(n).isdigit() else str(n)
    path = os.path.join(GOLDEN, name + ".npz")
    z, prog = np.load(path), Program.load(path)
    c = fhe.Circuit.from_program(prog, configuration=fhe.Configuration(device=local, seed=77))
    c.keygen()
    ex = c.executor(rank=rank, world=world, device=local)
    x, want = z["golden_inputs"].astype(np.int64)[0], z["golden_outputs"].astype(np.int64)[0]
    enc = c.encrypt(x)
    c.run(enc)                                    # warm-up
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.time()
    out = c.run(enc)
    torch.cuda.synchronize()
    wall = time.time() - t0
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = bool(np.array_equal(c.decrypt(out), want))
    return {"program": name, "n": int(z["meta_n"]), "qfloat_len": int(z["meta_qfloat_len"]), "wall_s": float(t.item()), "pbs": prog.n_pbs, "levels": len(prog.levels),
            "params": c.params.name, "digits_match_reference_clear_path": ok, "world": world}


if __name__ == "__main__":
    main()
