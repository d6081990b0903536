#!/usr/bin/env python
"""bench.py -- PBS/s per GPU on the QFloat microbench (BASELINE.json configs[1]) and encrypted 3x3 inversion wall time.

Workload (default): the reference's QFloat add, mul and div circuits (base 2, medium precision: length 31, 16 integer
digits; compiled programs traced from the unmodified reference, tests/golden/qf_{add,mul,div}_medium.npz) on
independent encrypted pairs drawn from the 4,096-pair set of BASELINE.json.  One STEP pushes `--pairs` of those pairs
per GPU (default 48, one batch lane each: 48 x 5,069 = 243,312 table lookups) through all three programs; every circuit
level is one batched lincomb -> keyswitch -> PBS launch group.  PBS/s = table lookups executed / time.  The full
4,096-pair pass is `--pairs 512` on 8 GPUs (profiles/).

  value     : inputs (ciphertexts) already resident in HBM, CUDA-event timed
  e2e       : same step through host buffers: HOST ciphertexts in (pinned), HOST ciphertexts out
  inversion : one encrypted 3x3 LU inversion (tests/golden/inv3_low_prefix.npz), wall time through Circuit.run();
              with N > 1 GPUs every level's lookups are sharded over the ranks and all-gathered (NCCL)
  --impl reference : the CPU restatement (oracle/tfhe_oracle_fast.c, all host threads) on a bounded sample
  --workload inv3_medium_batch --lanes K : BASELINE.json configs[4]: K independent 3x3 medium-precision inversions
              per GPU as batch lanes (replicas over GPUs), inversions/s and PBS/s
  --workload pbs_sweep : raw keyswitch+PBS batch-size sweep on every GPU

Multi-GPU: launched by torchrun, one process per GPU, keys replicated.  The microbench pairs are independent, so every
rank takes its own lanes with no data-path collective ("weak" scaling); the inversion is the path with a real exchange.
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
COUNTERS = os.path.join(ROOT, "profiles", "r2_pbs_kernel_counters.json")     # ncu-measured constants of the bootstrap kernels


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


# ---------------------------------------------------------------------------------------------- CPU arm
_CPU_CTX = {}


def cpu_context(progs):
    """keys and transform-domain keys of the CPU restatement, built once per process: the SAME parameter sets and the same
    pair blind rotation as the GPU arm, plus the one-GGSW-per-bit rotation (a CPU backend may prefer it)"""
    if not _CPU_CTX:
        from bounty_matrix_inversion_b200 import native, params as PR
        from oracle import oracle as orc
        orc.use_native_build()
        by_set = {}                                   # circuits that selected the same numbers share one key set
        for op in OPS:
            prog = progs[op][0]
            prm = PR.for_width(prog.width, prog.nu2, bsk_group=2)
            sig = (prm.n, prm.k, prm.N, prm.bsk_bl, prm.bsk_l, prm.ksk_bl, prm.ksk_l, prm.lwe_sigma, prm.glwe_sigma)
            if sig not in by_set:
                pk = native.ClientKeys(prm, seed=11, pairs=True)
                sk = native.ClientKeys(prm, seed=11, pairs=False)
                by_set[sig] = (pk, orc.Fast(prm, None, pk.ksk, bskp=pk.bskp), orc.Fast(prm, sk.bsk, sk.ksk))
            pk, fast_pairs, fast_single = by_set[sig]
            _CPU_CTX[op] = (prm, pk, fast_pairs, fast_single, prog.lut_polynomials(prm.N)[:1])
    return _CPU_CTX


def cpu_sample(progs, threads, seconds_hint=15.0):
    """time the CPU restatement (keyswitch + PBS, all host threads) on a bounded sample of the workload's lookups.
    Returns (PBS/s with the faster of the two blind rotations, lookups done, seconds, which rotation)"""
    from bounty_matrix_inversion_b200 import params as PR
    ctx = cpu_context(progs)
    total_pbs = sum(p.n_pbs for p, _, _ in progs.values())
    out = {}
    for kind in ("pairs", "single"):
        done, spent, parts = 0, 0.0, []
        for op in OPS:
            prog = progs[op][0]
            prm, keys, fast_pairs, fast_single, luts = ctx[op]
            fast = fast_pairs if kind == "pairs" else fast_single
            share = prog.n_pbs / total_pbs
            count = threads
            cts = keys.encrypt([PR.encode(i % 2, prog.width) for i in range(count)])
            t0 = time.time()
            fast.batch(luts, np.zeros(count, np.int32), cts, with_ks=True, threads=threads)     # calibrate
            dt = time.time() - t0
            reps = max(1, int(seconds_hint / 2 * share / max(dt, 1e-3)))
            count = threads * reps
            cts = np.tile(cts, (reps, 1))
            t0 = time.time()
            fast.batch(luts, np.zeros(count, np.int32), cts, with_ks=True, threads=threads)
            dt = time.time() - t0
            parts.append((prog.n_pbs, count / dt))
            done += count
            spent += dt
        # PBS/s on the workload's own mix of parameter sets (harmonic mean weighted by lookups per pair)
        out[kind] = (total_pbs / sum(n / r for n, r in parts), done, spent)
    best = max(out, key=lambda k: out[k][0])
    rate, done, spent = out[best]
    return rate, done + out["pairs" if best == "single" else "single"][1], spent + out["pairs" if best == "single" else "single"][2], \
        f"{best} blind rotation (pairs {out['pairs'][0]:.1f}, single {out['single'][0]:.1f} PBS/s)"


def workload_config(progs, pairs, world):
    from bounty_matrix_inversion_b200 import params as PR
    return {"workload": f"qfloat_microbench_base2_medium_{pairs}of4096pairs_per_gpu_step", "ops": "add,mul,div",
            "pairs_per_step_per_gpu": pairs, "pbs_per_pair": sum(p.n_pbs for p, _, _ in progs.values()),
            "levels_per_pair": sum(len(progs[op][0].levels) for op in OPS),
            "params": {op: PR.for_width(progs[op][0].width, progs[op][0].nu2, bsk_group=2).name for op in OPS},
            "parallelism": f"dp{world} (pairs sharded, keys replicated)"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    progs = load_programs()
    for _ in range(args.warmup):
        cpu_sample(progs, threads, seconds_hint=2.0)
    # each step is a bounded sample; its size shrinks if the requested number of steps would not fit the time budget
    left = args.time_budget - (time.time() - args.t_start) - 20.0
    per_step = max(3.0, min(args.cpu_seconds, left / max(args.steps, 1) - 3.0))
    t0 = time.time()
    rates, done, which = [], 0, ""
    for _ in range(args.steps):
        r, d, _s, which = cpu_sample(progs, threads, seconds_hint=per_step)
        rates.append(r)
        done += d
    wall = time.time() - t0
    value = float(np.mean(rates))
    cfg = workload_config(progs, args.pairs, world)              # identical to the GPU arm's `config`
    note = ("concrete-python (the reference's FHE runtime) is not installable here; this arm times the CPU restatement "
            "of the same keyswitch+PBS path (oracle/tfhe_oracle_fast.c, -O3 -march=native, built on this host) on the "
            "workload's parameter sets, each step a bounded sample of the workload's lookups")
    line = {"impl": "reference", "note": note, "metric": METRIC, "value": value, "unit": "PBS/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 (mod 2^64-2^32+1)", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "PBS/s", "cores": threads, "kind": "port",
                             "sample": f"{done} keyswitch+PBS over {args.steps} steps, mixed as in the workload; {which}"},
            "e2e": {"value": value, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def kernel_counters():
    try:
        return json.load(open(COUNTERS))
    except (OSError, ValueError):
        return {}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="qfloat_microbench", choices=["qfloat_microbench", "inv3_medium_batch", "pbs_sweep"])
    ap.add_argument("--pairs", type=int, default=48, help="QFloat pairs (batch lanes) per step and per GPU")
    ap.add_argument("--lanes", type=int, default=32, help="inv3_medium_batch: independent inversions per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer pass (profile runs of the full 4,096-pair workload)")
    ap.add_argument("--time-budget", type=float, default=540.0, help="seconds the whole run aims to stay within: the host-buffer "
                    "(e2e) pass repeats the step as often as the remaining time allows (at least twice, at most --steps)")
    ap.add_argument("--inversion", default="auto", help="also time one encrypted inversion: 2 | 3 | 4 (low precision), a "
                                                       "compiled program in tests/golden (e.g. inv3_medium, inv4_high_prefix), "
                                                       "none, or auto = inv3_low_prefix (levels sharded over the GPUs)")
    args = ap.parse_args()
    args.t_start = time.time()
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
    try:
        if args.workload == "inv3_medium_batch":
            return run_inversion_batch(args, fhe, PR, torch, dist, local, rank, world)
        if args.workload == "pbs_sweep":
            return run_pbs_sweep(args, PR, torch, dist, local, rank, world)
        return run_microbench(args, fhe, PR, torch, dist, local, rank, world)
    finally:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()


def run_microbench(args, fhe, PR, torch, dist, local, rank, world):
    progs = load_programs()
    P = args.pairs
    circuits, inputs = {}, {}
    pairs = synthetic_pairs(4096, seed=2026)
    my = pairs[(rank * P) % 4096: (rank * P) % 4096 + P]
    if len(my) < P:
        my = np.concatenate([my, pairs[: P - len(my)]])
    for op in OPS:
        prog = progs[op][0]
        # fixed seeds: reproducible benchmark keys; eager launches (the PBS launches are bracketed by events below)
        c = fhe.Circuit.from_program(prog, configuration=fhe.Configuration(device=local, seed=100 + len(circuits), cuda_graphs=False))
        c.keygen()
        ex = c.executor(device=local, lanes=P)
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

    # the host-buffer pass repeats the same step; how often is bounded by the time the run has left (the inversion and
    # the CPU sample still need about 130 s), identically on every rank
    left = args.time_budget - (time.time() - args.t_start) - 130.0
    e2e_steps = 1 if args.no_e2e else int(max(2, min(args.steps, left / max(dev_ms / args.steps * 1e-3, 1e-3))))
    if world > 1:
        agreed = torch.tensor([e2e_steps], dtype=torch.int64, device="cuda")
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN)
        e2e_steps = int(agreed.item())
    barrier()
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        outs = step_e2e()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.time() - t0) * 1e3) * (args.steps / e2e_steps)

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
    bench_params = {op: circuits[op].params.name for op in OPS}
    h2d = int(sum(pinned[op].numel() * 8 for op in OPS)) * world
    d2h = int(sum(circuits[op]._executor.outs.numel() * 8 for op in OPS)) * world
    # free the microbench's device memory before the inversion builds its own engine
    for op in OPS:
        circuits[op]._executor.eng.close()
    circuits.clear()
    torch.cuda.empty_cache()

    inv = None
    which = args.inversion if args.inversion != "auto" else "inv3_low_prefix"
    if which and which != "none":
        try:                                       # the second half of the headline metric; never at the price of the line
            inv = time_inversion(which, fhe, PR, local, rank, world, dist if world > 1 else None)
        except Exception as e:                     # noqa: BLE001
            if world > 1:
                raise                              # a rank that gave up would leave the others inside a collective
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
        pbs_jobs = sum(prof[op]["pbs_jobs"] for op in OPS)
        alg_bytes = sum(prof[op]["alg_bytes"] for op in OPS)
        sm_mhz = clocks["sm_mhz"] or 1965.0
        # Dominant kernel: the throughput build of the bootstrap (pbs_cluster_kernel).  Its binding resource is the
        # integer issue/pipes, not HBM (the key is shared in L2): achieved = warp-instructions per second, from the
        # instructions one bootstrap executes (ncu smsp__inst_executed.sum / bootstraps of the launch, a constant of
        # the kernel build recorded in profiles/) x the bootstraps per second measured HERE with CUDA events;
        # peak = 148 SMs x 4 schedulers x SM clock (one warp-instruction per scheduler and clock).
        kc = kernel_counters().get("pbs_cluster_pairs_throughput_N2048", {})
        inst_per_pbs = kc.get("warp_inst_per_pbs")
        peak_issue = 148 * 4 * sm_mhz * 1e6
        achieved = inst_per_pbs * pbs_jobs / (pbs_ms * 1e-3) if inst_per_pbs and pbs_ms else None
        line = {
            "metric": METRIC, "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 (mod 2^64-2^32+1)", "data": "synthetic",
            "config": dict(workload_config(progs, P, world), params=bench_params),
            "l2_note": "bootstrapping keys (74 MB per op) plus keyswitch keys and value store exceed what stays resident across "
                       "the three programs; each step re-streams all three key sets",
            "e2e": {"value": total / e2e_ms * 1e3, "unit": "PBS/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps_timed": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "int_pipe", "kernel": "pbs_cluster_kernel<11,3,4,1,0,1> (throughput build, pair rotation)",
                         "achieved": achieved / 1e9 if achieved else None, "peak": peak_issue / 1e9, "unit": "Gwarp-inst/s",
                         "frac": achieved / peak_issue if achieved else None,
                         "traffic": kc.get("dram_bytes_per_launch"), "traffic_launch_pbs": kc.get("pbs_per_launch"),
                         "peak_source": "148 SMs x 4 schedulers x SM clock under load (clocks.sm_mhz)",
                         "warp_inst_per_pbs": inst_per_pbs, "counters_from": kc.get("source"),
                         "pipe_pct_ncu": kc.get("pipe_pct"), "launches": n_launch, "pbs_busy_ms": pbs_ms,
                         "pbs_share_of_step": pbs_ms / dev_ms,
                         "hbm": {"achieved_algorithmic_GBs": alg_bytes / (pbs_ms * 1e-3) / 1e9 if pbs_ms else None, "peak_GBs": hbm_peak,
                                 "peak_source": peak_src,
                                 "note": "algorithmic key bytes per bootstrap x bootstraps / PBS time; the key is served from L2 "
                                         "(DRAM traffic per launch = `traffic`), so HBM is not the roof of this kernel"}},
            "parity_check": check,
            "blind_rotation": {op: "pairs" for op in OPS},
        }
        if inv:
            line["inversion"] = inv
        if not args.no_cpu:
            rate, done, spent, which_rot = cpu_sample(progs, os.cpu_count() or 1, args.cpu_seconds)
            line["cpu_baseline"] = {"value": rate, "unit": "PBS/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"{done} keyswitch+PBS in {spent:.1f}s, mixed as in the workload; {which_rot}"}
        print(json.dumps(line), flush=True)


def time_inversion(n, fhe, PR, local, rank, world, dist):
    """one encrypted n x n QFloat inversion, levels sharded over the ranks when world > 1; wall time of Circuit.run()
    (host ciphertexts in and out), then one instrumented eager pass for the split into keyswitch / bootstrap / exchange"""
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
    c.run(enc)                                    # warm-up (captures the CUDA graph)
    torch.cuda.synchronize()
    walls = []
    for _ in range(2):
        if dist:
            dist.barrier()
        t0 = time.time()
        out = c.run(enc)
        torch.cuda.synchronize()
        walls.append(time.time() - t0)
    t = torch.tensor([min(walls)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = bool(np.array_equal(c.decrypt(out), want))
    # where the time goes: eager pass with CUDA events around every level's phases
    ex.level_events = []
    ex._ensure(1)
    ex.vals[: ex.prog.n_inputs].copy_(torch.from_numpy(np.ascontiguousarray(enc.cts[:, None, :]).view(np.int64)))
    ex.run_device(1)
    lt = ex.collect_level_times()
    ex.level_events = None
    sizes = np.array([r[0] for r in lt])
    rec = {"program": name, "n": int(z["meta_n"]), "qfloat_len": int(z["meta_qfloat_len"]), "wall_s": float(t.item()),
           "pbs": ex.prog.n_pbs, "levels": len(ex.prog.levels), "params": c.params.name,
           "digits_match_reference_clear_path": ok, "world": world, "cuda_graph": bool(ex._graph and ex._graph[1] is not None),
           "level_capacity": ex.level_capacity, "levels_above_capacity": int((sizes > max(ex.level_capacity, 1)).sum()),
           "eager_pass": {"keyswitch_s": sum(r[1] for r in lt) / 1e3, "pbs_s": sum(r[2] for r in lt) / 1e3,
                          "exchange_s": sum(r[3] for r in lt) / 1e3}}
    rec["other_s"] = max(0.0, rec["wall_s"] - rec["eager_pass"]["pbs_s"])
    ex.eng.close()
    return rec


def run_inversion_batch(args, fhe, PR, torch, dist, local, rank, world):
    """BASELINE.json configs[4]: independent 3x3 medium-precision inversions, `--lanes` per GPU as batch lanes, replicas
    over the GPUs (no collective).  (The 256-inversion workload is 32 lanes x 8 GPUs.)"""
    from bounty_matrix_inversion_b200.fhe.program import Program
    path = os.path.join(GOLDEN, "inv3_medium.npz")
    z, prog = np.load(path), Program.load(path)
    K = args.lanes
    c = fhe.Circuit.from_program(prog, configuration=fhe.Configuration(device=local, seed=300 + rank))
    c.keygen()
    xs, wants = z["golden_inputs"].astype(np.int64), z["golden_outputs"].astype(np.int64)
    lanes = [xs[(rank * K + i) % len(xs)] for i in range(K)]
    enc = c.encrypt_batch([(row,) for row in lanes])
    c.run(enc)                                    # warm-up
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    walls = []
    for _ in range(max(1, args.steps)):
        if world > 1:
            dist.barrier()
        t0 = time.time()
        out = c.run(enc)
        torch.cuda.synchronize()
        walls.append(time.time() - t0)
    clocks = sampler.stop()
    got = np.stack(c.decrypt(out))
    ok = bool(np.array_equal(got, np.stack([wants[(rank * K + i) % len(xs)] for i in range(K)])))
    t = torch.tensor([float(np.mean(walls)), 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        wall, bad = t.tolist()
        ex = c._executor
        line = {"metric": METRIC, "value": ex.prog.n_pbs * K * world / wall, "unit": "PBS/s", "n_gpus": world, "steps": len(walls),
                "warmup": 1, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64 (mod 2^64-2^32+1)", "data": "synthetic",
                "config": {"workload": f"inv3_medium_batch_{K}lanes_per_gpu", "inversions": K * world, "pbs_per_inversion": ex.prog.n_pbs,
                           "levels": len(ex.prog.levels), "params": c.params.name, "parallelism": f"replicas x{world}, {K} batch lanes each"},
                "inversions_per_s": K * world / wall, "seconds_per_inversion_amortised": wall / (K * world),
                "digits_match_reference_clear_path": bad == 0.0, "clocks": clocks,
                "e2e": {"value": ex.prog.n_pbs * K * world / wall, "unit": "PBS/s",
                        "h2d_bytes_per_step": int(enc.cts.nbytes) * world, "d2h_bytes_per_step": int(out.cts.nbytes) * world},
                "gpu_launches": int(ex.eng.launch_count)}
        print(json.dumps(line), flush=True)


def run_pbs_sweep(args, PR, torch, dist, local, rank, world):
    """raw keyswitch+PBS batch-size sweep on every GPU (replicas): aggregate PBS/s per batch size"""
    from bounty_matrix_inversion_b200 import native
    prm = PR.for_width(4, 400.0, bsk_group=2)
    keys = native.ClientKeys(prm, seed=5, pairs=True)
    eng = native.Engine(prm, local)
    eng.load_keys(None, keys.ksk, bskp=keys.bskp)
    w = 4
    eng.load_luts(np.stack([PR.lut_polynomial([PR.encode(t, w) for t in range(16)], w, prm.N)]))
    rows = []
    for count in (1, 8, 36, 74, 148, 296, 592, 1184, 2368, 4736):
        cts = np.tile(keys.encrypt([PR.encode(i % 16, w) for i in range(16)]), (count // 16 + 1, 1))[:count]
        big = torch.from_numpy(cts.view(np.int64)).cuda()
        small = torch.zeros((count, prm.n + 1), dtype=torch.int64, device="cuda")
        out = torch.zeros((count, prm.big_dim + 1), dtype=torch.int64, device="cuda")
        idx = torch.arange(count, dtype=torch.int32, device="cuda")
        lut = torch.zeros(count, dtype=torch.int32, device="cuda")
        for _ in range(2):
            eng.keyswitch(big, small, count)
            eng.pbs(small, idx, lut, idx, out, count)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 3
        for _ in range(reps):
            eng.keyswitch(big, small, count)
            eng.pbs(small, idx, lut, idx, out, count)
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dec = [PR.decode(int(p), w) for p in keys.phase(out[:4].cpu().numpy().view(np.uint64))]
        rows.append({"count_per_gpu": count, "ms": t.item(), "pbs_per_s_all_gpus": count * world / t.item() * 1e3, "decrypts": dec == [0, 1, 2, 3]})
    if rank == 0:
        print(json.dumps({"metric": METRIC, "unit": "PBS/s", "n_gpus": world, "config": {"workload": "raw_ks_pbs_batch_sweep", "params": prm.name},
                          "sweep": rows}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
