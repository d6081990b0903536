"""`fhe.Compiler` / `fhe.Configuration` / `fhe.Circuit`: the compile-encrypt-run-decrypt surface the
reference drives (/root/reference/matrix_inversion/main.py:53-116,
qfloat_matrix_inversion.py:978-1052), implemented on the B200 engine."""
from __future__ import annotations

import dataclasses
import inspect
import time

import numpy as np

from .. import params as PR
from . import tracing
from .program import Program, lower


class Configuration:
    """accepts Concrete's keyword options; the ones this engine understands:
    tfhe_params (explicit TfheParams), p_error_sigmas, slack_bits, device,
    seed (None = OS entropy, like Concrete's keygen(); an integer = fixed keys for tests and benchmarks only),
    multiplication ("auto" | "quarter_square", see tracing.Trace),
    split_wide ("auto" | True | False), split_guard and collapse_borrows (see program.lower),
    blind_rotation ("pairs" | "single"): two key bits per bootstrap step with the pair key (default wherever the
    parameter set has one decomposition level and an even LWE dimension), or one GGSW per key bit,
    level_capacity ("auto" | int | 0): lookups per circuit level the run-time scheduler aims for (fhe/schedule.py;
    auto = what the engine bootstraps at minimum latency, per GPU and batch lane), cuda_graphs (replay the program as
    one CUDA graph, default True)"""

    def __init__(self, **options):
        self.options = dict(options)
        self.tfhe_params = options.get("tfhe_params")
        self.p_error_sigmas = options.get("p_error_sigmas", 6.5)
        self.slack_bits = options.get("slack_bits", 0)
        self.seed = options.get("seed")          # None: keys and encryptions from OS entropy; int: fixed test seed
        self.device = options.get("device", 0)
        self.multiplication = options.get("multiplication", "auto")
        self.split_wide = options.get("split_wide", "auto")
        self.split_guard = options.get("split_guard")
        self.collapse_borrows = options.get("collapse_borrows", True)
        self.blind_rotation = options.get("blind_rotation", "pairs")
        self.level_capacity = options.get("level_capacity", "auto")
        self.cuda_graphs = options.get("cuda_graphs", True)
        if self.blind_rotation not in ("pairs", "single"):
            raise ValueError("blind_rotation must be 'pairs' or 'single'")

    def fork(self, **options):
        merged = dict(self.options)
        merged.update(options)
        return Configuration(**merged)


class EncryptedData:
    """ciphertexts crossing the client/server boundary (fhe.PublicArguments / fhe.PublicResult)"""

    def __init__(self, cts: np.ndarray, batch: bool):
        self.cts, self.batch = cts, batch


PublicArguments = EncryptedData
PublicResult = EncryptedData


class Compiler:
    def __init__(self, function, parameter_encryption_statuses):
        self.function = function
        names = list(inspect.signature(function).parameters)
        if set(parameter_encryption_statuses) != set(names):
            raise ValueError(f"encryption statuses {sorted(parameter_encryption_statuses)} do not match parameters {names}")
        for n, st in parameter_encryption_statuses.items():
            if st != "encrypted":
                raise NotImplementedError(f"parameter '{n}': only 'encrypted' inputs are supported")
        self.names = names

    def trace(self, inputset, multiplication="auto"):
        samples = [s if isinstance(s, (tuple, list)) else (s,) for s in inputset]
        if not samples:
            raise ValueError("inputset is empty")
        for s in samples:
            if len(s) != len(self.names):
                raise ValueError("inputset sample arity does not match the function")
        cols = [np.stack([np.asarray(s[i], dtype=np.int64) for s in samples]) for i in range(len(self.names))]
        trace = tracing.Trace(len(samples), multiplication)
        with trace:
            args = []
            for col in cols:
                shape = col.shape[1:]
                flat = col.reshape(len(samples), -1)
                arr = np.empty(flat.shape[1], dtype=object)
                for e in range(flat.shape[1]):
                    arr[e] = trace.new_input(flat[:, e])
                args.append(tracing.Tracer(arr.reshape(shape)))
            result = self.function(*args)
            outs = result if isinstance(result, (tuple, list)) else (result,)
            flat_out, shapes = [], []
            for o in outs:
                a = tracing._obj(o)
                shapes.append(a.shape)
                for v in a.reshape(-1):
                    v = tracing._as_scalar(v)
                    flat_out.append(v if isinstance(v, int) else tracing._aff(v))
        return trace, flat_out, shapes, [c.shape[1:] for c in cols]

    def compile(self, inputset=None, configuration=None, verbose=False, **options):
        cfg = configuration or Configuration()
        if options:
            cfg = cfg.fork(**options)
        t0 = time.time()
        trace, flat_out, out_shapes, in_shapes = self.trace(inputset, cfg.multiplication)
        t1 = time.time()
        n_out = len(flat_out)
        prog = lower(trace, flat_out, (n_out,), slack_bits=cfg.slack_bits, split_wide=cfg.split_wide,
                     split_guard=cfg.split_guard, collapse_borrows=cfg.collapse_borrows)
        t2 = time.time()
        prm = None if cfg.tfhe_params == "deferred" else (cfg.tfhe_params or _select_params(prog, cfg))
        prog.stats.update(trace_s=round(t1 - t0, 3), lower_s=round(t2 - t1, 3), params_s=round(time.time() - t2, 3))
        if verbose:
            print("compiled:", prog.stats, prm)
        return Circuit(prog, prm, in_shapes, out_shapes, cfg)


def _select_params(program: Program, cfg: Configuration) -> PR.TfheParams:
    return PR.for_width(program.width, program.nu2, bsk_group=2 if cfg.blind_rotation == "pairs" else 1)


class Circuit:
    def __init__(self, program: Program, params: PR.TfheParams, in_shapes, out_shapes, cfg: Configuration):
        if params is not None:
            # explicit parameter sets follow the configuration too, where the pair blind rotation can run them
            pairs = cfg.blind_rotation == "pairs" and params.bsk_l == 1 and params.n % 2 == 0
            params = dataclasses.replace(params, bsk_group=2 if pairs else 1)
        self.program, self.params, self.cfg = program, params, cfg
        self.in_shapes, self.out_shapes = [tuple(s) for s in in_shapes], [tuple(s) for s in out_shapes]
        self.keys = None
        self._executor = None

    @classmethod
    def from_program(cls, program: Program, params=None, in_shapes=None, out_shapes=None, configuration=None):
        """a compiled Program (e.g. Program.load of a .npz traced elsewhere) -> runnable circuit"""
        cfg = configuration or Configuration()
        prm = params or cfg.tfhe_params or _select_params(program, cfg)
        in_shapes = in_shapes or [(program.n_inputs,)]
        out_shapes = out_shapes or [tuple(program.out_shape)]
        return cls(program, prm, in_shapes, out_shapes, cfg)

    # ---- introspection
    @property
    def statistics(self):
        return dict(self.program.stats, params=self.params.name if self.params else None)

    # ---- client
    def keygen(self, force=False, seed=None):
        from ..native import ClientKeys
        if self.keys is None or force:
            self.keys = ClientKeys(self.params, self.cfg.seed if seed is None else seed, pairs=self.params.bsk_group == 2)
            self._executor = None
        return self.keys

    def _flatten_args(self, args):
        if len(args) != len(self.in_shapes):
            raise ValueError(f"expected {len(self.in_shapes)} arguments, got {len(args)}")
        flat = []
        for a, shp in zip(args, self.in_shapes):
            a = np.asarray(a, dtype=np.int64)
            if a.shape != shp:
                raise ValueError(f"argument shape {a.shape} != {shp}")
            flat.append(a.reshape(-1))
        return np.concatenate(flat)

    def encrypt(self, *args):
        self.keygen()
        msg = self._flatten_args(args)
        W = self.program.width
        cts = self.keys.encrypt([PR.encode(int(m), W) for m in msg])
        return EncryptedData(cts, batch=False)

    def encrypt_batch(self, list_of_args):
        """independent evaluations of the same circuit, executed as one batch (one lane each)"""
        self.keygen()
        W = self.program.width
        msgs = np.stack([self._flatten_args(a) for a in list_of_args])
        cts = self.keys.encrypt([PR.encode(int(m), W) for m in msgs.reshape(-1)])
        return EncryptedData(cts.reshape(msgs.shape[0], msgs.shape[1], -1), batch=True)

    def decrypt(self, result: EncryptedData):
        W = self.program.width
        ph = self.keys.phase(result.cts.reshape(-1, self.params.big_dim + 1))
        vals = np.array([PR.decode_signed(int(p), W) for p in ph], dtype=np.int64)
        if result.batch:
            vals = vals.reshape(result.cts.shape[0], -1)
            return [self._unflatten(v) for v in vals]
        return self._unflatten(vals)

    def _unflatten(self, vals):
        outs, pos = [], 0
        for shp in self.out_shapes:
            n = int(np.prod(shp)) if shp else 1
            outs.append(vals[pos: pos + n].reshape(shp))
            pos += n
        return outs[0] if len(outs) == 1 else tuple(outs)

    # ---- server
    def executor(self, rank=0, world=1, group=None, device=None, lanes=1):
        """lanes: independent evaluations per run (batch lanes); the level scheduler divides the engine's launch
        capacity by it"""
        if self._executor is None:
            from ..native import Engine
            from .executor import Executor
            self.keygen()
            dev = self.cfg.device if device is None else device
            eng = Engine(self.params, dev)
            eng.load_keys(self.keys.bsk, self.keys.ksk, bskp=self.keys.bskp)
            capacity = self.cfg.level_capacity
            if capacity == "auto":
                capacity = eng.pbs_capacity * world // max(1, lanes)
            self._executor = Executor(self.program, self.params, eng, dev, rank, world, group, level_capacity=capacity,
                                      graphs=self.cfg.cuda_graphs)
        return self._executor

    def run(self, encrypted: EncryptedData) -> EncryptedData:
        out = self.executor(lanes=encrypted.cts.shape[0] if encrypted.batch else 1).run(encrypted.cts)
        return EncryptedData(out, encrypted.batch)

    def encrypt_run_decrypt(self, *args):
        return self.decrypt(self.run(self.encrypt(*args)))

    def simulate(self, *args):
        return self._unflatten(self.program.evaluate_clear(self._flatten_args(args)))
