"""ctypes binding of libbmi_tfhe.so (include/bmi_tfhe.h).

There is no fallback: if the shared library is missing the import of anything that
computes raises, and nothing in this package routes work to a CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build
from .params import P, TfheParams

U64P = C.c_void_p


class BmiParams(C.Structure):
    _fields_ = [("n", C.c_int32), ("k", C.c_int32), ("N", C.c_int32), ("bsk_bl", C.c_int32), ("bsk_l", C.c_int32),
                ("ksk_bl", C.c_int32), ("ksk_l", C.c_int32), ("lwe_sigma", C.c_double), ("glwe_sigma", C.c_double)]

    @classmethod
    def of(cls, p: TfheParams):
        return cls(p.n, p.k, p.N, p.bsk_bl, p.bsk_l, p.ksk_bl, p.ksk_l, p.lwe_sigma, p.glwe_sigma)


class NativeError(RuntimeError):
    pass


_lib = None

_SIGNATURES = {
    "bmi_version": (C.c_char_p, []),
    "bmi_last_error": (C.c_char_p, []),
    "bmi_random_seed": (C.c_int, [C.c_char_p]),
    "bmi_rng_words": (C.c_int, [C.c_char_p, C.c_uint64, C.c_uint64, U64P, C.c_int64]),
    "bmi_keygen_lwe": (C.c_int, [C.c_void_p, C.c_char_p, U64P]),
    "bmi_keygen_glwe": (C.c_int, [C.c_void_p, C.c_char_p, U64P]),
    "bmi_keygen_bsk": (C.c_int, [C.c_void_p, C.c_char_p, U64P, U64P, U64P, C.c_int]),
    "bmi_keygen_bsk_pairs": (C.c_int, [C.c_void_p, C.c_char_p, U64P, U64P, U64P, C.c_int]),
    "bmi_keygen_ksk": (C.c_int, [C.c_void_p, C.c_char_p, U64P, U64P, U64P, C.c_int]),
    "bmi_lwe_encrypt": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, U64P, U64P, C.c_int64, U64P]),
    "bmi_lwe_phase": (C.c_int, [U64P, C.c_int32, U64P, C.c_int64, U64P]),
    "bmi_ctx_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bmi_ctx_destroy": (C.c_int, [C.c_void_p]),
    "bmi_ctx_load_bsk": (C.c_int, [C.c_void_p, U64P]),
    "bmi_ctx_load_bsk_pairs": (C.c_int, [C.c_void_p, U64P]),
    "bmi_ctx_load_ksk": (C.c_int, [C.c_void_p, U64P]),
    "bmi_ctx_load_luts": (C.c_int, [C.c_void_p, U64P, C.c_int32]),
    "bmi_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "bmi_ctx_set_pbs_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "bmi_ctx_set_tma_stage": (C.c_int, [C.c_void_p, C.c_int32]),
    "bmi_scatter_rows": (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_void_p]),
    "bmi_ctx_pbs_capacity": (C.c_int32, [C.c_void_p]),
    "bmi_lincomb": (C.c_int, [C.c_void_p] * 7 + [C.c_int32, C.c_int32, C.c_void_p]),
    "bmi_keyswitch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "bmi_pbs": (C.c_int, [C.c_void_p] * 6 + [C.c_int32, C.c_int32, C.c_void_p]),
    "bmi_ks_pbs_host": (C.c_int, [C.c_void_p, U64P, C.c_void_p, U64P, C.c_int64]),
    "bmi_polymul_host": (C.c_int, [C.c_void_p, U64P, U64P, U64P, C.c_int32]),
}


def exported_symbols():
    """every symbol include/bmi_tfhe.h declares"""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_build.LIB):
            raise NativeError(
                f"{_build.LIB} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                "this package has no CPU fallback")
        _lib = C.CDLL(_build.LIB)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


def _check(rc):
    if rc != 0:
        raise NativeError(f"bmi error {rc}: {lib().bmi_last_error().decode()}")


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


# ------------------------------------------------------------------ client
def random_seed() -> bytes:
    """32 bytes of OS entropy (getrandom) through the C ABI"""
    buf = C.create_string_buffer(32)
    _check(lib().bmi_random_seed(buf))
    return buf.raw


def rng_words(seed: bytes, stream: int, ctr0: int, count: int) -> np.ndarray:
    """raw ChaCha20 keystream words of the client generator (known-answer tests)"""
    out = np.zeros(count, np.uint64)
    _check(lib().bmi_rng_words(bytes(seed), stream, ctr0, _p(out), count))
    return out


def _test_seeds(seed: int):
    """three fixed 32-byte seeds from an integer: reproducible keys for tests and benchmarks ONLY"""
    import hashlib
    return [hashlib.sha256(b"bmi-fixed-seed/%s/%d" % (tag, int(seed))).digest() for tag in (b"secret", b"evaluation", b"encrypt")]


class ClientKeys:
    """secret and evaluation keys of one circuit (host memory)"""

    def __init__(self, params: TfheParams, seed=None, threads: int = 0, evaluation_keys: bool = True,
                 pairs: bool = False):
        """seed: None (default) = every key and every encryption draws from OS entropy, the secret keys from a seed of
        their own; an integer = fixed, publicly derivable seeds for reproducible tests and benchmarks, never for
        real data.
        pairs: generate the pair bootstrapping key (two key bits per blind-rotation step, `bskp`) instead of the
        one-GGSW-per-bit key (`bsk`)"""
        self.params = params
        self.deterministic = seed is not None
        if seed is None:
            self._sk_seed, self._ev_seed, self._enc_seed = random_seed(), random_seed(), random_seed()
        else:
            self._sk_seed, self._ev_seed, self._enc_seed = _test_seeds(seed)
        self._enc_counter = 0            # ciphertexts encrypted so far under _enc_seed: a counter value is never reused
        bp = BmiParams.of(params)
        L = lib()
        threads = threads or (os.cpu_count() or 1)
        self.s = np.zeros(params.n, np.uint64)
        self.S = np.zeros(params.big_dim, np.uint64)
        _check(L.bmi_keygen_lwe(C.byref(bp), self._sk_seed, _p(self.s)))
        _check(L.bmi_keygen_glwe(C.byref(bp), self._sk_seed, _p(self.S)))
        self.bsk = self.bskp = self.ksk = None
        if evaluation_keys:
            rows = (params.k + 1) * params.bsk_l
            self.ksk = np.zeros((params.big_dim, params.ksk_l, params.n + 1), np.uint64)
            if pairs:
                self.bskp = np.zeros((params.n // 2, 3, rows, params.k + 1, params.N), np.uint64)
                _check(L.bmi_keygen_bsk_pairs(C.byref(bp), self._ev_seed, _p(self.s), _p(self.S), _p(self.bskp), threads))
            else:
                self.bsk = np.zeros((params.n, rows, params.k + 1, params.N), np.uint64)
                _check(L.bmi_keygen_bsk(C.byref(bp), self._ev_seed, _p(self.s), _p(self.S), _p(self.bsk), threads))
            _check(L.bmi_keygen_ksk(C.byref(bp), self._ev_seed, _p(self.s), _p(self.S), _p(self.ksk), threads))

    def encrypt(self, plaintexts, ct_index0=None) -> np.ndarray:
        """big-key LWE encryptions of field-element plaintexts -> [count][kN+1].  Mask and noise come from the next
        unused positions of this key set's encryption keystream; `ct_index0` pins the position instead and is accepted
        only with fixed test seeds (reusing a position under one seed reveals the difference of the two messages)."""
        pt = np.atleast_1d(np.array(plaintexts, dtype=np.uint64))
        if ct_index0 is None:
            ct_index0 = self._enc_counter
            self._enc_counter += pt.size
        elif not self.deterministic:
            raise ValueError("ct_index0 is a test hook: it needs ClientKeys(seed=<int>)")
        out = np.zeros((pt.size, self.params.big_dim + 1), np.uint64)
        bp = BmiParams.of(self.params)
        _check(lib().bmi_lwe_encrypt(C.byref(bp), self._enc_seed, int(ct_index0), _p(self.S), _p(pt), pt.size, _p(out)))
        return out

    def phase(self, cts, small: bool = False) -> np.ndarray:
        key = self.s if small else self.S
        cts = _u64(cts).reshape(-1, key.size + 1)
        out = np.zeros(cts.shape[0], np.uint64)
        _check(lib().bmi_lwe_phase(_p(key), key.size, _p(cts), cts.shape[0], _p(out)))
        return out


# ------------------------------------------------------------------ server
class Engine:
    """one GPU's execution context: keys and LUTs resident in HBM, kernels behind the C ABI"""

    def __init__(self, params: TfheParams, device: int = 0):
        self.params = params
        self.device = device
        self._h = C.c_void_p()
        bp = BmiParams.of(params)
        _check(lib().bmi_ctx_create(C.byref(bp), device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().bmi_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_keys(self, bsk, ksk: np.ndarray, bskp=None):
        """bsk: one GGSW per key bit; bskp: pair key (then bootstraps rotate two key bits per step); either may be None"""
        if bsk is not None:
            _check(lib().bmi_ctx_load_bsk(self._h, _p(_u64(bsk))))
        if bskp is not None:
            _check(lib().bmi_ctx_load_bsk_pairs(self._h, _p(_u64(bskp))))
        _check(lib().bmi_ctx_load_ksk(self._h, _p(_u64(ksk))))

    def load_luts(self, luts: np.ndarray):
        luts = _u64(luts).reshape(-1, self.params.N)
        _check(lib().bmi_ctx_load_luts(self._h, _p(luts), luts.shape[0]))
        self.n_luts = luts.shape[0]

    def set_pbs_mode(self, mode: int):
        """bootstrap kernel build: 0 automatic per launch size, 1 latency build, 2 throughput build, 3 8-CTA split kernel,
        5 the 8-CTA kernel with two points per thread and warp-shuffle stages (select before load_keys)"""
        _check(lib().bmi_ctx_set_pbs_mode(self._h, mode))

    def set_tma_stage(self, on: bool):
        """GGSW rows through TMA bulk copies into shared memory, for both key kinds (defaults: on for the pair rotation's
        latency build, where it measured 2-3 % faster; off for the one-GGSW-per-bit key, 4-8 % slower)"""
        _check(lib().bmi_ctx_set_tma_stage(self._h, int(bool(on))))

    @property
    def launch_count(self) -> int:
        return int(lib().bmi_ctx_launch_count(self._h))

    # device-pointer entry points (torch tensors own the memory)
    def lincomb(self, vals, row_ptr, idx, coef, konst, out, njobs, batch=1, stream=None):
        _check(lib().bmi_lincomb(self._h, vals.data_ptr(), row_ptr.data_ptr(), idx.data_ptr(), coef.data_ptr(),
                                 konst.data_ptr(), out.data_ptr(), njobs, batch, _stream(stream)))

    def scatter_rows(self, src, dst_row, dst, count, batch=1, stream=None):
        _check(lib().bmi_scatter_rows(self._h, src.data_ptr(), dst_row.data_ptr(), dst.data_ptr(), count, batch, _stream(stream)))

    @property
    def pbs_capacity(self) -> int:
        """ciphertexts one bootstrap launch handles at its minimum latency (0: no low-latency kernel for this set)"""
        return int(lib().bmi_ctx_pbs_capacity(self._h))

    def keyswitch(self, big, small, count, stream=None):
        _check(lib().bmi_keyswitch(self._h, big.data_ptr(), small.data_ptr(), count, _stream(stream)))

    def pbs(self, small, job_in, job_lut, job_out, out, njobs, batch=1, stream=None):
        _check(lib().bmi_pbs(self._h, small.data_ptr(), job_in.data_ptr(), job_lut.data_ptr(), job_out.data_ptr(),
                             out.data_ptr(), njobs, batch, _stream(stream)))

    # host-buffer entry points
    def ks_pbs_host(self, big_cts: np.ndarray, lut_idx) -> np.ndarray:
        big_cts = _u64(big_cts).reshape(-1, self.params.big_dim + 1)
        lut_idx = np.ascontiguousarray(lut_idx, np.int32)
        out = np.empty_like(big_cts)
        _check(lib().bmi_ks_pbs_host(self._h, _p(big_cts), _p(lut_idx), _p(out), big_cts.shape[0]))
        return out

    def polymul_host(self, a, b) -> np.ndarray:
        a, b = _u64(a).reshape(-1, self.params.N), _u64(b).reshape(-1, self.params.N)
        c = np.empty_like(a)
        _check(lib().bmi_polymul_host(self._h, _p(a), _p(b), _p(c), a.shape[0]))
        return c


def _stream(stream):
    if stream is None:
        return None
    return C.c_void_p(getattr(stream, "cuda_stream", stream))
