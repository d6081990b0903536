"""Tracing front-end: the subset of `concrete.fhe` tracing the reference exercises.

The reference's QFloat code (/root/reference/matrix_inversion/qfloat.py, base_p_arrays.py)
is ordinary numpy-style Python that runs on `fhe.tracing.tracer.Tracer` objects while
`fhe.Compiler.compile` traces it.  This module provides that Tracer.  Unlike Concrete's
tensor-level graph, every scalar of a traced tensor is kept as

  * an affine form  sum_t coef_t * base_t + const  over "base" ciphertexts (circuit inputs and
    table-lookup outputs)                                          -> leveled ops are free to compose
  * or a pending univariate function of one affine form               -> becomes ONE table lookup

so a chain like `(np.abs(curr) // base)` (qfloat.py:619) or `1 - (temp < 0)` fuses into a
single programmable bootstrap, exactly as Concrete's table-lookup fusing does, and every
lookup input reaches the engine as one sparse linear combination (one bmi_lincomb row).

Each scalar also carries its clear values on the compile-time inputset (concolic tracing):
they give the value ranges that fix the message width, as Concrete's inputset evaluation does.
"""
from __future__ import annotations

import numbers
import threading

import numpy as np

_state = threading.local()


def current():
    return getattr(_state, "trace", None)


class Trace:
    """collects the table lookups emitted while the circuit function runs"""

    def __init__(self, n_samples: int, multiplication: str = "auto"):
        """multiplication: how ciphertext * ciphertext is lowered.  "quarter_square" = two lookups
        ((a+b)^2/4 - (a-b)^2/4, Concrete's lowering); "auto" = ONE lookup on the packed value when one factor is a
        bit and the other spans at most four values on the inputset (three quarters of the reference's lookups are
        such products), quarter squares otherwise.  Same results either way."""
        if multiplication not in ("auto", "quarter_square"):
            raise ValueError("multiplication must be 'auto' or 'quarter_square'")
        self.multiplication = multiplication
        self.n_samples = n_samples
        self.n_inputs = 0
        self.jobs = []          # Job objects, index = base_id - n_inputs once inputs are frozen
        self.n_bases = 0
        self.input_ranges = []  # (min, max) of every input over the inputset

    def __enter__(self):
        if current() is not None:
            raise RuntimeError("nested tracing is not supported")
        _state.trace = self
        return self

    def __exit__(self, *exc):
        _state.trace = None

    def new_input(self, vals):
        base = self.n_bases
        self.n_bases += 1
        self.n_inputs += 1
        assert not self.jobs, "inputs must be declared before any lookup"
        vals = np.asarray(vals, dtype=np.int64)
        self.input_ranges.append((int(vals.min()), int(vals.max())))
        return Aff({base: 1}, 0, vals)

    def new_lookup(self, src: "Aff", fn, group: "Group", vals):
        base = self.n_bases
        self.n_bases += 1
        group.see(src.vals)
        self.jobs.append(Job(base, src.terms, src.const, fn, group, src.vals, vals))
        return Aff({base: 1}, 0, vals)


class Group:
    """lookups created by one tensor-level operation share one input range, like one Concrete node"""
    __slots__ = ("lo", "hi", "kids")

    def __init__(self):
        self.lo, self.hi, self.kids = None, None, None

    def child(self, name):
        """sub-node of a composite operation (e.g. the two lookups of a ciphertext product)"""
        if self.kids is None:
            self.kids = {}
        if name not in self.kids:
            self.kids[name] = Group()
        return self.kids[name]

    def see(self, vals):
        lo, hi = int(vals.min()), int(vals.max())
        self.lo = lo if self.lo is None else min(self.lo, lo)
        self.hi = hi if self.hi is None else max(self.hi, hi)


class Job:
    __slots__ = ("base", "terms", "const", "fn", "group", "src_vals", "out_vals")

    def __init__(self, base, terms, const, fn, group, src_vals, out_vals):
        self.base, self.terms, self.const, self.fn, self.group = base, terms, const, fn, group
        # source and result on every inputset sample (the lowering derives ranges of combinations from them)
        self.src_vals, self.out_vals = _compact(src_vals), _compact(out_vals)


def _compact(vals):
    v = np.asarray(vals)
    return v.astype(np.int16) if v.size and -32768 <= v.min() and v.max() <= 32767 else v.astype(np.int64)


# ------------------------------------------------------------------ scalars
class Aff:
    """affine form over base ciphertexts; `terms` empty means a known constant"""
    __slots__ = ("terms", "const", "vals")

    def __init__(self, terms, const, vals):
        self.terms, self.const, self.vals = terms, int(const), vals

    def is_const(self):
        return not self.terms


class Lazy:
    """fn(src) not yet turned into a lookup; composing more univariate steps keeps it lazy"""
    __slots__ = ("src", "fn", "vals", "group", "_aff")

    def __init__(self, src: Aff, fn, group: Group, vals=None):
        self.src, self.fn, self.group = src, fn, group
        self.vals = np.asarray(fn(src.vals), dtype=np.int64) if vals is None else vals
        self._aff = None

    def then(self, g, group):
        f = self.fn
        return Lazy(self.src, lambda x, f=f, g=g: g(f(x)), group)

    def aff(self) -> Aff:
        if self._aff is None:
            self._aff = current().new_lookup(self.src, self.fn, self.group, self.vals)
        return self._aff


def _const(c) -> Aff:
    tr = current()
    n = tr.n_samples if tr is not None else 1
    return Aff({}, int(c), np.full(n, int(c), dtype=np.int64))


def _as_scalar(x):
    """python / numpy integers stay ints; known-constant affine forms collapse to ints"""
    if isinstance(x, (Aff, Lazy)):
        if isinstance(x, Aff) and x.is_const():
            return x.const
        return x
    if isinstance(x, (numbers.Integral, np.integer, bool, np.bool_)):
        return int(x)
    if isinstance(x, (float, np.floating)) and float(x).is_integer():
        return int(x)
    raise TypeError(f"unsupported operand in an encrypted circuit: {type(x).__name__}")


def _aff(x) -> Aff:
    return x.aff() if isinstance(x, Lazy) else x


def _lin(a: Aff, b: Aff, sb: int) -> Aff:
    """a + sb * b"""
    terms = dict(a.terms)
    for k, c in b.terms.items():
        v = terms.get(k, 0) + sb * c
        if v:
            terms[k] = v
        else:
            terms.pop(k, None)
    return Aff(terms, a.const + sb * b.const, a.vals + sb * b.vals)


def _scale(a: Aff, c: int) -> Aff:
    if c == 0:
        return 0
    return Aff({k: v * c for k, v in a.terms.items()}, a.const * c, a.vals * c)


def s_add(a, b, group=None):
    a, b = _as_scalar(a), _as_scalar(b)
    if isinstance(a, int) and isinstance(b, int):
        return a + b
    if isinstance(b, int):
        a, b = b, a
    if isinstance(a, int):          # constant + symbolic
        if a == 0:
            return b
        if isinstance(b, Lazy):
            return b.then(lambda x, c=a: x + c, group or Group())
        return Aff(b.terms, b.const + a, b.vals + a)
    return _lin(_aff(a), _aff(b), 1)


def s_neg(a, group=None):
    a = _as_scalar(a)
    if isinstance(a, int):
        return -a
    if isinstance(a, Lazy):
        return a.then(lambda x: -x, group or Group())
    return _scale(a, -1)


def s_sub(a, b, group=None):
    a, b = _as_scalar(a), _as_scalar(b)
    if isinstance(b, int):
        return s_add(a, -b, group)
    if isinstance(a, int):
        return s_add(s_neg(b, group), a, group)
    return _lin(_aff(a), _aff(b), -1)


def _sq4(x):
    return (x * x) // 4


def s_mul(a, b, group=None):
    a, b = _as_scalar(a), _as_scalar(b)
    if isinstance(a, int) and isinstance(b, int):
        return a * b
    if isinstance(b, int):
        a, b = b, a
    if isinstance(a, int):          # plaintext scalar times ciphertext: leveled
        if a == 0:
            return 0
        if a == 1:
            return b
        if isinstance(b, Lazy):
            return b.then(lambda x, c=a: x * c, group or Group())
        return _scale(b, a)
    g = group or Group()
    fused = _fuse_same_source(a, b, lambda u, v: u * v, g)
    if fused is not None:
        return fused
    a, b = _aff(a), _aff(b)
    packed = _packed_product(a, b, g)
    if packed is not None:
        return packed
    # ciphertext * ciphertext = ((a+b)^2 - (a-b)^2) / 4: two table lookups, as Concrete lowers it
    plus = Lazy(_lin(a, b, 1), _sq4, g.child("sum")).aff()
    minus = Lazy(_lin(a, b, -1), _sq4, g.child("difference")).aff()
    return _lin(plus, minus, -1)


def _same_form(x: Aff, y: Aff) -> bool:
    return x is y or (x.const == y.const and x.terms == y.terms)


def _fuse_same_source(a, b, op, g: Group):
    """op(a, b) when both operands are functions of ONE affine form (two pending lookups on the same input, or a
    pending lookup and its own input): still a univariate function, so one lookup -- what Concrete's fusing does with
    `(np.abs(curr) // base) * np.sign(curr)` (qfloat.py:619) or `array * (array >= 0)` (qfloat.py:663)."""
    ident = lambda x: x
    if isinstance(a, Lazy) and isinstance(b, Lazy) and _same_form(a.src, b.src):
        src, fa, fb = a.src, a.fn, b.fn
    elif isinstance(a, Lazy) and isinstance(b, Aff) and _same_form(a.src, b):
        src, fa, fb = a.src, a.fn, ident
    elif isinstance(b, Lazy) and isinstance(a, Aff) and _same_form(b.src, a):
        src, fa, fb = b.src, ident, b.fn
    else:
        return None
    return Lazy(src, lambda x, fa=fa, fb=fb: op(np.asarray(fa(x)), np.asarray(fb(x))), g)


_PACK_SPAN = 8      # packed products stay within 3 message bits, below every circuit's width


def _packed_product(a: Aff, b: Aff, g: Group):
    """bit * small value as ONE lookup on x = (v - lo) + bit * R, R = span of v: f(x) = x >= R ? x - R + lo : 0"""
    tr = current()
    if tr is None or tr.multiplication != "auto":
        return None
    is_bit = lambda t: t.vals.min() >= 0 and t.vals.max() <= 1
    if is_bit(b):
        v, bit = a, b
    elif is_bit(a):
        v, bit = b, a
    else:
        return None
    lo, hi = int(v.vals.min()), int(v.vals.max())
    if 2 * (hi - lo + 1) > _PACK_SPAN:
        return None
    if 2 * (hi - lo + 3) <= _PACK_SPAN:        # room for a guard value on each side of the observed range
        lo, hi = lo - 1, hi + 1
    R = hi - lo + 1
    x = _lin(Aff(v.terms, v.const - lo, v.vals - lo), _scale(bit, R), 1)
    grp = g.child("packed")
    grp.see(np.array([0, 2 * R - 1]))           # the declared domain, not only what the inputset happened to hit
    return Lazy(x, lambda t, R=R, lo=lo: np.where(t >= R, t - R + lo, 0), grp).aff()


def s_univariate(a, fn, group=None):
    """fn: vectorised int -> int"""
    a = _as_scalar(a)
    if isinstance(a, int):
        return int(np.asarray(fn(np.asarray([a], dtype=np.int64)))[0])
    g = group or Group()
    if isinstance(a, Lazy):
        return a.then(fn, g)
    return Lazy(a, fn, g)


def s_compare(a, b, op, group=None):
    a, b = _as_scalar(a), _as_scalar(b)
    if isinstance(a, int) and isinstance(b, int):
        return int(op(a, b))
    if isinstance(b, int):
        return s_univariate(a, lambda x, c=b: op(x, c).astype(np.int64), group)
    if isinstance(a, int):
        return s_univariate(b, lambda x, c=a: op(c, x).astype(np.int64), group)
    fused = _fuse_same_source(a, b, lambda u, v: op(u, v).astype(np.int64), group or Group())
    if fused is not None:
        return fused
    return s_univariate(_lin(_aff(a), _aff(b), -1), lambda x: op(x, 0).astype(np.int64), group)


def _bits(v: int) -> int:
    return max(1, int(v).bit_length())


def s_bitwise(a, b, op, group=None):
    a, b = _as_scalar(a), _as_scalar(b)
    if isinstance(a, int) and isinstance(b, int):
        return int(op(a, b))
    if isinstance(b, int):
        return s_univariate(a, lambda x, c=b: op(x, c), group)
    if isinstance(a, int):
        return s_univariate(b, lambda x, c=a: op(c, x), group)
    fused = _fuse_same_source(a, b, lambda u, v: op(u, v), group or Group())
    if fused is not None:
        return fused
    if min(a.vals.min(), b.vals.min()) < 0:
        raise NotImplementedError("bitwise operations between signed encrypted values")
    sh = _bits(b.vals.max())        # pack both operands into one lookup input, as Concrete's chunked bitwise does
    packed = _lin(_scale(_aff(a), 1 << sh), _aff(b), 1)
    return s_univariate(packed, lambda x, sh=sh: op(x >> sh, x & ((1 << sh) - 1)), group)


# ------------------------------------------------------------------ tensors
def _obj(x):
    """anything -> numpy object array of scalars"""
    if isinstance(x, Tracer):
        return x.arr
    a = np.asarray(x)
    if a.dtype != object:
        if a.dtype.kind == "f":
            if not np.all(a == np.floor(a)):
                raise TypeError("non-integer constant in an encrypted circuit")
            a = a.astype(np.int64)
        a = a.astype(object)
    return a


def _elementwise(fn, *arrays):
    group = Group()
    out = np.frompyfunc(lambda *xs: fn(*xs, group), len(arrays), 1)(*[_obj(a) for a in arrays])
    return Tracer(out if isinstance(out, np.ndarray) else np.array(out, dtype=object))


class Tracer:
    """symbolic integer tensor (encrypted once compiled)"""
    __array_priority__ = 1000

    def __init__(self, arr):
        if not isinstance(arr, np.ndarray) or arr.dtype != object:
            arr = np.array(arr, dtype=object)
        self.arr = arr

    # ---- shape protocol
    @property
    def shape(self):
        return self.arr.shape

    @property
    def size(self):
        return self.arr.size

    @property
    def ndim(self):
        return self.arr.ndim

    @property
    def T(self):
        return Tracer(self.arr.T)

    def __len__(self):
        return self.arr.shape[0]

    def reshape(self, *shape):
        if len(shape) == 1 and not isinstance(shape[0], numbers.Integral):
            shape = tuple(shape[0])
        return Tracer(self.arr.reshape(shape))

    def flatten(self):
        return Tracer(self.arr.flatten())

    def astype(self, _dtype):
        return self

    def copy(self):
        return Tracer(self.arr.copy())

    def __getitem__(self, key):
        out = self.arr[key]
        if isinstance(out, np.ndarray):
            return Tracer(out.copy())        # Concrete's indexing yields a new value, never a view
        return Tracer(np.array(out, dtype=object))

    def __setitem__(self, key, value):
        v = _obj(value)
        self.arr[key] = v.item() if v.ndim == 0 else v

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __bool__(self):
        raise TypeError("the truth value of an encrypted tensor is not known at trace time")

    # ---- arithmetic
    def __add__(self, o): return _elementwise(s_add, self, o)
    def __radd__(self, o): return _elementwise(s_add, o, self)
    def __sub__(self, o): return _elementwise(s_sub, self, o)
    def __rsub__(self, o): return _elementwise(s_sub, o, self)
    def __mul__(self, o): return _elementwise(s_mul, self, o)
    def __rmul__(self, o): return _elementwise(s_mul, o, self)
    def __neg__(self): return _elementwise(s_neg, self)
    def __pos__(self): return self
    def __abs__(self): return _uni(self, np.abs)

    def __floordiv__(self, o):
        return _binary_const(self, o, lambda x, c: x // c, "//")

    def __mod__(self, o):
        return _binary_const(self, o, lambda x, c: x % c, "%")

    def __lshift__(self, o):
        return _binary_const(self, o, lambda x, c: x << c, "<<")

    def __rshift__(self, o):
        return _binary_const(self, o, lambda x, c: x >> c, ">>")

    def __truediv__(self, o):
        raise TypeError("true division of encrypted integers is not supported; use //")

    # ---- comparisons
    def __lt__(self, o): return _cmp(self, o, np.less)
    def __le__(self, o): return _cmp(self, o, np.less_equal)
    def __gt__(self, o): return _cmp(self, o, np.greater)
    def __ge__(self, o): return _cmp(self, o, np.greater_equal)
    def __eq__(self, o): return _cmp(self, o, np.equal)
    def __ne__(self, o): return _cmp(self, o, np.not_equal)
    __hash__ = None

    # ---- bitwise
    def __and__(self, o): return _bit(self, o, np.bitwise_and)
    def __rand__(self, o): return _bit(o, self, np.bitwise_and)
    def __or__(self, o): return _bit(self, o, np.bitwise_or)
    def __ror__(self, o): return _bit(o, self, np.bitwise_or)
    def __xor__(self, o): return _bit(self, o, np.bitwise_xor)
    def __rxor__(self, o): return _bit(o, self, np.bitwise_xor)

    # ---- numpy protocols
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method == "__call__" and not kwargs:
            if ufunc in _UFUNC_BINARY and len(inputs) == 2:
                return _UFUNC_BINARY[ufunc](*inputs)
            if ufunc in _UFUNC_UNARY and len(inputs) == 1:
                return _UFUNC_UNARY[ufunc](inputs[0])
        if method == "reduce" and ufunc is np.add:
            return _sum(inputs[0], kwargs.get("axis", 0))
        return NotImplemented

    def __array_function__(self, func, types, args, kwargs):
        if func in _ARRAY_FUNCTIONS:
            return _ARRAY_FUNCTIONS[func](*args, **kwargs)
        return NotImplemented


def _uni(x, fn):
    return _elementwise(lambda a, group: s_univariate(a, fn, group), x)


def _binary_const(x, c, fn, name):
    if isinstance(c, Tracer):
        raise NotImplementedError(f"{name} with an encrypted right operand")
    c = np.asarray(c)
    if c.ndim == 0:
        cv = int(c)
        return _uni(x, lambda v, cv=cv: fn(v, cv))
    return _elementwise(lambda a, b, group: s_univariate(a, lambda v, b=int(b): fn(v, b), group), x, c)


def _cmp(a, b, op):
    return _elementwise(lambda x, y, group: s_compare(x, y, op, group), a, b)


def _bit(a, b, op):
    return _elementwise(lambda x, y, group: s_bitwise(x, y, op, group), a, b)


def _sum(x, axis=None, **_):
    arr = _obj(x)
    if axis is None:
        flat = arr.reshape(-1)
        out = 0
        for v in flat:
            out = s_add(out, v)
        return Tracer(np.array(out, dtype=object))
    moved = np.moveaxis(arr, axis, 0)
    out = np.empty(moved.shape[1:], dtype=object)
    flat_in = moved.reshape(moved.shape[0], -1)
    flat_out = out.reshape(-1)
    for j in range(flat_in.shape[1]):
        acc = 0
        for v in flat_in[:, j]:
            acc = s_add(acc, v)
        flat_out[j] = acc
    return Tracer(out)


def _concatenate(arrays, axis=0, **_):
    return Tracer(np.concatenate([np.atleast_1d(_obj(a)) for a in arrays], axis=axis))


def _reshape(a, newshape=None, shape=None, **_):
    return Tracer(_obj(a).reshape(newshape if newshape is not None else shape))


_UFUNC_BINARY = {
    np.add: lambda a, b: _elementwise(s_add, a, b),
    np.subtract: lambda a, b: _elementwise(s_sub, a, b),
    np.multiply: lambda a, b: _elementwise(s_mul, a, b),
    np.less: lambda a, b: _cmp(a, b, np.less),
    np.less_equal: lambda a, b: _cmp(a, b, np.less_equal),
    np.greater: lambda a, b: _cmp(a, b, np.greater),
    np.greater_equal: lambda a, b: _cmp(a, b, np.greater_equal),
    np.equal: lambda a, b: _cmp(a, b, np.equal),
    np.not_equal: lambda a, b: _cmp(a, b, np.not_equal),
    np.bitwise_and: lambda a, b: _bit(a, b, np.bitwise_and),
    np.bitwise_or: lambda a, b: _bit(a, b, np.bitwise_or),
    np.bitwise_xor: lambda a, b: _bit(a, b, np.bitwise_xor),
    np.floor_divide: lambda a, b: a.__floordiv__(b) if isinstance(a, Tracer) else NotImplemented,
    np.remainder: lambda a, b: a.__mod__(b) if isinstance(a, Tracer) else NotImplemented,
}
_UFUNC_UNARY = {
    np.absolute: lambda a: _uni(a, np.abs),
    np.sign: lambda a: _uni(a, np.sign),
    np.negative: lambda a: _elementwise(s_neg, a),
    np.positive: lambda a: a,
    np.square: lambda a: _elementwise(s_mul, a, a),
}
_ARRAY_FUNCTIONS = {
    np.sum: _sum,
    np.concatenate: _concatenate,
    np.reshape: _reshape,
    np.abs: lambda a: _uni(a, np.abs),
    np.sign: lambda a: _uni(a, np.sign),
    np.copy: lambda a, **_: Tracer(_obj(a).copy()),
    np.flip: lambda a, axis=None: Tracer(np.flip(_obj(a), axis=axis)),
    np.transpose: lambda a, axes=None: Tracer(np.transpose(_obj(a), axes)),
    np.where: lambda c, a, b: _elementwise(s_add, _elementwise(s_mul, c, a),
                                           _elementwise(s_mul, _elementwise(s_sub, 1, c), b)),
}


# ------------------------------------------------------------ module level API
def _shape(shape):
    return (shape,) if isinstance(shape, numbers.Integral) else tuple(shape)


def zeros(shape):
    """fhe.zeros: an encrypted tensor of zeros while tracing, a plain int array otherwise
    (reference: base_p_arrays.py:97, qfloat.py:391)"""
    if current() is None:
        return np.zeros(_shape(shape), dtype=np.int64)
    out = np.empty(_shape(shape), dtype=object)
    out.fill(0)
    return Tracer(out)


def ones(shape):
    if current() is None:
        return np.ones(_shape(shape), dtype=np.int64)
    out = np.empty(_shape(shape), dtype=object)
    out.fill(1)
    return Tracer(out)


def zero():
    return zeros(())


def one():
    return ones(())


def univariate(function):
    """fhe.univariate(f)(x): one table lookup computing f elementwise (reference: base_p_arrays.py:365)"""
    def apply(x):
        if not isinstance(x, Tracer):
            return function(x)
        return _uni(x, lambda v: np.asarray(function(v), dtype=np.int64))
    return apply
