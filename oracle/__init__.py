"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front end of oracle/build/liboracle.so: the C restatement of the
reference's forward-BWT path (/root/reference/src/saca.rs:22-384 +
`compress::bwt::TransformIterator`, call sites src/block/dc.rs:45-50,
src/block/raw.rs:39-44), the SURVEY.md App. D generators and the §8(d) LCP
profiler.  Parity is PINNED by the reference's known-answer test
(saca.rs:409-413): see tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package — as the checker or as the timed
CPU baseline, never as the product.  dark_b200/ never imports it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liboracle.so")
_lib = None

MAX_DEPTH = 32


class Trace(ctypes.Structure):
    _fields_ = [("depth", ctypes.c_int),
                ("n1", ctypes.c_uint64 * MAX_DEPTH),
                ("names", ctypes.c_uint64 * MAX_DEPTH)]


class Profile(ctypes.Structure):
    _fields_ = [("b", ctypes.c_uint32), ("P", ctypes.c_uint32), ("R", ctypes.c_uint32),
                ("m", ctypes.c_uint64 * 64), ("sum_m", ctypes.c_uint64),
                ("max_lcp", ctypes.c_uint64), ("mean_lcp", ctypes.c_double),
                ("b_alg", ctypes.c_double)]


class OracleError(RuntimeError):
    """One of the reference's assert!/panic paths would have fired."""


def build():
    """Compile the oracle (gcc, a few seconds).  Building the checker is not using it."""
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p, u32p, u64 = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64
        L.oracle_arena_words.restype = u64
        L.oracle_arena_words.argtypes = [u64]
        L.oracle_saca.argtypes = [u8p, u64, u32p, ctypes.POINTER(Trace)]
        L.oracle_saca_arena.argtypes = [u8p, u64, u32p, u64, ctypes.POINTER(Trace)]
        L.oracle_bwt_emit.argtypes = [u8p, u64, u32p, u8p, ctypes.POINTER(u64)]
        L.oracle_bwt_emit.restype = None
        L.oracle_bwt_forward.argtypes = [u8p, u64, u8p, ctypes.POINTER(u64), u32p]
        L.oracle_sort_direct.argtypes = [u8p, u64, u32p]
        L.oracle_sort_direct.restype = None
        L.oracle_bwt_decode.argtypes = [u8p, u64, u64, u8p]
        L.oracle_verify_sa.argtypes = [u8p, u64, u32p]
        L.oracle_sm64.restype = u64
        L.oracle_sm64.argtypes = [u64]
        L.oracle_gen.argtypes = [ctypes.c_char_p, u64, u8p, u64]
        L.oracle_profile_lcp.argtypes = [u8p, u64, u32p, ctypes.POINTER(Profile)]
        _lib = L
    return _lib


def _u8(a):
    a = np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a,
                             dtype=np.uint8)
    return a


def _check(rc, what):
    if rc != 0:
        raise OracleError(f"{what}: oracle error {rc} (the reference would panic here)")


def gen(kind, seed, n):
    """SURVEY App. D generator -> np.uint8[n]."""
    out = np.empty(n, dtype=np.uint8)
    if lib().oracle_gen(kind.encode(), seed, out.ctypes.data, n):
        raise ValueError(f"unknown generator {kind!r}")
    return out


def saca(text, trace=False):
    """saca::Constructor::new(len).compute(text) -> np.uint32[n]  (saca.rs:351-378)."""
    t = _u8(text)
    sa = np.empty(t.size, dtype=np.uint32)
    tr = Trace()
    _check(lib().oracle_saca(t.ctypes.data, t.size, sa.ctypes.data, ctypes.byref(tr)), "saca")
    if trace:
        return sa, [(tr.n1[i], tr.names[i]) for i in range(tr.depth)]
    return sa


def bwt_emit(text, sa):
    """TransformIterator -> (bwt bytes np.uint8[n], origin)."""
    t = _u8(text)
    s = np.ascontiguousarray(sa, dtype=np.uint32)
    out = np.empty(t.size, dtype=np.uint8)
    origin = ctypes.c_uint64(0)
    lib().oracle_bwt_emit(t.ctypes.data, t.size, s.ctypes.data, out.ctypes.data, ctypes.byref(origin))
    return out, origin.value


def bwt_forward(text, want_sa=False):
    """block/dc.rs:45-50: (bwt, origin[, sa])."""
    t = _u8(text)
    out = np.empty(t.size, dtype=np.uint8)
    sa = np.empty(t.size, dtype=np.uint32) if want_sa else None
    origin = ctypes.c_uint64(0)
    _check(lib().oracle_bwt_forward(t.ctypes.data, t.size, out.ctypes.data, ctypes.byref(origin),
                                    sa.ctypes.data if want_sa else None), "bwt_forward")
    return (out, origin.value, sa) if want_sa else (out, origin.value)


def sort_direct(text):
    """saca.rs:25-35, the reference's own specification (brute force)."""
    t = _u8(text)
    sa = np.empty(t.size, dtype=np.uint32)
    lib().oracle_sort_direct(t.ctypes.data, t.size, sa.ctypes.data)
    return sa


def bwt_decode(bwt, origin):
    b = _u8(bwt)
    out = np.empty(b.size, dtype=np.uint8)
    _check(lib().oracle_bwt_decode(b.ctypes.data, b.size, origin, out.ctypes.data), "bwt_decode")
    return out


def verify_sa(text, sa):
    t = _u8(text)
    s = np.ascontiguousarray(sa, dtype=np.uint32)
    return lib().oracle_verify_sa(t.ctypes.data, t.size, s.ctypes.data) == 0


def profile(text, sa):
    """dict with b, P, R, m[r], sum_m, b_alg."""
    t = _u8(text)
    s = np.ascontiguousarray(sa, dtype=np.uint32)
    p = Profile()
    _check(lib().oracle_profile_lcp(t.ctypes.data, t.size, s.ctypes.data, ctypes.byref(p)), "profile")
    return {"b": p.b, "P": p.P, "R": p.R, "m": [p.m[i] for i in range(p.R)], "sum_m": p.sum_m,
            "max_lcp": p.max_lcp, "mean_lcp": p.mean_lcp, "b_alg": p.b_alg}


class Arena:
    """A reusable `Constructor` (arena allocated once, like Constructor::new) for timing loops."""

    def __init__(self, max_n):
        self.n = max_n
        self.words = lib().oracle_arena_words(max_n)
        self.arena = np.zeros(self.words, dtype=np.uint32)
        self.bwt = np.empty(max_n, dtype=np.uint8)

    def bwt_forward(self, text):
        t = _u8(text)
        assert t.size == self.n
        _check(lib().oracle_saca_arena(t.ctypes.data, t.size, self.arena.ctypes.data, self.words, None), "saca")
        origin = ctypes.c_uint64(0)
        lib().oracle_bwt_emit(t.ctypes.data, t.size, self.arena.ctypes.data, self.bwt.ctypes.data,
                              ctypes.byref(origin))
        return self.bwt, origin.value


# ---- distance coding + MTF (dc_oracle.c): PARITY UNPINNED, restated from memory of upstream rust-compress ------------
def dc_encode(bwt):
    """`bwt::dc::encode(&output, suf, &mut mtf)`: (distances u32[n] with filler n, init[256], mtf_symbols[256], num_unique)."""
    b = _u8(bwt)
    L = lib()
    dist = np.empty(b.size, dtype=np.uint32)
    init = (ctypes.c_uint64 * 256)()
    mtf = (ctypes.c_uint8 * 256)()
    nu = ctypes.c_uint32(0)
    L.oracle_dc_encode.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32)]
    _check(L.oracle_dc_encode(b.ctypes.data, b.size, dist.ctypes.data, init, mtf, ctypes.byref(nu)), "dc::encode")
    return dist, np.array(list(init), dtype=np.uint64), np.array(list(mtf), dtype=np.uint8), int(nu.value)


def dc_stream(bwt, dist, init):
    """The (distance, Context) items the reference's EncodeIterator yields: (pos, dist, symbol, last_rank) arrays;
    distance_limit = n - pos."""
    b = _u8(bwt)
    L = lib()
    n = b.size
    pos, d = np.empty(n, dtype=np.uint32), np.empty(n, dtype=np.uint32)
    sym, rk = np.empty(n, dtype=np.uint8), np.empty(n, dtype=np.uint8)
    ini = (ctypes.c_uint64 * 256)(*[int(x) for x in init])
    L.oracle_dc_stream.restype = ctypes.c_uint64
    L.oracle_dc_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p] + [ctypes.c_void_p] * 4
    dd = np.ascontiguousarray(dist, dtype=np.uint32)
    c = L.oracle_dc_stream(b.ctypes.data, dd.ctypes.data, n, ini, pos.ctypes.data, d.ctypes.data, sym.ctypes.data, rk.ctypes.data)
    return pos[:c].copy(), d[:c].copy(), sym[:c].copy(), rk[:c].copy()


def dc_decode(init, stream_dist, n):
    """`bwt::dc::decode(init, &mut out, &mut mtf, |ctx| next distance)` collected."""
    L = lib()
    out = np.empty(n, dtype=np.uint8)
    ini = (ctypes.c_uint64 * 256)(*[int(x) for x in init])
    sd = np.ascontiguousarray(stream_dist, dtype=np.uint32)
    L.oracle_dc_decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64]
    _check(L.oracle_dc_decode(ini, sd.ctypes.data, sd.size, out.ctypes.data, n), "dc::decode")
    return out
