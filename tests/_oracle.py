"""ctypes access to the CHECKERS (test infrastructure only).

* ``liborc``  -- oracle/_build/liborc.so, our C restatement (oracle/rhj_oracle.c)
* ``libref``  -- oracle/_ref/libref_rhj.so, the unmodified reference compiled from
  /root/reference by oracle/Makefile (present only if it was built in the dev container;
  it travels to the GPU box with the snapshot).

Nothing in the product package imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_SO = os.path.join(ROOT, "oracle", "_build", "liborc.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_rhj.so")
REF_T16_SO = os.path.join(ROOT, "oracle", "_ref", "libref_rhj_t16.so")      # the same sources with NUM_OF_THREADS = 16
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

TUPLE = np.dtype([("key", "<u8"), ("payload", "<u8")])      # structs.h:33-36
PAIR = np.dtype([("keyR", "<u8"), ("keyS", "<u8")])         # Result.h:9-12

_u64p = ctypes.POINTER(ctypes.c_uint64)


def build_oracle():
    """Compile oracle/rhj_oracle.c (gcc) if the .so is missing or stale."""
    src = os.path.join(ROOT, "oracle", "rhj_oracle.c")
    if (not os.path.exists(ORC_SO)) or os.path.getmtime(ORC_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return ORC_SO


_orc = None


def liborc():
    global _orc
    if _orc is None:
        lib = ctypes.CDLL(build_oracle())
        lib.orc_next_prime.restype = ctypes.c_size_t
        lib.orc_next_prime.argtypes = [ctypes.c_size_t]
        lib.orc_multi_radix_hash_join.restype = ctypes.c_int
        lib.orc_multi_radix_hash_join.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                                  ctypes.POINTER(ctypes.c_void_p), _u64p]
        lib.orc_free.argtypes = [ctypes.c_void_p]
        lib.orc_hash_relation.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p,
                                          ctypes.c_void_p]
        lib.orc_mix64.restype = ctypes.c_uint64
        lib.orc_mix64.argtypes = [ctypes.c_uint64]
        lib.orc_pairs_digest.argtypes = [ctypes.c_void_p, ctypes.c_uint64, _u64p, _u64p]
        lib.orc_filter.restype = ctypes.c_uint64
        lib.orc_filter.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p]
        lib.orc_gather_tuples.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
        lib.orc_column_sum.restype = ctypes.c_uint64
        lib.orc_column_sum.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
        _orc = lib
    return _orc


_ref = {}


def have_ref(threads=8):
    return os.path.exists(REF_SO if threads == 8 else REF_T16_SO)


def libref(threads=8):
    """the reference as shipped (8 workers per JobScheduler), or its 16-worker build (oracle/Makefile)"""
    assert threads in (8, 16)
    if threads not in _ref:
        lib = ctypes.CDLL(REF_SO if threads == 8 else REF_T16_SO)
        lib.ref_num_threads.restype = ctypes.c_int
        lib.ref_multi_radix_hash_join.restype = ctypes.c_int
        lib.ref_multi_radix_hash_join.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64,
                                                  ctypes.POINTER(ctypes.c_void_p), _u64p,
                                                  ctypes.POINTER(ctypes.c_double), ctypes.c_int]
        lib.ref_free.argtypes = [ctypes.c_void_p]
        assert lib.ref_num_threads() == threads
        _ref[threads] = lib
    return _ref[threads]


def as_tuples(keys, payloads):
    t = np.empty(len(keys), dtype=TUPLE)
    t["key"] = keys
    t["payload"] = payloads
    return t


def _take(ptr, n, free):
    if n == 0:
        return np.empty(0, dtype=PAIR)
    buf = (ctypes.c_uint64 * (2 * n)).from_address(ptr.value)
    out = np.frombuffer(buf, dtype=PAIR).copy()
    free(ptr)
    return out


def oracle_join(R, S):
    """orc_multi_radix_hash_join -> PAIR array in the reference's page-walk order."""
    R = np.ascontiguousarray(R, dtype=TUPLE)
    S = np.ascontiguousarray(S, dtype=TUPLE)
    out = ctypes.c_void_p()
    cnt = ctypes.c_uint64()
    rc = liborc().orc_multi_radix_hash_join(R.ctypes.data, len(R), S.ctypes.data, len(S),
                                            ctypes.byref(out), ctypes.byref(cnt))
    assert rc == 0
    return _take(out, cnt.value, liborc().orc_free)


def reference_join(R, S, want_pairs=True, threads=8):
    """The real reference (oracle/_ref).  Returns (pairs, seconds); pairs is None if not wanted."""
    R = np.ascontiguousarray(R, dtype=TUPLE)
    S = np.ascontiguousarray(S, dtype=TUPLE)
    out = ctypes.c_void_p()
    cnt = ctypes.c_uint64()
    sec = ctypes.c_double()
    rc = libref(threads).ref_multi_radix_hash_join(R.ctypes.data, len(R), S.ctypes.data, len(S), ctypes.byref(out),
                                            ctypes.byref(cnt), ctypes.byref(sec), 1 if want_pairs else 0)
    assert rc == 0
    if not want_pairs:
        return cnt.value, sec.value
    return _take(out, cnt.value, libref(threads).ref_free), sec.value


def oracle_partition(T, fanout=256):
    T = np.ascontiguousarray(T, dtype=TUPLE)
    out = np.empty_like(T)
    hist = np.zeros(fanout, dtype=np.uint64)
    liborc().orc_hash_relation(T.ctypes.data, len(T), fanout, out.ctypes.data, hist.ctypes.data)
    return out, hist


def mix64(x):
    """numpy splitmix64 finalizer, same as orc_mix64."""
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xbf58476d1ce4e5b9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94d049bb133111eb)
        x ^= x >> np.uint64(31)
    return x


def pairs_digest(pairs):
    """(count, sum, xor) of mix64(keyR*0x100000001b3 + keyS) -- order independent."""
    pairs = np.ascontiguousarray(pairs, dtype=PAIR)
    with np.errstate(over="ignore"):
        h = mix64(pairs["keyR"] * np.uint64(0x100000001b3) + pairs["keyS"])
        s = int(np.add.reduce(h, dtype=np.uint64)) if len(h) else 0
        x = int(np.bitwise_xor.reduce(h)) if len(h) else 0
    return len(pairs), s, x


def sort_pairs(pairs):
    pairs = np.ascontiguousarray(pairs, dtype=PAIR)
    return np.sort(pairs, order=["keyR", "keyS"])
