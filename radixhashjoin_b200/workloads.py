"""Deterministic synthetic join workloads (BASELINE.json configs 2-5).

Integer-only generators written with plain torch ops so that the same code runs on the CPU (oracle-
sized parity tests) and on the GPU (full-size benchmark inputs), chunked to bound temporary memory.
Every workload knows its expected result in closed form, so a full-size run can be verified by
count + an order-independent digest without materialising an oracle result:

    digest = (count, sum, xor) over pairs of mix64(rowidR * 0x100000001b3 + rowidS)   (mod 2^64)

Relations are int64 tensors of shape (n, 2) = (row id, value) holding the bit patterns of the
reference's u64 fields (structs.h:33-36; the value is `payload`).
"""
import torch

SEED = 42
A_MUL = 0x9E3779B97F4A7C15  # odd => j -> (j*A + B) mod 2^k is a bijection
B_ADD = 12345
_M64 = (1 << 64) - 1
_CHUNK = 1 << 24


def _s64(x):
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def lsr(x, k):
    """logical shift right of int64 bit patterns"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def mix64(x):
    """splitmix64 finalizer on int64 bit patterns (== oracle orc_mix64)."""
    x = x ^ lsr(x, 30)
    x = x * _s64(0xbf58476d1ce4e5b9)
    x = x ^ lsr(x, 27)
    x = x * _s64(0x94d049bb133111eb)
    x = x ^ lsr(x, 31)
    return x


def pair_hash(r, s):
    return mix64(r * _s64(0x100000001b3) + s)


def _fill(n, device, fn, first=0):
    """out[(n,2)] with out[i] = (first + i, fn(first + i)), chunked."""
    out = torch.empty((n, 2), dtype=torch.int64, device=device)
    for lo in range(0, n, _CHUNK):
        hi = min(n, lo + _CHUNK)
        idx = torch.arange(first + lo, first + hi, dtype=torch.int64, device=device)
        out[lo:hi, 0] = idx
        out[lo:hi, 1] = fn(idx)
    return out


def _shard(n, rank, world):
    """rows [lo, hi) of an n-row relation that rank `rank` of `world` owns before the join"""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def _xor_reduce(h):
    """xor of all elements (torch has no bitwise_xor reduction): fold in halves."""
    while h.numel() > 1:
        if h.numel() & 1:
            h = torch.cat([h, torch.zeros(1, dtype=torch.int64, device=h.device)])
        half = h.numel() // 2
        h = h[:half] ^ h[half:]
    return h[0]


def _digest_range(lo, hi, device, row_of):
    """digest of {(row_of(j), j) : lo <= j < hi}"""
    s, x = 0, 0
    for c0 in range(lo, hi, _CHUNK):
        c1 = min(hi, c0 + _CHUNK)
        j = torch.arange(c0, c1, dtype=torch.int64, device=device)
        h = pair_hash(row_of(j), j)
        s = (s + int(h.sum().item())) & _M64
        x ^= int(_xor_reduce(h).item()) & _M64
    return hi - lo, s, x


def _digest_of(n, device, row_of):
    return _digest_range(0, n, device, row_of)


class Workload:
    def __init__(self, name, R, S, expected_digest, description):
        self.name = name
        self.R = R
        self.S = S
        self.expected = expected_digest  # (count, sum, xor)
        self.description = description


def uniform_unique(log2n, device="cpu", seed=SEED, row_offset=0, log2_global=None):
    """Config 2 / 5: unique u64 keys on both sides, exactly one match per probe tuple.
    R[i] = (i, mix64(i+seed)); S[j] = (j, mix64(pi(j)+seed)), pi(j) = (j*A + B) mod N.
    With row_offset/log2_global the call generates rows [row_offset, row_offset + 2^log2n) of a
    global 2^log2_global relation pair (multi-GPU sharding; pi is taken mod the global size)."""
    n = 1 << log2n
    gbits = log2_global if log2_global is not None else log2n
    mask = (1 << gbits) - 1
    a, b = _s64(A_MUL), B_ADD

    def pi(j):
        return (j * a + b) & mask

    R = _fill(n, device, lambda i: mix64(i + row_offset + seed))
    S = _fill(n, device, lambda j: mix64(pi(j + row_offset) + seed))
    if row_offset:
        R[:, 0] += row_offset
        S[:, 0] += row_offset
    exp = _digest_of(n, device, lambda j: pi(j + row_offset)) if not row_offset and gbits == log2n else None
    return Workload(f"uniform_unique_2^{log2n}x2^{log2n}", R, S, exp,
                    "unique uniform u64 keys, 1:1 match (BASELINE config 2)")


def uniform_unique_global_digest(log2_global, device="cpu", lo=0, hi=None):
    """digest of {(pi(j), j) : lo <= j < hi} of the global uniform workload: the expected result of
    the probe rows [lo, hi); per-rank digests combine by (sum of counts, sum mod 2^64, xor)."""
    n = 1 << log2_global
    hi = n if hi is None else hi
    a, b = _s64(A_MUL), B_ADD
    return _digest_range(lo, hi, device, lambda j: (j * a + b) & (n - 1))


def foreign_key(log2_build, log2_probe, device="cpu", seed=SEED, rank=0, world=1):
    """Config 3: unique build keys, every probe tuple references one build row.
    S[j] = (j, mix64(row(j)+seed)), row(j) = (mix64(j ^ 0xabcdef) >> 1) mod N_R.
    With world > 1 the call returns rank `rank`'s row ranges of both relations and the digest of what ITS probe
    rows produce (the global expectation is the (sum, sum, xor) of the ranks' digests)."""
    nR, nS = 1 << log2_build, 1 << log2_probe

    def row(j):
        return lsr(mix64(j ^ 0xabcdef), 1) & (nR - 1)

    r0, r1 = _shard(nR, rank, world)
    s0, s1 = _shard(nS, rank, world)
    R = _fill(r1 - r0, device, lambda i: mix64(i + seed), r0)
    S = _fill(s1 - s0, device, lambda j: mix64(row(j) + seed), s0)
    return Workload(f"fk_2^{log2_build}x2^{log2_probe}", R, S, _digest_range(s0, s1, device, row),
                    "foreign-key join, unique build keys (BASELINE config 3)")


def zipf_probe(log2n, device="cpu", seed=SEED, rank=0, world=1):
    """Config 4: unique build keys, probe keys ~ Zipf(theta = 1) by an integer-only inverse CDF:
    k = (mix64(j^0x5151) >> 8) mod log2n, rank r = 2^k + (mix64(j^0x7171) & (2^k - 1)) in [1, N-1]
    => P(r) = 1 / (log2n * 2^floor(log2 r)) ~ 1/r; the hottest key draws 1/log2n of the probe side.
    build row = ((r-1) * A) mod N.  log2n is the GLOBAL size; world > 1 returns rank `rank`'s row ranges
    and the digest of what its probe rows produce."""
    n = 1 << log2n
    a = _s64(A_MUL)

    def row(j):
        k = lsr(mix64(j ^ 0x5151), 8) % log2n
        r = (1 << k) + (mix64(j ^ 0x7171) & ((1 << k) - 1))
        return ((r - 1) * a) & (n - 1)

    lo, hi = _shard(n, rank, world)
    R = _fill(hi - lo, device, lambda i: mix64(i + seed), lo)
    S = _fill(hi - lo, device, lambda j: mix64(row(j) + seed), lo)
    return Workload(f"zipf_2^{log2n}x2^{log2n}", R, S, _digest_range(lo, hi, device, row),
                    "Zipf(1.0) probe keys over unique build keys (BASELINE config 4)")


def duplicates(nR, nS, domain, device="cpu", seed=SEED):
    """N:M join like the contest columns (505-9741 distinct values per column):
    values = mix64(i + seed) mod domain on both sides.  No closed form; compare with the oracle."""
    R = _fill(nR, device, lambda i: lsr(mix64(i + seed), 1) % domain)
    S = _fill(nS, device, lambda j: lsr(mix64(j + seed + 0x1234567), 1) % domain)
    return Workload(f"dup_{nR}x{nS}_dom{domain}", R, S, None, "duplicate-heavy N:M join")


def to_numpy_tuples(t):
    """(n,2) int64 CPU tensor -> numpy structured TUPLE array (zero copy)."""
    import numpy as np
    from .api import TUPLE_DTYPE
    return t.contiguous().numpy().view(np.uint64).reshape(-1, 2).view(TUPLE_DTYPE).reshape(-1)
