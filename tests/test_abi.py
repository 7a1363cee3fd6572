"""CPU checks of the drop-in boundary: librhj.so loads without a GPU and exports every symbol
include/rhj.h declares; layouts match the reference's PODs; without a device the library refuses
to create a context instead of falling back to anything."""
import ctypes
import os
import re

import numpy as np
import pytest

from radixhashjoin_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rhj.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rhj_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"librhj.so does not export {n}"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_ctypes_signatures_have_the_arity_of_the_prototypes():
    """every prototype of include/rhj.h and its ctypes binding (radixhashjoin_b200/_lib.py) take the same number of
    arguments -- an argument added on one side only would shift every later one silently"""
    src = open(os.path.join(ROOT, "include", "rhj.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = dict(re.findall(r"\b(rhj_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", src))
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    for name, args in protos.items():
        args = args.strip()
        n = 0 if args in ("", "void") else args.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), f"{name}: {n} arguments in rhj.h, {len(_lib.SIGNATURES[name][1])} in _lib.SIGNATURES"


def test_pod_layouts_match_reference():
    # tuple {u64 key; u64 payload} structs.h:33-36 ; key_tuple {u64 keyR; u64 keyS} Result.h:9-12
    assert api.TUPLE_DTYPE.itemsize == 16 and api.TUPLE_DTYPE.fields["payload"][1] == 8
    assert api.PAIR_DTYPE.itemsize == 16 and api.PAIR_DTYPE.fields["keyS"][1] == 8
    assert api.PAGE_CAPACITY == 8191          # (128 KiB - 8) / 16, Result.cpp:7,11
    assert _lib.load().rhj_version() == b"0.1.0"


def test_pages_materialisation_matches_reference_format():
    """rhj_pairs_to_pages builds Result's page list (Result.cpp:21-35): newest page first, only the
    head partial, each page = [next*][8191 pairs] in 128 KiB."""
    lib = _lib.load()
    n = 8191 * 2 + 5
    pairs = np.empty(n, dtype=api.PAIR_DTYPE)
    pairs["keyR"] = np.arange(n)
    pairs["keyS"] = np.arange(n) * 3
    head_size = ctypes.c_uint64()
    head = lib.rhj_pairs_to_pages(pairs.ctypes.data, n, ctypes.byref(head_size))
    assert head and head_size.value == 5
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    seen, page, first = [], head, True
    while page:
        nxt = ctypes.c_void_p.from_address(page).value
        cnt = head_size.value if first else 8191
        buf = (ctypes.c_uint64 * (2 * cnt)).from_address(page + 8)
        seen.append(np.frombuffer(buf, dtype=api.PAIR_DTYPE).copy())
        libc.free(page)
        page, first = nxt, False
    assert [len(s) for s in seen] == [5, 8191, 8191]
    assert np.array_equal(np.concatenate(seen[::-1]), pairs)
    empty = lib.rhj_pairs_to_pages(None, 0, ctypes.byref(head_size))
    assert not empty and head_size.value == 8191   # Result(): size == capacity, head == nullptr


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.RhjError):
        api.RadixHashJoin(0)


def test_result_mirror_page_walk():
    r = api.Result(engine=object())
    r.pairs = np.zeros(8191 + 7, dtype=api.PAIR_DTYPE)
    r.pairs["keyR"] = np.arange(8191 + 7)
    pages = list(r.pages())
    assert [len(p) for p in pages] == [7, 8191] and r.size == 7 and not r.isEmpty()
    assert pages[0]["keyR"][0] == 8191
    assert api.Result(engine=object()).isEmpty()
