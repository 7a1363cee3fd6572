"""radixhashjoin_b200 -- B200-native radix hash join behind the reference's C++ surface.

The product is ``librhj.so`` (hand-written sm_100a CUDA kernels + the C ABI of ``include/rhj.h``)
and the C++ drop-in ``host/Result.cpp``.  This Python package is the thin host-side mirror used by
the tests and ``bench.py``: PyTorch supplies device memory, streams and ``torch.distributed``; all
compute goes through the C ABI.  Names follow the reference (relation / Result / row ids), see
``api.py``.
"""
from .api import (RadixHashJoin, RhjError, Relation, Result, TUPLE_DTYPE, PAIR_DTYPE, EMIT_FUSED,
                  EMIT_COUNT_THEN_WRITE, DIGIT_RAW, DIGIT_HASH)

__all__ = ["RadixHashJoin", "RhjError", "Relation", "Result", "TUPLE_DTYPE", "PAIR_DTYPE", "EMIT_FUSED",
           "EMIT_COUNT_THEN_WRITE", "DIGIT_RAW", "DIGIT_HASH"]
__version__ = "0.1.0"
