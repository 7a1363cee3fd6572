import os
import sys
import tarfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def small_dir(tmp_path_factory):
    """The contest `small` workload (14 binary relations + init/work/result), unpacked from the
    committed fixture tests/golden/small_relations.tar.xz."""
    d = tmp_path_factory.mktemp("work") / "small"      # the init file names ./small/rN
    d.mkdir()
    with tarfile.open(os.path.join(GOLDEN, "small_relations.tar.xz")) as tf:
        tf.extractall(d)
    for name in ("small.init", "small.work", "small.result"):
        with open(os.path.join(GOLDEN, name), "rb") as src, open(os.path.join(d, name), "wb") as dst:
            dst.write(src.read())
    return str(d)


@pytest.fixture(scope="session")
def small_joins_golden():
    """94 per-join records logged by the unmodified reference (tests/golden/make_small_joins.sh)."""
    with open(os.path.join(GOLDEN, "small_joins.txt")) as f:
        return sorted(tuple(int(v) for v in line.split()) for line in f if line.strip())


@pytest.fixture(scope="session")
def edge_dir(tmp_path_factory):
    """The `edge` workload (tests/golden/make_edge.py): six synthetic relations + init/work and the result the
    unmodified reference printed for them."""
    d = tmp_path_factory.mktemp("edgework") / "edge"    # the init file names ./edge/rN
    d.mkdir()
    with tarfile.open(os.path.join(GOLDEN, "edge_relations.tar.xz")) as tf:
        tf.extractall(d)
    for name in ("edge.init", "edge.work", "edge.result"):
        with open(os.path.join(GOLDEN, name), "rb") as src, open(os.path.join(d, name), "wb") as dst:
            dst.write(src.read())
    return str(d)


@pytest.fixture(scope="session")
def edge_joins_golden():
    """per-join records the unmodified reference logged for edge.work (tests/golden/make_edge.py)."""
    with open(os.path.join(GOLDEN, "edge_joins.txt")) as f:
        return sorted(tuple(int(v) for v in line.split()) for line in f if line.strip())


@pytest.fixture(scope="session")
def engine():
    """One rhj context on cuda:0.  Fails loudly (no skip, no fallback) if the CUDA library or the GPU
    is missing -- `-m gpu` tests are only meaningful on the GPU box."""
    import torch
    from radixhashjoin_b200 import RadixHashJoin
    assert torch.cuda.is_available(), "gpu-marked tests need a CUDA device"
    e = RadixHashJoin(0)
    yield e
    e.close()
