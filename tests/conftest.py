import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are SKIPPED (not failed) on a machine without a CUDA device, so a plain `pytest tests` on a CPU-only
    box reports the CPU suite's result."""
    try:
        from scrna_seq_qannealing_clustering_b200 import _lib
        have_gpu = _lib.load().qa_device_count() > 0
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product has no CPU path")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Native pieces are built in-tree by __graft_entry__.build() (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def gpu_ctx(built):
    from scrna_seq_qannealing_clustering_b200.engine import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
