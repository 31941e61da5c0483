import os
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native():
    from seismic_b200 import _native
    _native.lib()
    return _native


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()
    return oracle


def build_synth(n_docs, n_queries, dim=30522, seed=20260517, **build_params):
    from seismic_b200 import Dataset, HostIndex
    cfg = Dataset.synth_config(n_docs, dim=dim, seed=seed)
    docs = Dataset.synth_documents(cfg)
    queries = Dataset.synth_queries(cfg, n_queries)
    index = HostIndex.build(docs, **build_params)
    return docs, queries, index


@pytest.fixture(scope="session")
def synth_small():
    """20k docs, 300 queries: every list is short (nnz < dim*n_postings: nothing pruned)."""
    return build_synth(20000, 300)


@pytest.fixture(scope="session")
def synth_pruned():
    """30k docs over a 2k vocabulary with small n_postings: pruning, caps and many blocks per list are exercised."""
    return build_synth(30000, 300, dim=2000, n_postings=600, centroid_fraction=0.2, summary_energy=0.5, max_fraction=2.0)
