import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # A `gpu` test on a box without CUDA is skipped (the driver selects with -m "not gpu" here anyway).
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def vocab_dict(vocab_tokens, vocab_lens):
    """flat arrays -> the reference's f_gram_to_id dict (tuple -> id)."""
    return {tuple(int(t) for t in vocab_tokens[i, :vocab_lens[i]]): i for i in range(len(vocab_lens))}


@pytest.fixture(scope="session")
def golden():
    return load_golden
