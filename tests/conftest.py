import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _bounds_checked_build_is_clean():
    """With the bounds-checked library (KMB_LIB_PATH=...libkmer_mapper_b200_bounds.so, _build.py) the whole GPU
    session must end without a single device-side index out of range.  The product build reports -1: no checks."""
    yield
    try:
        import torch
        if not torch.cuda.is_available():
            return
        from kmer_mapper_b200 import _lib
        failures = _lib.get_option("bounds_failures")
    except Exception:
        return
    assert failures in (-1, 0), "%d device-side bounds checks failed (sites on stderr)" % failures


class GoldenIndex:
    """Duck-typed index (the six attributes mapper.pyx:22-29 reads) rebuilt from a golden fixture."""

    def __init__(self, z, name):
        self._hashes_to_index = z[name + "/hashes_to_index"]
        self._n_kmers = z[name + "/n_kmers"]
        self._nodes = z[name + "/nodes"]
        self._kmers = z[name + "/kmers"]
        self._frequencies = z[name + "/frequencies"]
        self._modulo = int(z[name + "/modulo"])

    def max_node_id(self):
        return int(self._nodes.max())


def golden_lookup_cases():
    z = np.load(os.path.join(GOLDEN, "golden_lookup.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


@pytest.fixture(scope="session")
def golden_lookup():
    z, names = golden_lookup_cases()
    out = {}
    for n in names:
        out[n] = dict(index=GoldenIndex(z, n), max_node_id=int(z[n + "/max_node_id"]),
                      cutoff=int(z[n + "/cutoff"]), queries=z[n + "/queries"],
                      ref_counts=z[n + "/ref_counts"], ref_member=z[n + "/ref_member"])
    return out


@pytest.fixture(scope="session")
def golden_encodings():
    return np.load(os.path.join(GOLDEN, "golden_encodings.npz"))


@pytest.fixture(scope="session")
def golden_hashing():
    return np.load(os.path.join(GOLDEN, "golden_hashing.npz"))
