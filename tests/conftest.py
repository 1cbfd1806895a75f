import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def load_maps():
    with open(os.path.join(GOLDEN, "maps.json")) as fh:
        return json.load(fh)


def pairs_to_dict(pairs):
    """maps.json stores dicts as [[key, value], ...] so int keys survive."""
    return {k: v for k, v in pairs}


def golden_lineage(name):
    """-> (branches, time dict, topology, data) with python-native branch names."""
    d = load_npz("lineage_%s.npz" % name)
    branches = [b.item() if hasattr(b, "item") else b for b in d["branches"]]
    time = {b: int(t) for b, t in zip(branches, d["times"])}
    topology = [[p.item(), c.item()] for p, c in d["topology"]]
    return branches, time, topology, d


@pytest.fixture(scope="session")
def has_cuda():
    import torch
    return torch.cuda.is_available()


def writer_inputs():
    """Deterministic inputs of the text writers: tests/golden/make_golden.py feeds them to the reference's
    writers (written_files.json), tests/test_host_maps.py to ours."""
    rng = np.random.RandomState(17)
    X = rng.negative_binomial(0.8, 0.2, size=(6, 5))
    uMs = {"A": rng.normal(size=(3, 5)), "B": rng.normal(size=(2, 5)) * 1e-7}
    H = rng.gamma(0.05, size=(2, 5))
    labs = np.array([0, 1, 2, 3, 3, 4])
    brns = np.array(["A", "A", "A", "B", "B", "B"])
    scal = np.exp(rng.normal(0, 0.7, size=6))
    gscale = np.exp(rng.normal(0.8, 1.0, size=5))
    alpha = np.exp(rng.normal(np.log(0.2), np.log(1.5), size=5))
    beta = np.exp(rng.normal(np.log(2.0), np.log(1.5), size=5)) + 1
    return X, uMs, H, labs, brns, scal, gscale, alpha, beta
