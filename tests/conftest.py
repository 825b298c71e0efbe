import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = ["ref_test", "syn12", "syn100", "cat40", "syn300"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


class Golden:
    """One golden case: inputs + what the UNMODIFIED reference produced (oracle/gen_golden.py)."""

    def __init__(self, name):
        self.name = name
        base = name[:-4] if name.endswith("_ccc") else name
        self.tree = os.path.join(GOLD, base + ".nwk")
        self.pa = os.path.join(GOLD, base + ".pa.txt")
        with open(os.path.join(GOLD, name + ".json")) as f:
            self.meta = json.load(f)
        self.z = np.load(os.path.join(GOLD, name + ".npz"))
        self.results = self.meta["results"]
        self.params = np.array([r["params"] for r in self.results])

    def ref3(self, k):
        """reference components regrouped as assignCat does: lik0r, LSE(lik1r, lik1), lik2"""
        c = self.z["components"][k]
        with np.errstate(invalid="ignore"):
            return np.stack([c[:, 0], np.logaddexp(c[:, 1], c[:, 2]), c[:, 3]], axis=1)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]

    return get


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)
