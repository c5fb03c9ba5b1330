import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden_names():
    g = os.path.join(ROOT, "tests", "golden")
    return sorted(f[:-4] for f in os.listdir(g) if f.endswith(".npz"))


def load_golden(name):
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["reads"] = str(d["reads"]).split("\n")
    d["k"] = int(d["k"])
    d["buckets"] = int(d["buckets"])
    d["tip_bound"] = int(d["tip_bound"])
    d["unitigs"] = [str(u) for u in d["unitigs"]]
    d["gfa"] = [str(l) for l in d["gfa"]]
    d["fastg"] = [str(l) for l in d["fastg"]]
    d["gfa_cov"] = [str(l) for l in d["gfa_cov"]]
    d["name"] = name
    return d


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)
