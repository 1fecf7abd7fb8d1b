import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def demo_corridor():
    """Inputs of the reference's formulation demo (faster/other/gurobi_continuous.cpp), see tests/golden/."""
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "corridor_continuous.json")))
    fx["polys"] = [(np.array(p["A"]), np.array(p["b"])) for p in fx["polys"]]
    return fx


@pytest.fixture(scope="session")
def built_lib():
    from faster_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def solver(built_lib):
    from faster_b200 import capi
    s = capi.Solver(0)
    yield s
    s.close()
