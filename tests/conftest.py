import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    import torch

    torch.set_default_dtype(torch.float64)


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def tables():
    return load_case("element_tables.npz")


MECH_CASES = ["hexa1", "hexa2", "tetra1", "tetra2", "quad1", "quad2", "tria1", "tria2",
              "hexa1_orphan"]
HEAT_CASES = ["heat_hexa1", "heat_tetra2", "heat_quad1", "heat_quad2"]
