import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def key_shapes():
    with open(os.path.join(GOLD, "state_dict_keys.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_hgqa():
    return dict(np.load(os.path.join(GOLD, "hgqa_b2.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def oracle_weights(key_shapes):
    from oracle import shg_oracle as so
    return so.make_state_dict(key_shapes, seed=0)


@pytest.fixture(scope="session")
def inputs_b2():
    from shg_vqa_b200.synthetic import make_inputs
    return make_inputs(2, seed=1234)
