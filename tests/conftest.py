"""pytest configuration: the `gpu` marker and shared fixture loaders."""
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
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_json(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def search_cases():
    return load_json("search_cases.json")


@pytest.fixture(scope="session")
def flag_cases():
    return load_json("flag_cases.json")


@pytest.fixture(scope="session")
def query_weight_cases():
    return load_json("query_weights.json")


@pytest.fixture(scope="session")
def known_answer():
    z = np.load(os.path.join(GOLD, "known_answer.npz"))
    d = {k: z[k] for k in z.files if k != "answers"}
    d["answers"] = json.loads(str(z["answers"]))
    return d
