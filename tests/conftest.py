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


def pytest_sessionstart(session):
    """The built libraries normally travel with the tree; if they are missing (fresh clone) build
    them once so the tests exercise the CUDA path instead of failing on import.  A failed build
    is left for the tests to report: nothing falls back to the CPU."""
    from multimodal_audio_search_b200 import _native, build
    if os.path.exists(_native.LIB_PATH) and os.path.exists(_native.TORCH_LIB_PATH):
        return
    try:
        build.build()
        build.build_torch_extension()
    except Exception as e:        # pragma: no cover
        print(f"[conftest] building the CUDA libraries failed: {e}", file=sys.stderr)


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
