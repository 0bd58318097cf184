import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); deselected on the CPU box")


@pytest.fixture(scope="session")
def golden():
    import torch
    path = os.path.join(ROOT, "tests", "golden", "wan_attention_golden.pt")
    return torch.load(path, map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def small_case(golden):
    """Seed-regenerated inputs of tests/golden/make_golden.py, verified against the stored checksums."""
    from tests.golden.make_golden import input_checksums, small_case as make
    case = make(0)
    got = input_checksums(case)
    for k, v in golden["input_checksums"].items():
        assert abs(got[k] - v) <= 1e-9 * max(1.0, abs(v)), f"RNG drift in golden input {k}"
    return case
