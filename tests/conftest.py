import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return torch.load(os.path.join(GOLDEN, name), map_location="cpu")
    return load


def relerr(a, b, name=None):
    """The parity metric of SURVEY.md 8d: max|a-b| / max|b|.  Named measurements are appended to
    gpurun_out/parity.jsonl so the achieved margins can be quoted."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    e = ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
    if name is not None:
        try:
            import json
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "parity.jsonl"), "a") as f:
                f.write(json.dumps({"name": name, "relerr": e}) + "\n")
        except OSError:
            pass
    return e
