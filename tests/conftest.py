import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = ["small64", "llama128", "ties128", "half64"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """One tests/golden/<name>.npz produced by make_golden.py from the imported reference."""

    def __init__(self, name: str):
        self.name = name
        self.z = np.load(GOLDEN_DIR / f"{name}.npz")
        self.B, self.Hq, self.Hkv, self.D, self.hidden, self.S, self.seed = (int(x) for x in self.z["meta"])
        self.dtype = torch.float16 if name.startswith("half") else torch.bfloat16
        self.ratios = [float(r) for r in self.z["ratios"]]

    def t(self, key: str) -> torch.Tensor:
        a = self.z[key]
        if a.dtype == np.uint16:
            return torch.from_numpy(a.copy()).view(self.dtype)
        return torch.from_numpy(a.copy())

    def __contains__(self, key):
        return key in self.z


@pytest.fixture(params=GOLDEN_CASES)
def golden(request) -> Golden:
    return Golden(request.param)


def ulp16_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Distance in units-in-the-last-place between two 16-bit float tensors of the same dtype."""
    def ordered(x):
        bits = x.contiguous().view(torch.int16).to(torch.int32) & 0xFFFF
        neg = bits >= 0x8000
        return torch.where(neg, 0x8000 - (bits & 0x7FFF), bits + 0x8000)
    return (ordered(a) - ordered(b)).abs()
