"""Test configuration: markers, import paths and shared fixtures.

``-m "not gpu"`` tests run on CPU (oracle vs golden vectors, host logic, C-ABI symbol table); ``-m gpu`` tests are the
parity tests proper and call the CUDA path through the C ABI.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "chessvision-3lc_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = ROOT / "tests" / "golden"
WEIGHTS = ROOT / "weights"
REFERENCE = Path(os.environ.get("CV_REFERENCE", "/root/reference"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout (only in the build container)")


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (the CUDA path has no fallback)")
    from chessvision import _native
    eng = _native.Engine(0, max_batch=8)
    yield eng
    eng.close()


def load_checkpoint(path):
    import torch
    blob = torch.load(path, map_location="cpu")
    for key in ("model_state_dict", "state_dict", "model"):
        if isinstance(blob, dict) and key in blob:
            blob = blob[key]
            break
    return {k: (v.float() if v.is_floating_point() else v) for k, v in blob.items()}
