import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # build the in-tree artefacts once (no-op when up to date; nvcc cross-compiles without a GPU)
    from vk_gaussian_splatting_b200 import build as B
    try:
        B.build_all()
    except Exception as e:  # pragma: no cover - surfaced by the tests that need the libraries
        print(f"[conftest] build failed: {e}", file=sys.stderr)


@pytest.fixture(scope="session")
def gpu_renderer():
    import vk_gaussian_splatting_b200 as g
    r = g.GaussianSplatting(0)
    yield r
    r.close()
