import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `pytest -m gpu`)")


def _ensure_built():
    # the product library and the C oracle are built in-tree; building is not using
    import importlib.util
    spec = importlib.util.spec_from_file_location("drb_build", os.path.join(ROOT, "dogeray_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from oracle import restated
    restated.build()


_ensure_built()

import dogeray_b200 as drb  # noqa: E402
from oracle import refhost, restated  # noqa: E402

SAMPLES = refhost.SAMPLES
HAVE_REF = refhost.available() and os.path.isdir(SAMPLES)
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built (needs /root/reference once: python oracle/make_ref.py)")


@pytest.fixture(scope="session")
def ref():
    if not HAVE_REF:
        pytest.skip("oracle/_ref not available")
    return refhost.RefHost()


@pytest.fixture(scope="session")
def have_gpu():
    return drb.device_count() > 0


def sample(name):
    return os.path.join(SAMPLES, name)


def all_sample_scenes():
    if not HAVE_REF:
        return []
    return sorted(n for n in os.listdir(SAMPLES) if n.endswith(".rts"))
