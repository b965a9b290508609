import importlib.util
import lzma
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def B():
    """ctypes binding of the product library (virgo-plus_b200/binding.py)."""
    return _load("vp_binding", os.path.join(ROOT, "virgo-plus_b200", "binding.py"))


@pytest.fixture(scope="session")
def O():
    """the CPU oracle (test infrastructure)."""
    return _load("gkr_oracle", os.path.join(ROOT, "oracle", "oracle.py"))


@pytest.fixture(scope="session")
def sha_pws_text():
    with lzma.open(os.path.join(GOLDEN, "SHA256_64.pws.xz"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def sha_circuit(B, sha_pws_text):
    return B.Circuit.from_pws_text(sha_pws_text)
