"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI symbol checks (CPU only).
`-m gpu`       : parity tests proper -- CUDA path through the C-ABI vs oracle + goldens.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "quantik-core-py_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# a fresh checkout has no built artefacts (they are git-ignored): build them once
if not os.path.exists(os.path.join(PKG, "quantik_b200", "_lib", "libquantik_b200.so")):
    import __graft_entry__
    __graft_entry__.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on CPU")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
