"""pytest configuration: registers the `gpu` marker and puts the repo root / tests on sys.path.

`-m "not gpu"`: oracle known-answers, host logic, C-ABI symbol checks (no CUDA calls).
`-m gpu`: parity of the CUDA path (through the C ABI) against the oracle.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def sim():
    """The product package (ctypes binding of libsslsim_b200.so).  The library is normally built by
    __graft_entry__.build(); if it is missing (fresh checkout) it is compiled here -- building is not a fallback,
    the calls still go to the CUDA library only."""
    mod = importlib.import_module("ssl-sim_b200")
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def orc():
    import oracle_binding
    oracle_binding.build()
    return oracle_binding
