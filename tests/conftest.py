"""Shared fixtures.  `-m "not gpu"` = oracle vs golden vectors, host logic, C-ABI exports (runs anywhere);
`-m gpu` = parity of the CUDA backend against the oracle, through the C-ABI (needs a B200)."""
from __future__ import annotations

import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    """Everything `__graft_entry__.build()` produces (idempotent: make decides what is stale)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle_port(built):
    from oracle import port
    port.lib()
    return port


@pytest.fixture(scope="session")
def ctx(built):
    """One CUDA context for the whole GPU session."""
    from simplepath_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def golden_names():
    return sorted(p.name[:-len(".flat.npz")] for p in GOLDEN.glob("*.flat.npz"))
