"""pytest configuration: `gpu` marker, import paths, shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "gevolution-1.2_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ora():
    """The plain-C restatement (always buildable: gcc only)."""
    import oracle
    oracle.build()
    return oracle.load_ora()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref); skipped where it was never built."""
    import oracle
    r = oracle.load_ref()
    if r is None:
        pytest.skip("oracle/_ref/libgevref.so not present (reference tree not available at build time)")
    return r


@pytest.fixture(scope="session")
def checker(ora):
    """Preferred CPU checker for GPU parity tests: compiled reference if present, else the C restatement."""
    import oracle
    return oracle.load_ref() or ora


@pytest.fixture(scope="session")
def gevb():
    import gevb as g
    g.lib()
    return g


@pytest.fixture(scope="session")
def gevb_host():
    """libgevb.so for its host-only entry points (settings reader, IC generator pieces, file writers): loads without a GPU"""
    import gevb as g
    g.lib()
    return g


@pytest.fixture()
def ctx(gevb, request):
    """A single-rank device context for lattice size given by the `ngrid` marker/param (default 16)."""
    made = []

    def make(N=16):
        c = gevb.Context(N, device=0)
        made.append(c)
        return c
    yield make
    for c in made:
        c.close()
