import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def acg():
    import arithmetic_circuits_b200 as m
    m._lib.lib()  # fail loudly if libacg.so is missing
    return m


@pytest.fixture(scope="session")
def _ctx_bn(acg):
    c = acg.Context(acg.BN254_FR, 0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def _ctx_bls(acg):
    c = acg.Context(acg.BLS12_381_FR, 0)
    yield c
    c.close()


@pytest.fixture(params=[0, 1, 2, 3, 4, 5, 6, 7, 8], ids=["tile128", "tile256", "tile64", "tile32", "tile128nat", "tile64nat", "tile128ip", "tile128ip1far", "tile128dense"])
def tile_variant(request):
    """Every GPU test runs under both tile geometries of the tiled kernel (bound at upload time)."""
    return request.param


@pytest.fixture
def ctx_bn(_ctx_bn, tile_variant):
    _ctx_bn.set_tiled_variant(tile_variant)
    return _ctx_bn


@pytest.fixture
def ctx_bls(_ctx_bls, tile_variant):
    _ctx_bls.set_tiled_variant(tile_variant)
    return _ctx_bls


@pytest.fixture
def ctxs(ctx_bn, ctx_bls):
    return {0: ctx_bn, 1: ctx_bls}
