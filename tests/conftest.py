import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand; prebuilt .so files are reused."""
    from oracle import pyoracle

    pyoracle.build()
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ddlo_lib():
    """The product's C-ABI library, built with nvcc if it is not there yet."""
    from dynamic_direct_lidar_odometry_b200 import binding, build

    build.build_library()
    return binding.load()


@pytest.fixture(scope="session")
def rt(ddlo_lib):
    from dynamic_direct_lidar_odometry_b200 import nano_gicp

    r = nano_gicp.Runtime(0)
    yield r
    r.close()
